/*
 * mnist_csv2.h -- drop-in replacement for the reference's lib/mnist_csv2.h: the MNIST CSV held in host memory and the draws
 * model/mnist_nn.c makes from it (csrc/host_io.cu).  Struct layouts are the reference's (model code
 * builds MnistCSV by initialiser, resets `sampled` / `num_sampled` itself and frees X, y and sampled).  The reference gives
 * this header the include guard of lib/mnist_csv.h, whose `struct MnistCSV` and `visualize_digit_data` differ: a program
 * uses one of the two, never both.
 */
#ifndef __mnist_csv_h__
#define __mnist_csv_h__
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/mnist_csv2.h:5-12 */
typedef struct MnistCSV {
	FILE* file;        /* open CSV of rows `label,p0,...,p783,`; consumed and closed by mnist_csv_init */
	float* X;          /* pixels, feature-major: X[example + feature * num_examples] */
	float* y;          /* labels */
	int num_examples;
	int num_sampled;   /* draws since the flags were last cleared */
	char* sampled;     /* one flag per example */
} MnistCSV;

/* lib/mnist_csv2.h:14-18   X points at the example's first pixel; pixel p is X[p * num_examples] */
typedef struct MnistExample {
	float* X;
	float y;
	int num_examples;
} MnistExample;

/* lib/mnist_csv2.c:13-34   fills every field from csv->file (parallel parse + blocked transpose) */
void mnist_csv_init(MnistCSV* csv);
/* lib/mnist_csv2.c:36-39   one uniform draw, with replacement */
MnistExample get_random_data_replace(MnistCSV* csv);
/* lib/mnist_csv2.c:41-62   one draw without replacement: same rand() stream and index rule, found on a Fenwick tree */
MnistExample get_random_data_take(MnistCSV* csv);
/* lib/mnist_csv2.c:64-79   28 x 28 characters on stdout */
void visualize_digit_data(MnistExample ex);

#ifdef __cplusplus
}
#endif
#endif
