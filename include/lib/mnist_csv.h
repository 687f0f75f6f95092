/*
 * mnist_csv.h -- drop-in replacement for the reference's lib/mnist_csv.h: the row-at-a-time MNIST CSV reader of
 * model/mnist_hinge.c (host_io/mnist_csv.c -> libbla_mnist_csv.so, linked BEFORE libbla.so).  It shares
 * its include guard, `struct MnistCSV` and `visualize_digit_data` names with lib/mnist_csv2.h, as in the reference: a
 * program uses one of the two.
 */
#ifndef __mnist_csv_h__
#define __mnist_csv_h__
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/mnist_csv.h:6-10   buffer holds one row: the label and 784 pixel values */
typedef struct MnistCSV {
	FILE* file;
	float* buffer;
	int num_lines;
} MnistCSV;

/* lib/mnist_csv.c:6-30   next row into csv->buffer; 1 (and "CSV file is empty" on stdout) once the stream is at its end */
int get_next_data(struct MnistCSV* csv);
/* lib/mnist_csv.c:32-47  28 x 28 characters on stdout from csv->buffer (thresholds 0.32 / 0.6) */
void visualize_digit_data(struct MnistCSV* csv);

#ifdef __cplusplus
}
#endif
#endif
