/*
 * util.h -- drop-in replacement for the reference's lib/util.h: activations (CUDA kernels) and the
 * two host-only helpers (CSV loader, Box-Muller sampler on libc rand()).
 */
#ifndef __util_h__
#define __util_h__

#include "matrix.h"
#include "csv.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lib/util.c:7-13 */
void relu(matrix_float_t* data, int num);
/* lib/util.c:15-34  softmax of every COLUMN of a rows x cols row-major array */
void softmax(matrix_float_t* data, int rows, int cols);
/* lib/util.c:36-55  softmax of every ROW */
void softmax_row_wise(matrix_float_t* data, int rows, int cols);
/* lib/util.c:57-65  host I/O */
void load_matrix_from_csv(Matrix* m, const char* filepath, int rows, int cols);
/* lib/util.c:68-95  host RNG (keeps the libc rand() sequence of the reference) */
double random_gaussian(unsigned int* seed);

#ifdef __cplusplus
}
#endif
#endif
