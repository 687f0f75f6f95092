/*
 * matrix.h -- drop-in replacement for the reference's lib/matrix.h (damians13/big-linear-algebra).
 *
 * Same type, same 19 function signatures, so the model programs relink unchanged against libbla.so.
 * Behind every function is a hand-written sm_100a CUDA kernel (big-linear-algebra_b200/csrc);
 * there is no CPU fallback: the first compute call prints an error and exit(1)s if no CUDA
 * device is usable.
 *
 * Element type: the reference HEAD says `typedef double matrix_float_t` (lib/matrix.h:4) but its
 * float-era sources (lib/layer.c, main.c, model/my_first_model.c, model/mnist_hinge.c) only compile
 * meaningfully with float; float is also the precision the B200 path
 * computes in (FP32 SIMT and 3xTF32 tensor path).  Parity is measured against the reference's
 * double build within the stated tolerances.
 *
 * `data` may point to (a) ordinary host memory (malloc, stack, CSV loader output): the call stages
 * it through HBM and is synchronous; (b) memory returned by this library to a host caller
 * (CUDA managed memory: host-dereferenceable, HBM-resident while kernels use it); (c) device
 * memory from bla_matrix_device()/cudaMalloc/torch: the call is asynchronous on the library
 * stream and results are device-resident (see include/bla.h).
 */
#ifndef __matrix_h__
#define __matrix_h__

#ifdef __cplusplus
extern "C" {
#endif

typedef float matrix_float_t;

/* Row-major dense matrix; replaces lib/matrix.h:6-11 */
typedef struct Matrix {
	int rows;
	int cols;
	matrix_float_t* data;
} Matrix;

/* lib/matrix.c:6-12   wraps (adopts) the caller's pointer; the struct itself is malloc'd */
struct Matrix* make_matrix(int rows, int cols, matrix_float_t* data);
/* lib/matrix.c:14-21  deep copy */
struct Matrix* clone_matrix(struct Matrix m);
/* lib/matrix.c:24-26  releases m->data (library memory or plain malloc memory) */
void free_matrix_data(struct Matrix* m);
/* lib/matrix.c:29-32  releases m->data and the struct */
void free_matrix(struct Matrix* m);
/* lib/matrix.c:35-44  a[m x n] . b[n x p]; prints and exit(1)s on a dimension mismatch */
struct Matrix* matrix_multiply(struct Matrix a, struct Matrix b);
/* lib/matrix.c:59-63  m *= f */
void matrix_scale(struct Matrix* m, matrix_float_t f);
/* lib/matrix.c:65-69  a += b over a's extent (no shape check) */
void matrix_add(struct Matrix* a, struct Matrix* b);
/* lib/matrix.c:71-89  text dump (synchronises) */
void print_matrix(struct Matrix m);
/* lib/matrix.c:91-93 */
void print_matrix_dim(struct Matrix m);
/* lib/matrix.c:95-103 a *= b; prints and exit(1)s on a shape mismatch */
void matrix_multiply_elementwise(struct Matrix* a, struct Matrix* b);
/* lib/matrix.c:105-118 transposed values land in the same buffer; rows/cols are swapped */
void matrix_transpose(struct Matrix* m);
/* lib/matrix.c:123-133 1 x cols column totals */
struct Matrix* matrix_row_sum(struct Matrix m);
/* lib/matrix.c:138-148 rows x 1; reproduces the reference's stride quirk unless BLA_QUIRKS=0 */
struct Matrix* matrix_col_sum(struct Matrix m);
/* lib/matrix.c:150-158 */
matrix_float_t frobenius_norm(struct Matrix m);
/* lib/matrix.c:160-168 */
matrix_float_t max_value(struct Matrix m);
/* lib/matrix.c:170-185 */
void matrix_z_score_normalize(Matrix* m);
/* lib/matrix.c:189-195 a[r][c] += b[r][c % b.cols] */
void matrix_add_tile_columns(struct Matrix* a, struct Matrix* b);
/* lib/matrix.c:199-205 a[r][c] += b[0][c] */
void matrix_add_tile_rows(struct Matrix* a, struct Matrix* b);
/* lib/matrix.c:47-57  c = a . b into caller storage; no dimension check, c must not alias */
void matrix_multiply_inplace(Matrix* a, Matrix* b, Matrix* c);

#ifdef __cplusplus
}
#endif
#endif
