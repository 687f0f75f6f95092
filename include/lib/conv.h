/*
 * conv.h -- drop-in replacement for the reference's lib/conv.h: SAME-padded conv2d forward and
 * backward on arrays of channel planes.  On the device the convolution is an implicit GEMM
 * (no im2col round trip through HBM); the caller's ConvData scratch matrices are still filled
 * because model code reads them back (lib/conv.c:221-227).
 *
 * reshape_channels_matrix / reshape_matrix_channels do what their NAMES and every call site say
 * (the reference's two bodies are swapped, lib/conv.c:174-203).
 */
#ifndef __conv_h__
#define __conv_h__

#include "matrix.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lib/conv.h:6-11: caller-allocated scratch */
typedef struct ConvData {
	Matrix* im2col;        /* [Ho*Wo x k*k*Cin] */
	Matrix* kernel_matrix; /* [k*k*Cin x Cout]  */
	Matrix* product;       /* [Ho*Wo x Cout]    */
	Matrix* output;        /* Cout planes of Ho x Wo */
} ConvData;

/* lib/conv.c:205-212 */
void conv(Matrix* X, Matrix** kernels, ConvData* data, int in_channels, int out_channels, int stride);
/* lib/conv.c:174-187 (intended direction): channels (C,H,W) -> matrix (H*W, C) */
void reshape_channels_matrix(Matrix* channels, Matrix* matrix);
/* lib/conv.c:190-203 (intended direction): matrix (H*W, C) -> channels (C,H,W) */
void reshape_matrix_channels(Matrix* matrix, Matrix* channels);
/* lib/conv.c:214-229 */
void conv_ddx(Matrix* del_Y, ConvData* data, ConvData* grad_data, Matrix** del_kernels, Matrix* del_input,
              int in_channels, int stride);

/* non-static helpers of lib/conv.c (:8, :80, :138, :156), exported for link compatibility */
void _im2col(Matrix* in, Matrix* out, int kernel_size, int in_channels, int stride);
void _col2im(Matrix* in, Matrix* out, int kernel_size, int out_channels, int stride);
void _reshape_kernels_matrix(Matrix** kernels, Matrix* matrix);
void _reshape_matrix_kernels(Matrix* matrix, Matrix** kernels);

#ifdef __cplusplus
}
#endif
#endif
