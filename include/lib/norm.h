/*
 * norm.h -- drop-in replacement for the reference's lib/norm.h (group normalisation over arrays
 * of channel planes).  `group_size` is the number of CHANNELS per group.  With BLA_QUIRKS=1
 * (default) the result is the reference's: the value stored in `stdevs` is the variance and the
 * output is (x - mean) / variance, because lib/norm.c:3 declares epsilon as `const int`
 * (= 0) and never takes the square root (lib/norm.c:36-44).
 * Note the argument order: (stdevs, means) here, (means, stdevs) in group_norm_ddx.
 */
#ifndef __norm_h__
#define __norm_h__

#include "matrix.h"

#ifdef __cplusplus
extern "C" {
#endif

/* lib/norm.c:5-50 */
void group_norm(Matrix* in, Matrix* out, matrix_float_t* stdevs, matrix_float_t* means, int channels, int group_size);
/* lib/norm.c:52-93 */
void group_norm_ddx(Matrix* source, Matrix* dest, Matrix* data, matrix_float_t* means, matrix_float_t* stdevs,
                    int channels, int group_size);

/* lib/norm.c:3 (data symbol, exported by the reference object) */
extern const int epsilon;

#ifdef __cplusplus
}
#endif
#endif
