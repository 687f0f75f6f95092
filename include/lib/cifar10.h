/*
 * cifar10.h -- drop-in replacement for the reference's lib/cifar10.h: the CIFAR-10 binary batch layout and the per-example
 * reader of model/cifar_unet.c (csrc/host_io.cu).  The device-resident pipeline (bla_cifar_* in bla.h)
 * keeps the same draw and the same row flip.
 */
#ifndef __cifar10_h__
#define __cifar10_h__
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/cifar10.c:6-11 */
extern const unsigned int CIFAR10_NUM_EXAMPLES_PER_FILE;   /* 10000 records */
extern const unsigned int CIFAR10_LINE_LENGTH;             /* 3073 = label byte + pixels */
extern const unsigned int CIFAR10_DATA_LENGTH;             /* 3072 pixel bytes */
extern const unsigned int CIFAR10_BATCH_FILE_SIZE;         /* 30730000 */
extern const unsigned int CIFAR10_NUM_PIXELS;              /* 1024 per colour plane */
extern const unsigned int CIFAR10_EXAMPLE_DIM;             /* 32 */

/* lib/cifar10.c:13-31   arr[3072] = one uniformly drawn record of the open batch file: red, green, blue planes, rows bottom-up */
void fill_random_data(int fd, uint8_t* arr);

#ifdef __cplusplus
}
#endif
#endif
