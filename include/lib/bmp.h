/*
 * bmp.h -- drop-in replacement for the reference's lib/bmp.h: the 24-bit BMP preview writer of model/cifar_unet.c
 * (csrc/host_io.cu).
 */
#ifndef __bmp_h__
#define __bmp_h__

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/bmp.h:6-12   three width x height planes, rows in file order (BMP stores the bottom row first) */
typedef struct BMPData {
	unsigned int width;
	unsigned int height;
	uint8_t* red;
	uint8_t* green;
	uint8_t* blue;
} BMPData;

/* lib/bmp.c:11-100 */
void write_bmp_data(const char* filepath, BMPData* data);

#ifdef __cplusplus
}
#endif
#endif
