/*
 * csv.h -- declarations of the reference's CSV codec (lib/csv.h).  The codec is host file I/O and
 * is OUT OF SCOPE of the B200 path: relinked model programs keep using the
 * reference's own csv.o.  libbla.so itself only needs a float reader for
 * load_weights_from_csv / load_matrix_from_csv and carries a private one.
 */
#ifndef __csv_h__
#define __csv_h__

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

float* read_csv_contents(const char* filepath);
float* read_csv_contents_file(FILE* f, int* num_values);
void write_csv_contents(const char* filepath, float* data, int cols, int rows);
int count_num_lines(FILE* f);

#ifdef __cplusplus
}
#endif
#endif
