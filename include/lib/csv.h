/*
 * csv.h -- drop-in replacement for the reference's lib/csv.h: the checkpoint / data text format either
 * side of every train / run.  Same four functions, same tokenisation and byte-identical output (csrc/csv_codec.cu); a
 * relinked program may keep the reference's own csv.o (its definitions win on the link line) or drop it.
 */
#ifndef __csv_h__
#define __csv_h__

#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* lib/csv.c:18-25   every value of the file, malloc'd (free() it, or hand it to make_matrix) */
float* read_csv_contents(const char* filepath);
/* lib/csv.c:28-54   as above from an open stream, which is closed; *num_values = number of commas */
float* read_csv_contents_file(FILE* f, int* num_values);
/* lib/csv.c:56-67   "%f," per value, a newline after every `cols` values */
void write_csv_contents(const char* filepath, float* data, int cols, int rows);
/* lib/csv.c:70-89   number of '\n' from the current position, -1 on a read error */
int count_num_lines(FILE* f);

#ifdef __cplusplus
}
#endif
#endif
