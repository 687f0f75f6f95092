/*
 * mnist_csv.c -- lib/mnist_csv.h for model/mnist_hinge.c: one CSV row per call.  Plain C, host only, built into
 * libbla_mnist_csv.so (`make -C big-linear-algebra_b200`), which goes on the link line before libbla.so: its `struct MnistCSV`
 * and `visualize_digit_data` are not the ones of lib/mnist_csv2.h (csrc/host_io.cu).
 *
 * The reference (lib/mnist_csv.c:6-30) reads with fgetc into a 4-byte token buffer: a value ends at ',' or at a '\n' that
 * follows at least one character, any other '\n' is skipped, and the row is complete after 785 values.  Same rule here on a
 * whole line at a time; a token longer than the reference's buffer (which it would overrun) is cut to what atof needs, and a
 * stream that ends inside a row (where the reference never returns) ends the row with 1.
 */
#include <stdio.h>
#include <stdlib.h>

#include "../../include/lib/mnist_csv.h"

#define ROW_VALUES 785

int get_next_data(struct MnistCSV* csv) {
	if (feof(csv->file)) {
		printf("CSV file is empty\n");
		return 1;
	}
	char token[64];
	int len = 0, have = 0;
	while (have < ROW_VALUES) {
		const int c = getc_unlocked(csv->file);
		if (c == EOF) return 1;
		if (c == ',' || (c == '\n' && len != 0)) {
			token[len] = '\0';
			csv->buffer[have++] = (float)atof(token);
			len = 0;
		} else if (c != '\n' && len < (int)sizeof(token) - 1) {
			token[len++] = (char)c;
		}
	}
	return 0;
}

void visualize_digit_data(struct MnistCSV* csv) {
	static const char rule[] = "============================\n";
	const float* px = csv->buffer + 1;
	fputs(rule, stdout);
	printf("Data for digit %.f:\n", csv->buffer[0]);
	for (int i = 0; i < 28; i++) {
		char line[30];
		for (int j = 0; j < 28; j++) {
			const float v = px[i * 28 + j];
			line[j] = v < 0.32 ? ' ' : v < 0.6 ? ':' : '#';
		}
		line[28] = '\n';
		line[29] = '\0';
		fputs(line, stdout);
	}
	fputs(rule, stdout);
}
