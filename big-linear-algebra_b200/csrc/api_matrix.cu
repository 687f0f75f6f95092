// api_matrix.cu -- the reference's lib/matrix.h and lib/util.h entry points (include/lib/*.h) on top
// of the sm_100a kernels.  Each function cites the reference code whose behaviour it keeps.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/lib/util.h"
#include "kernels.h"
#include "runtime.h"

using namespace bla;

namespace {

inline size_t elems(const Matrix& m) { return (size_t)m.rows * (size_t)m.cols; }

Matrix* new_struct(int rows, int cols, float* data) {
    // callers free() these structs with libc free (model/mnist_hinge.c:75, lib/layer.c:13-14)
    Matrix* p = (Matrix*)malloc(sizeof(Matrix));
    p->rows = rows;
    p->cols = cols;
    p->data = data;
    return p;
}

}  // namespace

extern "C" {

// lib/matrix.c:6-12
struct Matrix* make_matrix(int rows, int cols, matrix_float_t* data) { return new_struct(rows, cols, data); }

// lib/matrix.c:14-21
struct Matrix* clone_matrix(struct Matrix m) {
    CallScope sc;
    const size_t n = elems(m);
    const float* src = sc.in(m.data, n);
    float* dst = sc.new_result(n);
    k_copy(dst, src, n, sc.stream());
    return new_struct(m.rows, m.cols, dst);
}

// lib/matrix.c:24-26: the pointer may be ours (pooled) or anybody's malloc
void free_matrix_data(struct Matrix* m) {
    if (!m->data) return;
    if (!pool_free(m->data)) free(m->data);
}

// lib/matrix.c:29-32
void free_matrix(struct Matrix* m) {
    free_matrix_data(m);
    free(m);
}

// lib/matrix.c:47-57
void matrix_multiply_inplace(Matrix* a, Matrix* b, Matrix* c) {
    CallScope sc;
    GemmArgs g{};
    g.m = a->rows; g.k = a->cols; g.n = b->cols;
    g.a = sc.in(a->data, elems(*a)); g.lda = a->cols;
    g.b = sc.in(b->data, elems(*b)); g.ldb = b->cols;
    g.c = sc.out(c->data, (size_t)g.m * g.n); g.ldc = g.n;
    gemm(g, sc.stream());
}

// lib/matrix.c:35-44
struct Matrix* matrix_multiply(struct Matrix a, struct Matrix b) {
    if (a.cols != b.rows) {
        printf("Attempted to multiply %dx%d matrix by %dx%d matrix, exiting\n", a.rows, a.cols, b.rows, b.cols);
        exit(1);
    }
    CallScope sc;
    GemmArgs g{};
    g.m = a.rows; g.k = a.cols; g.n = b.cols;
    g.a = sc.in(a.data, elems(a)); g.lda = a.cols;
    g.b = sc.in(b.data, elems(b)); g.ldb = b.cols;
    g.c = sc.new_result((size_t)g.m * g.n); g.ldc = g.n;
    gemm(g, sc.stream());
    return new_struct(a.rows, b.cols, g.c);
}

// lib/matrix.c:59-63
void matrix_scale(struct Matrix* m, matrix_float_t f) {
    CallScope sc;
    const size_t n = elems(*m);
    k_scale(sc.inout(m->data, n), n, f, sc.stream());
}

// lib/matrix.c:65-69 (iterates a's extent, no shape check)
void matrix_add(struct Matrix* a, struct Matrix* b) {
    CallScope sc;
    const size_t n = elems(*a);
    const float* db = sc.in(b->data, n);
    k_add(sc.inout(a->data, n), db, n, sc.stream());
}

// lib/matrix.c:95-103
void matrix_multiply_elementwise(struct Matrix* a, struct Matrix* b) {
    if (a->cols != b->cols || a->rows != b->rows) {
        printf("Attempted to multiply elements of %dx%d matrix by %dx%d matrix, exiting\n", a->rows, a->cols, b->rows, b->cols);
        exit(1);
    }
    CallScope sc;
    const size_t n = elems(*a);
    const float* db = sc.in(b->data, n);
    k_hadamard(sc.inout(a->data, n), db, n, sc.stream());
}

// lib/matrix.c:105-118: same buffer, swapped dims
void matrix_transpose(struct Matrix* m) {
    CallScope sc;
    const size_t n = elems(*m);
    const int rows = m->rows, cols = m->cols;
    const float* src = sc.in(m->data, n);
    float* dst = sc.out(m->data, n);
    if ((const float*)dst == src) {
        // device/managed memory: transpose into scratch, copy back into the caller's buffer
        float* tmp = sc.scratch(n);
        k_transpose(src, tmp, rows, cols, sc.stream());
        k_copy(dst, tmp, n, sc.stream());
    } else {
        k_transpose(src, dst, rows, cols, sc.stream());   // host memory: staging in / staging out
    }
    m->rows = cols;
    m->cols = rows;
}

// lib/matrix.c:123-133
struct Matrix* matrix_row_sum(struct Matrix m) {
    CallScope sc;
    const float* src = sc.in(m.data, elems(m));
    float* out = sc.new_result((size_t)m.cols);
    k_row_sum(src, m.rows, m.cols, out, sc.stream());
    return new_struct(1, m.cols, out);
}

// lib/matrix.c:138-148 with its stride quirk unless quirks are off
struct Matrix* matrix_col_sum(struct Matrix m) {
    CallScope sc;
    const float* src = sc.in(m.data, elems(m));
    float* out = sc.new_result((size_t)m.rows);
    void* work = sc.scratch_bytes(reduce_workspace_bytes());
    const int quirk = rt().quirks;
    if (quirk && m.cols < m.rows) {
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "bla: matrix_col_sum on a %dx%d matrix: the reference reads out of bounds here "
                            "(lib/matrix.c:144); elements past the end count as 0\n", m.rows, m.cols);
        }
    }
    k_col_sum(src, m.rows, m.cols, out, quirk, work, sc.stream());
    return new_struct(m.rows, 1, out);
}

// lib/matrix.c:150-158
matrix_float_t frobenius_norm(struct Matrix m) {
    CallScope sc;
    const size_t n = elems(m);
    const float* src = sc.in(m.data, n);
    char* work = (char*)sc.scratch_bytes(reduce_workspace_bytes() + 16);
    double* dres = (double*)(work + reduce_workspace_bytes());
    k_sum_squares(src, n, dres, work, sc.stream());
    double h = 0;
    BLA_CUDA(cudaMemcpyAsync(&h, dres, sizeof(double), cudaMemcpyDeviceToHost, sc.stream()));
    BLA_CUDA(cudaStreamSynchronize(sc.stream()));
    return (matrix_float_t)sqrt(h);
}

// lib/matrix.c:160-168
matrix_float_t max_value(struct Matrix m) {
    CallScope sc;
    const size_t n = elems(m);
    const float* src = sc.in(m.data, n);
    char* work = (char*)sc.scratch_bytes(reduce_workspace_bytes() + 16);
    float* dres = (float*)(work + reduce_workspace_bytes());
    k_max(src, n, dres, work, sc.stream());
    float h = 0;
    BLA_CUDA(cudaMemcpyAsync(&h, dres, sizeof(float), cudaMemcpyDeviceToHost, sc.stream()));
    BLA_CUDA(cudaStreamSynchronize(sc.stream()));
    return h;
}

// lib/matrix.c:170-185
void matrix_z_score_normalize(Matrix* m) {
    CallScope sc;
    const size_t n = elems(*m);
    float* d = sc.inout(m->data, n);
    void* work = sc.scratch_bytes(reduce_workspace_bytes());
    k_zscore(d, n, work, sc.stream());
}

// lib/matrix.c:189-195
void matrix_add_tile_columns(struct Matrix* a, struct Matrix* b) {
    CallScope sc;
    const float* db = sc.in(b->data, (size_t)a->rows * b->cols);
    float* da = sc.inout(a->data, elems(*a));
    k_add_tile_columns(da, a->rows, a->cols, db, b->cols, sc.stream());
}

// lib/matrix.c:199-205
void matrix_add_tile_rows(struct Matrix* a, struct Matrix* b) {
    CallScope sc;
    const float* db = sc.in(b->data, (size_t)a->cols);
    float* da = sc.inout(a->data, elems(*a));
    k_add_tile_rows(da, a->rows, a->cols, db, sc.stream());
}

// lib/matrix.c:71-89: synchronises, brings the values to the host, same text format
void print_matrix(struct Matrix m) {
    const size_t n = elems(m);
    float* host = (float*)malloc((n ? n : 1) * sizeof(float));
    MemKind k = rt_initialised() ? classify(m.data) : kHost;
    if (k == kHost || k == kPinned) {
        memcpy(host, m.data, n * sizeof(float));
    } else {
        BLA_CUDA(cudaMemcpyAsync(host, m.data, n * sizeof(float), cudaMemcpyDefault, rt().stream));
        BLA_CUDA(cudaStreamSynchronize(rt().stream));
    }
    printf("%d x %d matrix\n", m.rows, m.cols);
    for (size_t i = 0; i < n; i++) {
        if (i % m.cols == 0) printf("[ ");
        if (host[i] == 0) printf("0 ");
        else if (host[i] < 0.01) printf("%.2e ", host[i]);
        else printf("%.2f ", host[i]);
        if ((i + 1) % m.cols == 0) printf("]\n");
    }
    printf("\n");
    free(host);
}

// lib/matrix.c:91-93
void print_matrix_dim(struct Matrix m) { printf("%d x %d matrix\n", m.rows, m.cols); }

// ---- lib/util.h --------------------------------------------------------------------------------

extern const double PI;
const double PI = 3.14159265358979323846;   // lib/util.c:5 (exported data symbol)

// lib/util.c:7-13
void relu(matrix_float_t* data, int num) {
    CallScope sc;
    k_relu(sc.inout(data, (size_t)num), (size_t)num, sc.stream());
}

// lib/util.c:15-34
void softmax(matrix_float_t* data, int rows, int cols) {
    CallScope sc;
    k_softmax_cols(sc.inout(data, (size_t)rows * cols), rows, cols, sc.stream());
}

// lib/util.c:36-55
void softmax_row_wise(matrix_float_t* data, int rows, int cols) {
    CallScope sc;
    k_softmax_rows(sc.inout(data, (size_t)rows * cols), rows, cols, sc.stream());
}

// model/mnist_nn.c:47-51
void bla_relu_ddx(matrix_float_t* data, int num) {
    CallScope sc;
    k_relu_ddx(sc.inout(data, (size_t)num), (size_t)num, sc.stream());
}

// model/cifar_unet.c:241-253
void bla_relu_backward(const matrix_float_t* source, const matrix_float_t* relu_result, matrix_float_t* dest, size_t n) {
    CallScope sc;
    const float* s = sc.in(source, n);
    const float* r = sc.in(relu_result, n);
    k_relu_backward(s, r, sc.out(dest, n), n, sc.stream());
}

// model/mnist_nn.c:234-268
void bla_softmax_xent(const float* logits, const float* expected, int classes, int batch, float* probs, float* grad, float grad_scale,
                      double* stats_device) {
    CallScope sc;
    const size_t n = (size_t)classes * batch;
    const float* dl = sc.in(logits, n);
    const float* de = sc.in(expected, n);
    float* dp = probs ? (probs == logits ? (float*)dl : sc.out(probs, n)) : nullptr;
    float* dg = grad ? (grad == logits ? (float*)dl : sc.out(grad, n)) : nullptr;
    k_softmax_xent(dl, de, classes, batch, dp, dg, grad_scale, stats_device, sc.stream());
}

// ---- host-only helpers of lib/util.c (kept on the host: file I/O and the libc rand() stream) ----

// Private CSV float reader with the reference's conventions (lib/csv.c:7-57): one value per comma
// (rows end with a trailing comma, as write_csv_contents emits them), newlines ignored.
float* bla_read_csv_floats(const char* filepath, int* count) {
    FILE* f = fopen(filepath, "r");
    if (!f) {
        printf("bla: cannot open CSV file %s, exiting\n", filepath);
        exit(1);
    }
    size_t cap = 1024, n = 0;
    float* vals = (float*)malloc(cap * sizeof(float));
    char tok[512];
    int len = 0, ch;
    while ((ch = fgetc(f)) != EOF) {
        if (ch == ',' || (ch == '\n' && len != 0)) {
            tok[len] = '\0';
            if (n == cap) { cap *= 2; vals = (float*)realloc(vals, cap * sizeof(float)); }
            vals[n++] = (float)atof(tok);
            len = 0;
        } else if (ch != '\n' && ch != '\r' && len < 511) {
            tok[len++] = (char)ch;
        }
    }
    fclose(f);
    if (count) *count = (int)n;
    return vals;
}

// lib/util.c:57-65
void load_matrix_from_csv(Matrix* m, const char* filepath, int rows, int cols) {
    int n = 0;
    float* v = bla_read_csv_floats(filepath, &n);
    if (n < rows * cols) {
        printf("bla: CSV file %s holds %d values, %d expected, exiting\n", filepath, n, rows * cols);
        exit(1);
    }
    CallScope sc;
    MemKind k = classify(m->data);
    if (k == kDevice) {
        BLA_CUDA(cudaMemcpyAsync(m->data, v, (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice, sc.stream()));
        BLA_CUDA(cudaStreamSynchronize(sc.stream()));
    } else {
        for (int i = 0; i < rows * cols; i++) m->data[i] = (matrix_float_t)v[i];
    }
    free(v);
    m->rows = rows;
    m->cols = cols;
}

// lib/util.c:68-95: Box-Muller on libc rand(); cached second variate, `seed` unused
double random_gaussian(unsigned int* seed) {
    (void)seed;
    static double spare;
    static int have_spare = 0;
    if (have_spare) {
        have_spare = 0;
        return spare;
    }
    double u1 = (double)rand() / RAND_MAX;
    while (u1 == 0) u1 = (double)rand() / RAND_MAX;
    double u2 = (double)rand() / RAND_MAX;
    double radius = sqrt(-2 * log(u1));
    double angle = 2 * PI * u2;
    spare = radius * sin(angle);
    have_spare = 1;
    return radius * cos(angle);
}

}  // extern "C"
