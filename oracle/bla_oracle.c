/*
 * bla_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the dense hot path of damians13/big-linear-algebra, used as the parity
 * checker for the CUDA product (libbla.so).  It is never linked into, imported by, or executed
 * from the product path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Parity is PINNED: every function here is compared (tests/test_oracle_pinned.py) against the
 * reference's own C compiled in oracle/_ref/ by oracle/build_ref.sh, and against the committed
 * golden vectors in tests/golden/ that were generated from that compiled reference by
 * tests/golden/make_golden.py.
 *
 * All functions work on flat row-major arrays (no struct Matrix) so numpy can call them through
 * ctypes.  The element type is REAL: the file is compiled twice, -DREAL=double (HEAD behaviour,
 * lib/matrix.h:4) and -DREAL=float (the float-era layer.c / model code).
 * Citations are file:line into the reference tree.
 *
 * The arithmetic ORDER of every accumulation follows the reference so that the f64 build agrees
 * with it bit for bit; loop nests are otherwise written independently.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef REAL
#define REAL double
#endif

#define IDX2(r, c, ld) ((size_t)(r) * (size_t)(ld) + (size_t)(c))

/* ------------------------------------------------------------------------------------------ */
/* lib/matrix.c                                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* matrix_multiply_inplace, lib/matrix.c:47-57: c[M x N] = a[M x K] . b[K x N]; every output is
 * one sequential-k sum held in REAL. (The reference iterates columns outermost; the order of
 * the outputs does not matter, the order inside each dot product does.) */
void orc_gemm(int M, int K, int N, const REAL* a, const REAL* b, REAL* c) {
    for (int r = 0; r < M; ++r) {
        const REAL* arow = a + IDX2(r, 0, K);
        for (int col = 0; col < N; ++col) {
            REAL acc = 0;
            for (int k = 0; k < K; ++k) acc += arow[k] * b[IDX2(k, col, N)];
            c[IDX2(r, col, N)] = acc;
        }
    }
}

/* matrix_scale, lib/matrix.c:59-63 */
void orc_scale(size_t n, REAL* m, REAL f) {
    for (size_t i = 0; i < n; ++i) m[i] *= f;
}

/* matrix_add, lib/matrix.c:65-69 (iterates a's extent; no shape check) */
void orc_add(size_t n, REAL* a, const REAL* b) {
    for (size_t i = 0; i < n; ++i) a[i] += b[i];
}

/* matrix_multiply_elementwise, lib/matrix.c:95-103 */
void orc_hadamard(size_t n, REAL* a, const REAL* b) {
    for (size_t i = 0; i < n; ++i) a[i] *= b[i];
}

/* matrix_transpose, lib/matrix.c:105-118: the transposed values land in the SAME buffer, the
 * caller swaps rows/cols.  rows/cols are the dims BEFORE the call. */
void orc_transpose(int rows, int cols, REAL* m) {
    size_t n = (size_t)rows * cols;
    REAL* tmp = (REAL*)malloc(n * sizeof(REAL));
    memcpy(tmp, m, n * sizeof(REAL));
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) m[IDX2(c, r, rows)] = tmp[IDX2(r, c, cols)];
    free(tmp);
}

/* matrix_row_sum, lib/matrix.c:123-133: out[1 x cols], column totals, summed top to bottom */
void orc_row_sum(int rows, int cols, const REAL* m, REAL* out) {
    for (int c = 0; c < cols; ++c) {
        REAL acc = 0;
        for (int r = 0; r < rows; ++r) acc += m[IDX2(r, c, cols)];
        out[c] = acc;
    }
}

/* matrix_col_sum, lib/matrix.c:138-148: row i of the result is the sum of `cols`
 * consecutive FLAT elements starting at i*rows (the stride is rows, not cols, :144).  For
 * cols >= rows every read is in bounds and this is exactly the reference.  For cols < rows the
 * reference reads past the buffer (UB); the contract of this project treats those elements as 0.
 * quirk == 0 gives the mathematically intended row totals instead. */
void orc_col_sum(int rows, int cols, const REAL* m, REAL* out, int quirk) {
    size_t n = (size_t)rows * cols;
    for (int r = 0; r < rows; ++r) {
        REAL acc = 0;
        size_t base = quirk ? (size_t)r * rows : (size_t)r * cols;
        for (int j = 0; j < cols; ++j) {
            size_t at = base + j;
            if (at < n) acc += m[at];
        }
        out[r] = acc;
    }
}

/* frobenius_norm, lib/matrix.c:150-158: the sum runs column by column (column-major order) */
REAL orc_frobenius(int rows, int cols, const REAL* m) {
    REAL acc = 0;
    for (int c = 0; c < cols; ++c)
        for (int r = 0; r < rows; ++r) {
            REAL v = m[IDX2(r, c, cols)];
            acc += v * v;
        }
    return (REAL)sqrt((double)acc);
}

/* max_value, lib/matrix.c:160-168 */
REAL orc_max(size_t n, const REAL* m) {
    REAL best = -INFINITY;
    for (size_t i = 0; i < n; ++i)
        if (m[i] > best) best = m[i];
    return best;
}

/* matrix_z_score_normalize, lib/matrix.c:170-185: one pass sum / sum of squares, the standard
 * deviation goes through sqrtf (single precision) even in the double build (:179, D7). */
void orc_zscore(size_t n, REAL* m) {
    REAL s = 0, ss = 0;
    for (size_t i = 0; i < n; ++i) {
        s += m[i];
        ss += m[i] * m[i];
    }
    REAL mean = s / (REAL)n;
    REAL sd = sqrtf(ss / (REAL)n - mean * mean);
    for (size_t i = 0; i < n; ++i) m[i] = (m[i] - mean) / sd;
}

/* matrix_add_tile_columns, lib/matrix.c:189-195: a[r][c] += b[r][c % bcols] */
void orc_add_tile_columns(int rows, int cols, REAL* a, int bcols, const REAL* b) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) a[IDX2(r, c, cols)] += b[IDX2(r, c % bcols, bcols)];
}

/* matrix_add_tile_rows, lib/matrix.c:199-205: a[r][c] += b[c] */
void orc_add_tile_rows(int rows, int cols, REAL* a, const REAL* b) {
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) a[IDX2(r, c, cols)] += b[c];
}

/* ------------------------------------------------------------------------------------------ */
/* lib/util.c and the model-local activation loops                                             */
/* ------------------------------------------------------------------------------------------ */

/* relu, lib/util.c:7-13 == model/mnist_nn.c:38-44 */
void orc_relu(size_t n, REAL* d) {
    for (size_t i = 0; i < n; ++i)
        if (d[i] < 0) d[i] = 0;
}

/* relu_ddx, model/mnist_nn.c:47-51 */
void orc_relu_ddx(size_t n, REAL* d) {
    for (size_t i = 0; i < n; ++i) d[i] = d[i] > 0 ? 1 : 0;
}

/* softmax over each COLUMN, lib/util.c:15-34 == model/mnist_nn.c:54-73 (max-subtracted, libm
 * double exp, then divide by the running sum accumulated top to bottom) */
void orc_softmax_cols(int rows, int cols, REAL* d) {
    for (int c = 0; c < cols; ++c) {
        REAL mx = -INFINITY;
        for (int r = 0; r < rows; ++r)
            if (d[IDX2(r, c, cols)] > mx) mx = d[IDX2(r, c, cols)];
        REAL tot = 0;
        for (int r = 0; r < rows; ++r) {
            REAL e = (REAL)exp((double)(d[IDX2(r, c, cols)] - mx));
            d[IDX2(r, c, cols)] = e;
            tot += e;
        }
        for (int r = 0; r < rows; ++r) d[IDX2(r, c, cols)] /= tot;
    }
}

/* softmax_row_wise, lib/util.c:36-55 */
void orc_softmax_rows(int rows, int cols, REAL* d) {
    for (int r = 0; r < rows; ++r) {
        REAL* row = d + IDX2(r, 0, cols);
        REAL mx = -INFINITY;
        for (int c = 0; c < cols; ++c)
            if (row[c] > mx) mx = row[c];
        REAL tot = 0;
        for (int c = 0; c < cols; ++c) {
            row[c] = (REAL)exp((double)(row[c] - mx));
            tot += row[c];
        }
        for (int c = 0; c < cols; ++c) row[c] /= tot;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* lib/layer.c -- per-sample dense layer with MSE back-propagation                             */
/* ------------------------------------------------------------------------------------------ */

/* Activation codes shared with the product's fingerprinting (include/bla.h BLA_ACT_*). */
enum { ORC_ACT_IDENTITY = 0, ORC_ACT_RELU = 1, ORC_ACT_SCALE = 2 };

static void orc_act(int code, double p, size_t n, REAL* d) {
    if (code == ORC_ACT_RELU) orc_relu(n, d);
    else if (code == ORC_ACT_SCALE)
        for (size_t i = 0; i < n; ++i) d[i] = (REAL)((double)d[i] * p); /* main.c:7-11: `*= 0.1` is a double multiply */
}

static void orc_act_ddx(int code, double p, size_t n, REAL* d) {
    if (code == ORC_ACT_RELU) orc_relu_ddx(n, d);
    else if (code == ORC_ACT_SCALE)
        for (size_t i = 0; i < n; ++i) d[i] = (REAL)p; /* main.c:13-17 */
    else
        for (size_t i = 0; i < n; ++i) d[i] = 1;
}

/* feed_forward, lib/layer.c:6-20: raw = W.x + b ; nodes = act(raw).  W is [n x n_prev]. */
void orc_dense_forward(int n, int n_prev, const REAL* W, const REAL* b, const REAL* x, int act,
                       double act_param, REAL* raw, REAL* nodes) {
    orc_gemm(n, n_prev, 1, W, x, raw);
    orc_add((size_t)n, raw, b);
    memcpy(nodes, raw, (size_t)n * sizeof(REAL));
    orc_act(act, act_param, (size_t)n, nodes);
}

/* back_propagate_errors + do_back_propagate_errors, lib/layer.c:48-107, for a chain of L dense
 * layers (index 0 is the first layer that has weights, L-1 the output layer).
 *   sizes[0..L]   node counts, sizes[0] = input width
 *   W[l], b[l]    parameters of layer l ([sizes[l+1] x sizes[l]], [sizes[l+1]]) -- updated in place
 *   raw[l], nodes[l]  outputs of orc_dense_forward for layer l;  x = network input
 * Every layer's delta is computed with PRE-update weights: the reference applies the updates on
 * the way out of the recursion (:72-73, :101-102). */
void orc_dense_backprop(int L, const int* sizes, REAL** W, REAL** b, REAL** raw, REAL** nodes,
                        const REAL* x, const int* act, const double* act_param,
                        const float* expectations, float learn_rate) {
    REAL** dW = (REAL**)calloc((size_t)L, sizeof(REAL*));
    REAL** db = (REAL**)calloc((size_t)L, sizeof(REAL*));
    int nout = sizes[L];
    /* dC/da of the output layer, :85-88 */
    REAL* dcda = (REAL*)malloc((size_t)nout * sizeof(REAL));
    for (int i = 0; i < nout; ++i) dcda[i] = 2 * (nodes[L - 1][i] - expectations[i]);
    for (int l = L - 1; l >= 0; --l) {
        int n = sizes[l + 1], np = sizes[l];
        /* bias step = -lr * act'(raw) (.) dC/da, :90-93 / :63-66 */
        db[l] = (REAL*)malloc((size_t)n * sizeof(REAL));
        memcpy(db[l], raw[l], (size_t)n * sizeof(REAL));
        orc_act_ddx(act[l], act_param[l], (size_t)n, db[l]);
        orc_hadamard((size_t)n, db[l], dcda);
        orc_scale((size_t)n, db[l], (REAL)(-learn_rate));
        /* weight step = bias step . prev_nodes^T, :95-97 / :68-70 */
        const REAL* prev = (l == 0) ? x : nodes[l - 1];
        dW[l] = (REAL*)malloc((size_t)n * np * sizeof(REAL));
        orc_gemm(n, 1, np, db[l], prev, dW[l]);
        if (l > 0) {
            /* dC/da of layer l-1 = W_l^T . (act'(raw_l) (.) dC/da_l), :53-59 */
            REAL* g = (REAL*)malloc((size_t)n * sizeof(REAL));
            memcpy(g, raw[l], (size_t)n * sizeof(REAL));
            orc_act_ddx(act[l], act_param[l], (size_t)n, g);
            orc_hadamard((size_t)n, g, dcda);
            REAL* next = (REAL*)malloc((size_t)np * sizeof(REAL));
            for (int j = 0; j < np; ++j) {
                REAL acc = 0;
                for (int i = 0; i < n; ++i) acc += W[l][IDX2(i, j, np)] * g[i];
                next[j] = acc;
            }
            free(g);
            free(dcda);
            dcda = next;
        }
    }
    free(dcda);
    for (int l = 0; l < L; ++l) {
        orc_add((size_t)sizes[l + 1] * sizes[l], W[l], dW[l]);
        orc_add((size_t)sizes[l + 1], b[l], db[l]);
        free(dW[l]);
        free(db[l]);
    }
    free(dW);
    free(db);
}

/* ------------------------------------------------------------------------------------------ */
/* lib/conv.c -- SAME-padded conv2d through im2col                                             */
/* ------------------------------------------------------------------------------------------ */

static void same_pad(int extent, int k, int stride, int* before, int* total) {
    /* lib/conv.c:12-24: pad = max(0, (ceil(extent/stride) - 1)*stride + k - extent);
     * floor(pad/2) goes in front, the rest behind */
    int o = (extent + stride - 1) / stride;
    int p = (o - 1) * stride + k - extent;
    if (p < 0) p = 0;
    *before = p / 2;
    *total = p;
}

/* _im2col, lib/conv.c:8-77.  x is C contiguous planes [C][H][W]; col is [Ho*Wo x C*k*k] with
 * column index c*k*k + ki*k + kj. */
void orc_im2col(int C, int H, int W, int k, int stride, const REAL* x, REAL* col) {
    int pt, pv, pl, ph;
    same_pad(H, k, stride, &pt, &pv);
    same_pad(W, k, stride, &pl, &ph);
    int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    int ck2 = C * k * k;
    for (int oi = 0; oi < Ho; ++oi)
        for (int oj = 0; oj < Wo; ++oj) {
            REAL* row = col + IDX2(oi * Wo + oj, 0, ck2);
            for (int c = 0; c < C; ++c)
                for (int ki = 0; ki < k; ++ki)
                    for (int kj = 0; kj < k; ++kj) {
                        int ii = oi * stride + ki - pt, jj = oj * stride + kj - pl;
                        REAL v = 0;
                        if (ii >= 0 && ii < H && jj >= 0 && jj < W) v = x[((size_t)c * H + ii) * W + jj];
                        row[c * k * k + ki * k + kj] = v;
                    }
        }
}

/* _col2im, lib/conv.c:80-135, the scatter-add adjoint of orc_im2col followed by the crop.
 * The reference iterates over the INPUT extent (:108-110) which equals the output extent only
 * for stride 1; this restatement iterates over the im2col rows (Ho*Wo), which is the
 * same thing for stride 1 and the mathematically correct adjoint for stride 2. */
void orc_col2im(int C, int H, int W, int k, int stride, const REAL* col, REAL* x) {
    int pt, pv, pl, ph;
    same_pad(H, k, stride, &pt, &pv);
    same_pad(W, k, stride, &pl, &ph);
    int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    int ck2 = C * k * k;
    memset(x, 0, (size_t)C * H * W * sizeof(REAL));
    for (int oi = 0; oi < Ho; ++oi)
        for (int oj = 0; oj < Wo; ++oj) {
            const REAL* row = col + IDX2(oi * Wo + oj, 0, ck2);
            for (int c = 0; c < C; ++c)
                for (int ki = 0; ki < k; ++ki)
                    for (int kj = 0; kj < k; ++kj) {
                        int ii = oi * stride + ki - pt, jj = oj * stride + kj - pl;
                        if (ii >= 0 && ii < H && jj >= 0 && jj < W)
                            x[((size_t)c * H + ii) * W + jj] += row[c * k * k + ki * k + kj];
                    }
        }
}

/* conv, lib/conv.c:205-212 with the INTENDED reshape direction (D3): y[F][Ho][Wo].
 * kernels is [F][C][k][k] contiguous.  The GEMM is orc_gemm so the accumulation order over
 * (c, ki, kj) equals the reference's. */
void orc_conv(int C, int H, int W, int F, int k, int stride, const REAL* x, const REAL* kernels,
              REAL* y) {
    int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    int ck2 = C * k * k, P = Ho * Wo;
    REAL* col = (REAL*)malloc((size_t)P * ck2 * sizeof(REAL));
    REAL* km = (REAL*)malloc((size_t)ck2 * F * sizeof(REAL));
    REAL* prod = (REAL*)malloc((size_t)P * F * sizeof(REAL));
    orc_im2col(C, H, W, k, stride, x, col);
    /* _reshape_kernels_matrix, lib/conv.c:138-153: (F,C,k,k) -> (C*k*k, F) */
    for (int f = 0; f < F; ++f)
        for (int q = 0; q < ck2; ++q) km[IDX2(q, f, F)] = kernels[IDX2(f, q, ck2)];
    orc_gemm(P, ck2, F, col, km, prod);
    /* reshape_matrix_channels (intended), lib/conv.c:189-203: (Ho*Wo, F) -> (F, Ho, Wo) */
    for (int f = 0; f < F; ++f)
        for (int p = 0; p < P; ++p) y[IDX2(f, p, P)] = prod[IDX2(p, f, F)];
    free(col);
    free(km);
    free(prod);
}

/* conv_ddx, lib/conv.c:214-229: dK[F][C][k][k] = im2col(x)^T . dQ ; dX = col2im(dQ . Kmat^T),
 * where dQ[(Ho*Wo) x F] is dy reshaped.  Accumulation orders follow orc_gemm on the transposed
 * operands exactly as the reference's materialised transposes do. */
void orc_conv_ddx(int C, int H, int W, int F, int k, int stride, const REAL* x,
                  const REAL* kernels, const REAL* dy, REAL* dkernels, REAL* dx) {
    int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    int ck2 = C * k * k, P = Ho * Wo;
    REAL* col = (REAL*)malloc((size_t)P * ck2 * sizeof(REAL));
    REAL* dq = (REAL*)malloc((size_t)P * F * sizeof(REAL));
    REAL* dcol = (REAL*)malloc((size_t)P * ck2 * sizeof(REAL));
    orc_im2col(C, H, W, k, stride, x, col);
    for (int f = 0; f < F; ++f)
        for (int p = 0; p < P; ++p) dq[IDX2(p, f, F)] = dy[IDX2(f, p, P)];
    /* wgrad: sum over output pixels p in increasing order */
    for (int q = 0; q < ck2; ++q)
        for (int f = 0; f < F; ++f) {
            REAL acc = 0;
            for (int p = 0; p < P; ++p) acc += col[IDX2(p, q, ck2)] * dq[IDX2(p, f, F)];
            dkernels[IDX2(f, q, ck2)] = acc;
        }
    /* dgrad: sum over filters f in increasing order, then scatter-add */
    for (int p = 0; p < P; ++p)
        for (int q = 0; q < ck2; ++q) {
            REAL acc = 0;
            for (int f = 0; f < F; ++f) acc += dq[IDX2(p, f, F)] * kernels[IDX2(f, q, ck2)];
            dcol[IDX2(p, q, ck2)] = acc;
        }
    orc_col2im(C, H, W, k, stride, dcol, dx);
    free(col);
    free(dq);
    free(dcol);
}

/* ------------------------------------------------------------------------------------------ */
/* lib/norm.c -- group normalisation (divides by the VARIANCE, epsilon is int 0)    */
/* ------------------------------------------------------------------------------------------ */

/* group_norm, lib/norm.c:5-50.  x,y are [C][HW]; group_size = channels per group.
 * stdevs[g] receives the variance (no sqrt, :36-37).  quirk == 0 gives textbook
 * (x-mean)/sqrt(var + 1e-8) instead. */
void orc_group_norm(int C, int HW, int group_size, const REAL* x, REAL* y, REAL* stdevs,
                    REAL* means, int quirk) {
    int G = (C + group_size - 1) / group_size;
    for (int g = 0; g < G; ++g) {
        int c0 = g * group_size;
        int nc = C - c0 < group_size ? C - c0 : group_size;
        size_t n = (size_t)nc * HW;
        const REAL* xs = x + (size_t)c0 * HW;
        REAL mean = 0;
        for (size_t i = 0; i < n; ++i) mean += xs[i];
        mean /= (int)n;
        means[g] = mean;
        REAL var = 0;
        for (size_t i = 0; i < n; ++i) {
            REAL d = xs[i] - mean;
            var += d * d;
        }
        var /= (int)n;
        REAL denom = quirk ? var : (REAL)sqrt((double)var + 1e-8);
        stdevs[g] = quirk ? var : denom;
        REAL* ys = y + (size_t)c0 * HW;
        for (size_t i = 0; i < n; ++i) ys[i] = (xs[i] - mean) / denom;
    }
}

/* group_norm_ddx, lib/norm.c:52-93.  dy = upstream gradient ("source"), x = forward input
 * ("data"), dx = "dest".  s = stdevs[g] + 0 (the stored variance under the quirk). */
void orc_group_norm_ddx(int C, int HW, int group_size, const REAL* dy, REAL* dx, const REAL* x,
                        const REAL* means, const REAL* stdevs) {
    int G = (C + group_size - 1) / group_size;
    for (int g = 0; g < G; ++g) {
        int c0 = g * group_size;
        int nc = C - c0 < group_size ? C - c0 : group_size;
        size_t n = (size_t)nc * HW;
        const REAL* xs = x + (size_t)c0 * HW;
        const REAL* gs = dy + (size_t)c0 * HW;
        REAL s = stdevs[g], mu = means[g];
        REAL gsum = 0, gwsum = 0;
        for (size_t i = 0; i < n; ++i) {
            REAL w = (xs[i] - mu) / s;
            gsum += gs[i];
            gwsum += w * gs[i];
        }
        gsum /= (int)n;
        gwsum /= (int)n;
        REAL* ds = dx + (size_t)c0 * HW;
        for (size_t i = 0; i < n; ++i) {
            REAL w = (xs[i] - mu) / s;
            ds[i] = (gs[i] - gsum - w * gwsum) / s;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* model/mnist_nn.c -- one mini-batch SGD step of the 784-256-128-10 MLP                        */
/* ------------------------------------------------------------------------------------------ */

/* Restates model/mnist_nn.c:218-315 for one batch of B samples laid out as columns.
 *   dims = {n0, n1, n2, n3}; W1[n1 x n0] b1[n1] W2[n2 x n1] b2[n2] W3[n3 x n2] b3[n3]
 *   X[n0 x B] are RAW pixel values (the 1/255.0F scaling of :218 is applied here);
 *   Y[n3 x B] one-hot.
 * Outputs: params updated in place, *loss_sum (the quirky cross-entropy of :249-256, which walks
 * the 10 x B matrices as if they were B x 10), *num_correct (:238-247), and -- if probs != NULL --
 * the softmax activations A3[n3 x B].
 * quirk selects the D2 col_sum behaviour for the bias gradients (:271,:282,:293). */
void orc_mlp_step(const int* dims, int B, REAL* W1, REAL* b1, REAL* W2, REAL* b2, REAL* W3,
                  REAL* b3, const REAL* Xraw, const REAL* Y, double lr_mult, int quirk,
                  double* loss_sum, int* num_correct, REAL* probs, int update) {
    int n0 = dims[0], n1 = dims[1], n2 = dims[2], n3 = dims[3];
    size_t sX = (size_t)n0 * B, s1 = (size_t)n1 * B, s2 = (size_t)n2 * B, s3 = (size_t)n3 * B;
    REAL* X = (REAL*)malloc(sX * sizeof(REAL));
    memcpy(X, Xraw, sX * sizeof(REAL));
    orc_scale(sX, X, (REAL)(1 / 255.0F)); /* :218 */
    REAL *Z1 = malloc(s1 * sizeof(REAL)), *A1 = malloc(s1 * sizeof(REAL));
    REAL *Z2 = malloc(s2 * sizeof(REAL)), *A2 = malloc(s2 * sizeof(REAL));
    REAL *Z3 = malloc(s3 * sizeof(REAL)), *A3 = malloc(s3 * sizeof(REAL));
    orc_gemm(n1, n0, B, W1, X, Z1);                 /* :221 */
    orc_add_tile_columns(n1, B, Z1, 1, b1);         /* :222 */
    memcpy(A1, Z1, s1 * sizeof(REAL));
    orc_relu(s1, A1);                               /* :224 */
    orc_gemm(n2, n1, B, W2, A1, Z2);                /* :226-229 */
    orc_add_tile_columns(n2, B, Z2, 1, b2);
    memcpy(A2, Z2, s2 * sizeof(REAL));
    orc_relu(s2, A2);
    orc_gemm(n3, n2, B, W3, A2, Z3);                /* :231-234 */
    orc_add_tile_columns(n3, B, Z3, 1, b3);
    memcpy(A3, Z3, s3 * sizeof(REAL));
    orc_softmax_cols(n3, B, A3);
    if (probs) memcpy(probs, A3, s3 * sizeof(REAL));

    /* :237-257 */
    int correct = 0;
    double loss = 0;
    for (int k = 0; k < B; ++k) {
        int pred = 0;
        REAL best = 0;
        for (int p = 0; p < n3; ++p)
            if (A3[IDX2(p, k, B)] > best) {
                best = A3[IDX2(p, k, B)];
                pred = p;
            }
        if (Y[IDX2(pred, k, B)] == 1) ++correct;
        /* cross_entropy_loss over the FLAT slice [k*n3, (k+1)*n3) of both matrices (:250-252,
         * :83-91); accumulated per sample in REAL, then added to the double batch loss */
        REAL l = 0;
        for (int i = 0; i < n3; ++i) {
            REAL a = A3[(size_t)k * n3 + i];
            REAL v = -1 * (Y[(size_t)k * n3 + i] * log(a + 1e-15));
            l += v;
        }
        loss += l;
    }
    if (loss_sum) *loss_sum = loss;
    if (num_correct) *num_correct = correct;

    if (update) {
        double scale = 1 / (double)n0; /* :260 (LAYER_INPUT_SIZE) */
        /* dZ3 = (A3 - Y) * scale, :263-268 */
        REAL* dZ3 = malloc(s3 * sizeof(REAL));
        for (size_t i = 0; i < s3; ++i) dZ3[i] = (REAL)((A3[i] + (-1.0F) * Y[i]) * (REAL)scale);
        REAL *dW3 = malloc((size_t)n3 * n2 * sizeof(REAL)), *db3 = malloc((size_t)n3 * sizeof(REAL));
        /* dW3 = dZ3 . A2^T : sum over samples in increasing order, :266-270 */
        for (int i = 0; i < n3; ++i)
            for (int j = 0; j < n2; ++j) {
                REAL acc = 0;
                for (int k = 0; k < B; ++k) acc += dZ3[IDX2(i, k, B)] * A2[IDX2(j, k, B)];
                dW3[IDX2(i, j, n2)] = acc;
            }
        orc_col_sum(n3, B, dZ3, db3, quirk); /* :271 */
        /* dA2 = W3^T . dZ3 ; dZ2 = relu'(Z2) (.) dA2, :273-278 */
        REAL* dZ2 = malloc(s2 * sizeof(REAL));
        for (int j = 0; j < n2; ++j)
            for (int k = 0; k < B; ++k) {
                REAL acc = 0;
                for (int i = 0; i < n3; ++i) acc += W3[IDX2(i, j, n2)] * dZ3[IDX2(i, k, B)];
                dZ2[IDX2(j, k, B)] = (Z2[IDX2(j, k, B)] > 0 ? (REAL)1 : (REAL)0) * acc;
            }
        REAL *dW2 = malloc((size_t)n2 * n1 * sizeof(REAL)), *db2 = malloc((size_t)n2 * sizeof(REAL));
        for (int i = 0; i < n2; ++i)
            for (int j = 0; j < n1; ++j) {
                REAL acc = 0;
                for (int k = 0; k < B; ++k) acc += dZ2[IDX2(i, k, B)] * A1[IDX2(j, k, B)];
                dW2[IDX2(i, j, n1)] = acc;
            }
        orc_col_sum(n2, B, dZ2, db2, quirk); /* :282 */
        REAL* dZ1 = malloc(s1 * sizeof(REAL));
        for (int j = 0; j < n1; ++j)
            for (int k = 0; k < B; ++k) {
                REAL acc = 0;
                for (int i = 0; i < n2; ++i) acc += W2[IDX2(i, j, n1)] * dZ2[IDX2(i, k, B)];
                dZ1[IDX2(j, k, B)] = (Z1[IDX2(j, k, B)] > 0 ? (REAL)1 : (REAL)0) * acc;
            }
        REAL *dW1 = malloc((size_t)n1 * n0 * sizeof(REAL)), *db1 = malloc((size_t)n1 * sizeof(REAL));
        for (int i = 0; i < n1; ++i)
            for (int j = 0; j < n0; ++j) {
                REAL acc = 0;
                for (int k = 0; k < B; ++k) acc += dZ1[IDX2(i, k, B)] * X[IDX2(j, k, B)];
                dW1[IDX2(i, j, n0)] = acc;
            }
        orc_col_sum(n1, B, dZ1, db1, quirk); /* :293 */
        /* scale by float(-lr) then add, :303-315 (epoch_learn_rate is a float, :186) */
        REAL lr = (REAL)(float)(-lr_mult);
        orc_scale((size_t)n3 * n2, dW3, lr); orc_scale((size_t)n3, db3, lr);
        orc_scale((size_t)n2 * n1, dW2, lr); orc_scale((size_t)n2, db2, lr);
        orc_scale((size_t)n1 * n0, dW1, lr); orc_scale((size_t)n1, db1, lr);
        orc_add((size_t)n3 * n2, W3, dW3); orc_add((size_t)n3, b3, db3);
        orc_add((size_t)n2 * n1, W2, dW2); orc_add((size_t)n2, b2, db2);
        orc_add((size_t)n1 * n0, W1, dW1); orc_add((size_t)n1, b1, db1);
        free(dZ3); free(dW3); free(db3); free(dZ2); free(dW2); free(db2);
        free(dZ1); free(dW1); free(db1);
    }
    free(X); free(Z1); free(A1); free(Z2); free(A2); free(Z3); free(A3);
}

/* ------------------------------------------------------------------------------------------ */
/* model/mnist_hinge.c -- one full-batch iteration of the 10 one-vs-rest hinge classifiers      */
/* ------------------------------------------------------------------------------------------ */

/* model/mnist_hinge.c:123-166 for N samples.  w[10][784] updated in place; Xraw[N][784] raw
 * pixels (scaled by 1/255.0F per sample, :135); labels[N].  grad[10][784] is the caller's
 * persistent gradient buffer: the reference clears only its first 784 BYTES (196 floats) each
 * iteration (memset(..., 784), :126), which is reproduced.  norms[10] receives
 * ||grad_p|| / N (:156).  Element type of this model is float regardless of REAL (it predates
 * matrix_float_t), so REAL=float is the meaningful build. */
void orc_hinge_iter(int N, int D, REAL* w, REAL* grad, const REAL* Xraw, const int* labels,
                    float learn_rate, float* norms) {
    for (int p = 0; p < 10; ++p) memset(grad + (size_t)p * D, 0, (size_t)D); /* bytes, :126 */
    REAL* xs = (REAL*)malloc((size_t)D * sizeof(REAL));
    for (int j = 0; j < N; ++j) {
        for (int k = 0; k < D; ++k) xs[k] = Xraw[(size_t)j * D + k] * (REAL)(1 / 255.0F);
        for (int p = 0; p < 10; ++p) {
            float y = (labels[j] == p) ? 1.0f : -1.0f;
            REAL dot = 0;
            for (int k = 0; k < D; ++k) dot += w[(size_t)p * D + k] * xs[k];
            float val = 1 - y * dot;
            if (val < 1)
                for (int k = 0; k < D; ++k) grad[(size_t)p * D + k] += -y * xs[k];
        }
    }
    free(xs);
    for (int p = 0; p < 10; ++p) {
        float tot = 0.0f;
        for (int k = 0; k < D; ++k) tot += grad[(size_t)p * D + k] * grad[(size_t)p * D + k];
        if (norms) norms[p] = (float)sqrt(tot) / N;
        orc_scale((size_t)D, grad + (size_t)p * D, (REAL)learn_rate);
        orc_add((size_t)D, w + (size_t)p * D, grad + (size_t)p * D);
    }
}

/* sizeof(REAL) so the Python side can assert it loaded the build it thinks it did */
int orc_real_size(void) { return (int)sizeof(REAL); }
