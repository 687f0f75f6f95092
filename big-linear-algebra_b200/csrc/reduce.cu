// reduce.cu -- reductions and softmax of the dense hot path.  HBM-bound:
// every input element is read once with coalesced (128-bit where aligned) loads, partial results
// are combined with warp shuffles, and the cross-block stage is a second tiny launch over a fixed
// partial buffer, so results are deterministic (no floating-point atomics).
#include <cfloat>
#include <cmath>
#include <cstdint>

#include "kernels.h"
#include "runtime.h"

namespace bla {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxPartials = 148 * 8;   // stage-1 blocks never exceed this

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum; result valid in thread 0
template <class T>
__device__ __forceinline__ T block_sum(T v) {
    __shared__ T sh[kThreads / 32];
    __syncthreads();   // protect sh across back-to-back calls
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : (T)0;
        r = warp_sum(r);
    }
    return r;
}
__device__ __forceinline__ float block_max(float v) {
    __shared__ float shm[kThreads / 32];
    __syncthreads();
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) shm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = -INFINITY;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kThreads / 32 ? shm[threadIdx.x] : -INFINITY;
        r = warp_max(r);
    }
    return r;
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

inline int stage1_blocks(size_t n) {
    size_t b = (n + (size_t)kThreads * 16 - 1) / ((size_t)kThreads * 16);
    size_t cap = (size_t)rt().num_sms * 8;
    if (cap > (size_t)kMaxPartials) cap = kMaxPartials;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- whole-array statistics: sum, sum of squares (double partials), max -----------------------
struct Stats { double sum, sumsq; float mx; };

__global__ void __launch_bounds__(kThreads) stats_stage1(const float* __restrict__ m, size_t n, bool vec, double* psum, double* psq,
                                                         float* pmax) {
    // per-thread partials in fp32 over short strided chains, promoted to double at block level
    float s = 0.f, q = 0.f, mx = -INFINITY;
    const size_t stride = (size_t)gridDim.x * kThreads;
    const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    double ds = 0.0, dq = 0.0;
    int chain = 0;
    if (vec) {
        const size_t n4 = n >> 2;
        for (size_t i = tid; i < n4; i += stride) {
            float4 v = *reinterpret_cast<const float4*>(m + 4 * i);
            s += (v.x + v.y) + (v.z + v.w);
            q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            if (++chain == 64) { ds += s; dq += q; s = q = 0.f; chain = 0; }
        }
        if (tid < (n & 3)) {
            float v = m[(n4 << 2) + tid];
            s += v; q += v * v; mx = fmaxf(mx, v);
        }
    } else {
        for (size_t i = tid; i < n; i += stride) {
            float v = m[i];
            s += v; q += v * v; mx = fmaxf(mx, v);
            if (++chain == 256) { ds += s; dq += q; s = q = 0.f; chain = 0; }
        }
    }
    ds += s; dq += q;
    double bs = block_sum<double>(ds);
    double bq = block_sum<double>(dq);
    float bm = block_max(mx);
    if (threadIdx.x == 0) { psum[blockIdx.x] = bs; psq[blockIdx.x] = bq; pmax[blockIdx.x] = bm; }
}

__global__ void __launch_bounds__(kThreads) stats_stage2(const double* psum, const double* psq, const float* pmax, int nparts, Stats* out) {
    double s = 0.0, q = 0.0; float mx = -INFINITY;
    for (int i = threadIdx.x; i < nparts; i += kThreads) { s += psum[i]; q += psq[i]; mx = fmaxf(mx, pmax[i]); }
    s = block_sum<double>(s);
    q = block_sum<double>(q);
    mx = block_max(mx);
    if (threadIdx.x == 0) { out->sum = s; out->sumsq = q; out->mx = mx; }
}

struct Workspace {
    double psum[kMaxPartials];
    double psq[kMaxPartials];
    float pmax[kMaxPartials];
    Stats stats;
    float partial_rows[1];  // col_sum partials follow (sized by reduce_workspace_bytes)
};
constexpr size_t kColPartialFloats = 1 << 20;

void run_stats(const float* m, size_t n, Workspace* w, cudaStream_t s) {
    int blocks = stage1_blocks(n);
    stats_stage1<<<blocks, kThreads, 0, s>>>(m, n, aligned16(m), w->psum, w->psq, w->pmax);
    BLA_LAUNCH_CHECK();
    stats_stage2<<<1, kThreads, 0, s>>>(w->psum, w->psq, w->pmax, blocks, &w->stats);
    BLA_LAUNCH_CHECK();
    count_launch(2);
}

__global__ void write_sumsq(const Stats* st, double* out) { *out = st->sumsq; }
__global__ void write_max(const Stats* st, float* out) { *out = st->mx; }

// (x - mean) / sd with mean/var from double sums and sd through sqrtf (lib/matrix.c:176-184)
__global__ void __launch_bounds__(kThreads) zscore_apply(float* m, size_t n, const Stats* st, bool vec) {
    const double mean = st->sum / (double)n;
    const float sd = sqrtf((float)(st->sumsq / (double)n - mean * mean));
    const size_t stride = (size_t)gridDim.x * kThreads;
    const size_t tid = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (vec) {
        const size_t n4 = n >> 2;
        for (size_t i = tid; i < n4; i += stride) {
            float4 v = *reinterpret_cast<float4*>(m + 4 * i);
            v.x = (float)((double)v.x - mean) / sd; v.y = (float)((double)v.y - mean) / sd;
            v.z = (float)((double)v.z - mean) / sd; v.w = (float)((double)v.w - mean) / sd;
            *reinterpret_cast<float4*>(m + 4 * i) = v;
        }
        if (tid < (n & 3)) { size_t j = (n4 << 2) + tid; m[j] = (float)((double)m[j] - mean) / sd; }
    } else {
        for (size_t i = tid; i < n; i += stride) m[i] = (float)((double)m[i] - mean) / sd;
    }
}

// ---- row_sum: out[c] = sum_r m[r][c]; 32 columns per block, 8 row lanes, smem combine ----------
__global__ void __launch_bounds__(kThreads) row_sum_kernel(const float* __restrict__ m, int rows, int cols, float* out) {
    __shared__ float sh[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int c0 = blockIdx.x * 32; c0 < cols; c0 += gridDim.x * 32) {
        const int c = c0 + tx;
        float acc = 0.f;
        if (c < cols)
            for (int r = ty; r < rows; r += 8) acc += m[(size_t)r * cols + c];
        sh[ty][tx] = acc;
        __syncthreads();
        if (ty == 0 && c < cols) {
            float t = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += sh[k][tx];
            out[c] = t;
        }
        __syncthreads();
    }
}

// ---- col_sum: out[i] = sum_{j<cols} flat[i*stride + j] (elements past the end count as 0) ------
// stride = rows reproduces lib/matrix.c:144; stride = cols is the intended row total.
// Grid = (chunks, rows): each block sums one contiguous chunk of one window.
__global__ void __launch_bounds__(kThreads) col_sum_stage1(const float* __restrict__ m, size_t total, int cols, size_t win_stride,
                                                           int chunk, float* partial, int nchunks) {
    const int row = blockIdx.y, ch = blockIdx.x;
    const size_t base = (size_t)row * win_stride;
    const int j0 = ch * chunk;
    int j1 = j0 + chunk;
    if (j1 > cols) j1 = cols;
    float acc = 0.f;
    // clip the window at the end of the buffer once, then stream it with 128-bit loads when it is aligned
    size_t lo = base + j0, hi = base + j1;
    if (hi > total) hi = total;
    if (lo < hi) {
        const float* ptr = m + lo;
        const size_t n = hi - lo;
        if ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) {
            const size_t n4 = n >> 2;
            float a0 = 0.f, a1 = 0.f;
            size_t i = threadIdx.x;
            const float4* p4 = reinterpret_cast<const float4*>(ptr);
            for (; i + 3 * kThreads < n4; i += 4 * kThreads) {   // four independent 128-bit loads in flight per thread
                const float4 u = __ldg(p4 + i), w = __ldg(p4 + i + kThreads), y = __ldg(p4 + i + 2 * kThreads), z = __ldg(p4 + i + 3 * kThreads);
                a0 += ((u.x + u.y) + (u.z + u.w)) + ((y.x + y.y) + (y.z + y.w));
                a1 += ((w.x + w.y) + (w.z + w.w)) + ((z.x + z.y) + (z.z + z.w));
            }
            for (; i + kThreads < n4; i += 2 * kThreads) {
                const float4 u = reinterpret_cast<const float4*>(ptr)[i], w = reinterpret_cast<const float4*>(ptr)[i + kThreads];
                a0 += (u.x + u.y) + (u.z + u.w);
                a1 += (w.x + w.y) + (w.z + w.w);
            }
            for (; i < n4; i += kThreads) {
                const float4 u = reinterpret_cast<const float4*>(ptr)[i];
                a0 += (u.x + u.y) + (u.z + u.w);
            }
            acc = a0 + a1;
            if (threadIdx.x < (n & 3)) acc += ptr[(n4 << 2) + threadIdx.x];
        } else {
            for (size_t i = threadIdx.x; i < n; i += kThreads) acc += ptr[i];
        }
    }
    acc = block_sum<float>(acc);
    if (threadIdx.x == 0) partial[(size_t)row * nchunks + ch] = acc;
}
__global__ void col_sum_stage2(const float* partial, int rows, int nchunks, float* out) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float acc = 0.f;
    for (int c = 0; c < nchunks; ++c) acc += partial[(size_t)r * nchunks + c];
    out[r] = acc;
}

// ---- softmax ------------------------------------------------------------------------------------
// columns: one thread per column, rows strided by cols (coalesced across the warp); rows is small
// (10 classes).  expf keeps 2 ulp, far inside the 1e-5 budget against the reference's double exp.
__global__ void __launch_bounds__(kThreads) softmax_cols_kernel(float* d, int rows, int cols) {
    for (int c = blockIdx.x * kThreads + threadIdx.x; c < cols; c += gridDim.x * kThreads) {
        float mx = -INFINITY;
        for (int r = 0; r < rows; ++r) mx = fmaxf(mx, d[(size_t)r * cols + c]);
        float tot = 0.f;
        for (int r = 0; r < rows; ++r) {
            float e = expf(d[(size_t)r * cols + c] - mx);
            d[(size_t)r * cols + c] = e;
            tot += e;
        }
        for (int r = 0; r < rows; ++r) d[(size_t)r * cols + c] /= tot;
    }
}

// rows: one warp per row, shuffle reductions for max and sum
__global__ void __launch_bounds__(kThreads) softmax_rows_kernel(float* d, int rows, int cols) {
    const int lane = threadIdx.x & 31;
    const int warps = kThreads / 32;
    for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
        float* row = d + (size_t)r * cols;
        float mx = -INFINITY;
        for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, row[c]);
        mx = warp_max(mx);
        float tot = 0.f;
        for (int c = lane; c < cols; c += 32) {
            float e = expf(row[c] - mx);
            row[c] = e;
            tot += e;
        }
        tot = warp_sum(tot);
        for (int c = lane; c < cols; c += 32) row[c] /= tot;
    }
}

// Fused tail of the MLP forward + head of its backward (model/mnist_nn.c:234-268): column
// softmax, argmax hit count, the reference's flat-slice cross-entropy, and dZ = (p - y) * scale.
// The loss of "sample k" in the reference reads the FLAT slice [k*classes, (k+1)*classes) of both
// the probability and the expectation matrix (:250-252), not column k; summed over all k that is
// simply sum over the whole matrix of -y*log(p + 1e-15), which is what is accumulated here.
__global__ void __launch_bounds__(kThreads) softmax_xent_kernel(const float* logits, const float* __restrict__ expected, int classes,
                                                                int batch, float* probs, float* grad, float grad_scale, double* stats) {
    double loss = 0.0;
    int correct = 0;
    for (int c = blockIdx.x * kThreads + threadIdx.x; c < batch; c += gridDim.x * kThreads) {
        float mx = -INFINITY;
        for (int r = 0; r < classes; ++r) mx = fmaxf(mx, logits[(size_t)r * batch + c]);
        float tot = 0.f;
        for (int r = 0; r < classes; ++r) tot += expf(logits[(size_t)r * batch + c] - mx);
        int pred = 0;
        float best = 0.f;
        float l = 0.f;
        for (int r = 0; r < classes; ++r) {
            const size_t at = (size_t)r * batch + c;
            float p = expf(logits[at] - mx) / tot;
            float y = expected[at];
            if (p > best) { best = p; pred = r; }
            l += -1.f * (y * logf(p + 1e-15f));
            if (probs) probs[at] = p;
            if (grad) grad[at] = (p + (-1.0f) * y) * grad_scale;
        }
        if (expected[(size_t)pred * batch + c] == 1.f) ++correct;
        loss += (double)l;
    }
    double bl = block_sum<double>(loss);
    double bc = block_sum<double>((double)correct);
    if (threadIdx.x == 0 && stats) {
        atomicAdd(&stats[0], bl);   // two scalars per step; order-insensitive at double precision
        atomicAdd(&stats[1], bc);
    }
}

}  // namespace

size_t reduce_workspace_bytes() { return sizeof(Workspace) + kColPartialFloats * sizeof(float); }

void k_row_sum(const float* m, int rows, int cols, float* out, cudaStream_t s) {
    if (cols <= 0) return;
    int blocks = (cols + 31) / 32;
    int cap = rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    row_sum_kernel<<<blocks, kThreads, 0, s>>>(m, rows, cols, out);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_col_sum(const float* m, int rows, int cols, float* out, int quirk, void* work, cudaStream_t s) {
    if (rows <= 0) return;
    Workspace* w = (Workspace*)work;
    // chunks: enough blocks to fill the machine, but partials must fit the workspace
    int target_blocks = rt().num_sms * 4;
    int nchunks = (target_blocks + rows - 1) / rows;
    int max_chunks = (cols + 2047) / 2048;   // at least 2048 elements per block
    if (nchunks > max_chunks) nchunks = max_chunks;
    if (nchunks < 1) nchunks = 1;
    while ((size_t)rows * nchunks > kColPartialFloats && nchunks > 1) --nchunks;
    if ((size_t)rows * nchunks > kColPartialFloats) die("bla: matrix_col_sum of %d rows exceeds the workspace, exiting", rows);
    int chunk = (cols + nchunks - 1) / nchunks;
    if (chunk < 1) chunk = 1;
    size_t total = (size_t)rows * cols;
    size_t win = quirk ? (size_t)rows : (size_t)cols;
    dim3 grid(nchunks, rows);
    if (rows > 65535) die("bla: matrix_col_sum supports at most 65535 rows, exiting");
    col_sum_stage1<<<grid, kThreads, 0, s>>>(m, total, cols, win, chunk, w->partial_rows, nchunks);
    BLA_LAUNCH_CHECK();
    col_sum_stage2<<<(rows + 127) / 128, 128, 0, s>>>(w->partial_rows, rows, nchunks, out);
    BLA_LAUNCH_CHECK();
    count_launch(2);
}

void k_sum_squares(const float* m, size_t n, double* out1, void* work, cudaStream_t s) {
    Workspace* w = (Workspace*)work;
    run_stats(m, n, w, s);
    write_sumsq<<<1, 1, 0, s>>>(&w->stats, out1);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_max(const float* m, size_t n, float* out1, void* work, cudaStream_t s) {
    Workspace* w = (Workspace*)work;
    run_stats(m, n, w, s);
    write_max<<<1, 1, 0, s>>>(&w->stats, out1);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_zscore(float* m, size_t n, void* work, cudaStream_t s) {
    if (!n) return;
    Workspace* w = (Workspace*)work;
    run_stats(m, n, w, s);
    size_t items = aligned16(m) ? (n >> 2) + 1 : n;
    size_t blocks = (items + kThreads * 4 - 1) / (kThreads * 4);
    size_t cap = (size_t)rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    zscore_apply<<<(int)blocks, kThreads, 0, s>>>(m, n, &w->stats, aligned16(m));
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_softmax_cols(float* d, int rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    int blocks = (cols + kThreads - 1) / kThreads;
    int cap = rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    softmax_cols_kernel<<<blocks, kThreads, 0, s>>>(d, rows, cols);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_softmax_rows(float* d, int rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    int blocks = (rows + 7) / 8;
    int cap = rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    softmax_rows_kernel<<<blocks, kThreads, 0, s>>>(d, rows, cols);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_softmax_xent(const float* logits, const float* expected, int classes, int batch, float* probs, float* grad, float grad_scale,
                    double* stats, cudaStream_t s) {
    if (batch <= 0) return;
    int blocks = (batch + kThreads - 1) / kThreads;
    int cap = rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    softmax_xent_kernel<<<blocks, kThreads, 0, s>>>(logits, expected, classes, batch, probs, grad, grad_scale, stats);
    BLA_LAUNCH_CHECK();
    count_launch();
}

}  // namespace bla
