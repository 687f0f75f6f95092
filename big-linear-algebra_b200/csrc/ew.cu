// ew.cu -- elementwise kernels of the dense hot path.  All are HBM-bound:
// one pass, 128-bit coalesced accesses, 4 independent 16-byte requests in flight per thread,
// grids sized in multiples of the SM count.  Bytes per element (fp32): scale 8, add 12,
// hadamard 12, copy 8, relu 8, add_tile 8 (+bias), transpose 8.
#include <cmath>
#include <cstdint>

#include "kernels.h"
#include "runtime.h"

namespace bla {

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// out[i] = op(x[i], y[i]); out may alias x.  NIN = number of inputs actually read (1 or 2).
template <int NIN, class Op>
__global__ void __launch_bounds__(kThreads) ew_vec_kernel(float* out, const float* x, const float* y, size_t n, Op op) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    const size_t n4 = n >> 2;
    const size_t stride = (size_t)gridDim.x * kThreads;
    size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    // main loop: kUnroll float4 per thread per trip, all loads issued before the first use
    for (; i + (kUnroll - 1) * stride < n4; i += kUnroll * stride) {
        float4 a[kUnroll], b[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            a[u] = ld4(x + 4 * (i + u * stride));
            if (NIN == 2) b[u] = ld4(y + 4 * (i + u * stride));
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            float4 r;
            r.x = op(a[u].x, b[u].x); r.y = op(a[u].y, b[u].y); r.z = op(a[u].z, b[u].z); r.w = op(a[u].w, b[u].w);
            st4(out + 4 * (i + u * stride), r);
        }
    }
    for (; i < n4; i += stride) {
        float4 a = ld4(x + 4 * i), b = make_float4(0, 0, 0, 0);
        if (NIN == 2) b = ld4(y + 4 * i);
        float4 r;
        r.x = op(a.x, b.x); r.y = op(a.y, b.y); r.z = op(a.z, b.z); r.w = op(a.w, b.w);
        st4(out + 4 * i, r);
    }
    // ragged tail (< 4 elements)
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        size_t j = (n4 << 2) + threadIdx.x;
        out[j] = op(x[j], NIN == 2 ? y[j] : 0.f);
    }
}

// fallback for pointers that are not 16-byte aligned (interior pointers such as buffer + 1)
template <int NIN, class Op>
__global__ void __launch_bounds__(kThreads) ew_scalar_kernel(float* out, const float* x, const float* y, size_t n, Op op) {
    pdl_trigger();
    pdl_wait();
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) out[i] = op(x[i], NIN == 2 ? y[i] : 0.f);
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

inline int grid_for(size_t work_items, int per_thread) {
    size_t blocks = (work_items + (size_t)kThreads * per_thread - 1) / ((size_t)kThreads * per_thread);
    size_t cap = (size_t)rt().num_sms * 8;  // 8 resident 256-thread CTAs per SM
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <int NIN, class Op>
void launch_ew(float* out, const float* x, const float* y, size_t n, Op op, cudaStream_t s) {
    if (n == 0) return;
    bool vec = aligned16(out) && aligned16(x) && (NIN == 1 || aligned16(y));
    if (vec) {
        BLA_CUDA(launch_pdl(ew_vec_kernel<NIN, Op>, dim3(grid_for(n >> 2 ? n >> 2 : 1, kUnroll)), dim3(kThreads), 0, s, 1, out, x, y, n, op));
    } else {
        BLA_CUDA(launch_pdl(ew_scalar_kernel<NIN, Op>, dim3(grid_for(n, 4)), dim3(kThreads), 0, s, 1, out, x, y, n, op));
    }
    count_launch();
}

struct OpScale { float f; __device__ float operator()(float a, float) const { return a * f; } };
struct OpAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMul { __device__ float operator()(float a, float b) const { return a * b; } };
struct OpCopy { __device__ float operator()(float a, float) const { return a; } };
struct OpAxpy { float alpha; __device__ float operator()(float y, float x) const { return y + alpha * x; } };
struct OpRelu { __device__ float operator()(float a, float) const { return a < 0.f ? 0.f : a; } };
struct OpReluDdx { __device__ float operator()(float a, float) const { return a > 0.f ? 1.f : 0.f; } };
// dest = relu_result <= 0 ? 0 : source   (x = source, y = relu_result)
struct OpReluBwd { __device__ float operator()(float src, float r) const { return r <= 0.f ? 0.f : src; } };

}  // namespace

void k_scale(float* m, size_t n, float f, cudaStream_t s) { launch_ew<1>(m, m, nullptr, n, OpScale{f}, s); }
void k_add(float* a, const float* b, size_t n, cudaStream_t s) { launch_ew<2>(a, a, b, n, OpAdd{}, s); }
void k_hadamard(float* a, const float* b, size_t n, cudaStream_t s) { launch_ew<2>(a, a, b, n, OpMul{}, s); }
void k_copy(float* dst, const float* src, size_t n, cudaStream_t s) { launch_ew<1>(dst, src, nullptr, n, OpCopy{}, s); }
void k_axpy(float* y, const float* x, float alpha, size_t n, cudaStream_t s) { launch_ew<2>(y, y, x, n, OpAxpy{alpha}, s); }
void k_relu(float* d, size_t n, cudaStream_t s) { launch_ew<1>(d, d, nullptr, n, OpRelu{}, s); }
void k_relu_ddx(float* d, size_t n, cudaStream_t s) { launch_ew<1>(d, d, nullptr, n, OpReluDdx{}, s); }
void k_relu_backward(const float* src, const float* relu_result, float* dst, size_t n, cudaStream_t s) {
    launch_ew<2>(dst, src, relu_result, n, OpReluBwd{}, s);
}

// ---------------------------------------------------------------------------------------------
// bias broadcasts
// ---------------------------------------------------------------------------------------------
namespace {

// a[r][c] += b[r][c % bcols].  One block row per matrix row; float4 along c when bcols == 1
// (the bias-column case of every dense layer, model/mnist_nn.c:222).
__global__ void __launch_bounds__(kThreads) add_tile_columns_kernel(float* a, int rows, int cols, const float* __restrict__ b,
                                                                    int bcols, bool vec) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        float* row = a + (size_t)r * cols;
        const float* brow = b + (size_t)r * bcols;
        if (bcols == 1 && vec) {
            const float bv = brow[0];
            const int c4 = cols >> 2;
            for (int i = blockIdx.x * kThreads + threadIdx.x; i < c4; i += gridDim.x * kThreads) {
                float4 v = ld4(row + 4 * i);
                v.x += bv; v.y += bv; v.z += bv; v.w += bv;
                st4(row + 4 * i, v);
            }
            // cols % 4 == 0 is part of `vec`
        } else {
            for (int c = blockIdx.x * kThreads + threadIdx.x; c < cols; c += gridDim.x * kThreads) row[c] += brow[c % bcols];
        }
    }
}

// a[r][c] += b[c]
__global__ void __launch_bounds__(kThreads) add_tile_rows_kernel(float* a, int rows, int cols, const float* __restrict__ b, bool vec) {
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        float* row = a + (size_t)r * cols;
        if (vec) {
            const int c4 = cols >> 2;
            for (int i = blockIdx.x * kThreads + threadIdx.x; i < c4; i += gridDim.x * kThreads) {
                float4 v = ld4(row + 4 * i);
                float4 w = ld4(b + 4 * i);
                v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
                st4(row + 4 * i, v);
            }
        } else {
            for (int c = blockIdx.x * kThreads + threadIdx.x; c < cols; c += gridDim.x * kThreads) row[c] += b[c];
        }
    }
}

dim3 grid2d(int rows, int cols_items) {
    int gx = (cols_items + kThreads - 1) / kThreads;
    if (gx < 1) gx = 1;
    int cap = rt().num_sms * 8;
    if (gx > cap) gx = cap;
    int gy = rows;
    int maxy = cap / gx;
    if (maxy < 1) maxy = 1;
    if (gy > maxy) gy = maxy;
    if (gy > 65535) gy = 65535;
    return dim3(gx, gy);
}

}  // namespace

void k_add_tile_columns(float* a, int rows, int cols, const float* b, int bcols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    bool vec = aligned16(a) && (cols % 4 == 0);
    add_tile_columns_kernel<<<grid2d(rows, vec && bcols == 1 ? cols / 4 : cols), kThreads, 0, s>>>(a, rows, cols, b, bcols, vec);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_add_tile_rows(float* a, int rows, int cols, const float* b, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    bool vec = aligned16(a) && aligned16(b) && (cols % 4 == 0);
    add_tile_rows_kernel<<<grid2d(rows, vec ? cols / 4 : cols), kThreads, 0, s>>>(a, rows, cols, b, vec);
    BLA_LAUNCH_CHECK();
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// transpose: 64x64 tiles through padded shared memory, 128-bit global accesses on both sides
// when the shape allows, scalar otherwise.  dst[c][r] = src[r][c]; src and dst must not alias.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kTile = 64;

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols,
                                                        bool vec) {
    __shared__ float tile[kTile][kTile + 1];
    const int tiles_c = (cols + kTile - 1) / kTile;
    const int tiles_r = (rows + kTile - 1) / kTile;
    const long long ntiles = (long long)tiles_c * tiles_r;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int r0 = (int)(t / tiles_c) * kTile, c0 = (int)(t % tiles_c) * kTile;
        if (vec) {
            // 256 threads: 16 float4 per tile row, 16 rows per pass, 4 passes
            const int q = threadIdx.x & 15, rr = threadIdx.x >> 4;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                int r = r0 + rr + 16 * p, c = c0 + 4 * q;
                float4 v = make_float4(0, 0, 0, 0);
                if (r < rows && c < cols) v = ld4(src + (size_t)r * cols + c);   // cols % 4 == 0
                tile[rr + 16 * p][4 * q + 0] = v.x; tile[rr + 16 * p][4 * q + 1] = v.y;
                tile[rr + 16 * p][4 * q + 2] = v.z; tile[rr + 16 * p][4 * q + 3] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                int c = c0 + rr + 16 * p, r = r0 + 4 * q;   // dst row = src col
                if (c < cols && r < rows) {
                    float4 v;
                    v.x = tile[4 * q + 0][rr + 16 * p]; v.y = tile[4 * q + 1][rr + 16 * p];
                    v.z = tile[4 * q + 2][rr + 16 * p]; v.w = tile[4 * q + 3][rr + 16 * p];
                    st4(dst + (size_t)c * rows + r, v);                          // rows % 4 == 0
                }
            }
        } else {
            const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
            for (int rr = ty; rr < kTile; rr += 4) {
                int r = r0 + rr, c = c0 + tx;
                tile[rr][tx] = (r < rows && c < cols) ? src[(size_t)r * cols + c] : 0.f;
            }
            __syncthreads();
            for (int cc = ty; cc < kTile; cc += 4) {
                int c = c0 + cc, r = r0 + tx;
                if (c < cols && r < rows) dst[(size_t)c * rows + r] = tile[tx][cc];
            }
        }
        __syncthreads();
    }
}

}  // namespace

void k_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t s) {
    if (rows <= 0 || cols <= 0) return;
    long long ntiles = (long long)((rows + kTile - 1) / kTile) * ((cols + kTile - 1) / kTile);
    long long cap = (long long)rt().num_sms * 8;
    int grid = (int)(ntiles < cap ? ntiles : cap);
    bool vec = aligned16(src) && aligned16(dst) && rows % 4 == 0 && cols % 4 == 0;
    transpose_kernel<<<grid, 256, 0, s>>>(src, dst, rows, cols, vec);
    BLA_LAUNCH_CHECK();
    count_launch();
}

// ---------------------------------------------------------------------------------------------
// synthetic data: counter-based generator (splitmix64 of seed + index), identical on host/device
// ---------------------------------------------------------------------------------------------
// uniform_at() lives in kernels.h (shared with the U-Net's dropout mask)

namespace {
__global__ void __launch_bounds__(kThreads) fill_uniform_kernel(float* dst, size_t n, unsigned long long seed, float lo, float hi) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) dst[i] = uniform_at(seed, i, lo, hi);
}
__global__ void __launch_bounds__(kThreads) u8_to_float_kernel(float* dst, const unsigned char* __restrict__ src, size_t n, float scale,
                                                               bool vec) {
    const size_t stride = (size_t)gridDim.x * kThreads;
    if (vec) {
        const size_t n4 = n >> 2;
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
            uchar4 u = reinterpret_cast<const uchar4*>(src)[i];
            st4(dst + 4 * i, make_float4(u.x * scale, u.y * scale, u.z * scale, u.w * scale));
        }
        if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
            size_t j = (n4 << 2) + threadIdx.x;
            dst[j] = src[j] * scale;
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) dst[i] = src[i] * scale;
    }
}
}  // namespace

void k_fill_uniform(float* dst, size_t n, unsigned long long seed, float lo, float hi, cudaStream_t s) {
    if (!n) return;
    fill_uniform_kernel<<<grid_for(n, 4), kThreads, 0, s>>>(dst, n, seed, lo, hi);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_u8_to_float(float* dst, const unsigned char* src, size_t n, float scale, cudaStream_t s) {
    if (!n) return;
    bool vec = aligned16(dst) && (((uintptr_t)src & 3) == 0);
    u8_to_float_kernel<<<grid_for(vec ? (n >> 2) + 1 : n, 4), kThreads, 0, s>>>(dst, src, n, scale, vec);
    BLA_LAUNCH_CHECK();
    count_launch();
}

}  // namespace bla

extern "C" {
void bla_fill_uniform(float* dst, size_t n, unsigned long long seed, float lo, float hi) {
    bla::k_fill_uniform(dst, n, seed, lo, hi, bla::rt().stream);
}
void bla_host_uniform(float* dst, size_t n, unsigned long long seed, float lo, float hi) {
    for (size_t i = 0; i < n; ++i) dst[i] = bla::uniform_at(seed, i, lo, hi);
}
void bla_u8_to_float(float* dst, const unsigned char* src, size_t n, float scale) {
    bla::k_u8_to_float(dst, src, n, scale, bla::rt().stream);
}
}
