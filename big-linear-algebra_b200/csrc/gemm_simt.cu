// gemm_simt.cu -- the FP32-exact GEMM path: C = op(A).op(B) with every product and
// sum in IEEE fp32 FMA on the SIMT pipe.  Replaces the reference's triple loop
// (lib/matrix.c:47-57) and, through the transpose flags and the fused epilogue, its
// transpose/multiply/transpose + add_tile + activation sequences (model/mnist_nn.c:221-293).
//
// Design (B200: 148 SMs x 128 FP32 lanes): 128x128x16 CTA tiles, 256 threads, 8x8 register
// micro-tiles (16 FFMA per 128-bit shared load), double-buffered shared memory with register
// prefetch of the next global tile, 2 CTAs per SM.  Skinny problems use 64x64 tiles; problems
// with few output tiles and a long K are split along K into a workspace and reduced by a second
// kernel (deterministic, no atomics), which also shortens the fp32 accumulation chains.
// Roofline: FP32 FMA peak = SMs x 128 x 2 x clock.
#include <cstdint>

#include "kernels.h"
#include "runtime.h"

namespace bla {

namespace {

constexpr int BK = 16;
constexpr int kThreads = 256;

struct SimtParams {
    int m, n, k;
    const float* a; int lda;
    const float* b; int ldb;
    float* c; int ldc;
    bla_epilogue epi;
    int k_chunk;       // K range per blockIdx.z (== k when not split)
    float* partial;    // split-K workspace [splits][m][n] or nullptr
    bool a_vec, b_vec, c_vec;
};

// One element-row of 4 consecutive elements of a row-major matrix [R][C] with leading dim ld,
// starting at (r, c0).  Out-of-range elements read as 0.
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int ld, int r, int c0, int R, int C, bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < R) {
        const float* p = base + (size_t)r * ld + c0;
        if (vec && c0 + 3 < C) {
            v = *reinterpret_cast<const float4*>(p);
        } else {
            if (c0 + 0 < C) v.x = p[0];
            if (c0 + 1 < C) v.y = p[1];
            if (c0 + 2 < C) v.z = p[2];
            if (c0 + 3 < C) v.w = p[3];
        }
    }
    return v;
}

__device__ __forceinline__ float epilogue_value(float acc, int i, int j, const SimtParams& p) {
    float v = acc;
    if (p.epi.alpha != 0.f) v *= p.epi.alpha;
    if (p.epi.bias_rows) v += p.epi.bias_rows[i];
    if (p.epi.bias_cols) v += p.epi.bias_cols[j];
    if (p.epi.pre_activation) p.epi.pre_activation[(size_t)i * p.ldc + j] = v;
    if (p.epi.activation == BLA_ACT_RELU) v = v < 0.f ? 0.f : v;
    if (p.epi.gate) v = p.epi.gate[(size_t)i * p.ldc + j] > 0.f ? v : 0.f;
    return v;
}

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(kThreads, 2) gemm_simt_kernel(const SimtParams p) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    constexpr int TM = BM / 16, TN = BN / 16;   // 8x8 (128) or 4x4 (64)
    constexpr int GM = TM / 4, GN = TN / 4;     // float4 groups per thread along m / n
    constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
    constexpr int A_V = BM * BK / 4 / kThreads;  // float4 per thread per A tile (2 or 1)
    constexpr int B_V = BN * BK / 4 / kThreads;

    __shared__ __align__(16) float As[2][BK][LDA_S];
    __shared__ __align__(16) float Bs[2][BK][LDB_S];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * p.k_chunk;
    int kend = kbeg + p.k_chunk;
    if (kend > p.k) kend = p.k;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_V], rb[B_V];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int v = 0; v < A_V; ++v) {
            const int idx = tid + v * kThreads;
            if (!TA) {   // A[m][k]: 4 consecutive k of one row
                const int row = idx / (BK / 4), kq = idx % (BK / 4);
                ra[v] = load4(p.a, p.lda, m0 + row, k0 + 4 * kq, p.m, kend, p.a_vec);
            } else {     // A stored [k][m]: 4 consecutive m of one k
                const int kr = idx / (BM / 4), mq = idx % (BM / 4);
                ra[v] = load4(p.a, p.lda, k0 + kr, m0 + 4 * mq, kend, p.m, p.a_vec);
            }
        }
#pragma unroll
        for (int v = 0; v < B_V; ++v) {
            const int idx = tid + v * kThreads;
            if (!TB) {   // B[k][n]: 4 consecutive n of one k
                const int kr = idx / (BN / 4), nq = idx % (BN / 4);
                rb[v] = load4(p.b, p.ldb, k0 + kr, n0 + 4 * nq, kend, p.n, p.b_vec);
            } else {     // B stored [n][k]: 4 consecutive k of one column
                const int col = idx / (BK / 4), kq = idx % (BK / 4);
                rb[v] = load4(p.b, p.ldb, n0 + col, k0 + 4 * kq, p.n, kend, p.b_vec);
            }
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int v = 0; v < A_V; ++v) {
            const int idx = tid + v * kThreads;
            if (!TA) {
                const int row = idx / (BK / 4), kq = idx % (BK / 4);
                As[buf][4 * kq + 0][row] = ra[v].x; As[buf][4 * kq + 1][row] = ra[v].y;
                As[buf][4 * kq + 2][row] = ra[v].z; As[buf][4 * kq + 3][row] = ra[v].w;
            } else {
                const int kr = idx / (BM / 4), mq = idx % (BM / 4);
                *reinterpret_cast<float4*>(&As[buf][kr][4 * mq]) = ra[v];
            }
        }
#pragma unroll
        for (int v = 0; v < B_V; ++v) {
            const int idx = tid + v * kThreads;
            if (!TB) {
                const int kr = idx / (BN / 4), nq = idx % (BN / 4);
                *reinterpret_cast<float4*>(&Bs[buf][kr][4 * nq]) = rb[v];
            } else {
                const int col = idx / (BK / 4), kq = idx % (BK / 4);
                Bs[buf][4 * kq + 0][col] = rb[v].x; Bs[buf][4 * kq + 1][col] = rb[v].y;
                Bs[buf][4 * kq + 2][col] = rb[v].z; Bs[buf][4 * kq + 3][col] = rb[v].w;
            }
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        fetch(kbeg);
        stash(0);
    }
    __syncthreads();

    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) fetch(k0 + BK);   // global loads in flight while this tile is multiplied
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float af[TM], bf[TN];
#pragma unroll
            for (int g = 0; g < GM; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][g * (BM / GM) + ty * 4]);
                af[4 * g + 0] = t.x; af[4 * g + 1] = t.y; af[4 * g + 2] = t.z; af[4 * g + 3] = t.w;
            }
#pragma unroll
            for (int g = 0; g < GN; ++g) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][g * (BN / GN) + tx * 4]);
                bf[4 * g + 0] = t.x; bf[4 * g + 1] = t.y; bf[4 * g + 2] = t.z; bf[4 * g + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

    // ---- store ----
    // The relu' gate is read for the whole micro-tile before the first store: the output may alias
    // the gate as far as the compiler knows, and interleaved load/store pairs would each wait a
    // full DRAM round trip.
    const bool early_gate = p.epi.gate && !p.partial && !p.epi.bias_rows && !p.epi.bias_cols && !p.epi.pre_activation &&
                            p.epi.activation == BLA_ACT_IDENTITY;   // the dgrad case: gate commutes with alpha
    if (early_gate) {
#pragma unroll
        for (int gi = 0; gi < GM; ++gi)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = m0 + gi * (BM / GM) + ty * 4 + ii;
#pragma unroll
                for (int gj = 0; gj < GN; ++gj) {
                    const int j = n0 + gj * (BN / GN) + tx * 4;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float gv = 1.f;
                        if (i < p.m && j + jj < p.n) gv = __ldg(p.epi.gate + (size_t)i * p.ldc + j + jj);
                        if (!(gv > 0.f)) acc[4 * gi + ii][4 * gj + jj] = 0.f;
                    }
                }
            }
    }
    SimtParams q = p;
    if (early_gate) q.epi.gate = nullptr;   // already applied: a zeroed accumulator stays zero through alpha
#pragma unroll
    for (int gi = 0; gi < GM; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = m0 + gi * (BM / GM) + ty * 4 + ii;
            if (i >= p.m) continue;
#pragma unroll
            for (int gj = 0; gj < GN; ++gj) {
                const int j = n0 + gj * (BN / GN) + tx * 4;
                if (j >= p.n) continue;
                float v[4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) v[jj] = acc[4 * gi + ii][4 * gj + jj];
                if (p.partial) {
                    float* dst = p.partial + ((size_t)blockIdx.z * p.m + i) * p.n + j;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (j + jj < p.n) dst[jj] = v[jj];
                } else {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (j + jj < p.n) v[jj] = epilogue_value(v[jj], i, j + jj, q);
                    float* dst = p.c + (size_t)i * p.ldc + j;
                    if (p.c_vec && j + 3 < p.n) {
                        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                            if (j + jj < p.n) dst[jj] = v[jj];
                    }
                }
            }
        }
}

// sums the split-K partials in a fixed order and applies the epilogue
__global__ void __launch_bounds__(kThreads) splitk_reduce_kernel(const SimtParams p, int splits) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    const size_t total = (size_t)p.m * p.n;
    for (size_t e = (size_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (size_t)gridDim.x * kThreads) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += p.partial[(size_t)z * total + e];
        const int i = (int)(e / p.n), j = (int)(e % p.n);
        p.c[(size_t)i * p.ldc + j] = epilogue_value(s, i, j, p);
    }
}

template <int BM, int BN>
void launch_tile(const SimtParams& p, bool ta, bool tb, dim3 grid, cudaStream_t s) {
    if (!ta && !tb) BLA_CUDA(launch_pdl(gemm_simt_kernel<BM, BN, false, false>, dim3(grid), dim3(kThreads), 0, s, 1, p));
    else if (!ta && tb) BLA_CUDA(launch_pdl(gemm_simt_kernel<BM, BN, false, true>, dim3(grid), dim3(kThreads), 0, s, 1, p));
    else if (ta && !tb) BLA_CUDA(launch_pdl(gemm_simt_kernel<BM, BN, true, false>, dim3(grid), dim3(kThreads), 0, s, 1, p));
    else BLA_CUDA(launch_pdl(gemm_simt_kernel<BM, BN, true, true>, dim3(grid), dim3(kThreads), 0, s, 1, p));
    BLA_LAUNCH_CHECK();
    count_launch();
}

inline bool al16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

void gemm_simt(const GemmArgs& g, cudaStream_t s) {
    if (g.m <= 0 || g.n <= 0) return;
    SimtParams p;
    p.m = g.m; p.n = g.n; p.k = g.k;
    p.a = g.a; p.lda = g.lda; p.b = g.b; p.ldb = g.ldb; p.c = g.c; p.ldc = g.ldc;
    p.epi = g.epi;
    p.a_vec = al16(g.a) && g.lda % 4 == 0;
    p.b_vec = al16(g.b) && g.ldb % 4 == 0;
    p.c_vec = al16(g.c) && g.ldc % 4 == 0;
    p.partial = nullptr;
    p.k_chunk = g.k > 0 ? g.k : 1;

    const int sms = rt().num_sms;
    const long long tiles128 = (long long)ceil_div(g.m, 128) * ceil_div(g.n, 128);
    const bool big = g.m > 64 && g.n > 64 && tiles128 >= sms;
    const int BM = big ? 128 : 64, BN = BM;
    const long long tiles = (long long)ceil_div(g.m, BM) * ceil_div(g.n, BN);

    int splits = 1;
    if (tiles < 2LL * sms && g.k >= 1024) {
        long long want = (2LL * sms + tiles - 1) / tiles;
        long long maxs = g.k / 512;
        splits = (int)(want < maxs ? want : maxs);
        if (splits > 128) splits = 128;
        if (splits < 1) splits = 1;
    }
    const bool own_ws = g.workspace != nullptr;   // a GEMM beside the library stream: the caller's scratch, never the pool
    if (own_ws && (size_t)g.m * g.n > 0) {
        const long long fit = (long long)(g.workspace_floats / ((size_t)g.m * g.n));
        if (splits > fit) splits = fit < 1 ? 1 : (int)fit;
    }
    float* ws = nullptr;
    if (splits > 1) {
        int chunk = ceil_div(g.k, splits);
        chunk = (chunk + BK - 1) / BK * BK;
        splits = ceil_div(g.k, chunk);
        p.k_chunk = chunk;
        if (splits > 1) {
            ws = own_ws ? g.workspace : (float*)pool_alloc(kDevice, (size_t)splits * g.m * g.n * sizeof(float));
            p.partial = ws;
        }
    }
    dim3 grid(ceil_div(g.n, BN), ceil_div(g.m, BM), splits);
    if (grid.y > 65535 || grid.z > 65535) die("bla: GEMM grid too large (%d x %d), exiting", g.m, g.n);
    if (BM == 128) launch_tile<128, 128>(p, g.ta, g.tb, grid, s);
    else launch_tile<64, 64>(p, g.ta, g.tb, grid, s);
    if (ws) {
        size_t total = (size_t)g.m * g.n;
        size_t blocks = (total + kThreads - 1) / kThreads;
        size_t cap = (size_t)sms * 8;
        if (blocks > cap) blocks = cap;
        BLA_CUDA(launch_pdl(splitk_reduce_kernel, dim3((int)blocks), dim3(kThreads), 0, s, 1, p, splits));
        BLA_LAUNCH_CHECK();
        count_launch();
        if (!own_ws) pool_free(ws);   // stream-ordered reuse: the pool only serves this one stream
    }
}

void gemm(const GemmArgs& g, cudaStream_t s) {
    const int path = rt().gemm_path;
    if (path != BLA_GEMM_FP32) {
        const bool forced = path == BLA_GEMM_3XTF32;
        // AUTO: tensor path where a 128x256 tcgen05 tile is not mostly padding, or where a skinny product is long enough that
        // even a mostly empty tile beats the FP32 FMA kernel (the U-Net's attention projections: 16384 x 48 x 256 and their
        // weight gradients 256 x 48 x 16384)
        static double skinny_min = -1.0;   // BLA_TC_SKINNY_MIN: m*n*k from which a skinny product takes the tensor path (tuning probe)
        if (skinny_min < 0) { const char* e = getenv("BLA_TC_SKINNY_MIN"); skinny_min = e ? atof(e) : 1.0e8; }
        const bool worthwhile = g.k >= 64 && ((g.m >= 128 && g.n >= 128) || (g.m >= 16 && g.n >= 16 && (double)g.m * g.n * g.k >= skinny_min));
        if ((forced || worthwhile) && gemm_3xtf32(g, s)) return;
    }
    if (g.mask_out && g.mask_written) *g.mask_written = false;   // only the tensor path writes the bit form of a ReLU mask
    gemm_simt(g, s);
}

}  // namespace bla

extern "C" {

void bla_gemm_ex(int trans_a, int trans_b, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c, int ldc,
                 const bla_epilogue* epi) {
    using namespace bla;
    if (m < 0 || n < 0 || k < 0) die("bla: bla_gemm with negative dimension %d x %d x %d, exiting", m, n, k);
    CallScope sc;
    const size_t a_elems = trans_a ? (size_t)(k - 1) * lda + m : (size_t)(m - 1) * lda + k;
    const size_t b_elems = trans_b ? (size_t)(n - 1) * ldb + k : (size_t)(k - 1) * ldb + n;
    const size_t c_elems = (size_t)(m - 1) * ldc + n;
    GemmArgs g{};
    g.ta = trans_a != 0; g.tb = trans_b != 0;
    g.m = m; g.n = n; g.k = k;
    g.a = sc.in(a, (m && k) ? a_elems : 0); g.lda = lda;
    g.b = sc.in(b, (n && k) ? b_elems : 0); g.ldb = ldb;
    if (epi) {
        g.epi = *epi;
        if (epi->bias_rows) g.epi.bias_rows = sc.in(epi->bias_rows, m);
        if (epi->bias_cols) g.epi.bias_cols = sc.in(epi->bias_cols, n);
        if (epi->gate) g.epi.gate = sc.in(epi->gate, c_elems);
        if (epi->pre_activation) g.epi.pre_activation = sc.out(epi->pre_activation, c_elems);
    }
    g.c = (ldc == n) ? sc.out(c, c_elems) : sc.inout(c, c_elems);   // strided C keeps the gaps
    g.ldc = ldc;
    gemm(g, sc.stream());
}

void bla_gemm(int trans_a, int trans_b, int m, int n, int k, const float* a, int lda, const float* b, int ldb, float* c, int ldc) {
    bla_gemm_ex(trans_a, trans_b, m, n, k, a, lda, b, ldb, c, ldc, nullptr);
}

}  // extern "C"
