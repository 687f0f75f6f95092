// conv_implicit.cu -- batched, device-resident conv2d forward / wgrad / dgrad as IMPLICIT GEMMs
// (SURVEY.md K10/K11): the im2col matrix of lib/conv.c:8-77 is never materialised; its elements
// are gathered from the NCHW tensors while the B tile of the GEMM is staged into shared memory.
//
//   fprop:  y [img][F][P]   = W[F][C*k*k] . col(x)^T            M = F, N = imgs*P,   K = C*k*k
//   wgrad:  dW[F][C*k*k]    = sum_img dy[img][F][P] . col(x)    M = F, N = C*k*k,    K = imgs*P   (split-K)
//   dgrad:  dx[img][C][H*W] = W^T[C][F*k*k] . col'(dy)          M = C, N = imgs*H*W, K = F*k*k
// with SAME padding exactly as the reference computes it (lib/conv.c:12-24) and, for dgrad, the
// true adjoint for any stride (the reference's col2im is only valid for stride 1).
// FP32 FMA on the SIMT pipe with the same 128x128x16 / 64x64x16 register-tiled core as gemm_simt.cu
// (<= 1e-5 vs the double reference); k ordering inside a dot product is the reference's
// (c, ki, kj).  Roofline: FP32 FMA peak; algorithmic work 2*M*N*K flop.
#include <cstdint>

#include "../../include/bla.h"
#include "kernels.h"

extern "C" int bla_tc_available(void);
#include "runtime.h"

namespace bla {

struct ConvGeom;   // api_conv.cu

namespace {

constexpr int BK = 16;
constexpr int kThreads = 256;

enum Mode { kFprop = 0, kWgrad = 1, kDgrad = 2 };

struct ConvP {
    int imgs, C, H, W, F, k, stride, Ho, Wo, pad_top, pad_left;
    int M, N, K;           // GEMM view
    const float* a;        // fprop: W [F][C*k*k]; wgrad: dy [img][F][P]; dgrad: Wt [C][F*k*k]
    const float* src;      // fprop/wgrad: x [img][C][H][W]; dgrad: dy [img][F][Ho][Wo]
    float* out;            // fprop: y [img][F][P]; wgrad: dW [F][C*k*k] (or partials); dgrad: dx [img][C][H*W]
    int k_chunk;           // split-K (wgrad)
    float* partial;
};

// element (row k, column n) of the implicit B matrix
template <int MODE>
__device__ __forceinline__ float b_elem(const ConvP& p, int kk, int n) {
    if (MODE == kFprop || MODE == kWgrad) {
        // fprop: kk = q = (c, ki, kj), n = (img, oi, oj);  wgrad: kk = (img, oi, oj), n = q
        const int q = MODE == kFprop ? kk : n;
        const int pix = MODE == kFprop ? n : kk;
        const int k2 = p.k * p.k;
        const int c = q / k2, r = q - c * k2;
        const int ki = r / p.k, kj = r - ki * p.k;
        const int P = p.Ho * p.Wo;
        const int img = pix / P, pp = pix - img * P;
        const int oi = pp / p.Wo, oj = pp - oi * p.Wo;
        const int ih = oi * p.stride + ki - p.pad_top, iw = oj * p.stride + kj - p.pad_left;
        if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) return 0.f;
        return __ldg(p.src + (((size_t)img * p.C + c) * p.H + ih) * p.W + iw);
    } else {
        // dgrad: kk = (f, ki, kj), n = (img, ih, iw): dy[img][f][(ih + pt - ki)/s][(iw + pl - kj)/s] when divisible
        const int k2 = p.k * p.k;
        const int f = kk / k2, r = kk - f * k2;
        const int ki = r / p.k, kj = r - ki * p.k;
        const int HW = p.H * p.W;
        const int img = n / HW, pp = n - img * HW;
        const int ih = pp / p.W, iw = pp - ih * p.W;
        const int th = ih + p.pad_top - ki, tw = iw + p.pad_left - kj;
        if (th < 0 || tw < 0) return 0.f;
        const int oi = th / p.stride, oj = tw / p.stride;
        if (oi * p.stride != th || oj * p.stride != tw || oi >= p.Ho || oj >= p.Wo) return 0.f;
        return __ldg(p.src + (((size_t)img * p.F + f) * p.Ho + oi) * p.Wo + oj);
    }
}

// element (row i, column kk) of A
template <int MODE>
__device__ __forceinline__ float a_elem(const ConvP& p, int i, int kk) {
    if (MODE == kWgrad) {   // dy [img][F][P], kk = (img, pixel)
        const int P = p.Ho * p.Wo;
        const int img = kk / P, pp = kk - img * P;
        return __ldg(p.a + ((size_t)img * p.F + i) * P + pp);
    }
    return __ldg(p.a + (size_t)i * p.K + kk);
}

template <int MODE>
__device__ __forceinline__ void c_store(const ConvP& p, int i, int j, float v, int z) {
    if (MODE == kWgrad) {
        float* dst = p.partial ? p.partial + (size_t)z * p.M * p.N : p.out;
        dst[(size_t)i * p.N + j] = v;
    } else {
        // batched planes: out[img][i][pixel], j = (img, pixel)
        const int PP = MODE == kFprop ? p.Ho * p.Wo : p.H * p.W;
        const int img = j / PP, pp = j - img * PP;
        p.out[((size_t)img * p.M + i) * PP + pp] = v;
    }
}

template <int BM, int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 2) conv_implicit_kernel(const ConvP p) {
    constexpr int TM = BM / 16, TN = BN / 16;
    constexpr int GM = TM / 4, GN = TN / 4;
    constexpr int A_E = BM * BK / kThreads, B_E = BN * BK / kThreads;   // elements per thread per tile
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * p.k_chunk;
    const int kend = min(p.K, kbeg + p.k_chunk);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float ra[A_E], rb[B_E];
    // A tile: thread covers row (tid / BK .. ) fixed k column; B tile: thread covers one k row, consecutive n
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < A_E; ++e) {
            const int idx = tid + e * kThreads;
            const int row = idx / BK, kk = idx % BK;       // consecutive threads walk k: contiguous for the dense A
            const int i = m0 + row, k = k0 + kk;
            ra[e] = (i < p.M && k < kend) ? a_elem<MODE>(p, i, k) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < B_E; ++e) {
            const int idx = tid + e * kThreads;
            const int kk = idx / BN, col = idx % BN;       // consecutive threads walk n: neighbouring pixels
            const int k = k0 + kk, n = n0 + col;
            rb[e] = (k < kend && n < p.N) ? b_elem<MODE>(p, k, n) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int e = 0; e < A_E; ++e) {
            const int idx = tid + e * kThreads;
            As[buf][idx % BK][idx / BK] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < B_E; ++e) {
            const int idx = tid + e * kThreads;
            Bs[buf][idx / BN][idx % BN] = rb[e];
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
        if (more) fetch(k0 + BK);
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
#pragma unroll
    for (int gi = 0; gi < GM; ++gi)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = m0 + gi * (BM / GM) + ty * 4 + ii;
            if (i >= p.M) continue;
#pragma unroll
            for (int gj = 0; gj < GN; ++gj)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = n0 + gj * (BN / GN) + tx * 4 + jj;
                    if (j < p.N) c_store<MODE>(p, i, j, acc[4 * gi + ii][4 * gj + jj], blockIdx.z);
                }
        }
}

__global__ void __launch_bounds__(256) conv_splitk_reduce(const float* __restrict__ partial, int splits, size_t count, float* out) {
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < count; e += (size_t)gridDim.x * 256) {
        float s = 0.f;
        for (int z = 0; z < splits; ++z) s += partial[(size_t)z * count + e];
        out[e] = s;
    }
}

// Wt[c][(f, ki, kj)] = W[f][c][ki][kj]
__global__ void __launch_bounds__(256) permute_weights_dgrad(const float* __restrict__ w, float* wt, int F, int C, int k2) {
    const size_t total = (size_t)F * C * k2;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        const int r = (int)(e % k2);
        const int f = (int)((e / k2) % F);
        const int c = (int)(e / ((size_t)k2 * F));
        wt[e] = w[((size_t)f * C + c) * k2 + r];
    }
}

// tensor-path weight layouts over Cp >= C (Fp >= F) channels, zero weights for the padding channels:
//   mode 0  taps [f][(ki*k + kj)*Cp + c] = W[f][c][ki][kj]              (fprop: 16 consecutive k' = 16 channels of one tap)
//   mode 1  flip [c][(ki*k + kj)*Fp + f] = W[f][c][k-1-ki][k-1-kj]      (dgrad as a forward conv of dy; pad = Fp)
//   mode 2  the inverse of mode 0: out [F][C][k2] <- taps [F][(tap, c)]   (weight gradient back to the reference layout)
__global__ void __launch_bounds__(256) permute_weights_taps(const float* __restrict__ w, float* out, int F, int C, int k2, int pad, int mode) {
    const size_t total = mode == 0 ? (size_t)F * k2 * pad : mode == 1 ? (size_t)C * k2 * pad : (size_t)F * C * k2;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        if (mode == 2) {
            const int tap = (int)(e % k2), c = (int)((e / k2) % C), f = (int)(e / ((size_t)C * k2));
            out[e] = w[((size_t)f * k2 + tap) * pad + c];
        } else if (mode == 0) {
            const int c = (int)(e % pad), tap = (int)((e / pad) % k2), f = (int)(e / ((size_t)pad * k2));
            out[e] = c < C ? w[((size_t)f * C + c) * k2 + tap] : 0.f;
        } else {
            const int f = (int)(e % pad), tap = (int)((e / pad) % k2), c = (int)(e / ((size_t)pad * k2));
            out[e] = f < F ? w[((size_t)f * C + c) * k2 + (k2 - 1 - tap)] : 0.f;
        }
    }
}

bool tensor_path_wanted() { return rt().gemm_path != BLA_GEMM_FP32 && bla_tc_available(); }

void launch_permute(const float* w, float* out, int F, int C, int k2, int pad, int mode, size_t total, cudaStream_t s) {
    size_t blocks = (total + 255) / 256, cap = (size_t)rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    permute_weights_taps<<<(int)blocks, 256, 0, s>>>(w, out, F, C, k2, pad, mode);
    BLA_LAUNCH_CHECK();
    count_launch();
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// every job of a table in one launch: grid (blocks per job, jobs)
__global__ void __launch_bounds__(256) permute_weights_batch(const ConvPermuteJob* __restrict__ jobs) {
    const ConvPermuteJob j = jobs[blockIdx.y];
    const size_t total = j.mode == 0 ? (size_t)j.F * j.k2 * j.pad : (size_t)j.C * j.k2 * j.pad;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        if (j.mode == 0) {
            const int c = (int)(e % j.pad), tap = (int)((e / j.pad) % j.k2), f = (int)(e / ((size_t)j.pad * j.k2));
            j.dst[e] = c < j.C ? j.src[((size_t)f * j.C + c) * j.k2 + tap] : 0.f;
        } else {
            const int f = (int)(e % j.pad), tap = (int)((e / j.pad) % j.k2), c = (int)(e / ((size_t)j.pad * j.k2));
            j.dst[e] = f < j.F ? j.src[((size_t)f * j.C + c) * j.k2 + (j.k2 - 1 - tap)] : 0.f;
        }
    }
}

ConvP base_params(int imgs, int C, int H, int W, int F, int k, int stride) {
    ConvP p{};
    p.imgs = imgs; p.C = C; p.H = H; p.W = W; p.F = F; p.k = k; p.stride = stride;
    p.Ho = (H + stride - 1) / stride;
    p.Wo = (W + stride - 1) / stride;
    int pv = (p.Ho - 1) * stride + k - H; if (pv < 0) pv = 0;     // lib/conv.c:12-24
    int ph = (p.Wo - 1) * stride + k - W; if (ph < 0) ph = 0;
    p.pad_top = pv / 2;
    p.pad_left = ph / 2;
    return p;
}

template <int MODE>
void launch(ConvP& p, int splits, cudaStream_t s) {
    const int sms = rt().num_sms;
    const long long tiles128 = (long long)ceil_div(p.M, 128) * ceil_div(p.N, 128) * splits;
    const bool big = p.M > 64 && p.N > 64 && tiles128 >= sms;
    if (big) {
        dim3 grid(ceil_div(p.N, 128), ceil_div(p.M, 128), splits);
        conv_implicit_kernel<128, 128, MODE><<<grid, kThreads, 0, s>>>(p);
    } else {
        dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), splits);
        conv_implicit_kernel<64, 64, MODE><<<grid, kThreads, 0, s>>>(p);
    }
    BLA_LAUNCH_CHECK();
    count_launch();
}

}  // namespace

size_t conv_taps_elems(int F, int C, int k) { return (size_t)F * k * k * round_up(C, 32); }
size_t conv_flip_elems(int F, int C, int k) { return (size_t)C * k * k * round_up(F, 16); }
ConvPermuteJob conv_taps_job(const float* w, float* dst, int F, int C, int k) { return {w, dst, F, C, k * k, round_up(C, 32), 0}; }
ConvPermuteJob conv_flip_job(const float* w, float* dst, int F, int C, int k) { return {w, dst, F, C, k * k, round_up(F, 16), 1}; }
bool conv_tensor_path_wanted() { return tensor_path_wanted(); }

void conv_permute_weights_batch(const ConvPermuteJob* jobs_device, int njobs, cudaStream_t s) {
    if (njobs <= 0) return;
    permute_weights_batch<<<dim3(64, njobs), 256, 0, s>>>(jobs_device);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void conv2d_forward(const float* x, const float* w, float* y, int imgs, int C, int H, int W, int F, int k, int stride, cudaStream_t s,
                    NhwcCache* cache, const float* w_taps, const float* bias, const float* addend) {
    ConvP p = base_params(imgs, C, H, W, F, k, stride);
    p.M = F; p.N = imgs * p.Ho * p.Wo; p.K = C * k * k;
    p.a = w; p.src = x; p.out = y; p.k_chunk = p.K;
    if (p.M <= 0 || p.N <= 0) return;
    if (tensor_path_wanted() && p.N >= 256 && (long long)F * p.K >= 2048) {
        // tcgen05 implicit GEMM: the im2col tile is gathered by 4-D TMA boxes (gemm_tc.cu, conv mode).  Channels are padded to a
        // multiple of 32 inside the NHWC copy (16 would do for the forward pass; 32 lets the weight gradient share the copy), rows
        // of the filter matrix beyond F are zero-filled by TMA: a 3-channel input or a 3-filter output conv runs here too.
        const int Cp = round_up(C, 32);
        float* wt = nullptr;
        if (!w_taps) {   // a caller that convolves with the same weights all step long permutes them once (conv_permute_weights_batch)
            wt = (float*)pool_alloc(kDevice, (size_t)F * k * k * Cp * sizeof(float));
            launch_permute(w, wt, F, C, k * k, Cp, 0, (size_t)F * k * k * Cp, s);
        }
        const bool done = conv2d_tc(x, w_taps ? w_taps : wt, y, imgs, C, Cp, H, W, 1, H, W, F, k, stride, p.pad_top, p.pad_left, cache, s,
                                    bias, addend);
        if (wt) pool_free(wt);
        if (done) return;
    }
    launch<kFprop>(p, 1, s);
    if (bias) k_add_tile_columns(y, imgs * F, p.Ho * p.Wo, bias, 1, s);
    if (addend) k_add(y, addend, (size_t)imgs * F * p.Ho * p.Wo, s);
}

void conv2d_wgrad(const float* x, const float* dy, float* dw, int imgs, int C, int H, int W, int F, int k, int stride, cudaStream_t s,
                  NhwcCache* cache) {
    ConvP p = base_params(imgs, C, H, W, F, k, stride);
    p.M = F; p.N = C * k * k; p.K = imgs * p.Ho * p.Wo;
    p.a = dy; p.src = x; p.out = dw;
    if (p.M <= 0 || p.N <= 0) return;
    if (tensor_path_wanted() && p.K >= 1024 && (long long)F * p.N >= 2048) {
        // tcgen05 implicit GEMM over the output pixels into the tap-major layout, then back to [F][C][k][k]
        const int Cp = round_up(C, 32);
        const size_t total = (size_t)F * Cp * k * k;
        float* taps = (float*)pool_alloc(kDevice, total * sizeof(float));
        bool final_written = false;
        const bool done = conv2d_wgrad_tc(x, dy, taps, dw, &final_written, imgs, C, Cp, H, W, F, k, stride, p.pad_top, p.pad_left, cache, s);
        if (done && !final_written) launch_permute(taps, dw, F, C, k * k, Cp, 2, (size_t)F * C * k * k, s);
        pool_free(taps);
        if (done) return;
    }
    const int sms = rt().num_sms;
    const long long tiles = (long long)ceil_div(p.M, 64) * ceil_div(p.N, 64);
    int splits = 1;
    if (tiles < 2LL * sms && p.K >= 1024) {
        long long want = (2LL * sms + tiles - 1) / tiles, maxs = p.K / 256;
        splits = (int)(want < maxs ? want : maxs);
        if (splits > 256) splits = 256;
        if (splits < 1) splits = 1;
    }
    int chunk = (ceil_div(p.K, splits) + BK - 1) / BK * BK;
    splits = ceil_div(p.K, chunk);
    p.k_chunk = chunk;
    float* ws = nullptr;
    if (splits > 1) {
        ws = (float*)pool_alloc(kDevice, (size_t)splits * p.M * p.N * sizeof(float));
        p.partial = ws;
    }
    launch<kWgrad>(p, splits, s);
    if (ws) {
        const size_t count = (size_t)p.M * p.N;
        size_t blocks = (count + 255) / 256, cap = (size_t)sms * 8;
        if (blocks > cap) blocks = cap;
        conv_splitk_reduce<<<(int)blocks, 256, 0, s>>>(ws, splits, count, dw);
        BLA_LAUNCH_CHECK();
        count_launch();
        pool_free(ws);
    }
}

void conv2d_dgrad(const float* dy, const float* w, float* dx, int imgs, int C, int H, int W, int F, int k, int stride, cudaStream_t s,
                  const float* w_flip) {
    ConvP p = base_params(imgs, C, H, W, F, k, stride);
    p.M = C; p.N = imgs * H * W; p.K = F * k * k;
    if (p.M <= 0 || p.N <= 0) return;
    if (tensor_path_wanted() && p.N >= 256 && (long long)C * p.K >= 2048) {
        // dgrad is a stride-1 convolution of dy -- spread out with stride-1 zeros between its pixels when the forward conv was
        // strided -- with the flipped, transposed filters: dx[i] = sum_ki' dyu[i - (k-1-pad) + ki'] . w[k-1-ki'].  (For stride 2
        // three quarters of the gathered values are zeros; still several times faster than the FP32 FMA kernel.)
        const int Fp = round_up(F, 16);
        float* wf = nullptr;
        if (!w_flip) {
            wf = (float*)pool_alloc(kDevice, (size_t)C * k * k * Fp * sizeof(float));
            launch_permute(w, wf, F, C, k * k, Fp, 1, (size_t)C * k * k * Fp, s);
        }
        const bool done = conv2d_tc(dy, w_flip ? w_flip : wf, dx, imgs, F, Fp, p.Ho, p.Wo, stride, H, W, C, k, 1, k - 1 - p.pad_top,
                                    k - 1 - p.pad_left, nullptr, s);
        if (wf) pool_free(wf);
        if (done) return;
    }
    float* wt = (float*)pool_alloc(kDevice, (size_t)F * C * k * k * sizeof(float));
    const size_t total = (size_t)F * C * k * k;
    size_t blocks = (total + 255) / 256, cap = (size_t)rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    permute_weights_dgrad<<<(int)blocks, 256, 0, s>>>(w, wt, F, C, k * k);
    BLA_LAUNCH_CHECK();
    count_launch();
    p.a = wt; p.src = dy; p.out = dx; p.k_chunk = p.K;
    launch<kDgrad>(p, 1, s);
    pool_free(wt);
}

}  // namespace bla

extern "C" {

// include/bla.h "batched conv2d": x [imgs][C][H][W], w [F][C][k][k], y [imgs][F][Ho][Wo]
void bla_conv2d_forward(const float* x, const float* w, float* y, int imgs, int channels, int height, int width, int filters,
                        int kernel_size, int stride) {
    bla::conv2d_forward(x, w, y, imgs, channels, height, width, filters, kernel_size, stride, bla::rt().stream);
}
void bla_conv2d_wgrad(const float* x, const float* dy, float* dw, int imgs, int channels, int height, int width, int filters,
                      int kernel_size, int stride) {
    bla::conv2d_wgrad(x, dy, dw, imgs, channels, height, width, filters, kernel_size, stride, bla::rt().stream);
}
void bla_conv2d_dgrad(const float* dy, const float* w, float* dx, int imgs, int channels, int height, int width, int filters,
                      int kernel_size, int stride) {
    bla::conv2d_dgrad(dy, w, dx, imgs, channels, height, width, filters, kernel_size, stride, bla::rt().stream);
}

}  // extern "C"
