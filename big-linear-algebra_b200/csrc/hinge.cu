// hinge.cu -- model/mnist_hinge.c:100-172 as a device-resident full-batch iteration (BASELINE.json configs[1], SURVEY 8(d)
// config 2: "batched 10 x 784 x N GEMV").
//
// The reference walks the training file once per iteration and, per sample and per one-vs-rest model, does a 1 x 784 . 784 x 1
// matrix_multiply and a 784-element gradient loop.  Over a device-resident store (data.cu, sample-major [N][784]) one iteration is
// ONE pass over the samples: a persistent CTA per SM pulls tiles of 28 samples (one contiguous 88 KB block: a single cp.async.bulk,
// double buffered behind an mbarrier) into shared memory and, from that one copy,
//   A  takes the ten dot products of every sample (a warp = 4 samples, the ten weight rows in shared memory, every 128-bit weight
//      read serves 4 dots) and turns them into the gradient coefficients m = -y/255 where (1 - y (w.x/255)) < 1, else 0
//      (:136-147, the reference's condition, kept as written)
//   B  accumulates G [10][784] += m^T . x: thread = 4 features x 10 models in registers, kept across the CTA's tiles
// then per-CTA partials, a deterministic fold, (data parallel: one all-reduce of the 10 x 784 gradient and the sample count) and
// the update: grad = (first 196 floats cleared) + sum;  norm = |grad| / N;  grad *= lr;  w += grad   (:124-127, :154-160).
// The sample matrix is read from HBM exactly once (round 1 made two passes: 1.1 TB/s of algorithmic bytes); at 20 flop per sample
// byte the pass sits on the FP32 FMA / HBM ridge.  The reference's partial clear of the gradient buffer (memset of 784 BYTES,
// SURVEY D8) is reproduced.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/bla.h"
#include "kernels.h"
#include "runtime.h"

using namespace bla;

const float* bla_mnist_x_device(const bla_mnist* m);
const float* bla_mnist_y_device(const bla_mnist* m);

struct bla_hinge {
    int features, classes, max_examples;
    float* w;        // [16][features], rows >= classes are zero
    float* grad;     // [classes][features], persistent (the reference only clears part of it per iteration)
    float* part;     // [ctas][16][features] per-CTA partial gradients of pass 2
    int ctas;
    float* fold;     // [kFold][16][features]: the per-CTA partials summed in kFold groups
    float* fold1;    // [16][features]: the local gradient of this iteration; element [classes][0] carries the local sample count
    float* norms;    // [classes]
};

namespace {
constexpr int kPad = 16, kThreads = 256;
constexpr int kPassThreads = 512;             // the one-pass kernel: 16 warps on the SM's four schedulers

constexpr int kTile = 28;                    // samples per tile: 14 warps x 2 samples in phase A
constexpr int kSpw = 2;                      // samples per warp in phase A
constexpr int kMaskPad = 12;                 // ten coefficients per sample, padded to whole float4

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "HW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HW_DONE;\n\t"
        "bra HW_LOOP;\n\t"
        "HW_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// One pass: tiles [tile0, tile1) of kTile samples each belong to this CTA; part [ctas][16][features] receives its partial gradient.
template <int NC>
__global__ void __launch_bounds__(kPassThreads, 1) hinge_onepass_kernel(const float* __restrict__ x, const float* __restrict__ labels,
                                                                    const float* __restrict__ w, int n, int features, int tiles_per_cta,
                                                                    float* __restrict__ part) {
    extern __shared__ __align__(128) unsigned char hs[];
    const int F = features, f4 = F / 4;
    float* ws = reinterpret_cast<float*>(hs);                                   // [NC][F]
    float* xs = ws + (size_t)NC * F;                                            // [2][kTile][F]
    float* ms = xs + (size_t)2 * kTile * F;                                     // [kTile][kMaskPad]
    const uint32_t bars = smem_u32(ms + kTile * kMaskPad);                      // two mbarriers
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntiles = (n + kTile - 1) / kTile;
    const int tile0 = blockIdx.x * tiles_per_cta, tile1 = min(ntiles, tile0 + tiles_per_cta);
    auto issue = [&](int tile, int buf) {                                       // one thread: the tile is ONE contiguous block of the store
        const int j0 = tile * kTile, cnt = min(kTile, n - j0);
        const uint32_t bytes = (uint32_t)cnt * F * sizeof(float);
        mbar_expect_tx(bars + 8u * buf, bytes);
        bulk_load(smem_u32(xs + (size_t)buf * kTile * F), x + (size_t)j0 * F, bytes, bars + 8u * buf);
    };
    if (threadIdx.x == 0) {
        mbar_init(bars, 1); mbar_init(bars + 8u, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (tile0 < tile1) issue(tile0, 0);
        if (tile0 + 1 < tile1) issue(tile0 + 1, 1);
    }
    for (int e = threadIdx.x; e < NC * f4; e += kPassThreads) reinterpret_cast<float4*>(ws)[e] = __ldg(reinterpret_cast<const float4*>(w) + e);
    // phase B: thread (chunk, half) owns features 4*chunk..4*chunk+3 of every model over the tile's samples of parity `half`
    const int chunk = threadIdx.x < f4 ? threadIdx.x : threadIdx.x - f4, half = threadIdx.x < f4 ? 0 : 1;
    const bool in_b = threadIdx.x < 2 * f4;
    float4 acc[NC];
#pragma unroll
    for (int p = 0; p < NC; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    uint32_t phase[2] = {0u, 0u};
    for (int tile = tile0; tile < tile1; ++tile) {
        const int buf = (tile - tile0) & 1;
        const int j0 = tile * kTile, cnt = min(kTile, n - j0);
        const float* xt = xs + (size_t)buf * kTile * F;
        mbar_wait(bars + 8u * buf, phase[buf]);
        phase[buf] ^= 1u;
        // ---- A: dot products and coefficients of samples kSpw*warp .. of the tile ----
        if (kSpw * warp < cnt) {
            float d[kSpw * NC];                                                 // d[s * NC + p]
#pragma unroll
            for (int i = 0; i < kSpw * NC; ++i) d[i] = 0.f;
            for (int c4 = lane; c4 < f4; c4 += 32) {
                float4 xv[kSpw];
#pragma unroll
                for (int s_ = 0; s_ < kSpw; ++s_)
                    xv[s_] = kSpw * warp + s_ < cnt ? reinterpret_cast<const float4*>(xt + (size_t)(kSpw * warp + s_) * F)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int p = 0; p < NC; ++p) {
                    const float4 wv = reinterpret_cast<const float4*>(ws + p * F)[c4];
#pragma unroll
                    for (int s_ = 0; s_ < kSpw; ++s_) {
                        float t = d[s_ * NC + p];
                        t = fmaf(xv[s_].x, wv.x, t); t = fmaf(xv[s_].y, wv.y, t); t = fmaf(xv[s_].z, wv.z, t); t = fmaf(xv[s_].w, wv.w, t);
                        d[s_ * NC + p] = t;
                    }
                }
            }
            // kSpw*NC = 20 sums over the warp with halving exchanges: after the two steps a lane holds 5 values, original index
            // 4i + 2*b3 + b4 (b_k = bit k of the lane); three plain steps finish them
            static_assert(kSpw == 2 && NC % 2 == 0, "the halving reduction is written for 2 samples per warp");
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const bool up = lane & 16;
                const float send = up ? d[2 * i] : d[2 * i + 1], keep = up ? d[2 * i + 1] : d[2 * i];
                d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < NC / 2; ++i) {
                const bool up = lane & 8;
                const float send = up ? d[2 * i] : d[2 * i + 1], keep = up ? d[2 * i + 1] : d[2 * i];
                d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int i = 0; i < NC / 2; ++i) {
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 4);
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 2);
                d[i] += __shfl_xor_sync(0xffffffffu, d[i], 1);
            }
            if ((lane & 7) == 0) {
                const int low = ((lane >> 3) & 1) * 2 + ((lane >> 4) & 1);
#pragma unroll
                for (int i = 0; i < NC / 2; ++i) {
                    const int idx = 4 * i + low, s_ = idx / NC, p = idx - s_ * NC;
                    float m = 0.f;
                    if (kSpw * warp + s_ < cnt) {
                        const float y = ((int)labels[j0 + kSpw * warp + s_] == p) ? 1.f : -1.f;   // mnist_hinge.c:133-134
                        const float val = 1.f - y * (d[i] * (1 / 255.0F));                       // :136, :141
                        if (val < 1.f) m = -y * (1 / 255.0F);                                    // :146-147 with the 1/255 of :136 folded in
                    }
                    ms[(kSpw * warp + s_) * kMaskPad + p] = m;
                }
            }
        }
        __syncthreads();
        // ---- B: G[p][4c..4c+3] += m[s][p] * x[s][4c..4c+3] over the samples of this thread's parity ----
        if (in_b) {
#pragma unroll 2
            for (int s_ = half; s_ < cnt; s_ += 2) {
                const float4 xv = reinterpret_cast<const float4*>(xt + (size_t)s_ * F)[chunk];
                float m[kMaskPad];
#pragma unroll
                for (int q = 0; q < kMaskPad / 4; ++q) {
                    const float4 mv = reinterpret_cast<const float4*>(ms + s_ * kMaskPad)[q];
                    m[4 * q] = mv.x; m[4 * q + 1] = mv.y; m[4 * q + 2] = mv.z; m[4 * q + 3] = mv.w;
                }
#pragma unroll
                for (int p = 0; p < NC; ++p) {
                    acc[p].x = fmaf(m[p], xv.x, acc[p].x); acc[p].y = fmaf(m[p], xv.y, acc[p].y);
                    acc[p].z = fmaf(m[p], xv.z, acc[p].z); acc[p].w = fmaf(m[p], xv.w, acc[p].w);
                }
            }
        }
        __syncthreads();                                                        // every reader of this buffer (and of ms) is done
        if (threadIdx.x == 0 && tile + 2 < tile1) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic reads before the async-proxy overwrite
            issue(tile + 2, buf);
        }
    }
    // the two parity halves meet in shared memory (the tile buffers are free now), then one partial per CTA
    __syncthreads();
    float4* comb = reinterpret_cast<float4*>(xs);
    if (in_b && half == 1) {
#pragma unroll
        for (int p = 0; p < NC; ++p) comb[p * f4 + chunk] = acc[p];
    }
    __syncthreads();
    if (in_b && half == 0) {
#pragma unroll
        for (int p = 0; p < NC; ++p) {
            const float4 o = comb[p * f4 + chunk];
            reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * kPad + p) * F)[chunk] =
                make_float4(acc[p].x + o.x, acc[p].y + o.y, acc[p].z + o.z, acc[p].w + o.w);
        }
    }
}

// per-CTA partial gradients summed in CTA order, two deterministic levels: [ctas] -> [kFold] groups (enough CTAs to pull the
// partials at bandwidth: one level on 13 CTAs took 44 us), then [kFold] -> 1 inside the update kernel (or, data parallel, by a
// second small launch ahead of the all-reduce)
constexpr int kFold = 8;
__global__ void __launch_bounds__(kThreads) hinge_fold_kernel(const float* __restrict__ part, int nparts, size_t count4, float* __restrict__ out,
                                                              float count, size_t count_at) {
    const size_t e = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= count4) return;
    const int per = (nparts + (int)gridDim.y - 1) / (int)gridDim.y, c0 = blockIdx.y * per, c1 = min(nparts, c0 + per);   // gridDim.y groups
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int c = c0; c < c1; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(part) + (size_t)c * count4 + e);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (count_at / 4 == e && gridDim.y == 1) (&acc.x)[count_at & 3] = count;   // the local sample count rides in an unused row
    reinterpret_cast<float4*>(out)[(size_t)blockIdx.y * count4 + e] = acc;
}

// one CTA per model: the gradient buffer's first `cleared` floats start from zero, the rest accumulates (:126); norm, scale, add
__global__ void __launch_bounds__(kThreads) hinge_update_kernel(float* __restrict__ w, float* __restrict__ grad, const float* __restrict__ part,
                                                                int nparts, int features, int cleared, const float* __restrict__ count, float n_local, float lr,
                                                                float* __restrict__ norms) {
    __shared__ float red[kThreads / 32];
    const int p = blockIdx.x;
    float tot = 0.f;
    for (int k = threadIdx.x; k < features; k += kThreads) {
        float fresh = 0.f;
        for (int c = 0; c < nparts; ++c) fresh += part[((size_t)c * kPad + p) * features + k];      // fixed order: deterministic
        const float g = (k < cleared ? 0.f : grad[(size_t)p * features + k]) + fresh;
        tot = fmaf(g, g, tot);
        grad[(size_t)p * features + k] = g * lr;                          // matrix_scale(&gradients[j], learn_rate), :159
        w[(size_t)p * features + k] += g * lr;                            // matrix_add, :160
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) t += red[i];
        norms[p] = sqrtf(t) / (count ? *count : n_local);                                 // :156 (the sample count of ALL ranks when data parallel)
    }
}
}  // namespace

namespace bla {
bool comm_active();
}
extern "C" void bla_allreduce_sum_f32(float* buf, size_t n);

extern "C" {

bla_hinge* bla_hinge_create(int features, int classes, int max_examples) {
    rt();
    if (classes != 10) die("bla: bla_hinge is built for the reference's 10 one-vs-rest models (mnist_hinge.c:103), exiting");
    if (features % 4) die("bla: bla_hinge needs a multiple of 4 features, exiting");
    bla_hinge* h = new bla_hinge{features, classes, max_examples, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr};
    h->ctas = rt().num_sms * 4;
    h->w = (float*)pool_alloc(kDevice, (size_t)kPad * features * sizeof(float));
    h->grad = (float*)pool_alloc(kDevice, (size_t)classes * features * sizeof(float));
    h->part = (float*)pool_alloc(kDevice, (size_t)h->ctas * kPad * features * sizeof(float));
    h->fold = (float*)pool_alloc(kDevice, (size_t)kFold * kPad * features * sizeof(float));
    h->fold1 = (float*)pool_alloc(kDevice, (size_t)kPad * features * sizeof(float));
    h->norms = (float*)pool_alloc(kDevice, kPad * sizeof(float));
    cudaStream_t s = rt().stream;
    BLA_CUDA(cudaMemsetAsync(h->w, 0, (size_t)kPad * features * sizeof(float), s));
    BLA_CUDA(cudaMemsetAsync(h->grad, 0, (size_t)classes * features * sizeof(float), s));
    BLA_CUDA(cudaMemsetAsync(h->part, 0, (size_t)h->ctas * kPad * features * sizeof(float), s));   // rows >= classes stay zero
    return h;
}

void bla_hinge_destroy(bla_hinge* h) {
    if (!h) return;
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
    for (float* p : {h->w, h->grad, h->part, h->fold, h->fold1, h->norms}) pool_free(p);
    delete h;
}

// weights [classes][features]: the ten 1 x 784 rows of data/mnist_hinge/weights_<p>.csv (mnist_hinge.c:104-109)
void bla_hinge_set_weights(bla_hinge* h, const float* w) {
    BLA_CUDA(cudaMemcpyAsync(h->w, w, (size_t)h->classes * h->features * sizeof(float), cudaMemcpyDefault, rt().stream));
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
}
void bla_hinge_get_weights(bla_hinge* h, float* w) {
    BLA_CUDA(cudaMemcpyAsync(w, h->w, (size_t)h->classes * h->features * sizeof(float), cudaMemcpyDefault, rt().stream));
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
}

// One iteration of mnist_hinge.c:123-166 over every example of the store.  norms_host (may be NULL: the call then stays
// asynchronous) receives |gradient_p| / N per model, the numbers the reference prints every tenth iteration (:156-158).
void bla_hinge_iteration(bla_hinge* h, bla_mnist* data, float learn_rate, float* norms_host) {
    const int n = bla_mnist_num_examples(data), F = h->features;
    if (n > h->max_examples) die("bla: bla_hinge_iteration over %d examples exceeds max_examples %d, exiting", n, h->max_examples);
    if (n <= 0) return;
    cudaStream_t s = rt().stream;
    const float* X = bla_mnist_x_device(data);
    if (h->classes != 10) die("bla: bla_hinge_iteration is built for the reference's 10 one-vs-rest models, exiting");
    const size_t smem = ((size_t)h->classes * F + (size_t)2 * kTile * F + (size_t)kTile * kMaskPad) * sizeof(float) + 64;
    if (smem > 227 * 1024) die("bla: bla_hinge with %d features does not fit the one-pass kernel's shared memory, exiting", F);
    static bool attr = false;
    if (!attr) { BLA_CUDA(cudaFuncSetAttribute(hinge_onepass_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
    const int ntiles = ceil_div(n, kTile);
    int ctas = std::min(rt().num_sms, ntiles);
    const int tiles_per_cta = ceil_div(ntiles, ctas);
    ctas = ceil_div(ntiles, tiles_per_cta);
    hinge_onepass_kernel<10><<<ctas, kPassThreads, smem, s>>>(X, bla_mnist_y_device(data), h->w, n, F, tiles_per_cta, h->part);
    BLA_LAUNCH_CHECK();
    // per-CTA partials -> kFold group sums, in CTA order (deterministic); rows >= classes of a partial are never written by the pass
    // (zero since creation)
    const size_t count4 = (size_t)kPad * F / 4;
    const unsigned fold_blocks = (unsigned)ceil_div((long long)count4, kThreads);
    hinge_fold_kernel<<<dim3(fold_blocks, kFold), kThreads, 0, s>>>(h->part, ctas, count4, h->fold, 0.f, 0);
    BLA_LAUNCH_CHECK();
    if (comm_active()) {
        // data parallel: every rank has walked its own shard of the samples; the gradient is a plain sum over samples
        // (mnist_hinge.c:129-150), so one all-reduce of the [classes + 1][features] block -- element [classes][0] carries the sample
        // count -- gives every rank the full-batch gradient and N
        hinge_fold_kernel<<<dim3(fold_blocks, 1), kThreads, 0, s>>>(h->fold, kFold, count4, h->fold1, (float)n, (size_t)h->classes * F);
        BLA_LAUNCH_CHECK();
        count_launch();
        bla_allreduce_sum_f32(h->fold1, (size_t)(h->classes + 1) * F);
        hinge_update_kernel<<<h->classes, kThreads, 0, s>>>(h->w, h->grad, h->fold1, 1, F, 784 / 4, h->fold1 + (size_t)h->classes * F, (float)n,
                                                            learn_rate, h->norms);
    } else {
        hinge_update_kernel<<<h->classes, kThreads, 0, s>>>(h->w, h->grad, h->fold, kFold, F, 784 / 4, nullptr, (float)n, learn_rate,
                                                            h->norms);   // memset(.., 784): 196 floats
    }
    BLA_LAUNCH_CHECK();
    count_launch(3);
    if (norms_host) {
        BLA_CUDA(cudaMemcpyAsync(norms_host, h->norms, h->classes * sizeof(float), cudaMemcpyDeviceToHost, s));
        rt().d2h_bytes += h->classes * sizeof(float);
        BLA_CUDA(cudaStreamSynchronize(s));
    }
}

}  // extern "C"
