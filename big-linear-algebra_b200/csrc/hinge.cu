// hinge.cu -- model/mnist_hinge.c:100-172 as a device-resident full-batch iteration (BASELINE.json configs[1], SURVEY 8(d)
// config 2: "batched 10 x 784 x N GEMV", HBM-bound at ~5 flop per byte).
//
// The reference walks the training file once per iteration and, per sample and per one-vs-rest model, does a 1 x 784 . 784 x 1
// matrix_multiply and a 784-element gradient loop.  Over a device-resident store (data.cu, sample-major [N][784]) one iteration is
//   pass 1  M [N][16] = -y/255 where (1 - y (w.x/255)) < 1 else 0   (:136-147, the reference's condition, kept as written): the ten
//           weight rows sit in shared memory, a warp takes 4 samples at a time so that every 128-bit weight read serves 4 dots
//   pass 2  G [10][784] = M^T . X: a CTA streams a slice of the samples, thread = 4 features x 10 models in registers,
//           per-CTA partials
//   update  grad = (first 196 floats cleared) + sum of the partials;  norm = |grad| / N;  grad *= lr;  w += grad   (:124-127, :154-160)
// -- the sample matrix is read from HBM exactly twice (the gradient needs the whole dot product of its sample first), everything
// else is a few MB; the reference's partial clear of the gradient buffer (memset of 784 BYTES) is reproduced.
#include <cmath>
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
    float* fold;     // [kFold][16][features]
    int ctas;
    float* mask;     // [max_examples][16]
    float* norms;    // [classes]
};

namespace {
constexpr int kPad = 16, kThreads = 256;

// pass 1: one warp = 4 samples at a time; lane owns the float4 feature chunks lane, lane + 32, ...
template <int NC>
__global__ void __launch_bounds__(kThreads, 3) hinge_mask_kernel(const float* __restrict__ x, const float* __restrict__ labels, const float* __restrict__ w,
                                                                 int n, int features, float* __restrict__ mask) {
    constexpr int classes = NC;
    extern __shared__ __align__(16) float ws[];            // [classes][features]
    for (int e = threadIdx.x; e < classes * features / 4; e += kThreads)
        reinterpret_cast<float4*>(ws)[e] = reinterpret_cast<const float4*>(w)[e];
    __syncthreads();
    const int f4 = features / 4, lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * kThreads + threadIdx.x) >> 5, nwarps = (gridDim.x * kThreads) >> 5;
    for (int j0 = warp * 4; j0 < n; j0 += nwarps * 4) {
        float dots[4][NC];
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_)
#pragma unroll
            for (int p = 0; p < NC; ++p) dots[s_][p] = 0.f;
        for (int c4 = lane; c4 < f4; c4 += 32) {
            float4 xv[4];
#pragma unroll
            for (int s_ = 0; s_ < 4; ++s_)
                xv[s_] = j0 + s_ < n ? __ldg(reinterpret_cast<const float4*>(x + (size_t)(j0 + s_) * features) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int p = 0; p < NC; ++p) {
                const float4 wv = reinterpret_cast<const float4*>(ws + p * features)[c4];
#pragma unroll
                for (int s_ = 0; s_ < 4; ++s_)
                    dots[s_][p] += (xv[s_].x * wv.x + xv[s_].y * wv.y) + (xv[s_].z * wv.z + xv[s_].w * wv.w);
            }
        }
#pragma unroll
        for (int s_ = 0; s_ < 4; ++s_)
#pragma unroll
            for (int p = 0; p < NC; ++p) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dots[s_][p] += __shfl_xor_sync(0xffffffffu, dots[s_][p], o);
            }
        if (lane < kPad) {
#pragma unroll
            for (int s_ = 0; s_ < 4; ++s_) {
                if (j0 + s_ >= n) break;
                float m = 0.f;
                float d = 0.f;
#pragma unroll
                for (int p = 0; p < NC; ++p) d = lane == p ? dots[s_][p] : d;
                if (lane < classes) {
                    const float y = ((int)labels[j0 + s_] == lane) ? 1.f : -1.f;    // mnist_hinge.c:133-134
                    const float val = 1.f - y * (d * (1 / 255.0F));                   // :136, :141
                    if (val < 1.f) m = -y * (1 / 255.0F);                             // :146-147 with the 1/255 of :136 folded in
                }
                mask[(size_t)(j0 + s_) * kPad + lane] = m;
            }
        }
    }
}

// pass 2: CTA = a slice of the samples, thread t < features/4 owns 4 features x `classes` accumulators; part [ctas][16][features]
template <int NC>
__global__ void __launch_bounds__(kThreads) hinge_grad_kernel(const float* __restrict__ x, const float* __restrict__ mask, int n, int features,
                                                              int per_cta, float* __restrict__ part) {
    const int f4 = features / 4;
    const int jbeg = blockIdx.x * per_cta, jend = min(n, jbeg + per_cta);
    for (int t = threadIdx.x; t < f4; t += kThreads) {
        float4 acc[NC];
#pragma unroll
        for (int p = 0; p < NC; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int j = jbeg; j < jend; ++j) {
            const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)j * features) + t);
            float m[kPad];
#pragma unroll
            for (int q = 0; q < (NC + 3) / 4; ++q) {
                const float4 mv = __ldg(reinterpret_cast<const float4*>(mask + (size_t)j * kPad) + q);
                m[4 * q] = mv.x; m[4 * q + 1] = mv.y; m[4 * q + 2] = mv.z; m[4 * q + 3] = mv.w;
            }
#pragma unroll
            for (int p = 0; p < NC; ++p) {
                acc[p].x = fmaf(m[p], xv.x, acc[p].x); acc[p].y = fmaf(m[p], xv.y, acc[p].y);
                acc[p].z = fmaf(m[p], xv.z, acc[p].z); acc[p].w = fmaf(m[p], xv.w, acc[p].w);
            }
        }
#pragma unroll
        for (int p = 0; p < NC; ++p) reinterpret_cast<float4*>(part + ((size_t)blockIdx.x * kPad + p) * features)[t] = acc[p];
    }
}

// partial gradients of pass 2 folded in two deterministic levels: [ctas] -> [kFold] here, [kFold] -> 1 in the update kernel
constexpr int kFold = 8;
__global__ void __launch_bounds__(kThreads) hinge_fold_kernel(const float* __restrict__ part, int nparts, size_t count4, float* __restrict__ out) {
    const size_t e = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (e >= count4) return;
    const int per = (nparts + kFold - 1) / kFold, c0 = blockIdx.y * per, c1 = min(nparts, c0 + per);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int c = c0; c < c1; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(part) + (size_t)c * count4 + e);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[(size_t)blockIdx.y * count4 + e] = acc;
}

// one CTA per model: the gradient buffer's first `cleared` floats start from zero, the rest accumulates (:126); norm, scale, add
__global__ void __launch_bounds__(kThreads) hinge_update_kernel(float* __restrict__ w, float* __restrict__ grad, const float* __restrict__ part,
                                                                int nparts, int features, int cleared, int n, float lr, float* __restrict__ norms) {
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
        norms[p] = sqrtf(t) / (float)n;                                   // :156
    }
}
}  // namespace

extern "C" {

bla_hinge* bla_hinge_create(int features, int classes, int max_examples) {
    rt();
    if (classes != 10) die("bla: bla_hinge is built for the reference's 10 one-vs-rest models (mnist_hinge.c:103), exiting");
    if (features % 4) die("bla: bla_hinge needs a multiple of 4 features, exiting");
    bla_hinge* h = new bla_hinge{features, classes, max_examples, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr};
    h->ctas = rt().num_sms * 4;
    h->w = (float*)pool_alloc(kDevice, (size_t)kPad * features * sizeof(float));
    h->grad = (float*)pool_alloc(kDevice, (size_t)classes * features * sizeof(float));
    h->part = (float*)pool_alloc(kDevice, (size_t)h->ctas * kPad * features * sizeof(float));
    h->fold = (float*)pool_alloc(kDevice, (size_t)kFold * kPad * features * sizeof(float));
    h->mask = (float*)pool_alloc(kDevice, (size_t)max_examples * kPad * sizeof(float));
    h->norms = (float*)pool_alloc(kDevice, kPad * sizeof(float));
    cudaStream_t s = rt().stream;
    BLA_CUDA(cudaMemsetAsync(h->w, 0, (size_t)kPad * features * sizeof(float), s));
    BLA_CUDA(cudaMemsetAsync(h->grad, 0, (size_t)classes * features * sizeof(float), s));
    return h;
}

void bla_hinge_destroy(bla_hinge* h) {
    if (!h) return;
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
    for (float* p : {h->w, h->grad, h->part, h->fold, h->mask, h->norms}) pool_free(p);
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
    const size_t smem = (size_t)h->classes * F * sizeof(float);
    if (h->classes != 10) die("bla: bla_hinge_iteration is built for the reference's 10 one-vs-rest models, exiting");
    static bool attr = false;
    if (!attr) { BLA_CUDA(cudaFuncSetAttribute(hinge_mask_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    if (smem > 64 * 1024) die("bla: bla_hinge weights (%zu bytes) do not fit the kernel's shared memory, exiting", smem);
    hinge_mask_kernel<10><<<rt().num_sms * 6, kThreads, smem, s>>>(X, bla_mnist_y_device(data), h->w, n, F, h->mask);
    BLA_LAUNCH_CHECK();
    int ctas = h->ctas;
    const int per_cta = ceil_div(n, ctas);
    ctas = ceil_div(n, per_cta);
    hinge_grad_kernel<10><<<ctas, kThreads, 0, s>>>(X, h->mask, n, F, per_cta, h->part);
    BLA_LAUNCH_CHECK();
    const size_t count4 = (size_t)kPad * F / 4;
    hinge_fold_kernel<<<dim3((unsigned)ceil_div((long long)count4, kThreads), kFold), kThreads, 0, s>>>(h->part, ctas, count4, h->fold);
    BLA_LAUNCH_CHECK();
    hinge_update_kernel<<<h->classes, kThreads, 0, s>>>(h->w, h->grad, h->fold, kFold, F, 784 / 4, n, learn_rate, h->norms);   // memset(.., 784): 196 floats
    BLA_LAUNCH_CHECK();
    count_launch(4);
    if (norms_host) {
        BLA_CUDA(cudaMemcpyAsync(norms_host, h->norms, h->classes * sizeof(float), cudaMemcpyDeviceToHost, s));
        rt().d2h_bytes += h->classes * sizeof(float);
        BLA_CUDA(cudaStreamSynchronize(s));
    }
}

}  // extern "C"
