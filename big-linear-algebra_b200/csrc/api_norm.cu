// api_norm.cu -- lib/norm.h: group normalisation forward / backward.
// HBM-bound: forward 8 B/elem (read x, write y), backward 12 B/elem (read dy, x; write dx).
// One CTA owns one (image, group): the group's slab (group_size channels x H*W, contiguous in the
// [C][H*W] layout) is read ONCE from HBM into shared memory with 128-bit loads, the mean and the
// centred second moment are reduced there (two passes over shared memory, like the reference's two
// passes over DRAM, lib/norm.c:13-37), and the normalised values are written straight out.  Slabs
// larger than shared memory fall back to re-reading global memory (L2-resident at these sizes).
#include <cooperative_groups.h>

#include <cstdint>

#include "../../include/lib/norm.h"
#include "kernels.h"
#include "planes.h"
#include "runtime.h"

using namespace bla;

namespace bla {

namespace {

constexpr int kThreads = 512;
constexpr size_t kSmemSlabFloats = 48 * 1024;   // 192 KB of the 227 KB a CTA may use

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// every thread receives the block total
__device__ __forceinline__ float block_total(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
        t = warp_sum(t);
        if (threadIdx.x == 0) sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    return t;
}

struct GnParams {
    int C, HW, group_size, G;   // per image
    int quirk;
    // fused neighbours (kernels.h GnFuse): forward y = dropout(relu(norm(x))), backward dy is gated by the same two masks
    int relu;
    float drop_rate;
    unsigned long long drop_seed;
    const float* addend;   // backward only: dx += addend (same layout as dx): the residual branch's gradient
};

// forward epilogue on element `idx` of the whole [images][C][HW] tensor
__device__ __forceinline__ float gn_post(float q, size_t idx, const GnParams& p) {
    if (p.relu) q = q > 0.f ? q : 0.f;                                                     // multi_channel_relu, cifar_unet.c:235
    if (p.drop_rate > 0.f && uniform_at(p.drop_seed, idx, 0.f, 1.f) < p.drop_rate) q = 0.f;   // _dropout, cifar_unet.c:1032
    return q;
}
__device__ __forceinline__ float4 gn_post4(float4 q, size_t idx, const GnParams& p) {
    if (p.relu | (p.drop_rate > 0.f)) { q.x = gn_post(q.x, idx, p); q.y = gn_post(q.y, idx + 1, p); q.z = gn_post(q.z, idx + 2, p); q.w = gn_post(q.w, idx + 3, p); }
    return q;
}
// backward: the upstream gradient passes where the fused forward let the activation through.  relu(norm(x)) > 0 <=> x > mean
// (the divisor is positive), so the ReLU mask needs no saved activation; the dropout mask is regenerated from its counter.
__device__ __forceinline__ float gn_gate(float d, float xv, float mu, size_t idx, const GnParams& p) {
    if (p.relu && !(xv > mu)) d = 0.f;                                                      // multi_channel_relu_ddx, cifar_unet.c:241
    if (p.drop_rate > 0.f && uniform_at(p.drop_seed, idx, 0.f, 1.f) < p.drop_rate) d = 0.f;   // _dropout_mask, cifar_unet.c:1170
    return d;
}
// backward epilogue: + the gradient that reaches the same tensor through the residual connection (cifar_unet.c:1218-1220)
__device__ __forceinline__ float4 gn_add4(float4 o, size_t idx, const GnParams& p) {
    if (p.addend) { const float4 a = __ldg(reinterpret_cast<const float4*>(p.addend + idx)); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
    return o;
}
__device__ __forceinline__ float4 gn_gate4(float4 d, float4 xv, float mu, size_t idx, const GnParams& p) {
    if (p.relu | (p.drop_rate > 0.f)) {
        d.x = gn_gate(d.x, xv.x, mu, idx, p); d.y = gn_gate(d.y, xv.y, mu, idx + 1, p);
        d.z = gn_gate(d.z, xv.z, mu, idx + 2, p); d.w = gn_gate(d.w, xv.w, mu, idx + 3, p);
    }
    return d;
}

// grid = (G, images).  x,y: [images][C][HW]; means/vars: [images][G]
template <bool IN_SMEM>
__global__ void __launch_bounds__(kThreads) group_norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* vars, float* means,
                                                                  GnParams p) {
    extern __shared__ __align__(16) float slab[];
    __shared__ float red[kThreads / 32];
    const int g = blockIdx.x, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const size_t n = (size_t)nc * p.HW;
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float* xs = x + off;
    float* ys = y + off;
    const bool vec = ((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(ys)) & 15) == 0 && (n & 3) == 0;

    float s = 0.f;
    if (vec) {
        for (size_t i = threadIdx.x; i < n / 4; i += kThreads) {
            float4 v = reinterpret_cast<const float4*>(xs)[i];
            if (IN_SMEM) reinterpret_cast<float4*>(slab)[i] = v;
            s += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (size_t i = threadIdx.x; i < n; i += kThreads) {
            float v = xs[i];
            if (IN_SMEM) slab[i] = v;
            s += v;
        }
    }
    const float mean = block_total(s, red) / (float)(int)n;    // lib/norm.c:23
    const float* src = IN_SMEM ? slab : xs;
    float q = 0.f;
    for (size_t i = threadIdx.x; i < n; i += kThreads) {
        float d = src[i] - mean;
        q += d * d;
    }
    const float var = block_total(q, red) / (float)(int)n;     // lib/norm.c:36
    // D5: the reference divides by the variance (+ an integer epsilon that is 0); quirks off = textbook
    const float denom = p.quirk ? var : sqrtf(var + 1e-8f);
    if (threadIdx.x == 0) {
        means[(size_t)img * p.G + g] = mean;
        vars[(size_t)img * p.G + g] = p.quirk ? var : denom;
    }
    if (vec) {
        for (size_t i = threadIdx.x; i < n / 4; i += kThreads) {
            float4 v = IN_SMEM ? reinterpret_cast<const float4*>(slab)[i] : reinterpret_cast<const float4*>(xs)[i];
            v.x = (v.x - mean) / denom; v.y = (v.y - mean) / denom; v.z = (v.z - mean) / denom; v.w = (v.w - mean) / denom;
            reinterpret_cast<float4*>(ys)[i] = gn_post4(v, off + 4 * i, p);
        }
    } else {
        for (size_t i = threadIdx.x; i < n; i += kThreads) ys[i] = gn_post((src[i] - mean) / denom, off + i, p);
    }
}

// dx = (dy - mean(dy) - w * mean(w*dy)) / s,  w = (x - mu)/s,  s = stored "stdev"   lib/norm.c:62-90
template <bool IN_SMEM>
__global__ void __launch_bounds__(kThreads) group_norm_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, const float* __restrict__ x,
                                                                  const float* __restrict__ means, const float* __restrict__ stdevs, GnParams p) {
    extern __shared__ __align__(16) float slab[];   // [0,n): w ; [n,2n): dy   (IN_SMEM only)
    __shared__ float red[kThreads / 32];
    const int g = blockIdx.x, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const size_t n = (size_t)nc * p.HW;
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float mu = means[(size_t)img * p.G + g], sd = stdevs[(size_t)img * p.G + g];
    float gs = 0.f, gw = 0.f;
    for (size_t i = threadIdx.x; i < n; i += kThreads) {
        float w = (x[off + i] - mu) / sd;
        float d = gn_gate(dy[off + i], x[off + i], mu, off + i, p);
        if (IN_SMEM) { slab[i] = w; slab[n + i] = d; }
        gs += d;
        gw += w * d;
    }
    const float mean_g = block_total(gs, red) / (float)(int)n;
    const float mean_gw = block_total(gw, red) / (float)(int)n;
    for (size_t i = threadIdx.x; i < n; i += kThreads) {
        float w = IN_SMEM ? slab[i] : (x[off + i] - mu) / sd;
        float d = IN_SMEM ? slab[n + i] : gn_gate(dy[off + i], x[off + i], mu, off + i, p);
        dx[off + i] = (d - mean_g - w * mean_gw) / sd + (p.addend ? p.addend[off + i] : 0.f);
    }
}


// ---- fast paths: 1024-thread CTAs, two per SM, 128-bit accesses -------------------------------------
constexpr int kFastThreads = 1024;
constexpr int kFastF4 = 8;   // float4 per thread held in registers: slabs up to 1024*8*4 = 32768 elements (32 ch x 32 x 32)

__device__ __forceinline__ float block_total_1024(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = threadIdx.x < 32 ? sh[threadIdx.x] : 0.f;
    if (threadIdx.x < 32) {
        t = warp_sum(t);
        if (threadIdx.x == 0) sh[0] = t;
    }
    __syncthreads();
    t = sh[0];
    return t;
}

// forward: the group's slab is read from HBM exactly once into REGISTERS (8 x 128-bit loads per thread issued
// back to back = 128 KB in flight per SM, no shared-memory round trip), reduced twice, and written out.
__global__ void __launch_bounds__(kFastThreads, 1) group_norm_fwd_regs(const float* __restrict__ x, float* __restrict__ y, float* vars,
                                                                       float* means, GnParams p) {
    __shared__ float red[32];
    const int g = blockIdx.x, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const int n4 = (nc * p.HW) >> 2;
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float4* xs = reinterpret_cast<const float4*>(x + off);
    float4* ys = reinterpret_cast<float4*>(y + off);
    float4 v[kFastF4];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < kFastF4; ++u) {
        const int i = threadIdx.x + u * kFastThreads;
        v[u] = i < n4 ? xs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    const float inv_n = 1.f / (float)(4 * n4);
    const float mean = block_total_1024(s, red) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < kFastF4; ++u) {
        const int i = threadIdx.x + u * kFastThreads;
        if (i < n4) {
            const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float var = block_total_1024(q, red) * inv_n;
    const float denom = p.quirk ? var : sqrtf(var + 1e-8f);
    if (threadIdx.x == 0) {
        means[(size_t)img * p.G + g] = mean;
        vars[(size_t)img * p.G + g] = p.quirk ? var : denom;
    }
#pragma unroll
    for (int u = 0; u < kFastF4; ++u) {
        const int i = threadIdx.x + u * kFastThreads;
        if (i < n4) {
            float4 o;
            o.x = (v[u].x - mean) / denom; o.y = (v[u].y - mean) / denom; o.z = (v[u].z - mean) / denom; o.w = (v[u].w - mean) / denom;
            ys[i] = gn_post4(o, off + 4 * (size_t)i, p);
        }
    }
}

// backward: pass 1 reads x and dy (HBM) for the two group sums, pass 2 re-reads them (the CTA's 2 x 128 KB were
// just touched, they come from L2) and writes dx -- 12 B/elem of DRAM traffic, 128-bit accesses throughout.
__global__ void __launch_bounds__(kFastThreads, 1) group_norm_bwd_vec(const float* __restrict__ dy, float* __restrict__ dx,
                                                                      const float* __restrict__ x, const float* __restrict__ means,
                                                                      const float* __restrict__ stdevs, GnParams p) {
    __shared__ float red[32];
    const int g = blockIdx.x, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const int n4 = (nc * p.HW) >> 2;
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float4* xs = reinterpret_cast<const float4*>(x + off);
    const float4* gs4 = reinterpret_cast<const float4*>(dy + off);
    float4* ds = reinterpret_cast<float4*>(dx + off);
    const float mu = means[(size_t)img * p.G + g], sd = stdevs[(size_t)img * p.G + g];
    float gs = 0.f, gw = 0.f;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += kFastThreads) {
        const float4 a = xs[i];
        const float4 d = gn_gate4(gs4[i], a, mu, off + 4 * (size_t)i, p);
        gs += (d.x + d.y) + (d.z + d.w);
        gw += ((a.x - mu) / sd * d.x + (a.y - mu) / sd * d.y) + ((a.z - mu) / sd * d.z + (a.w - mu) / sd * d.w);
    }
    const float inv_n = 1.f / (float)(4 * n4);
    const float mean_g = block_total_1024(gs, red) * inv_n;
    const float mean_gw = block_total_1024(gw, red) * inv_n;
#pragma unroll 4
    for (int i = threadIdx.x; i < n4; i += kFastThreads) {
        const float4 a = xs[i];
        const float4 d = gn_gate4(gs4[i], a, mu, off + 4 * (size_t)i, p);
        float4 o;
        o.x = (d.x - mean_g - (a.x - mu) / sd * mean_gw) / sd; o.y = (d.y - mean_g - (a.y - mu) / sd * mean_gw) / sd;
        o.z = (d.z - mean_g - (a.z - mu) / sd * mean_gw) / sd; o.w = (d.w - mean_g - (a.w - mu) / sd * mean_gw) / sd;
        ds[i] = gn_add4(o, off + 4 * (size_t)i, p);
    }
}

inline bool al16(const void* q) { return ((uintptr_t)q & 15) == 0; }

// ---- cluster paths: one thread-block CLUSTER per (image, group) --------------------------------------------
// The slab is split over the CTAs of a cluster; every CTA keeps its share in registers (read from HBM exactly
// once) and the partial sums are exchanged through distributed shared memory.  Small CTAs (512 threads, <= 48
// registers) mean two or more clusters' CTAs are resident per SM, so one group's loads overlap another's
// reductions and stores -- which a single 1024-thread CTA per SM cannot do.
namespace cg = cooperative_groups;
constexpr int kClThreads = 512;

template <int CL>
__device__ __forceinline__ float cluster_total(float v, float* slot /* one float of this CTA's smem, 2 slots alternate */,
                                               float* red) {
    // CTA total
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < kClThreads / 32 ? red[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) *slot = t;
    }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();                                   // partials visible cluster-wide
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) tot += *cluster.map_shared_rank(slot, r);   // same order in every CTA -> identical totals
    return tot;
}

template <int CL>
__global__ void __launch_bounds__(kClThreads, 2) group_norm_fwd_cluster(const float* __restrict__ x, float* __restrict__ y, float* vars,
                                                                        float* means, GnParams p) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    constexpr int F4 = 8;
    __shared__ float red[kClThreads / 32];
    __shared__ float slots[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int g = blockIdx.x / CL, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const int n4 = (nc * p.HW) >> 2;
    const int per = (n4 + CL - 1) / CL;                // float4 per CTA
    const int beg = rank * per, end = min(n4, beg + per);
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float4* xs = reinterpret_cast<const float4*>(x + off);
    float4* ys = reinterpret_cast<float4*>(y + off);
    float4 v[F4];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = beg + threadIdx.x + u * kClThreads;
        v[u] = i < end ? xs[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    const float inv_n = 1.f / (float)(4 * n4);
    const float mean = cluster_total<CL>(s, &slots[0], red) * inv_n;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = beg + threadIdx.x + u * kClThreads;
        if (i < end) {
            const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    const float var = cluster_total<CL>(q, &slots[1], red) * inv_n;
    const float denom = p.quirk ? var : sqrtf(var + 1e-8f);
    if (rank == 0 && threadIdx.x == 0) {
        means[(size_t)img * p.G + g] = mean;
        vars[(size_t)img * p.G + g] = p.quirk ? var : denom;
    }
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = beg + threadIdx.x + u * kClThreads;
        if (i < end) {
            float4 o;
            o.x = (v[u].x - mean) / denom; o.y = (v[u].y - mean) / denom; o.z = (v[u].z - mean) / denom; o.w = (v[u].w - mean) / denom;
            ys[i] = gn_post4(o, off + 4 * (size_t)i, p);
        }
    }
    cluster.sync();   // no CTA may exit while a peer can still read its shared memory
}

template <int CL>
__global__ void __launch_bounds__(kClThreads, 2) group_norm_bwd_cluster(const float* __restrict__ dy, float* __restrict__ dx,
                                                                        const float* __restrict__ x, const float* __restrict__ means,
                                                                        const float* __restrict__ stdevs, GnParams p) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    constexpr int F4 = 4;
    __shared__ float red[kClThreads / 32];
    __shared__ float slots[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int g = blockIdx.x / CL, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const int n4 = (nc * p.HW) >> 2;
    const int per = (n4 + CL - 1) / CL;
    const int beg = rank * per, end = min(n4, beg + per);
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float4* xs = reinterpret_cast<const float4*>(x + off);
    const float4* gs4 = reinterpret_cast<const float4*>(dy + off);
    float4* ds = reinterpret_cast<float4*>(dx + off);
    const float mu = means[(size_t)img * p.G + g], sd = stdevs[(size_t)img * p.G + g];
    float4 w[F4], d[F4];
    float gs = 0.f, gw = 0.f;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = beg + threadIdx.x + u * kClThreads;
        w[u] = i < end ? xs[i] : make_float4(mu, mu, mu, mu);
        d[u] = i < end ? gn_gate4(gs4[i], w[u], mu, off + 4 * (size_t)i, p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        w[u].x = (w[u].x - mu) / sd; w[u].y = (w[u].y - mu) / sd; w[u].z = (w[u].z - mu) / sd; w[u].w = (w[u].w - mu) / sd;
        gs += (d[u].x + d[u].y) + (d[u].z + d[u].w);
        gw += (w[u].x * d[u].x + w[u].y * d[u].y) + (w[u].z * d[u].z + w[u].w * d[u].w);
    }
    const float inv_n = 1.f / (float)(4 * n4);
    const float mean_g = cluster_total<CL>(gs, &slots[0], red) * inv_n;
    const float mean_gw = cluster_total<CL>(gw, &slots[1], red) * inv_n;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = beg + threadIdx.x + u * kClThreads;
        if (i < end) {
            float4 o;
            o.x = (d[u].x - mean_g - w[u].x * mean_gw) / sd; o.y = (d[u].y - mean_g - w[u].y * mean_gw) / sd;
            o.z = (d[u].z - mean_g - w[u].z * mean_gw) / sd; o.w = (d[u].w - mean_g - w[u].w * mean_gw) / sd;
            ds[i] = gn_add4(o, off + 4 * (size_t)i, p);
        }
    }
    cluster.sync();
}

// ---- persistent forward with a TMA bulk-copy prefetch ring ------------------------------------------------
// One CTA per SM walks over (image, group) slabs.  The slab buffer is cut into four quarters, each with its own mbarrier and
// its own eight warps: as soon as a quarter's warps have moved it from shared memory into registers they re-arm it with the
// SAME quarter of the CTA's next slab (one cp.async.bulk), so the copy engine always has up to four requests in flight and HBM
// never waits for the block reductions, the normalisation or the stores.  (Round 1 refilled the whole buffer only after all
// 1024 threads had emptied it: 71 % of the measured HBM bandwidth.)
__device__ __forceinline__ uint32_t gn_smem_u32(const void* q) { return (uint32_t)__cvta_generic_to_shared(q); }
constexpr int kQuarterF4 = kFastThreads / 4 * kFastF4;      // 2048 float4 = 32 KB per quarter

__global__ void __launch_bounds__(kFastThreads, 1) group_norm_fwd_tma(const float* __restrict__ x, float* __restrict__ y, float* vars,
                                                                      float* means, GnParams p, int images) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    extern __shared__ __align__(128) float slab[];     // up to 32768 floats
    __shared__ float red[32];
    __shared__ __align__(8) unsigned long long bar[4];
    const int slabs = p.G * images;
    const int quarter = threadIdx.x >> 8, qt = threadIdx.x & 255;      // this thread's quarter of the slab and its place in it
    const uint32_t bar_a = gn_smem_u32(&bar[quarter]), slab_a = gn_smem_u32(slab) + (uint32_t)quarter * kQuarterF4 * 16u;
    auto slab_info = [&](int sidx, size_t& off, int& n4) {
        const int g = sidx % p.G, img = sidx / p.G;
        const int c0 = g * p.group_size;
        const int nc = min(p.group_size, p.C - c0);
        n4 = (nc * p.HW) >> 2;
        off = ((size_t)img * p.C + c0) * p.HW;
    };
    // float4 [quarter * 2048, min(n4, (quarter + 1) * 2048)) of slab sidx -> this quarter of the buffer; true if there is any
    auto quarter_f4 = [&](int n4) { return max(0, min(n4 - quarter * kQuarterF4, kQuarterF4)); };
    auto issue = [&](int sidx) {       // one thread per quarter
        size_t off; int n4;
        slab_info(sidx, off, n4);
        const uint32_t bytes = (uint32_t)quarter_f4(n4) * 16u;
        if (!bytes) return;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(slab_a), "l"(reinterpret_cast<const char*>(x + off) + (size_t)quarter * kQuarterF4 * 16u), "r"(bytes), "r"(bar_a) : "memory");
    };
    if (qt == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int sidx = blockIdx.x;
    if (qt == 0 && sidx < slabs) issue(sidx);
    uint32_t phase = 0;
    const float4* mine = reinterpret_cast<const float4*>(slab) + quarter * kQuarterF4;
    for (; sidx < slabs; sidx += gridDim.x) {
        size_t off; int n4;
        slab_info(sidx, off, n4);
        const int q4 = quarter_f4(n4);
        float4 v[kFastF4];
        float s = 0.f;
        if (q4 > 0) {
            // wait for this quarter's bytes
            asm volatile(
                "{\n\t.reg .pred p;\n\tGN_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra GN_DONE;\n\tbra GN_WAIT;\n\tGN_DONE:\n\t}"
                ::"r"(bar_a), "r"(phase) : "memory");
            phase ^= 1;
        }
#pragma unroll
        for (int u = 0; u < kFastF4; ++u) {
            const int i = qt + u * 256;
            v[u] = i < q4 ? mine[i] : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
        // this quarter's eight warps have left it: refill it with the same quarter of the next slab
        asm volatile("bar.sync %0, 256;" ::"r"(quarter + 1) : "memory");
        const int next = sidx + gridDim.x;
        if (qt == 0 && next < slabs) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(next);
        }
        const float inv_n = 1.f / (float)(4 * n4);
        const float mean = block_total_1024(s, red) * inv_n;
        float q = 0.f;
#pragma unroll
        for (int u = 0; u < kFastF4; ++u) {
            const int i = qt + u * 256;
            if (i < q4) {
                const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float var = block_total_1024(q, red) * inv_n;
        const float denom = p.quirk ? var : sqrtf(var + 1e-8f);
        if (threadIdx.x == 0) {
            means[sidx] = mean;                            // [img][g] == sidx
            vars[sidx] = p.quirk ? var : denom;
        }
        float4* ys = reinterpret_cast<float4*>(y + off) + quarter * kQuarterF4;
        const size_t e0 = off + 4 * (size_t)quarter * kQuarterF4;
#pragma unroll
        for (int u = 0; u < kFastF4; ++u) {
            const int i = qt + u * 256;
            if (i < q4) {
                float4 o;
                o.x = (v[u].x - mean) / denom; o.y = (v[u].y - mean) / denom; o.z = (v[u].z - mean) / denom; o.w = (v[u].w - mean) / denom;
                ys[i] = gn_post4(o, e0 + 4 * (size_t)i, p);
            }
        }
    }
}

// ---- small slabs (<= 8192 elements: 32 channels of a 16 x 16, 8 x 8 or 4 x 4 image) ---------------------------------------
// One 256-thread CTA per (image, group), the whole slab of x and dy in registers (read once, 12 B/elem), one shuffle + shared
// memory reduction, no cluster: many small CTAs per SM hide each other's latency.  (The persistent cluster kernel below
// spends ~7 us per slab on its barrier chain whatever the slab size: 50 us for 512 slabs of 2 KB.)
constexpr int kSmallThreads = 256;
template <int F4>
__global__ void __launch_bounds__(kSmallThreads) group_norm_bwd_small(const float* __restrict__ dy, float* __restrict__ dx,
                                                                      const float* __restrict__ x, const float* __restrict__ means,
                                                                      const float* __restrict__ stdevs, GnParams p) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    __shared__ float red[2][kSmallThreads / 32];
    const int g = blockIdx.x, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const int n4 = (nc * p.HW) >> 2;
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float4* xs = reinterpret_cast<const float4*>(x + off);
    const float4* gs4 = reinterpret_cast<const float4*>(dy + off);
    float4* ds = reinterpret_cast<float4*>(dx + off);
    const float mu = means[(size_t)img * p.G + g], sd = stdevs[(size_t)img * p.G + g];
    float4 w[F4], d[F4];
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = threadIdx.x + u * kSmallThreads;
        const bool ok = i < n4;
        w[u] = ok ? __ldg(xs + i) : make_float4(mu, mu, mu, mu);
        d[u] = ok ? gn_gate4(__ldg(gs4 + i), w[u], mu, off + 4 * (size_t)i, p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float gs = 0.f, gw = 0.f;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        w[u].x = (w[u].x - mu) / sd; w[u].y = (w[u].y - mu) / sd; w[u].z = (w[u].z - mu) / sd; w[u].w = (w[u].w - mu) / sd;
        gs += (d[u].x + d[u].y) + (d[u].z + d[u].w);
        gw += (w[u].x * d[u].x + w[u].y * d[u].y) + (w[u].z * d[u].z + w[u].w * d[u].w);
    }
    gs = warp_sum(gs); gw = warp_sum(gw);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = gs; red[1][threadIdx.x >> 5] = gw; }
    __syncthreads();
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int k = 0; k < kSmallThreads / 32; ++k) { t0 += red[0][k]; t1 += red[1][k]; }   // same order in every thread
    const float inv_n = 1.f / (float)(nc * p.HW);
    const float mean_g = t0 * inv_n, mean_gw = t1 * inv_n;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = threadIdx.x + u * kSmallThreads;
        if (i < n4) {
            float4 o;
            o.x = (d[u].x - mean_g - w[u].x * mean_gw) / sd; o.y = (d[u].y - mean_g - w[u].y * mean_gw) / sd;
            o.z = (d[u].z - mean_g - w[u].z * mean_gw) / sd; o.w = (d[u].w - mean_g - w[u].w * mean_gw) / sd;
            ds[i] = gn_add4(o, off + 4 * (size_t)i, p);
        }
    }
}

// forward twin of group_norm_bwd_small: the slab of x in registers (read once, 8 B/elem), mean and centred second moment by two
// shuffle + shared-memory reductions, ReLU / dropout fused on the way out
template <int F4>
__global__ void __launch_bounds__(kSmallThreads) group_norm_fwd_small(const float* __restrict__ x, float* __restrict__ y, float* vars, float* means,
                                                                      GnParams p) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    __shared__ float red[2][kSmallThreads / 32];
    const int g = blockIdx.x, img = blockIdx.y;
    const int c0 = g * p.group_size;
    const int nc = min(p.group_size, p.C - c0);
    const int n4 = (nc * p.HW) >> 2;
    const size_t off = ((size_t)img * p.C + c0) * p.HW;
    const float4* xs = reinterpret_cast<const float4*>(x + off);
    float4* ys = reinterpret_cast<float4*>(y + off);
    float4 v[F4];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = threadIdx.x + u * kSmallThreads;
        v[u] = i < n4 ? __ldg(xs + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    const float inv_n = 1.f / (float)(4 * n4);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kSmallThreads / 32; ++k) t += red[0][k];
    const float mean = t * inv_n;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = threadIdx.x + u * kSmallThreads;
        if (i < n4) {
            const float a = v[u].x - mean, b = v[u].y - mean, c = v[u].z - mean, d = v[u].w - mean;
            q += (a * a + b * b) + (c * c + d * d);
        }
    }
    q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = q;
    __syncthreads();
    t = 0.f;
#pragma unroll
    for (int k = 0; k < kSmallThreads / 32; ++k) t += red[1][k];
    const float var = t * inv_n;
    const float denom = p.quirk ? var : sqrtf(var + 1e-8f);
    if (threadIdx.x == 0) {
        means[(size_t)img * p.G + g] = mean;
        vars[(size_t)img * p.G + g] = p.quirk ? var : denom;
    }
#pragma unroll
    for (int u = 0; u < F4; ++u) {
        const int i = threadIdx.x + u * kSmallThreads;
        if (i < n4) {
            float4 o;
            o.x = (v[u].x - mean) / denom; o.y = (v[u].y - mean) / denom; o.z = (v[u].z - mean) / denom; o.w = (v[u].w - mean) / denom;
            ys[i] = gn_post4(o, off + 4 * (size_t)i, p);
        }
    }
}

// ---- persistent backward: 2-CTA clusters + TMA bulk prefetch + DSMEM sums ---------------------------------------
// A cluster owns one (image, group) slab at a time; each CTA handles half of it.  The NEXT slab's halves of x and dy
// are fetched into shared memory by cp.async.bulk while the current ones -- already in registers -- are reduced
// (two sums, exchanged between the CTAs through distributed shared memory) and dx is written: x and dy are read from
// HBM exactly once (12 B/elem) and the reductions never leave HBM idle.
__global__ void __launch_bounds__(kFastThreads, 1) group_norm_bwd_tma(const float* __restrict__ dy, float* __restrict__ dx,
                                                                      const float* __restrict__ x, const float* __restrict__ means,
                                                                      const float* __restrict__ stdevs, GnParams p, int images) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    constexpr int F4 = 4;                                  // float4 per thread per tensor: 2 CTAs x 1024 x 4 x 4 = 32768 elements
    extern __shared__ __align__(128) float stage[];        // [x half | dy half], each up to 16384 floats
    __shared__ float red2[2][32];
    __shared__ __align__(16) float inbox[2][2][2];         // [parity][source CTA][sum]: delivered by st.async, see the loop
    __shared__ __align__(8) unsigned long long bar[4];
    __shared__ __align__(8) unsigned long long xbar[2];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int slabs = p.G * images;
    constexpr int kHalfFloats = kFastThreads * F4 * 4;     // 16384
    constexpr int kQF4 = kFastThreads / 4 * F4;            // 1024 float4 of x and of dy per quarter
    // the CTA's share is cut into four quarters (x and dy pieces side by side in time: one mbarrier per quarter); the eight warps
    // that own a quarter refill it with the same quarter of the cluster's next slab as soon as they hold it in registers
    const int quarter = threadIdx.x >> 8, qt = threadIdx.x & 255;
    const uint32_t bar_a = gn_smem_u32(&bar[quarter]), st_a = gn_smem_u32(stage) + (uint32_t)quarter * kQF4 * 16u;
    int par = 0;
    auto slab_info = [&](int sidx, size_t& off, int& beg, int& end) {
        const int g = sidx % p.G, img = sidx / p.G;
        const int c0 = g * p.group_size;
        const int nc = min(p.group_size, p.C - c0);
        const int n4 = (nc * p.HW) >> 2;
        const int per = (n4 + 1) / 2;
        beg = rank * per;
        end = min(n4, beg + per);
        off = ((size_t)img * p.C + c0) * p.HW;
    };
    auto quarter_f4 = [&](int beg, int end) { return max(0, min(end - beg - quarter * kQF4, kQF4)); };
    auto issue = [&](int sidx) {       // one thread per quarter
        size_t off; int beg, end;
        slab_info(sidx, off, beg, end);
        const uint32_t bytes = (uint32_t)quarter_f4(beg, end) * 16u;
        if (!bytes) return;
        const size_t src = ((size_t)beg + (size_t)quarter * kQF4) * 16;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(2u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(st_a), "l"(reinterpret_cast<const char*>(x + off) + src), "r"(bytes), "r"(bar_a) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(st_a + (uint32_t)kHalfFloats * 4u), "l"(reinterpret_cast<const char*>(dy + off) + src), "r"(bytes), "r"(bar_a) : "memory");
    };
    if (qt == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        if (quarter < 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gn_smem_u32(&xbar[quarter])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    int it = 0;                                            // slabs done: inbox / xbar [it & 1], in phase (it >> 1) & 1
    cluster.sync();                                        // the peer's xbar is initialised before anything is delivered to it
    const int cid = blockIdx.x / 2, ncl = gridDim.x / 2;
    int sidx = cid;
    if (qt == 0 && sidx < slabs) issue(sidx);
    uint32_t phase = 0;
    const float4* xq = reinterpret_cast<const float4*>(stage) + quarter * kQF4;
    const float4* dq = reinterpret_cast<const float4*>(stage + kHalfFloats) + quarter * kQF4;
    for (; sidx < slabs; sidx += ncl) {
        size_t off; int beg, end;
        slab_info(sidx, off, beg, end);
        const int q4 = quarter_f4(beg, end);
        if (q4 > 0) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tGNB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra GNB_DONE;\n\tbra GNB_WAIT;\n\tGNB_DONE:\n\t}"
                ::"r"(bar_a), "r"(phase) : "memory");
            phase ^= 1;
        }
        const float mu = means[sidx], sd = stdevs[sidx];
        const int qbeg = beg + quarter * kQF4;                 // slab float4 index of this quarter's first element
        float4 w[F4], d[F4];
#pragma unroll
        for (int u = 0; u < F4; ++u) {
            const int i = qt + u * 256;
            const bool ok = i < q4;
            w[u] = ok ? xq[i] : make_float4(mu, mu, mu, mu);
            d[u] = ok ? gn_gate4(dq[i], w[u], mu, off + 4 * (size_t)(qbeg + i), p) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("bar.sync %0, 256;" ::"r"(quarter + 1) : "memory");   // this quarter is free again: prefetch the next slab's
        const int next = sidx + ncl;
        if (qt == 0 && next < slabs) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(next);
        }
        float gs = 0.f, gw = 0.f;
#pragma unroll
        for (int u = 0; u < F4; ++u) {
            w[u].x = (w[u].x - mu) / sd; w[u].y = (w[u].y - mu) / sd; w[u].z = (w[u].z - mu) / sd; w[u].w = (w[u].w - mu) / sd;
            gs += (d[u].x + d[u].y) + (d[u].z + d[u].w);
            gw += (w[u].x * d[u].x + w[u].y * d[u].y) + (w[u].z * d[u].z + w[u].w * d[u].w);
        }
        // CTA totals of both sums in ONE shuffle + shared-memory pass.  Lane 0 of warp k (k = 0: sum dy, 1: sum w.dy) then DELIVERS its
        // total into inbox[par][rank][k] of BOTH CTAs of the cluster with st.async, which completes 4 bytes on the receiving CTA's
        // mbarrier xbar[par]: every thread waits for the 16 bytes of its own CTA's inbox and adds the two CTAs' totals in rank order
        // (same bits in both).  No cluster barrier and no release fence in the loop -- round 2's barrier.cluster.arrive.release made
        // all 1024 threads wait for their dx stores of the slab before to drain (ncu: 16 % of the kernel's stall samples on that one
        // ERRBAR / UCGABAR_ARV pair).  The inboxes alternate: a CTA can be at most one slab ahead of its peer (it needs the peer's
        // totals of slab k to leave slab k), so inbox[par] is rewritten only after both CTAs have read it two slabs earlier.
        float tot[2];
        gs = warp_sum(gs); gw = warp_sum(gw);
        if ((threadIdx.x & 31) == 0) { red2[0][threadIdx.x >> 5] = gs; red2[1][threadIdx.x >> 5] = gw; }
        __syncthreads();
        if (threadIdx.x < 64) {
            float v = red2[threadIdx.x >> 5][threadIdx.x & 31];
            v = warp_sum(v);
            if ((threadIdx.x & 31) == 0) {
                const int k = threadIdx.x >> 5;
                if (k == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gn_smem_u32(&xbar[par])), "r"(16u) : "memory");
                const uint32_t slot = gn_smem_u32(&inbox[par][rank][k]), bar_l = gn_smem_u32(&xbar[par]);
#pragma unroll
                for (uint32_t dst = 0; dst < 2; ++dst) {
                    uint32_t slot_c, bar_c;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(slot_c) : "r"(slot), "r"(dst));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar_c) : "r"(bar_l), "r"(dst));
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                                 ::"r"(slot_c), "r"(__float_as_uint(v)), "r"(bar_c) : "memory");
                }
            }
        }
        asm volatile(
            "{\n\t.reg .pred p;\n\tGNX_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra GNX_DONE;\n\tbra GNX_WAIT;\n\tGNX_DONE:\n\t}"
            ::"r"(gn_smem_u32(&xbar[par])), "r"((uint32_t)(it >> 1) & 1u) : "memory");
        ++it;
        tot[0] = *(volatile float*)&inbox[par][0][0] + *(volatile float*)&inbox[par][1][0];
        tot[1] = *(volatile float*)&inbox[par][0][1] + *(volatile float*)&inbox[par][1][1];
        par ^= 1;
        const int g = sidx % p.G;
        const int nc = min(p.group_size, p.C - g * p.group_size);
        const float inv_n = 1.f / (float)(nc * p.HW);
        const float mean_g = tot[0] * inv_n, mean_gw = tot[1] * inv_n;
        float4* ds = reinterpret_cast<float4*>(dx + off) + qbeg;
#pragma unroll
        for (int u = 0; u < F4; ++u) {
            const int i = qt + u * 256;
            if (i < q4) {
                float4 o;
                o.x = (d[u].x - mean_g - w[u].x * mean_gw) / sd; o.y = (d[u].y - mean_g - w[u].y * mean_gw) / sd;
                o.z = (d[u].z - mean_g - w[u].z * mean_gw) / sd; o.w = (d[u].w - mean_g - w[u].w * mean_gw) / sd;
                ds[i] = gn_add4(o, off + 4 * (size_t)(qbeg + i), p);
            }
        }
    }
    cluster.sync();
}

template <int CL, class... KArgs, class... Args>
void launch_cluster(void (*kernel)(KArgs...), dim3 grid, cudaStream_t s, Args... args) {
    BLA_CUDA(launch_pdl(kernel, grid, dim3(kClThreads), 0, s, CL, args...));
}


}  // namespace

void k_group_norm_fwd(const float* x, float* y, float* vars, float* means, int images, int C, int HW, int group_size, int quirk,
                      cudaStream_t s, const GnFuse* fuse) {
    GnParams p{C, HW, group_size, (C + group_size - 1) / group_size, quirk, 0, 0.f, 0ull, nullptr};
    if (fuse) { p.relu = fuse->relu; p.drop_rate = fuse->drop_rate; p.drop_seed = fuse->drop_seed; }
    if (images <= 0 || C <= 0 || HW <= 0) return;
    const size_t slab = (size_t)min(group_size, C) * HW;
    dim3 grid(p.G, images);
    // every group slab starts 16-byte aligned and has a multiple of 4 elements?
    const bool vec_ok = al16(x) && al16(y) && (HW % 4 == 0) && (C % group_size == 0 || ((size_t)(C % group_size) * HW) % 4 == 0);
    if (vec_ok && slab <= (size_t)kSmallThreads * 8 * 4 && (long long)p.G * images >= 2LL * rt().num_sms) {
        // many small slabs: one small register-resident CTA each (the persistent kernel's per-slab barrier chain costs more than
        // these slabs' transfer time)
        if (slab <= (size_t)kSmallThreads * 2 * 4) BLA_CUDA(launch_pdl(group_norm_fwd_small<2>, dim3(grid), dim3(kSmallThreads), 0, s, 1, x, y, vars, means, p));
        else BLA_CUDA(launch_pdl(group_norm_fwd_small<8>, dim3(grid), dim3(kSmallThreads), 0, s, 1, x, y, vars, means, p));
    } else if (vec_ok && slab <= (size_t)kFastThreads * kFastF4 * 4 && (long long)p.G * images >= 2LL * rt().num_sms) {
        // many slabs: persistent CTAs with a TMA prefetch of the next slab
        static bool attr = false;
        if (!attr) {
            BLA_CUDA(cudaFuncSetAttribute(group_norm_fwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastThreads * kFastF4 * 16));
            attr = true;
        }
        BLA_CUDA(launch_pdl(group_norm_fwd_tma, dim3(rt().num_sms), dim3(kFastThreads), slab * sizeof(float), s, 1, x, y, vars, means, p, images));
    } else if (vec_ok && slab <= (size_t)2 * kClThreads * 8 * 4) {
        // cluster of 2 CTAs x 512 threads x 8 float4 covers 32768 elements (32 channels x 32 x 32)
        launch_cluster<2>(group_norm_fwd_cluster<2>, dim3(2 * p.G, images), s, x, y, vars, means, p);
    } else if (vec_ok && slab <= (size_t)kFastThreads * kFastF4 * 4) {
        group_norm_fwd_regs<<<grid, kFastThreads, 0, s>>>(x, y, vars, means, p);
    } else if (slab <= kSmemSlabFloats) {
        static bool attr = false;
        if (!attr) {
            BLA_CUDA(cudaFuncSetAttribute(group_norm_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemSlabFloats * 4)));
            attr = true;
        }
        group_norm_fwd_kernel<true><<<grid, kThreads, slab * sizeof(float), s>>>(x, y, vars, means, p);
    } else {
        group_norm_fwd_kernel<false><<<grid, kThreads, 0, s>>>(x, y, vars, means, p);
    }
    BLA_LAUNCH_CHECK();
    count_launch();
}

void k_group_norm_bwd(const float* dy, float* dx, const float* x, const float* means, const float* stdevs, int images, int C, int HW,
                      int group_size, cudaStream_t s, const GnFuse* fuse) {
    GnParams p{C, HW, group_size, (C + group_size - 1) / group_size, 1, 0, 0.f, 0ull, nullptr};
    if (fuse) { p.relu = fuse->relu; p.drop_rate = fuse->drop_rate; p.drop_seed = fuse->drop_seed; p.addend = fuse->addend; }
    if (images <= 0 || C <= 0 || HW <= 0) return;
    const size_t slab = (size_t)min(group_size, C) * HW;
    dim3 grid(p.G, images);
    const bool vec_ok = al16(dy) && al16(dx) && al16(x) && al16(p.addend) && (HW % 4 == 0);
    const bool tail_ok = (C % group_size == 0) || (((size_t)(C % group_size) * HW) % 4 == 0);
    if (vec_ok && tail_ok && slab <= (size_t)kSmallThreads * 8 * 4 && (long long)p.G * images >= 2LL * rt().num_sms) {
        if (slab <= (size_t)kSmallThreads * 2 * 4) BLA_CUDA(launch_pdl(group_norm_bwd_small<2>, dim3(grid), dim3(kSmallThreads), 0, s, 1, dy, dx, x, means, stdevs, p));
        else BLA_CUDA(launch_pdl(group_norm_bwd_small<8>, dim3(grid), dim3(kSmallThreads), 0, s, 1, dy, dx, x, means, stdevs, p));
    } else if (vec_ok && tail_ok && slab <= (size_t)2 * kFastThreads * 4 * 4 && slab % 8 == 0 && (long long)p.G * images >= (long long)rt().num_sms) {
        // many slabs: persistent 2-CTA clusters with TMA prefetch (each CTA's half slab must be a whole number of float4)
        static bool attr = false;
        const int smem = 2 * kFastThreads * 4 * 16;   // 128 KB
        if (!attr) {
            BLA_CUDA(cudaFuncSetAttribute(group_norm_bwd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            attr = true;
        }
        BLA_CUDA(launch_pdl(group_norm_bwd_tma, dim3((rt().num_sms / 2) * 2), dim3(kFastThreads), (size_t)smem, s, 2, dy, dx, x, means, stdevs, p, images));
    } else if (vec_ok && tail_ok && slab <= (size_t)4 * kClThreads * 4 * 4) {
        // cluster of 4 CTAs x 512 threads x (4 + 4) float4: x and dy are read from HBM exactly once
        launch_cluster<4>(group_norm_bwd_cluster<4>, dim3(4 * p.G, images), s, dy, dx, x, means, stdevs, p);
    } else if (vec_ok) {
        group_norm_bwd_vec<<<grid, kFastThreads, 0, s>>>(dy, dx, x, means, stdevs, p);
    } else if (2 * slab <= kSmemSlabFloats) {
        static bool attr = false;
        if (!attr) {
            BLA_CUDA(cudaFuncSetAttribute(group_norm_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSmemSlabFloats * 4)));
            attr = true;
        }
        group_norm_bwd_kernel<true><<<grid, kThreads, 2 * slab * sizeof(float), s>>>(dy, dx, x, means, stdevs, p);
    } else {
        group_norm_bwd_kernel<false><<<grid, kThreads, 0, s>>>(dy, dx, x, means, stdevs, p);
    }
    BLA_LAUNCH_CHECK();
    count_launch();
}

}  // namespace bla

extern "C" {

const int epsilon = 1e-8;   // lib/norm.c:3 -- an int, hence 0 (exported data symbol of the reference object)

// lib/norm.c:5-50
void group_norm(Matrix* in, Matrix* out, matrix_float_t* stdevs, matrix_float_t* means, int channels, int group_size) {
    if (channels <= 0) return;
    CallScope sc;
    const int HW = in[0].rows * in[0].cols;
    const int G = (channels + group_size - 1) / group_size;
    PlaneSet xin(sc, in, channels, true);
    PlaneSet yout(sc, out, channels, false);
    float* dsd = sc.out(stdevs, G);
    float* dmu = sc.out(means, G);
    k_group_norm_fwd(xin.dev(), yout.dev(), dsd, dmu, 1, channels, HW, group_size, rt().quirks, sc.stream());
    yout.write_back();
}

// lib/norm.c:52-93
void group_norm_ddx(Matrix* source, Matrix* dest, Matrix* data, matrix_float_t* means, matrix_float_t* stdevs, int channels,
                    int group_size) {
    if (channels <= 0) return;
    CallScope sc;
    const int HW = source[0].rows * source[0].cols;
    const int G = (channels + group_size - 1) / group_size;
    PlaneSet dy(sc, source, channels, true);
    PlaneSet x(sc, data, channels, true);
    PlaneSet dx(sc, dest, channels, false);
    const float* dmu = sc.in(means, G);
    const float* dsd = sc.in(stdevs, G);
    k_group_norm_bwd(dy.dev(), dx.dev(), x.dev(), dmu, dsd, 1, channels, HW, group_size, sc.stream());
    dx.write_back();
}

// additive, batched, device-resident: x,y [images][C][HW]; vars/means [images][G]  (include/bla.h)
void bla_group_norm(const float* x, float* y, float* vars, float* means, int images, int channels, int hw, int group_size) {
    k_group_norm_fwd(x, y, vars, means, images, channels, hw, group_size, rt().quirks, rt().stream);
}
void bla_group_norm_ddx(const float* dy, float* dx, const float* x, const float* means, const float* vars, int images, int channels,
                        int hw, int group_size) {
    k_group_norm_bwd(dy, dx, x, means, vars, images, channels, hw, group_size, rt().stream);
}

}  // extern "C"
