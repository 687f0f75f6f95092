// gemm_tc.cu -- the 3xTF32 tensor-core GEMM path: C = op(A).op(B) with FP32-class
// accuracy from TF32 tcgen05.mma instructions, accumulating in FP32 in tensor memory (TMEM).
//
// Arithmetic.  Each fp32 operand x is split as x = hi + lo with hi = x truncated to TF32 (the
// tensor core itself ignores the low 13 mantissa bits of a 32-bit operand, so the RAW fp32 tile is
// the hi tile) and lo = x - hi (exact subtraction, 13 significant bits, of which the tensor core keeps 11).
// Every product is issued three times, a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, into one TMEM
// accumulator; the dropped lo.lo term and the rounding of lo are O(2^-21) relative.
//
// Dataflow per CTA (persistent, one CTA per SM, 14 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of raw A (128 x 16) and B (16 x 256) fp32
//               into swizzled shared memory, 4-stage ring, mbarrier complete_tx.
//   warps 6-13  splitters: read the raw tile from shared memory, write the lo tile next to it
//               (same swizzled positions, so the split is a flat elementwise pass), then
//               fence.proxy.async and arrive on the stage's "split" barrier.
//   warp 1      MMA issuer (one lane): 6 tcgen05.mma.kind::tf32 128x256x8 per stage
//               (2 k-steps x 3 products), tcgen05.commit -> frees the stage / publishes the tile.
//   warps 2-5   epilogue: tcgen05.ld the 128 x 256 fp32 accumulator (TMEM is double-buffered, so
//               the next tile's MMAs overlap), transpose through swizzled shared memory so that
//               global stores / gate loads are full 128-byte rows, fused alpha / bias / ReLU /
//               pre-activation copy / relu' gate, or raw partials for split-K.
// Operand layouts are expressed only through the TMA box and the UMMA descriptor: K-major
// (A row-major, B stored [n][k]) and MN-major (A stored [k][m], B row-major) -- no transposes.
#include <cuda.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "runtime.h"

namespace bla {

namespace {

constexpr int BM = 128, BN = 256, BK = 16;   // BN = widest tile (TMEM stage); the tile width actually used is p.bn
constexpr int kAccStages = 2;
constexpr int kSplitWarps = 8;   // 4 warps needed ~700 cycles per 24 KB stage, as long as its MMAs: the split was co-critical
constexpr int kThreads = 32 * (6 + kSplitWarps);   // TMA, MMA, 4 epilogue, kSplitWarps splitters
constexpr uint32_t kABytes = BM * BK * 4;               // 8 KB
// Per-CTA stage geometry.  One CTA per tile: raw A 8 KB + raw B 16 KB, mirrored by the lo tiles = 48 KB, 4 stages.
// CTA pair: every CTA stages only half of B: 2 x (8 + 8) KB = 32 KB, 6 stages (the pair's barrier round trips cross
// the cluster, so the ring is deeper).
template <int CL> struct Geo {
    static constexpr uint32_t kBBytes = (BN / CL) * BK * 4;
    static constexpr uint32_t kRawBytes = kABytes + kBBytes;
    static constexpr uint32_t kStageBytes = 2 * kRawBytes;
    static constexpr int kStages = CL == 2 ? 6 : 4;
};
constexpr uint32_t kStagingBytes = 4 * 2 * 32 * 32 * 4; // 32 KB: two 32x32 fp32 tiles per epilogue warp
constexpr uint32_t kSmemBytes = 4 * Geo<1>::kStageBytes + kStagingBytes + 1024 + 1024;   // + barriers/flags + alignment slack
static_assert(6 * Geo<2>::kStageBytes == 4 * Geo<1>::kStageBytes, "both geometries fill the same 192 KB");

struct TcParams {
    int m, n, k;
    int m_tiles, n_tiles, splits, kblocks_per_split, kblocks;
    int bn;                  // tile width: multiple of 16 (B K-major) or 32 (B MN-major), <= 256
    uint32_t stage_tx_bytes; // bytes TMA delivers per stage
    bool a_kmajor, b_kmajor;
    bool a_box3, b_box3;     // MN-major operand fetched as ONE 3-D box {32 columns, BK rows, atoms} per k-block instead of one 2-D box per atom
    float* c; int ldc;
    float* partial;          // [splits][m][n] when splits > 1
    bla_epilogue epi;
    uint32_t* mask_out; int mask_ld;            // bit j % 32 of mask_out[i * mask_ld + j / 32] = (C[i][j] > 0)
    const uint32_t* gate_bits; int gate_ld;     // epi.gate in the same bit form
    bool c_vec;              // 16-byte aligned rows
    bool tma_store;          // epilogue writes C (or the split-K partials) with cp.async.bulk.tensor stores
    int pdl_early;           // trigger the stream's next kernel at the start (runtime.h) or leave it to this kernel's exit
    int cluster;             // 1, or 2: CTA PAIRS (tcgen05 cta_group::2): 256 x bn tiles, each CTA stages only HALF of B
    // implicit-GEMM convolution (conv != 0): B[k'][n] is gathered from x [img][C][H][W] by 4-D TMA boxes, k' = (ki, kj, c)
    // and n = (img, oi, oj); C is y [img][F][P] written by 3-D TMA boxes.  See conv2d_forward_tc().
    // conv == 2 is the weight gradient: A = dy [img][F][P] (3-D boxes, K-major, k = pixel), B[k = pixel][n = (ki, kj, c)] gathered from
    // the same padded NHWC input as MN-major atoms of 32 channels x 16 pixels; C = dW [F][(ki, kj, c)] through the plain GEMM epilogue.
    int conv, cv_P, cv_Wo, cv_stride, cv_pad_top, cv_pad_left, cv_k, cv_cblocks, cv_bw, cv_bh, cv_C;
    const float* cv_bias;     // conv == 1: + bias[img][filter]   (the U-Net's time-embedding add, cifar_unet.c:1024-1030)
    const float* cv_add;      // conv == 1: + addend[img][filter][pixel]   (the residual connection, cifar_unet.c:1067-1071)
    int cv_final, cv_Creal;   // conv == 2 with split-K: the reduce kernel writes dW in the reference layout [F][C][k][k] (C = cv_Creal)
    int ring;                // 0 = all stages of the geometry
    int debug;               // BLA_TC_DEBUG: 1 = no split (1xTF32: hi.hi only), 4 = never skip a zero lo tile, 5 = no split work but
                             // three products -- bottleneck experiments
};

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int x, int y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y), "r"(src)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"(src) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "h"(mask) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
// Arrive on a barrier of the leader CTA.  Default (CTA-scope release) semantics on purpose: what the leader's MMA
// lane consumes is shared memory that each SM's own threads wrote and fenced to the async proxy; a cluster-scope
// release/acquire pair compiles to MEMBAR.GPU + CCTL.IVALL per k-block and halved the kernel's speed.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// One 32-bit word into the shared memory of another CTA of the cluster, counted as 4 bytes of transaction on a barrier of THAT CTA:
// whoever sees the barrier's phase complete sees the word (the async-proxy store and its complete_tx are ordered) -- a cross-CTA
// message without any cluster-scope fence.
__device__ __forceinline__ void st_async_u32(uint32_t cluster_addr, uint32_t value, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(value),
                 "r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITC_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAITC_DONE;\n\t"
        "bra WAITC_LOOP;\n\t"
        "WAITC_DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)0x3) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// true in exactly one lane of the (converged) warp; ptxas knows the branch it guards runs on a single lane
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
        " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride
// byte offsets (all >> 4), version 1 (Blackwell), SWIZZLE_128B.
// layout_type: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B -- the only layout the tensor
// core accepts for MN-major 32-bit operands (32-byte swizzle atoms, 4-row groups).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version
    d |= (uint64_t)layout_type << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, FP32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn_major, bool b_mn_major, int bn, int m = BM) {
    return (1u << 4)                       // c_format  = F32
           | (2u << 7)                     // a_format  = TF32
           | (2u << 10)                    // b_format  = TF32
           | ((a_mn_major ? 1u : 0u) << 15)
           | ((b_mn_major ? 1u : 0u) << 16)
           | ((uint32_t)(bn >> 3) << 17)   // n_dim
           | ((uint32_t)(m >> 4) << 24);   // m_dim (256 for a CTA pair)
}

__device__ __forceinline__ float epilogue_value(float acc, int i, int j, const TcParams& p) {
    float v = acc;
    if (p.epi.alpha != 0.f) v *= p.epi.alpha;
    if (p.epi.bias_rows) v += p.epi.bias_rows[i];
    if (p.epi.bias_cols) v += p.epi.bias_cols[j];
    if (p.epi.pre_activation) p.epi.pre_activation[(size_t)i * p.ldc + j] = v;
    if (p.epi.activation == BLA_ACT_RELU) v = v < 0.f ? 0.f : v;
    if (p.epi.gate) v = p.epi.gate[(size_t)i * p.ldc + j] > 0.f ? v : 0.f;
    return v;
}

template <int CL>   // 1: one CTA per tile (cta_group::1); 2: CTA pairs (cta_group::2) -- separate kernels so that the
                    // single-CTA kernel carries no cluster-scoped instruction and can be launched without a cluster
__global__ void __launch_bounds__(kThreads, 1)
gemm_3xtf32_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tma_c, const TcParams p) {
    constexpr int kStages = Geo<CL>::kStages;
    const int nstages = p.ring > 0 && p.ring < kStages ? p.ring : kStages;   // BLA_TC_STAGES: a shorter ring, for latency experiments
    constexpr uint32_t kRawBytes = Geo<CL>::kRawBytes, kStageBytes = Geo<CL>::kStageBytes;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t staging_base = smem_base + kStages * kStageBytes;
    const uint32_t bar_base = staging_base + kStagingBytes;
    // barrier slots (8 bytes each)
    auto bar_full = [&](int s) { return bar_base + 8u * s; };
    auto bar_split = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto bar_empty = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto bar_acc_full = [&](int a) { return bar_base + 8u * (3 * kStages + a); };
    auto bar_acc_empty = [&](int a) { return bar_base + 8u * (3 * kStages + kAccStages + a); };
    const uint32_t tmem_slot = bar_base + 8u * (3 * kStages + 2 * kAccStages);
    // per stage, per splitter warp: does the lo tile of A / of B contain any non-zero?  (uint32 [kStages][4][2])
    const uint32_t lo_flags_u32 = bar_base + 256u;   // [kStages][2 * kSplitWarps] uint32: bit 0 = A lo tile non-zero, bit 1 = B
    volatile uint32_t* lo_flags = reinterpret_cast<volatile uint32_t*>(smem_gen + (lo_flags_u32 - smem_base));
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform by construction

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_split(s), kSplitWarps * CL);   // one arrival per splitter warp (of both CTAs of a pair)
            mbar_init(bar_empty(s), 1);      // tcgen05.commit (multicast to both CTAs of a pair)
        }
        for (int a = 0; a < kAccStages; ++a) {
            mbar_init(bar_acc_full(a), 1);   // tcgen05.commit
            mbar_init(bar_acc_empty(a), 4 * CL);  // one arrival per epilogue warp (of both CTAs of a pair)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // whole warp: allocate all 512 TMEM columns (2 accumulator stages x 256)
        if (CL == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {   // same warp id, same slot address in both CTAs of the pair
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;
    // Programmatic dependent launch (runtime.h): everything above touched only shared memory, TMEM and the kernel parameters, and may
    // have run while the previous kernel of the stream was still draining; from here on global memory is read and written.
    if (p.pdl_early) pdl_trigger();
    pdl_wait();

    // Work units.  cluster == 1: a unit is one 128 x bn tile.  cluster == 2: a unit is a 256 x bn tile computed by a CTA
    // PAIR with tcgen05.mma.cta_group::2: CTA r owns rows 128r.. (its own A tile and TMEM accumulator) and stages only
    // columns [r*bn/2, (r+1)*bn/2) of B; the tensor cores of both SMs read both halves.  What bounds the one-CTA kernel
    // is the rate at which TMA can fill one SM's shared memory with fp32 operands (measured ~24 B/clk/SM: a 1xTF32
    // run with a third of the MMAs was barely faster); the pair stages 16 KB instead of 24 KB per k-block and SM.
    constexpr int cl = CL;
    const uint32_t rank = cl == 2 ? cluster_ctarank() : 0u;
    if (cl == 2) cluster_sync_all();   // the peer's mbarriers are initialised before anything can signal them
    const int m_units = p.m_tiles / cl;
    const int total_tiles = m_units * p.n_tiles * p.splits;          // units
    const int unit0 = (int)blockIdx.x / cl, unit_stride = (int)gridDim.x / cl;
    auto decode = [&](int unit, int& split, int& m0, int& n0) {
        split = unit / (m_units * p.n_tiles);
        const int mn = unit % (m_units * p.n_tiles);
        m0 = ((mn % m_units) * cl + (int)rank) * BM;
        n0 = (mn / m_units) * p.bn;
    };

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (p.conv == 1) {
            // Implicit im2col, whole warp: lane `at` owns the at-th 32-pixel atom of the tile's B columns.  Its pixel
            // coordinates are decoded ONCE per tile (three integer divisions); per k-block = 16 channels of one filter tap
            // (ki, kj) every lane only adds the tap to its box origin and issues ONE 4-D box {16 c, bw, bh, images} of the
            // padded NHWC input = 32 K-major rows of 64 bytes.  (A single issuing lane doing the divisions per atom and
            // k-block was the bottleneck of the first version: 62 TFLOP/s.)
            const int ncols = p.bn / cl, natoms = ncols / 32;
            int stage = 0; uint32_t phase = 0;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
                int split, m0, n0;
                decode(tile, split, m0, n0);
                const int kb0 = split * p.kblocks_per_split, kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
                const int n = n0 + (int)rank * ncols * (cl - 1) + 32 * lane;
                const int img = n / p.cv_P, pix = n - img * p.cv_P;
                const int oi = pix / p.cv_Wo, oj = pix - oi * p.cv_Wo;
                const int x0 = oj * p.cv_stride - p.cv_pad_left, y0 = oi * p.cv_stride - p.cv_pad_top;
                int tap = kb0 / p.cv_cblocks, cb = kb0 - tap * p.cv_cblocks;
                int ki = tap / p.cv_k, kj = tap - ki * p.cv_k;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    mbar_wait(bar_empty(stage), phase ^ 1);
                    if (elect_one()) {   // one elected lane: its operands stay in uniform registers (see the plain GEMM branch)
                        mbar_arrive_expect_tx(bar_full(stage), p.stage_tx_bytes);
                        tma_load_2d(sa, &tma_a, kb * BK, m0, bar_full(stage));                     // weights [F][(ki, kj, c)]: K-major
                    }
                    __syncwarp();
                    if (lane < natoms && p.debug != 2)   // debug 2: no gather (the host drops the B bytes from the expected count)
                        tma_load_4d(sb + lane * (BK * 128), &tma_b, cb * BK, x0 + kj, y0 + ki, img, bar_full(stage));
                    if (++cb == p.cv_cblocks) { cb = 0; if (++kj == p.cv_k) { kj = 0; ++ki; } }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (p.conv == 2) {
            // Weight gradient, whole warp: the contraction runs over the output pixels.  k-block = 16 consecutive pixels of
            // one image; lane `at` owns the at-th 32-column atom of the tile = 32 channels of one filter tap (decoded once per
            // tile) and issues ONE box {32 c, bw, bh, 1} of the padded NHWC input shifted by its tap = 16 MN-major rows of
            // 128 bytes; lane 0 also loads the 128 x 16 tile of dy (3-D box {16 pixels, 128 filters, 1 image}).
            const int ncols = p.bn / cl, natoms = ncols / 32;
            const int blocks_per_img = p.cv_P / BK;
            int stage = 0; uint32_t phase = 0;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
                int split, m0, n0;
                decode(tile, split, m0, n0);
                const int kb0 = split * p.kblocks_per_split, kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
                const int nn = n0 + (int)rank * ncols * (cl - 1) + 32 * lane;
                const int tap = nn / p.cv_C, c0 = nn - tap * p.cv_C;
                const int ki = tap / p.cv_k, kj = tap - ki * p.cv_k;
                const int dx = kj - p.cv_pad_left, dy = ki - p.cv_pad_top;
                int img = kb0 / blocks_per_img, pix0 = (kb0 - img * blocks_per_img) * BK;
                int oi0 = pix0 / p.cv_Wo, oj0 = pix0 - oi0 * p.cv_Wo;
                for (int kb = kb0; kb < kb1; ++kb) {
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    mbar_wait(bar_empty(stage), phase ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_full(stage), p.stage_tx_bytes);
                        tma_load_3d(sa, &tma_a, pix0, m0, img, bar_full(stage));
                    }
                    __syncwarp();
                    if (lane < natoms)
                        tma_load_4d(sb + lane * (BK * 128), &tma_b, c0, oj0 * p.cv_stride + dx, oi0 * p.cv_stride + dy, img, bar_full(stage));
                    pix0 += BK; oj0 += BK;
                    while (oj0 >= p.cv_Wo) { oj0 -= p.cv_Wo; ++oi0; }
                    if (pix0 == p.cv_P) { pix0 = 0; oi0 = 0; oj0 = 0; ++img; }
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        } else {
            // Plain GEMM.  The whole warp walks the (tile, k-block) sequence and ONE ELECTED lane issues: coordinates and
            // barrier addresses then live in uniform registers (a `lane == 0` branch makes the compiler wrap every TMA operand in
            // a vote / broadcast retry loop).  L2 prefetch ahead of the loads was measured in round 1 (cp.async.bulk.prefetch:
            // -8..20 %, prefetch.global.L2 from a helper warp: -10 %) and again in round 2 on this loop (0..-5 %,
            // profiles/r02_prefetch_probe.txt): the feed is not bound by DRAM latency, and it is gone.
            int stage = 0; uint32_t phase = 0;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
                int split, m0, n0;
                decode(tile, split, m0, n0);
                const int kb0 = split * p.kblocks_per_split, kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
                const int half = p.bn / cl, nh = n0 + (int)rank * half * (cl - 1);   // a pair's CTA stages only ITS half of the B columns
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(bar_empty(stage), phase ^ 1);
                    if (elect_one()) {
                        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                        const int k0 = kb * BK;
                        mbar_arrive_expect_tx(bar_full(stage), p.stage_tx_bytes);
                        if (p.a_kmajor) {
                            tma_load_2d(sa, &tma_a, k0, m0, bar_full(stage));                              // box {16 k, 128 m}
                        } else if (p.a_box3) {
                            tma_load_3d(sa, &tma_a, 0, k0, m0 >> 5, bar_full(stage));                      // box {32 m, 16 k, 4 atoms}
                        } else {
#pragma unroll
                            for (int at = 0; at < BM / 32; ++at)                                            // box {32 m, 16 k} per atom
                                tma_load_2d(sa + at * (BK * 128), &tma_a, m0 + 32 * at, k0, bar_full(stage));
                        }
                        if (p.b_kmajor) {
                            tma_load_2d(sb, &tma_b, k0, nh, bar_full(stage));                              // box {16 k, bn (/2) n}
                        } else if (p.b_box3) {
                            tma_load_3d(sb, &tma_b, 0, k0, nh >> 5, bar_full(stage));                      // box {32 n, 16 k, atoms}
                        } else {
                            for (int at = 0; at < half / 32; ++at)
                                tma_load_2d(sb + at * (BK * 128), &tma_b, nh + 32 * at, k0, bar_full(stage));
                        }
                    }
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer =======================================
        // The WHOLE warp walks the loop (uniform control flow: barrier waits, flag reads and descriptor arithmetic stay in
        // uniform registers) and one elected lane issues.  Measured with ncu on the layer-1 GEMM: as a `lane == 0` branch the
        // compiler wrapped every tcgen05.mma operand in a vote / elect / broadcast retry loop, ~170 dependent instructions per
        // k-block executed by a single thread -- the issuer never waited for data, it WAS the bound (1,150 cycles per k-block
        // with the tensor pipe 45 % busy).
        if (rank == 0) {   // in a pair only the leader CTA issues (for both SMs)
            const uint32_t idesc = make_idesc(!p.a_kmajor, !p.b_kmajor, p.bn, cl == 2 ? 2 * BM : BM);
            // K-major (rows of BK floats = 64 B, SWIZZLE_64B): 8-row groups 512 B apart (SBO), k-step = 32 B
            // inside the swizzle row.
            // MN-major (SWIZZLE_128B_BASE32B): 32-wide atoms BK*128 B apart (LBO), 4-k groups 512 B apart
            // (SBO), k-step of 8 = 1024 B.
            const uint32_t a_lbo = p.a_kmajor ? 16u : BK * 128u, b_lbo = p.b_kmajor ? 16u : BK * 128u;
            const uint32_t a_lt = p.a_kmajor ? 4u : 1u, b_lt = p.b_kmajor ? 4u : 1u;
            const uint32_t a_kstep = (p.a_kmajor ? 32u : 1024u) >> 4, b_kstep = (p.b_kmajor ? 32u : 1024u) >> 4;
            // descriptor = constant fields | (shared address >> 4): built once, then only the 14-bit address field moves
            const uint64_t a_desc0 = make_desc(0u, a_lbo, 512u, a_lt), b_desc0 = make_desc(0u, b_lbo, 512u, b_lt);
            const bool skip_ok = p.debug != 4 && p.debug != 5;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
                int split, m0_, n0_;
                decode(tile, split, m0_, n0_);
                const int kb0 = split * p.kblocks_per_split;
                const int kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
                if (cl == 1) mbar_wait(bar_acc_empty(acc), acc_phase ^ 1);      // epilogue has drained this accumulator
                else mbar_wait_cluster(bar_acc_empty(acc), acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                uint32_t accumulate = 0;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (cl == 1) mbar_wait(bar_split(stage), phase);            // raw tile landed AND lo tile written
                    else mbar_wait_cluster(bar_split(stage), phase);            // ... in both CTAs of the pair
                    tcgen05_fence_after();
                    // lo-tile-is-zero shortcut: one word per splitter warp (in a pair the peer's arrive by st.async), read as
                    // 128-bit broadcasts
                    uint32_t f = 0;
#pragma unroll
                    for (int w = 0; w < kSplitWarps * cl / 4; ++w) {
                        uint32_t x, y, z, t;
                        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(t)
                                     : "r"(lo_flags_u32 + 4u * (uint32_t)(stage * 2 * kSplitWarps + 4 * w)));
                        f |= x | y | z | t;
                    }
                    const bool a_lo_nz = !skip_ok || (f & 1u), b_lo_nz = !skip_ok || (f & 2u);
                    const uint32_t sa = (smem_base + stage * kStageBytes) >> 4, sb = sa + (kABytes >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const uint64_t a_hi = a_desc0 + (sa + ks * a_kstep), b_hi = b_desc0 + (sb + ks * b_kstep);
                            const uint64_t a_lo = a_hi + (kRawBytes >> 4), b_lo = b_hi + (kRawBytes >> 4);
                            if (cl == 1) {
                                if (a_lo_nz) { umma_tf32(tmem_d, a_lo, b_hi, idesc, accumulate); accumulate = 1u; }
                                if (b_lo_nz) { umma_tf32(tmem_d, a_hi, b_lo, idesc, accumulate); accumulate = 1u; }
                                umma_tf32(tmem_d, a_hi, b_hi, idesc, accumulate);
                            } else {
                                if (a_lo_nz) { umma_tf32_pair(tmem_d, a_lo, b_hi, idesc, accumulate); accumulate = 1u; }
                                if (b_lo_nz) { umma_tf32_pair(tmem_d, a_hi, b_lo, idesc, accumulate); accumulate = 1u; }
                                umma_tf32_pair(tmem_d, a_hi, b_hi, idesc, accumulate);
                            }
                            accumulate = 1u;
                        }
                        // stage reusable once these MMAs retire; with a cluster the peer fills half of this stage and this CTA
                        // fills half of the peer's, so both CTAs' empty barriers hear from both MMA issuers
                        if (cl == 1) tcgen05_commit(bar_empty(stage));
                        else tcgen05_commit_pair(bar_empty(stage));
                    }
                    accumulate = 1u;
                    __syncwarp();
                    if (++stage == nstages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) {
                    if (cl == 1) tcgen05_commit(bar_acc_full(acc));    // accumulator complete -> epilogue (of both CTAs)
                    else tcgen05_commit_pair(bar_acc_full(acc));
                }
                __syncwarp();
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 6) {
        // ===================================== splitters ========================================
        const int t = threadIdx.x - 6 * 32;   // 0 .. 32*kSplitWarps-1
        int stage = 0; uint32_t phase = 0;
        for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
            int split, m0_, n0_;
            decode(tile, split, m0_, n0_);
            const int kb0 = split * p.kblocks_per_split;
            const int kb1 = min(p.kblocks, kb0 + p.kblocks_per_split);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bar_full(stage), phase);
                float4* raw = reinterpret_cast<float4*>(smem_gen + stage * kStageBytes);
                float4* lo = reinterpret_cast<float4*>(smem_gen + stage * kStageBytes + kRawBytes);
                const int n_f4 = (int)(kABytes / 16) + (p.bn / cl) * (BK * 4 / 16);   // A tile + the staged part of the B tile
                uint32_t nz_a = 0, nz_b = 0;
                if (p.debug != 1 && p.debug != 5)
#pragma unroll 4
                for (int i = t; i < n_f4; i += 32 * kSplitWarps) {
                    const float4 v = raw[i];
                    float4 r;
                    // lo = x - trunc_tf32(x), exact (at most 13 significant bits); the tensor core reads trunc_tf32(x) from the raw
                    // tile and truncates lo to its top 11 bits: |error| <= 2^-22 |x|, the size of the dropped lo.lo term.  (Rounding
                    // lo with cvt.rna.tf32 first -- round 1 -- compiles to a 4-instruction emulation per element; the splitter warps
                    // were then the bound of the kernel: ncu, 950 cycles per k-block against 512 of MMA.)
                    r.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                    r.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                    r.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                    r.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                    uint32_t any = (__float_as_uint(r.x) | __float_as_uint(r.y)) | (__float_as_uint(r.z) | __float_as_uint(r.w));
                    lo[i] = r;
                    any &= 0x7FFFFFFFu;                       // -0.0 is still zero
                    if (i < (int)(kABytes / 16)) nz_a |= any; else nz_b |= any;
                }
                // Operands that are exactly representable in TF32 (pixel values, one-hot labels, small integers) have an
                // all-zero lo tile: the MMA issuer skips the product with it -- bit-identical result, a third less work.
                const bool wa = __any_sync(0xffffffffu, nz_a != 0), wb = __any_sync(0xffffffffu, nz_b != 0);
                const int slot = stage * 2 * kSplitWarps + (int)rank * kSplitWarps + (warp - 6);
                const uint32_t fl = (wa ? 1u : 0u) | (wb ? 2u : 0u);
                fence_proxy_async();      // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if (cl == 1) {
                        lo_flags[slot] = fl;
                        mbar_arrive(bar_split(stage));
                    } else if (rank == 0) {   // leader CTA (its MMA lane issues for the pair): the first warp also arms the barrier
                        lo_flags[slot] = fl;  // for the peer's kSplitWarps flag words
                        if (warp == 6) mbar_arrive_expect_tx(bar_split(stage), 4u * kSplitWarps);
                        else mbar_arrive(bar_split(stage));
                    } else {   // peer CTA: flag word and arrival both go to the leader's barrier
                        const uint32_t lbar = mapa_shared(bar_split(stage), 0);
                        st_async_u32(mapa_shared(lo_flags_u32 + 4u * slot, 0), fl, lbar);
                        mbar_arrive_cluster(lbar);
                    }
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================== epilogue (warps 2-5) =============================
        const int q = warp & 3;                      // TMEM lane quarter this warp may access
        float* stg = reinterpret_cast<float*>(smem_gen + kStages * kStageBytes + (warp - 2) * 8192);
        const uint32_t stg_u32 = staging_base + (warp - 2) * 8192;
        int sbuf = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
            int split, m0, n0;
            decode(tile, split, m0, n0);
            const int n_end = min(p.n, n0 + p.bn);
            const int row_base = m0 + 32 * q;
            if (p.epi.gate && !p.gate_bits && !p.partial && row_base + lane < p.m) {
                // The relu' gate is one 128-byte row segment per lane and 32-column chunk.  Loaded on demand it exposes a DRAM
                // round trip per chunk (the K = 128 dgrad of the MLP spent 21 us per tile here); asked for now, while this
                // tile's MMAs are still running, it is an L2 hit by the time the accumulator is ready.
                const float* grow = p.epi.gate + (size_t)(row_base + lane) * p.ldc;
                for (int col0 = n0; col0 < n_end; col0 += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(grow + col0));
            }
            // bit-form gate: one word per lane (row) and 32-column chunk, fetched one chunk ahead -- the first one NOW, behind the tile's MMAs
            const uint32_t* gate_row = p.gate_bits ? p.gate_bits + (size_t)min(row_base + lane, p.m - 1) * p.gate_ld + (n0 >> 5) : nullptr;
            uint32_t gate_next = (gate_row && !p.partial) ? __ldg(gate_row) : 0u;
            mbar_wait(bar_acc_full(acc), acc_phase);
            tcgen05_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < BN / 32; ++ch) {
                const int col0 = n0 + 32 * ch;
                if (col0 >= n_end) break;            // warp-uniform
                const uint32_t gate_w = gate_next;
                if (gate_row && !p.partial && col0 + 32 < n_end) gate_next = __ldg(gate_row + ch + 1);
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN + 32 * ch), v);
                if (p.tma_store) {
                    // Fast path: the epilogue runs on the accumulator row each lane already holds (its bias is a
                    // per-lane constant, its gate row is 128 contiguous bytes), the 32x32 tile is parked in
                    // shared memory in the SWIZZLE_128B pattern and ONE bulk tensor store writes it out
                    // (clipped at the matrix edge by the tensor map).  Two staging tiles per warp.
                    const int i = row_base + lane;
                    const bool row_ok = i < p.m;
                    float o[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[e] = __uint_as_float(v[e]);
                    if (!p.partial) {
                        if (p.epi.alpha != 0.f) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] *= p.epi.alpha;
                        }
                        if (p.epi.bias_rows) {
                            const float brow = row_ok ? __ldg(p.epi.bias_rows + i) : 0.f;
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] += brow;
                        }
                        if (p.epi.activation == BLA_ACT_RELU) {
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = o[e] < 0.f ? 0.f : o[e];
                        }
                        if (p.mask_out && row_ok) {   // the sign pattern of this row's 32 columns, for a later relu' gate
                            uint32_t w = 0;
#pragma unroll
                            for (int e = 0; e < 32; ++e) w |= (o[e] > 0.f ? 1u : 0u) << e;
                            p.mask_out[(size_t)i * p.mask_ld + (col0 >> 5)] = w;
                        }
                        if (p.gate_bits && row_ok) {   // 4 bytes per row and chunk instead of 128
                            const uint32_t w = gate_w;
#pragma unroll
                            for (int e = 0; e < 32; ++e) o[e] = (w >> e) & 1u ? o[e] : 0.f;
                        } else if (p.epi.gate && row_ok) {
                            const float* gp = p.epi.gate + (size_t)i * p.ldc + col0;
                            if (col0 + 31 < p.n) {
                                float4 g4[8];
#pragma unroll
                                for (int c4 = 0; c4 < 8; ++c4) g4[c4] = __ldg(reinterpret_cast<const float4*>(gp) + c4);
#pragma unroll
                                for (int c4 = 0; c4 < 8; ++c4) {
                                    o[4 * c4 + 0] = g4[c4].x > 0.f ? o[4 * c4 + 0] : 0.f; o[4 * c4 + 1] = g4[c4].y > 0.f ? o[4 * c4 + 1] : 0.f;
                                    o[4 * c4 + 2] = g4[c4].z > 0.f ? o[4 * c4 + 2] : 0.f; o[4 * c4 + 3] = g4[c4].w > 0.f ? o[4 * c4 + 3] : 0.f;
                                }
                            } else {
#pragma unroll
                                for (int e = 0; e < 32; ++e)
                                    if (col0 + e < p.n) o[e] = __ldg(gp + e) > 0.f ? o[e] : 0.f;
                            }
                        }
                    }
                    if (p.conv == 1 && !p.partial && (p.cv_bias || p.cv_add) && row_ok) {
                        // fused neighbours of the U-Net's convolutions: y += bias[img][filter] and / or y += addend[img][filter][pixel]
                        // (every group of 4 columns lies inside one image: P is a multiple of 4)
#pragma unroll
                        for (int half = 0; half < 2; ++half) {   // four 128-bit addend loads in flight at a time (register budget)
                            float4 a4[4];
                            if (p.cv_add) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int col = col0 + 4 * (4 * half + q), img = col / p.cv_P, px = col - img * p.cv_P;
                                    a4[q] = col < p.n ? __ldg(reinterpret_cast<const float4*>(p.cv_add + ((size_t)img * p.m + i) * p.cv_P + px))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int c4 = 4 * half + q, col = col0 + 4 * c4, img = col / p.cv_P;
                                const float tb = (p.cv_bias && col < p.n) ? __ldg(p.cv_bias + (size_t)img * p.m + i) : 0.f;
                                o[4 * c4] += tb; o[4 * c4 + 1] += tb; o[4 * c4 + 2] += tb; o[4 * c4 + 3] += tb;
                                if (p.cv_add) { o[4 * c4] += a4[q].x; o[4 * c4 + 1] += a4[q].y; o[4 * c4 + 2] += a4[q].z; o[4 * c4 + 3] += a4[q].w; }
                            }
                        }
                    }
                    if (p.conv == 1 && !p.partial && p.cv_P < 32) {
                        // small images (P = 4, 8 or 16 output pixels): the 32 columns of this chunk are 32/P whole images; lane =
                        // filter row writes P contiguous floats per image straight from its registers (64-byte runs for 4 x 4)
                        if (row_ok) {
                            const int img0 = col0 / p.cv_P;
#pragma unroll
                            for (int c4 = 0; c4 < 8; ++c4) {
                                const int col = 4 * c4, ii = col / p.cv_P, px = col - ii * p.cv_P;
                                if (col0 + col < p.n)
                                    *reinterpret_cast<float4*>(p.c + ((size_t)(img0 + ii) * p.m + i) * p.cv_P + px) =
                                        make_float4(o[col], o[col + 1], o[col + 2], o[col + 3]);
                            }
                        }
                        continue;
                    }
                    if (lane == 0) tma_store_wait_read1();     // the store that last used this staging tile has been read
                    __syncwarp();
                    float* sb = stg + sbuf * 1024;
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4)
                        *reinterpret_cast<float4*>(sb + lane * 32 + ((c4 ^ (lane & 7)) << 2)) =
                            make_float4(o[4 * c4], o[4 * c4 + 1], o[4 * c4 + 2], o[4 * c4 + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.conv != 1 || p.partial) {
                            tma_store_2d(&tma_c, col0, (p.partial ? split * p.m : 0) + row_base, stg_u32 + sbuf * 4096);
                        } else {
                            const int img = col0 / p.cv_P, pix = col0 - img * p.cv_P;
                            if (p.debug == 3) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }   // experiment: no store
                            else tma_store_3d(&tma_c, pix, row_base, img, stg_u32 + sbuf * 4096);   // y as {P, F, img}
                        }
                    }
                    sbuf ^= 1;
                    continue;
                }
                // lane r holds row r, columns col0..col0+31: park it in shared memory with the 16-byte
                // chunks XOR-swizzled by the row so that both the column-wise writes here and the
                // row-wise reads below are bank-conflict free
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float4 f = make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]), __uint_as_float(v[4 * c4 + 2]),
                                           __uint_as_float(v[4 * c4 + 3]));
                    *reinterpret_cast<float4*>(stg + lane * 32 + ((c4 ^ (lane & 7)) << 2)) = f;
                }
                __syncwarp();
                // now 8 lanes cover one 128-byte row: 4 rows per pass, 8 passes.  All shared-memory reads
                // and all gate loads are issued before the first global store: the output may alias the
                // gate as far as the compiler knows, and interleaving would serialise on DRAM latency.
                const int c4 = lane & 7;
                const int j = col0 + 4 * c4;
                float4 f[8], gt[8];
#pragma unroll
                for (int pass = 0; pass < 8; ++pass) {
                    const int r = pass * 4 + (lane >> 3);
                    f[pass] = *reinterpret_cast<const float4*>(stg + r * 32 + ((c4 ^ (r & 7)) << 2));
                }
                const bool full4 = j + 3 < n_end;
                if (p.epi.gate && !p.partial) {
#pragma unroll
                    for (int pass = 0; pass < 8; ++pass) {
                        const int i = row_base + pass * 4 + (lane >> 3);
                        gt[pass] = make_float4(1.f, 1.f, 1.f, 1.f);
                        if (i < p.m && j < n_end) {
                            const float* gp = p.epi.gate + (size_t)i * p.ldc + j;
                            if (p.c_vec && full4) {
                                gt[pass] = __ldg(reinterpret_cast<const float4*>(gp));
                            } else {
                                gt[pass].x = gp[0];
                                if (j + 1 < n_end) gt[pass].y = gp[1];
                                if (j + 2 < n_end) gt[pass].z = gp[2];
                                if (j + 3 < n_end) gt[pass].w = gp[3];
                            }
                        }
                    }
                }
#pragma unroll
                for (int pass = 0; pass < 8; ++pass) {
                    const int i = row_base + pass * 4 + (lane >> 3);
                    if (i < p.m && j < n_end) {
                        float o[4] = {f[pass].x, f[pass].y, f[pass].z, f[pass].w};
                        if (p.partial) {
                            float* dst = p.partial + ((size_t)split * p.m + i) * p.n + j;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (j + jj < n_end) dst[jj] = o[jj];
                        } else {
                            const float g4[4] = {gt[pass].x, gt[pass].y, gt[pass].z, gt[pass].w};
                            const float brow = p.epi.bias_rows ? p.epi.bias_rows[i] : 0.f;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                float v = o[jj];
                                if (p.epi.alpha != 0.f) v *= p.epi.alpha;
                                v += brow;
                                if (p.epi.bias_cols && j + jj < n_end) v += p.epi.bias_cols[j + jj];
                                if (p.epi.pre_activation && j + jj < n_end) p.epi.pre_activation[(size_t)i * p.ldc + j + jj] = v;
                                if (p.epi.activation == BLA_ACT_RELU) v = v < 0.f ? 0.f : v;
                                if (p.epi.gate) v = g4[jj] > 0.f ? v : 0.f;
                                o[jj] = v;
                            }
                            float* dst = p.c + (size_t)i * p.ldc + j;
                            if (p.c_vec && full4) {
                                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                            } else {
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj)
                                    if (j + jj < n_end) dst[jj] = o[jj];
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (cl == 1) mbar_arrive(bar_acc_empty(acc));
                else mbar_arrive_cluster(mapa_shared(bar_acc_empty(acc), 0));   // the leader's MMA lane waits on it
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (p.tma_store && lane == 0) tma_store_wait_all();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (cl == 2) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or signal its barriers
    if (warp == 1) {
        tcgen05_fence_after();
        if (cl == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// sums split-K partials in a fixed order and applies the epilogue.  Block = 64 float4 columns x 4 split groups:
// every thread keeps 8 independent 128-bit loads in flight, the four group sums are combined through shared memory.
__global__ void __launch_bounds__(256) tc_splitk_reduce_kernel(const TcParams p) {
    if (p.pdl_early) pdl_trigger();
    pdl_wait();   // launched with programmatic serialisation right behind the GEMM whose partials it folds
    const size_t total = (size_t)p.m * p.n;
    if ((total & 3) == 0 && (p.n & 3) == 0 && (p.c_vec || p.conv == 1 || p.cv_final)) {
        __shared__ float4 part[4][64];
        const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
        const size_t total4 = total >> 2;
        for (size_t base = (size_t)blockIdx.x * 64; base < total4; base += (size_t)gridDim.x * 64) {
            const size_t e4 = base + col;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e4 < total4) {
                const float4* src = reinterpret_cast<const float4*>(p.partial) + e4;
#pragma unroll 8
                for (int z = grp; z < p.splits; z += 4) {
                    const float4 v = src[(size_t)z * total4];
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
            part[grp][col] = acc;
            __syncthreads();
            uint32_t nib = 0;
            if (grp == 0 && e4 < total4) {
                float4 t = part[0][col];
#pragma unroll
                for (int g2 = 1; g2 < 4; ++g2) { t.x += part[g2][col].x; t.y += part[g2][col].y; t.z += part[g2][col].z; t.w += part[g2][col].w; }
                const size_t e = e4 << 2;
                const int i = (int)(e / p.n), j = (int)(e % p.n);
                if (p.conv == 1) {   // row i = filter, column j = (image, pixel) -> y [img][F][P]; P is a multiple of 4
                    const int img = j / p.cv_P, pix = j - img * p.cv_P;
                    const size_t at = ((size_t)img * p.m + i) * p.cv_P + pix;
                    if (p.cv_bias) { const float tb = p.cv_bias[(size_t)img * p.m + i]; t.x += tb; t.y += tb; t.z += tb; t.w += tb; }
                    if (p.cv_add) { const float4 a = *reinterpret_cast<const float4*>(p.cv_add + at); t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w; }
                    *reinterpret_cast<float4*>(p.c + at) = t;
                } else if (p.cv_final) {   // weight gradient: column j = (tap, padded channel) -> dW [F][C][k*k], padding channels dropped
                    const int tap = j / p.cv_C, c = j - tap * p.cv_C, k2 = p.cv_k * p.cv_k;
                    float* dst = p.c + ((size_t)i * p.cv_Creal + c) * k2 + tap;
                    if (c < p.cv_Creal) dst[0] = t.x;
                    if (c + 1 < p.cv_Creal) dst[k2] = t.y;
                    if (c + 2 < p.cv_Creal) dst[2 * k2] = t.z;
                    if (c + 3 < p.cv_Creal) dst[3 * k2] = t.w;
                } else {
                    float4 o;
                    o.x = epilogue_value(t.x, i, j, p); o.y = epilogue_value(t.y, i, j + 1, p);
                    o.z = epilogue_value(t.z, i, j + 2, p); o.w = epilogue_value(t.w, i, j + 3, p);
                    *reinterpret_cast<float4*>(p.c + (size_t)i * p.ldc + j) = o;
                    if (p.mask_out) nib = (o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) | (o.w > 0.f ? 8u : 0u);
                }
            }
            if (p.mask_out && grp == 0) {
                // n % 32 == 0 (checked by the host): the 8 lanes that share a 32-column word sit in one row; their nibbles meet by shuffles
                uint32_t w = nib << (4 * (col & 7));
                w |= __shfl_xor_sync(0xffffffffu, w, 1); w |= __shfl_xor_sync(0xffffffffu, w, 2); w |= __shfl_xor_sync(0xffffffffu, w, 4);
                if ((col & 7) == 0 && e4 < total4) {
                    const size_t e = e4 << 2;
                    p.mask_out[(size_t)(e / p.n) * p.mask_ld + (e % p.n >> 5)] = w;
                }
            }
            __syncthreads();
        }
        return;
    }
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (size_t)gridDim.x * 256) {
        float s = 0.f;
        for (int z = 0; z < p.splits; ++z) s += p.partial[(size_t)z * total + e];
        const int i = (int)(e / p.n), j = (int)(e % p.n);
        p.c[(size_t)i * p.ldc + j] = epilogue_value(s, i, j, p);
    }
}

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)sym;
        else
            cudaGetLastError();
    }
    return fn;
}

// 2-D fp32 tensor map over a row-major [rows][cols] array with leading dimension ld.
//   K-major operand : box = {BK cols (64 B), box_rows}, SWIZZLE_64B
//   MN-major operand: box = {32 cols (128 B), BK rows}, SWIZZLE_128B with 32-byte atoms
bool make_map(CUtensorMap* map, const float* base, int rows, int cols, int ld, int box_rows, bool mn_major) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {mn_major ? 32u : (cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t elem[2] = {1u, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// MN-major operand, all atoms of a tile in one box.  The row-major [rows = k][cols = m or n] array is described as a 3-D tensor
// {32 columns of an atom, k, atom} (strides ld and 32 floats): a box {32, BK, atoms} lands as `atoms` consecutive [BK][32] blocks,
// exactly what `atoms` 2-D boxes produce -- but as ONE request.  The TMA engine costs ~105 cycles per box plus ~86 B/clk
// (profiles/probes/tma_feed_probe.cu): eight 2 KB boxes take 1030 cycles, one 16 KB box 300; the short-K GEMMs of the MLP were bound
// by exactly that.  The map pretends the last atom is whole: the caller checks that the bytes behind a ragged last row exist.
bool make_map_mn3(CUtensorMap* map, const float* base, int rows, int cols, int ld, int atoms) {
    cuuint64_t dims[3] = {32u, (cuuint64_t)rows, (cuuint64_t)((cols + 31) / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 32u * sizeof(float)};
    cuuint32_t box[3] = {32u, (cuuint32_t)BK, (cuuint32_t)atoms};
    cuuint32_t elem[3] = {1u, 1u, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}
// may the 3-D form read the [rows][cols] array at `base`?  Columns past `cols` in the last atom are read from the next row (harmless:
// they only feed output rows / columns that are clipped) -- for the LAST row they must still be inside the allocation.
bool mn3_readable(const float* base, int rows, int cols, int ld) {
    const int padded = (cols + 31) / 32 * 32;
    if (padded <= ld) return true;
    const size_t need = ((size_t)(rows - 1) * ld + padded) * sizeof(float);
    return owned_bytes_from(base) >= need;
}

// fp32 tensor map for the epilogue's 32x32 tile stores (SWIZZLE_128B, clipped at the matrix edge)
bool make_map_c(CUtensorMap* map, float* base, long long rows, int cols, int ld) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t elem[2] = {1u, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 4-D map over the conv input in padded NHWC form xp [img][Hp][Wp][C]: a box {16 c, bw*s, bh*s, 32/(bw*bh)} with traversal
// strides {1, s, s, 1} lands in shared memory as 32 (output pixels) rows of 16 channels = 64 bytes: 32 rows of a K-major
// B tile of the implicit im2col matrix.  (Channels must be innermost: a TMA box has to start on a 16-byte boundary of the
// innermost dimension, so a one-pixel filter-tap shift is impossible with W innermost -- profiles/probes/tma4d_probe.cu.)
bool make_map_conv_in(CUtensorMap* map, const ConvTc& cv, int bw, int bh) {
    const int nimg = 32 / (bw * bh);
    cuuint64_t dims[4] = {(cuuint64_t)cv.C, (cuuint64_t)cv.W, (cuuint64_t)cv.H, (cuuint64_t)cv.imgs};
    cuuint64_t strides[3] = {(cuuint64_t)cv.C * 4, (cuuint64_t)cv.W * cv.C * 4, (cuuint64_t)cv.H * cv.W * cv.C * 4};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(bw * cv.stride), (cuuint32_t)(bh * cv.stride), (cuuint32_t)nimg};
    cuuint32_t elem[4] = {1u, (cuuint32_t)cv.stride, (cuuint32_t)cv.stride, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)cv.in, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Weight gradient, B operand: the same padded NHWC input as MN-major atoms -- a box {32 c, bw*s, bh*s, 1} with traversal strides
// {1, s, s, 1} lands as 16 (output pixels = k) rows of 32 channels = 128 bytes, 32-byte swizzle atoms (the layout the tensor core
// wants for MN-major 32-bit operands).
bool make_map_conv_in_mn(CUtensorMap* map, const ConvTc& cv) {
    const int bw = cv.Wo < BK ? cv.Wo : BK, bh = BK / bw;
    cuuint64_t dims[4] = {(cuuint64_t)cv.C, (cuuint64_t)cv.W, (cuuint64_t)cv.H, (cuuint64_t)cv.imgs};
    cuuint64_t strides[3] = {(cuuint64_t)cv.C * 4, (cuuint64_t)cv.W * cv.C * 4, (cuuint64_t)cv.H * cv.W * cv.C * 4};
    cuuint32_t box[4] = {32u, (cuuint32_t)(bw * cv.stride), (cuuint32_t)(bh * cv.stride), 1u};
    cuuint32_t elem[4] = {1u, (cuuint32_t)cv.stride, (cuuint32_t)cv.stride, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)cv.in, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// Weight gradient, A operand: dy [img][F][P] as {P, F, img}; a box {16 pixels, 128 filters, 1} is a K-major 128 x 16 tile.
bool make_map_conv_dy(CUtensorMap* map, const float* dy, const ConvTc& cv) {
    const int P = cv.Ho * cv.Wo;
    cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)cv.F, (cuuint64_t)cv.imgs};
    cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)cv.F * P * 4};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1u};
    cuuint32_t elem[3] = {1u, 1u, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)dy, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 3-D map over the conv output y [img][F][P] as {P, F, img} for the epilogue's 32 (filters) x 32 (pixels) tile stores.  Images
// with fewer than 32 pixels are written straight from registers instead (a tile then spans several images).
bool make_map_conv_out(CUtensorMap* map, const ConvTc& cv) {
    const int P = cv.Ho * cv.Wo;
    const int Pd = P >= 32 ? P : 32;      // the map is not used for small images; keep it encodable
    cuuint64_t dims[3] = {(cuuint64_t)Pd, (cuuint64_t)cv.F, (cuuint64_t)cv.imgs};
    cuuint64_t strides[2] = {(cuuint64_t)P * 4, (cuuint64_t)cv.F * P * 4};
    cuuint32_t box[3] = {32u, 32u, 1u}, elem[3] = {1u, 1u, 1u};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)cv.out, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

bool g_tc_broken = false;
unsigned long long g_tc_launches = 0;

}  // namespace

bool gemm_3xtf32(const GemmArgs& g, cudaStream_t s) {
    if (g_tc_broken || !encode_fn()) return false;
    if (g.m <= 0 || g.n <= 0 || g.k <= 0) return false;
    // TMA needs 16-byte aligned bases and row pitches
    const int cmode = g.conv ? g.conv->mode : 0;   // 0 plain GEMM, 1 conv forward, 2 conv weight gradient
    if (((uintptr_t)g.a & 15) || (cmode != 2 && (g.lda & 3))) return false;
    if (!g.conv && (((uintptr_t)g.b & 15) || (g.ldb & 3))) return false;

    TcParams p{};
    p.m = g.m; p.n = g.n; p.k = g.k;
    p.a_kmajor = !g.ta;           // A row-major [m][k]  -> K-major;  stored [k][m] -> MN-major
    p.b_kmajor = g.conv ? cmode == 1 : g.tb;   // B stored [n][k] (and the forward conv's NHWC gather) -> K-major;  row-major [k][n] -> MN-major
    p.c = g.c; p.ldc = g.ldc; p.epi = g.epi;
    p.mask_out = g.mask_out; p.mask_ld = g.mask_ld; p.gate_bits = g.epi.gate ? g.gate_bits : nullptr; p.gate_ld = g.gate_ld;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("BLA_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.debug = dbg; }
    p.pdl_early = pdl_early() ? 1 : 0;
    { static int ring = -1; if (ring < 0) { const char* e = getenv("BLA_TC_STAGES"); ring = e ? atoi(e) : 0; } p.ring = ring; }
    p.c_vec = (((uintptr_t)g.c & 15) == 0) && (g.ldc % 4 == 0);
    p.m_tiles = ceil_div(g.m, BM);
    p.n_tiles = ceil_div(g.n, BN);
    p.kblocks = ceil_div(g.k, BK);
    // Tile width: among the legal UMMA widths pick the one with the least (waves x width) so that the
    // last wave of the persistent grid is not mostly idle (470 tiles on 148 SMs = 4 waves at width 256,
    // but also 4 waves at width 224), then shrink it to the narrowest width with the same tile count
    // (784 columns -> 4 x 208 instead of 4 x 256).
    // Clusters of 2 (TMA multicast of the B tile) need an even number of m tiles; each CTA then fetches half of
    // the B tile, so the width must split into two whole 32-column atoms.
    { static int en = -1; if (en < 0) { const char* e = getenv("BLA_TC_CLUSTER"); en = e ? atoi(e) : 1; }
      // short contractions stay on the single-CTA kernel: the pair's cluster start-up and cross-SM barrier hops cost
      // more than the halved staging saves (K = 128: 62 us as a pair, 33 us alone)
      p.cluster = (en && p.m_tiles >= 2 && p.m_tiles % 2 == 0 && g.k >= 512) ? 2 : 1; }
    if (g.conv) {
        const ConvTc& cv = *g.conv;
        p.cv_bias = cmode == 1 ? cv.bias : nullptr; p.cv_add = cmode == 1 ? cv.addend : nullptr;
        p.conv = cmode; p.cv_C = cv.C; p.cv_P = cv.Ho * cv.Wo; p.cv_Wo = cv.Wo; p.cv_stride = cv.stride; p.cv_pad_top = cv.pad_top; p.cv_pad_left = cv.pad_left;
        p.cv_k = cv.k; p.cv_cblocks = cv.C / BK;
        p.cv_bw = cv.Wo < 32 ? cv.Wo : 32;
        p.cv_bh = 32 / p.cv_bw < cv.Ho ? 32 / p.cv_bw : cv.Ho;
    }
    const int gran = p.cluster == 2 ? 64 : 32;   // whole 32-column epilogue chunks / MN-major atoms (per CTA half)
    bool force_no_split = false;
    int conv_splits = 0;   // > 0: chosen together with the tile width below (conv forward only: measured on B200, the same model
                           // LOSES for the weight gradient (U-Net step 13.7 -> 13.9 ms) and for small plain GEMMs (1024^3: 70 -> 61 TFLOP/s))
    if (cmode == 1) {
        // Convolutions over small images have few output columns (64 images of 4 x 4 pixels: 1024) but a long contraction
        // (k*k*C up to 4608): pick tile width AND split-K together from a cycle model of one k-block -- the 3 x 2 MMAs of a
        // 128 x bn x 8 tile (185 cycles at bn = 256) against the TMA fill of the stage (~24 B/clk/SM) -- plus the cost of
        // writing and re-reading the partial sums.
        const int slots = rt().num_sms / p.cluster;
        const int m_units = p.m_tiles / p.cluster;
        double best = -1.0;
        int best_bn = BN, best_sp = 1;
        for (int bn = BN; bn >= gran; bn -= gran) {
            const long long tiles_ = (long long)m_units * ceil_div(g.n, bn);
            const double mma = 6.0 * 185.0 * bn / 256.0, fill = (kABytes + (double)(bn / p.cluster) * BK * 4) / 24.0;
            const double per_kb = (mma > fill ? mma : fill) + 40.0;
            for (int sp = 1; sp <= 32; ++sp) {
                if (sp > 1 && (p.kblocks / sp < 8 || (cmode == 1 && g.m % BM != 0))) break;
                const long long waves = (tiles_ * sp + slots - 1) / slots;
                const int kb = ceil_div(p.kblocks, sp);
                double cost = (double)waves * (kb * per_kb + 1500.0);                                  // + tile prologue / epilogue
                if (sp > 1) cost += (double)(sp + 1) * g.m * g.n * 4.0 / 1500.0 + 4000.0;               // partials out and back, 2nd launch
                if (best < 0 || cost < best) { best = cost; best_bn = bn; best_sp = sp; }
            }
        }
        p.n_tiles = ceil_div(g.n, best_bn);
        p.bn = best_bn;
        conv_splits = best_sp;
    } else {
        const int slots = rt().num_sms / p.cluster;
        const int m_units = p.m_tiles / p.cluster;
        double best_cost = -1.0;
        int best_bn = BN;
        for (int bn = BN; bn >= 192; bn -= gran) {
            const int nt = ceil_div(g.n, bn);
            int w = (ceil_div(g.n, nt) + gran - 1) / gran * gran;             // the width this choice ends up with (shrunk below)
            if (w > BN) w = BN;
            const long long units_ = (long long)m_units * nt;
            // split-K as decided further down: few tiles and a long contraction -> up to 64 slices, one wave
            long long sp = 1;
            if (units_ * p.cluster * 2 <= rt().num_sms && p.kblocks >= 32) {
                sp = slots / units_;
                if (sp > p.kblocks / 16) sp = p.kblocks / 16;
                if (sp > 64) sp = 64;
                if (sp < 1) sp = 1;
            }
            const long long waves = (units_ * sp + slots - 1) / slots;
            const long long kb = (p.kblocks + sp - 1) / sp;
            // measured on B200 (square 4096): a 192-wide tile costs 1.33x more per column than a 256-wide one
            // (the A tile and its split are amortised over fewer columns).  Short contractions (K <= 256: the MLP's layer-2
            // forward and dgrad) are bound by the epilogue, whose cost is proportional to the width: no penalty there.
            // With split-K the slices per tile and the 64-slice cap enter too: 256 x 784 x 60000 is faster as 4 x 256 columns
            // x 18 slices (124 us) than as 5 x 192 x 14 (148 us), 128 x 256 x 60000 faster as 2 x 128 x 64 than as 1 x 256 x 64.
            const double cost = (double)waves * (double)kb * w * (p.kblocks <= 16 ? 1.0 : 1.0 + 0.33 * (BN - w) / 64.0);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_bn = bn; }
        }
        { static int force = -1; if (force < 0) { const char* e = getenv("BLA_TC_BN"); force = e ? atoi(e) : 0; }   // tuning probe
          if (force >= gran && force <= BN && force % gran == 0) best_bn = force; }
        p.n_tiles = ceil_div(g.n, best_bn);
        p.bn = (ceil_div(g.n, p.n_tiles) + gran - 1) / gran * gran;
        if (p.bn > BN) p.bn = BN;
        // Less than one wave of tiles (the tail of the MLP's layer-1 GEMM; every GEMM of a 7,500-column data-parallel shard): wide
        // tiles leave SMs idle -- or, for a long contraction, take the split-K route with its second launch and (splits + 1) passes
        // over the result.  The alternative is ONE wave of narrow tiles on all SMs.  Cycle model of one k-block per CTA
        // (profiles/r02_tma_feed_probe.txt: 105 cycles per TMA box + 86 B/clk; one 128 x 256 x 8 MMA = 128 cycles; ~13 splitter
        // instructions per float4 on 8 warps) plus ~600 cycles of epilogue per 32-column chunk.
        static int narrow_on = -1;
        if (narrow_on < 0) { const char* e = getenv("BLA_TC_NARROW"); narrow_on = e ? atoi(e) : 1; }
        const long long units_full = (long long)m_units * ceil_div(g.n, BN);
        if (narrow_on && cmode == 0 && (long long)m_units * p.n_tiles < slots && slots / m_units >= 1) {
            int bn_n = (ceil_div(g.n, slots / m_units) + gran - 1) / gran * gran;
            if (bn_n < gran) bn_n = gran;
            if (bn_n < p.bn) {
                auto kb_cycles = [&](int bn) {
                    const double cols = (double)bn / p.cluster;
                    const double mma = 6.0 * 128.0 * bn / 256.0, tma = 2 * 105.0 + (kABytes + cols * BK * 4) / 86.0;
                    const double split = (kABytes + cols * BK * 4) / 16.0 * 13.0 / 64.0;
                    double t = mma > tma ? mma : tma;
                    if (split > t) t = split;
                    return t + 120.0;
                };
                double t_now;
                if (units_full * p.cluster * 2 <= rt().num_sms && p.kblocks >= 32) {   // today: split-K over full-width tiles
                    long long sp = slots / units_full;
                    if (sp > p.kblocks / 16) sp = p.kblocks / 16;
                    if (sp > 64) sp = 64;
                    if (sp < 1) sp = 1;
                    t_now = (double)ceil_div(p.kblocks, sp) * kb_cycles(BN) + (sp > 1 ? (double)(sp + 1) * g.m * g.n * 4.0 / 3000.0 + 6000.0 : 600.0 * BN / 32);
                } else {
                    t_now = (double)p.kblocks * kb_cycles(p.bn) + 600.0 * p.bn / 32;
                }
                const double t_narrow = (double)p.kblocks * kb_cycles(bn_n) + 600.0 * bn_n / 32;
                if (t_narrow < t_now) {
                    p.bn = bn_n;
                    p.n_tiles = ceil_div(g.n, bn_n);
                    force_no_split = true;
                }
            }
        }
    }
    p.stage_tx_bytes = kABytes + (uint32_t)(p.bn / p.cluster) * BK * 4;   // per CTA: its A tile + its share of the B tile
    if (p.conv && p.debug == 2) p.stage_tx_bytes = kABytes;

    CUtensorMap ma, mb;
    bool ok;
    { static int en = -1; if (en < 0) { const char* e = getenv("BLA_TC_BOX3"); en = e ? atoi(e) : 1; }   // A/B switch
      p.a_box3 = en && cmode == 0 && !p.a_kmajor && mn3_readable(g.a, g.k, g.m, g.lda);
      p.b_box3 = en && cmode == 0 && !p.b_kmajor && mn3_readable(g.b, g.k, g.n, g.ldb); }
    if (cmode == 2) ok = make_map_conv_dy(&ma, g.a, *g.conv);
    else if (p.a_kmajor) ok = make_map(&ma, g.a, g.m, g.k, g.lda, BM, false);
    else {
        if (p.a_box3 && !make_map_mn3(&ma, g.a, g.k, g.m, g.lda, BM / 32)) p.a_box3 = false;
        ok = p.a_box3 || make_map(&ma, g.a, g.k, g.m, g.lda, BK, true);
    }
    if (cmode == 1) ok = ok && make_map_conv_in(&mb, *g.conv, p.cv_bw, p.cv_bh);
    else if (cmode == 2) ok = ok && make_map_conv_in_mn(&mb, *g.conv);
    else if (p.b_kmajor) ok = ok && make_map(&mb, g.b, g.n, g.k, g.ldb, p.bn / p.cluster, false);
    else {
        if (p.b_box3 && !make_map_mn3(&mb, g.b, g.k, g.n, g.ldb, p.bn / p.cluster / 32)) p.b_box3 = false;
        ok = ok && (p.b_box3 || make_map(&mb, g.b, g.k, g.n, g.ldb, BK, true));
    }
    if (!ok) return false;

    const int sms = rt().num_sms;
    const long long tiles = (long long)p.m_tiles * p.n_tiles;
    static int tail_on = -1;
    if (tail_on < 0) { const char* e = getenv("BLA_TC_TAIL"); tail_on = e ? atoi(e) : 1; }
    if (cmode == 0 && tail_on && !g.no_tail_split && p.kblocks >= 32) {
        // Wave quantisation: 235 units on 74 CTA pairs are 3.18 waves but cost 4.  When the last wave would be less than half
        // full, the columns of the full waves and the tail are issued as two GEMMs: the tail then has few tiles and takes the
        // split-K path, so its SMs are all busy for a fraction of a wave instead of 18 % of them for a whole one.
        const long long slots = sms / p.cluster, units = tiles / p.cluster, m_units = p.m_tiles / p.cluster;
        const long long full = units / slots, rem = units - full * slots;
        const long long main_tiles = m_units > 0 ? full * slots / m_units : 0;
        if (full >= 1 && rem > 0 && rem * 2 <= slots && main_tiles >= 1 && main_tiles < p.n_tiles) {
            const int n_main = (int)(main_tiles * p.bn);
            if (g.plan_main_columns) { *g.plan_main_columns = n_main; return true; }
            GemmArgs lo = g, hi = g;
            lo.no_tail_split = hi.no_tail_split = true;
            lo.n = n_main;
            hi.n = g.n - n_main;
            hi.b = g.tb ? g.b + (size_t)n_main * g.ldb : g.b + n_main;
            hi.c = g.c + n_main;
            if (g.epi.bias_cols) hi.epi.bias_cols = g.epi.bias_cols + n_main;
            if (g.epi.gate) hi.epi.gate = g.epi.gate + n_main;
            if (g.gate_bits) hi.gate_bits = g.gate_bits + n_main / 32;     // n_main is a multiple of the tile width, itself a multiple of 32
            if (g.mask_out) hi.mask_out = g.mask_out + n_main / 32;
            if (g.epi.pre_activation) hi.epi.pre_activation = g.epi.pre_activation + n_main;
            if (gemm_3xtf32(lo, s)) {
                if (!gemm_3xtf32(hi, s)) {
                    if (hi.mask_out && hi.mask_written) *hi.mask_written = false;
                    gemm_simt(hi, s);
                }
                return true;
            }
        }
    }
    if (g.plan_main_columns) { *g.plan_main_columns = g.n; return true; }
    int splits = 1;
    if (conv_splits > 0) {
        splits = conv_splits;
    } else if (!force_no_split && tiles * 2 <= sms && p.kblocks >= 32) {
        long long want = (sms / p.cluster) / (tiles / p.cluster);          // one wave: units * splits <= cluster slots
        long long maxs = p.kblocks / 16;
        splits = (int)(want < maxs ? want : maxs);
        if (splits < 1) splits = 1;
        if (splits > 64) splits = 64;
    }
    p.kblocks_per_split = ceil_div(p.kblocks, splits);
    splits = ceil_div(p.kblocks, p.kblocks_per_split);
    p.splits = splits;
    float* ws = nullptr;
    const bool own_ws = g.workspace != nullptr;   // the caller's scratch (a GEMM beside the library stream), never the pool then
    if (own_ws && splits > 1 && (size_t)splits * g.m * g.n > g.workspace_floats) return false;
    if (splits > 1) {
        ws = own_ws ? g.workspace : (float*)pool_alloc(kDevice, (size_t)splits * g.m * g.n * sizeof(float));
        p.partial = ws;
        if (cmode == 2 && g.conv->dw_final && g.n % 4 == 0) {
            p.cv_final = 1; p.cv_Creal = g.conv->C_real; p.c = g.conv->dw_final;
            if (g.conv->wrote_final) *g.conv->wrote_final = true;
        }
    }

    CUtensorMap mc;
    memset(&mc, 0, sizeof(mc));
    p.tma_store = p.c_vec && !g.epi.pre_activation && !g.epi.bias_cols &&
                  (splits == 1 || (g.m % BM == 0 && g.n % 4 == 0));
    if (g.mask_out && g.mask_written) {   // which of this product's launches can write the sign mask?
        const bool can = splits == 1 ? p.tma_store : ((((size_t)g.m * g.n) & 3) == 0 && g.n % 32 == 0 && p.c_vec);
        if (!can) { *g.mask_written = false; p.mask_out = nullptr; }
    }
    if (!p.tma_store) p.gate_bits = nullptr;   // the register-store epilogue reads the float gate
    if (cmode == 1 && splits == 1) {
        p.tma_store = make_map_conv_out(&mc, *g.conv);
        if (!p.tma_store) { if (ws && !own_ws) pool_free(ws); return false; }
    } else if (cmode == 1) {   // split-K partials [splits][F][imgs*P] (F a multiple of 128 here); the reduce kernel writes y
        p.tma_store = make_map_c(&mc, ws, (long long)splits * g.m, g.n, g.n);
        if (!p.tma_store) { if (ws && !own_ws) pool_free(ws); return false; }
    } else if (p.tma_store) {
        p.tma_store = splits == 1 ? make_map_c(&mc, g.c, g.m, g.n, g.ldc) : make_map_c(&mc, ws, (long long)splits * g.m, g.n, g.n);
    }

    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_3xtf32_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_3xtf32_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            g_tc_broken = true;
            if (ws && !own_ws) pool_free(ws);
            return false;
        }
        attr_set = true;
    }
    const long long units = tiles / p.cluster * splits;
    const int slots = sms / p.cluster;
    const int grid = (int)(units < slots ? units : slots) * p.cluster;
    if (p.cluster == 1) BLA_CUDA(launch_pdl(gemm_3xtf32_kernel<1>, dim3(grid), dim3(kThreads), kSmemBytes, s, 1, ma, mb, mc, p));
    else BLA_CUDA(launch_pdl(gemm_3xtf32_kernel<2>, dim3(grid), dim3(kThreads), kSmemBytes, s, 2, ma, mb, mc, p));
    count_launch();
    ++g_tc_launches;
    if (ws) {
        size_t totalc = (size_t)g.m * g.n;
        size_t blocks = (totalc / 4 + 63) / 64;   // 64 float4 columns per block on the vector path
        size_t cap = (size_t)sms * 8;
        if (blocks > cap) blocks = cap;
        BLA_CUDA(launch_pdl(tc_splitk_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, s, 1, p));
        count_launch();
        if (!own_ws) pool_free(ws);
    }
    return true;
}

// NCHW -> zero-padded (and optionally zero-dilated, channel-padded) NHWC: xp[img][pt + i*dil][pl + j*dil][c] = x[img][c][i][j], every
// other element of xp (border, dilation holes, channels C..Cp-1) = 0 -- the kernel writes the zeros itself, no memset pass.
// grid (tiles of 32 consecutive padded pixels, images); a block walks the channels in tiles of 32 through shared memory: loads are
// coalesced along the image row, stores are one float4 (4 channels) per thread = full 128-byte lines per pixel.
__global__ void __launch_bounds__(256) nchw_to_padded_nhwc_kernel(const float* __restrict__ x, float* __restrict__ xp, int C, int Cp, int H, int W,
                                                                  int Hp, int Wp, int pt, int pl, int dil) {
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    __shared__ float tile[32][33];   // [channel][pixel]
    const int img = blockIdx.y, q0 = blockIdx.x * 32, npix = Hp * Wp;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int q = q0 + tx;
    const int ip = q / Wp, jp = q - ip * Wp;
    const int di = ip - pt, dj = jp - pl;
    const int i = di / dil, j = dj / dil;
    const bool ok = q < npix && di >= 0 && dj >= 0 && i * dil == di && j * dil == dj && i < H && j < W;
    const float* src = x + (size_t)img * C * H * W + (size_t)i * W + j;
    const int sp = threadIdx.x >> 3, sc = (threadIdx.x & 7) * 4;   // store role: pixel, channel quad
    float* dst = xp + ((size_t)img * npix + q0 + sp) * Cp + sc;
    const bool st_ok = q0 + sp < npix;
    for (int c0 = blockIdx.z * 32; c0 < Cp; c0 += gridDim.z * 32) {   // small images: the channel tiles are spread over blockIdx.z too
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c = c0 + ty + 8 * cc;
            tile[ty + 8 * cc][tx] = (ok && c < C) ? __ldg(src + (size_t)c * H * W) : 0.f;
        }
        __syncthreads();
        if (st_ok && c0 + sc < Cp)
            *reinterpret_cast<float4*>(dst + c0) = make_float4(tile[sc][sp], tile[sc + 1][sp], tile[sc + 2][sp], tile[sc + 3][sp]);
        __syncthreads();
    }
}

void nhwc_cache_release(NhwcCache* c) {
    if (c && c->xp) pool_free(c->xp);
    if (c) *c = NhwcCache();
}

// the padded NHWC copy a tensor-path conv gathers from: from the cache when the same tensor was transformed before, else a fresh
// transform into the cache's buffer (or a temporary pool block when there is no cache -- the caller frees it, stream-ordered)
float* padded_nhwc(const float* in, int imgs, int C, int Cp, int H, int W, int Hp, int Wp, int pad_top, int pad_left, int dil, NhwcCache* cache,
                   cudaStream_t s) {
    const size_t need = (size_t)imgs * Hp * Wp * Cp;
    float* xp;
    if (cache) {
        if (cache->valid && cache->src == in && cache->imgs == imgs && cache->C == C && cache->Cp == Cp && cache->H == H && cache->W == W &&
            cache->Hp == Hp && cache->Wp == Wp && cache->pt == pad_top && cache->pl == pad_left && cache->dil == dil)
            return cache->xp;
        if (cache->cap < need) {
            if (cache->xp) pool_free(cache->xp);
            cache->xp = (float*)pool_alloc(kDevice, need * sizeof(float));
            cache->cap = need;
        }
        cache->src = in; cache->imgs = imgs; cache->C = C; cache->Cp = Cp; cache->H = H; cache->W = W; cache->Hp = Hp; cache->Wp = Wp;
        cache->pt = pad_top; cache->pl = pad_left; cache->dil = dil; cache->valid = true;
        xp = cache->xp;
    } else {
        xp = (float*)pool_alloc(kDevice, need * sizeof(float));
    }
    const int px_tiles = ceil_div(Hp * Wp, 32), c_tiles = ceil_div(Cp, 32);
    int gz = ceil_div(4LL * rt().num_sms, (long long)px_tiles * imgs);     // enough blocks for ~4 per SM
    if (gz > c_tiles) gz = c_tiles;
    if (gz < 1) gz = 1;
    BLA_CUDA(launch_pdl(nchw_to_padded_nhwc_kernel, dim3(px_tiles, imgs, gz), dim3(256), 0, s, 1, in, xp, C, Cp, H, W, Hp, Wp, pad_top, pad_left, dil));
    BLA_LAUNCH_CHECK();
    count_launch();
    return xp;
}

// out [imgs][F][Ho][Wo] = conv(in, w): `in` is [imgs][C][Hin][Win], placed at spacing `dil` (zeros between: the transposed conv of a
// strided dgrad) and offset (pad_top, pad_left) inside a logical image of H x W, then convolved with stride `stride`.
bool conv2d_tc(const float* in, const float* w_taps, float* out, int imgs, int C, int Cp, int Hin, int Win, int dil, int H, int W, int F,
               int k, int stride, int pad_top, int pad_left, NhwcCache* cache, cudaStream_t s, const float* bias, const float* addend) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int P = Ho * Wo;
    // eligibility: whole 16-channel k-blocks, 16-byte aligned rows, and 32-pixel atoms that tile the output exactly
    if (Cp % BK || Cp < C || ((uintptr_t)in & 15) || ((uintptr_t)out & 15) || (P % 4)) return false;
    const bool big = P >= 32 && P % 32 == 0 && Wo <= 32 && 32 % Wo == 0;
    const bool small = P < 32 && 32 % P == 0;
    if (!big && !small) return false;
    // The SAME padding is materialised once in a zero-padded NHWC copy of the input (8 B/elem, against k*k*F uses of every
    // element): every TMA box then starts inside the tensor, on a 16-channel (64-byte) boundary.
    const int Hp = (Ho - 1) * stride + k > H + pad_top ? (Ho - 1) * stride + k : H + pad_top;
    const int Wp = (Wo - 1) * stride + k > W + pad_left ? (Wo - 1) * stride + k : W + pad_left;
    if (pad_top + (Hin - 1) * dil + 1 > Hp || pad_left + (Win - 1) * dil + 1 > Wp) return false;
    float* xp = padded_nhwc(in, imgs, C, Cp, Hin, Win, Hp, Wp, pad_top, pad_left, dil, cache, s);
    ConvTc cv{1, xp, out, imgs, Cp, Hp, Wp, F, k, stride, Ho, Wo, 0, 0};
    cv.bias = bias; cv.addend = addend;
    if (addend && ((uintptr_t)addend & 15)) { if (!cache) pool_free(xp); return false; }
    GemmArgs g{};
    g.m = F; g.n = imgs * P; g.k = k * k * Cp;
    g.a = w_taps; g.lda = g.k;
    g.c = out; g.ldc = P;
    g.conv = &cv;
    const bool done = gemm_3xtf32(g, s);
    if (!cache) pool_free(xp);   // stream-ordered reuse
    return done;
}

// dw_taps [F][(ki, kj, c)] = sum over images and output pixels of dy[img][f][pixel] * x[img][c][pixel*stride + tap - pad]
bool conv2d_wgrad_tc(const float* x, const float* dy, float* dw_taps, float* dw_final, bool* wrote_final, int imgs, int C, int Cp, int H, int W,
                     int F, int k, int stride, int pad_top, int pad_left, NhwcCache* cache, cudaStream_t s) {
    *wrote_final = false;
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const int P = Ho * Wo;
    // eligibility: 32-channel atoms, 16-pixel k-blocks inside one image that are whole box rows
    if (Cp % 32 || Cp < C || P % BK || ((uintptr_t)x & 15) || ((uintptr_t)dy & 15) || ((uintptr_t)dw_taps & 15)) return false;
    if (!(Wo % BK == 0 || (Wo < BK && BK % Wo == 0))) return false;
    const int Hp = (Ho - 1) * stride + k > H + pad_top ? (Ho - 1) * stride + k : H + pad_top;
    const int Wp = (Wo - 1) * stride + k > W + pad_left ? (Wo - 1) * stride + k : W + pad_left;
    float* xp = padded_nhwc(x, imgs, C, Cp, H, W, Hp, Wp, pad_top, pad_left, 1, cache, s);
    ConvTc cv{2, xp, nullptr, imgs, Cp, Hp, Wp, F, k, stride, Ho, Wo, 0, 0};   // the padding is inside xp
    cv.dw_final = dw_final; cv.C_real = C; cv.wrote_final = wrote_final;
    GemmArgs g{};
    g.m = F; g.n = k * k * Cp; g.k = imgs * P;
    g.a = dy; g.lda = P;
    g.c = dw_taps; g.ldc = g.n;
    g.conv = &cv;
    const bool done = gemm_3xtf32(g, s);
    if (!cache) pool_free(xp);
    return done;
}

}  // namespace bla

extern "C" int bla_tc_available(void) {
    bla::rt();
    return bla::encode_fn() != nullptr && !bla::g_tc_broken;
}
extern "C" unsigned long long bla_tc_launch_count(void) { return bla::g_tc_launches; }
// include/bla.h: how many of the n columns of a row-major m x n x k product the FIRST tensor-kernel launch computes (the rest, if
// any, is the split-K tail launch; n when the product is not split or not eligible)
extern "C" int bla_tc_main_columns(int m, int n, int k) {
    using namespace bla;
    if (!bla_tc_available() || m <= 0 || n <= 0 || k <= 0) return n;
    int out = n;
    GemmArgs g{};
    g.m = m; g.n = n; g.k = k;
    g.a = reinterpret_cast<const float*>(uintptr_t(256)); g.lda = k;      // aligned dummies: the query never dereferences them
    g.b = reinterpret_cast<const float*>(uintptr_t(256)); g.ldb = n;
    g.c = reinterpret_cast<float*>(uintptr_t(256)); g.ldc = n;
    g.plan_main_columns = &out;
    if (!gemm_3xtf32(g, rt().stream)) return n;
    return out;
}
