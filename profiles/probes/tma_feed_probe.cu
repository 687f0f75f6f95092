// How fast can ONE SM's shared memory be filled by TMA, as a function of the box shape, the ring depth and where the data comes from?
// A persistent CTA per SM: lane 0 of warp 0 issues the boxes of a "stage" (mbarrier expect_tx), lane 0 of warp 1 waits for the stage and
// hands it straight back -- no consumer work at all, so the number is the feed ceiling of a TMA -> mbarrier ring.
//   nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_feed_probe tma_feed_probe.cu -lcuda ; ./tma_feed_probe
// Prints bytes per clock and SM (at the 1.965 GHz the B200 runs when it is not power capped; the printed GB/s do not depend on that).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW1:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D1;\n\tbra W1;\n\tD1:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y) : "memory");
}

struct Params {
    int stages, boxes_per_stage, box_bytes, iters;   // iters = stages (k-blocks) per CTA
    int box_cols, box_rows;                          // elements (fp32)
    int tensor_cols, tensor_rows;                    // the 2-D tensor the boxes walk over
    int walk;                                        // 0: boxes advance along the inner dimension (K-major operand), 1: along rows
};

__global__ void __launch_bounds__(64, 1) feed_kernel(const __grid_constant__ CUtensorMap map, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)p.boxes_per_stage * p.box_bytes;
    const uint32_t bars = base + (uint32_t)p.stages * stage_bytes;
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(bars + 16u * s, 1); mbar_init(bars + 16u * s + 8u, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int stage = 0; uint32_t phase = 0;
        // every CTA walks its own region of the tensor; plain 32-bit counters (a 64-bit division per box would be the bottleneck)
        const int nx = p.tensor_cols / p.box_cols, ny = p.tensor_rows / p.box_rows;
        int ix = (int)((blockIdx.x * 37u) % (unsigned)nx), iy = (int)((blockIdx.x * 11u) % (unsigned)ny);
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(bars + 16u * stage + 8u, phase ^ 1);
            mbar_expect(bars + 16u * stage, stage_bytes);
            for (int bx = 0; bx < p.boxes_per_stage; ++bx) {
                tma_load_2d(base + stage * stage_bytes + bx * p.box_bytes, &map, ix * p.box_cols, iy * p.box_rows, bars + 16u * stage);
                if (p.walk == 0) { if (++ix == nx) { ix = 0; if (++iy == ny) iy = 0; } }   // along the inner dimension (K-major operand)
                else { if (++iy == ny) { iy = 0; if (++ix == nx) ix = 0; } }                // down the rows (MN-major operand)
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(bars + 16u * stage, phase);
            mbar_arrive(bars + 16u * stage + 8u);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EncodeFn encode = (EncodeFn)sym;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaFuncSetAttribute(feed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // two sources: 1 GiB (streams from DRAM) and 32 MiB (stays in L2)
    float *big, *small;
    const size_t big_rows = 32768, big_cols = 8192, small_rows = 1024, small_cols = 8192;
    CK(cudaMalloc(&big, big_rows * big_cols * 4)); CK(cudaMemset(big, 0, big_rows * big_cols * 4));
    CK(cudaMalloc(&small, small_rows * small_cols * 4)); CK(cudaMemset(small, 0, small_rows * small_cols * 4));
    struct Case { const char* name; int cols, rows; CUtensorMapSwizzle sw; int boxes; int walk; };
    const Case cases[] = {
        {"64B x 128 rows  SW64  (K-major A, BK=16)", 16, 128, CU_TENSOR_MAP_SWIZZLE_64B, 2, 0},
        {"128B x 128 rows SW128 (K-major A, BK=32)", 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, 1, 0},
        {"128B x 256 rows SW128 (K-major, BK=32, 32 KB box)", 32, 256, CU_TENSOR_MAP_SWIZZLE_128B, 1, 0},
        {"128B x 16 rows  SW128_32B (MN-major atom, BK=16) x8", 32, 16, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 8, 1},
        {"128B x 32 rows  SW128_32B (MN-major atom, BK=32) x4", 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 4, 1},
        {"128B x 64 rows  SW128 x2", 32, 64, CU_TENSOR_MAP_SWIZZLE_128B, 2, 1},
        {"256B x 64 rows  no swizzle x1", 64, 64, CU_TENSOR_MAP_SWIZZLE_NONE, 1, 0},
    };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    printf("TMA feed ceiling per SM (no consumer work), %d SMs, every stage = 16 KB unless noted\n", sms);
    printf("%-56s %-6s %7s %10s %12s %10s\n", "box", "source", "stages", "GB/s chip", "GB/s per SM", "B/clk/SM");
    for (const Case& c : cases) {
        for (int src = 0; src < 2; ++src) {
            const size_t rows = src ? small_rows : big_rows, cols = src ? small_cols : big_cols;
            CUtensorMap map;
            cuuint64_t dims[2] = {cols, rows};
            cuuint64_t strides[1] = {cols * 4};
            cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows}, elem[2] = {1, 1};
            if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src ? small : big, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                printf("%-56s encode failed\n", c.name);
                continue;
            }
            for (int stages : {2, 4, 6, 8, 11}) {
                Params p{};
                p.stages = stages; p.boxes_per_stage = c.boxes; p.box_bytes = c.cols * c.rows * 4; p.iters = 4000;
                p.box_cols = c.cols; p.box_rows = c.rows; p.tensor_cols = (int)cols; p.tensor_rows = (int)rows; p.walk = c.walk;
                const size_t smem = (size_t)stages * p.boxes_per_stage * p.box_bytes + 1024 + 512;
                if (smem > 200 * 1024) continue;
                feed_kernel<<<sms, 64, smem>>>(map, p);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                feed_kernel<<<sms, 64, smem>>>(map, p);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double bytes = (double)sms * p.iters * p.boxes_per_stage * p.box_bytes;
                const double gbs = bytes / (ms * 1e-3) / 1e9;
                printf("%-56s %-6s %7d %10.0f %12.1f %10.1f\n", c.name, src ? "L2" : "DRAM", stages, gbs, gbs / sms, gbs / sms / 1.965);
            }
        }
    }
    return 0;
}
