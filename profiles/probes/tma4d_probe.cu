// Probe: which 4-D tiled TMA configurations are legal on sm_100a (box over x[img][C][H][W] viewed as {W,H,img,C}).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma4d_probe tma4d_probe.cu && ./tma4d_probe <variant>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int c0, int c1, int c2, int c3, int n) {
    extern __shared__ __align__(1024) float buf[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), dst = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(n * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    __syncthreads();
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(bar_a) : "memory");
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = variant >= 5 ? 36 : 32, H = variant >= 5 ? 34 : 32, C = 128, IM = 3;
    std::vector<float> h((size_t)IM * C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 512 * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    CUtensorMap map;
    cuuint64_t dims[4], strides[3]; cuuint32_t box[4], elem[4] = {1, 1, 1, 1};
    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    int c[4] = {0, 0, 0, 0};
    // variants: 0 = {W,H,img,C} swizzled, coord (-1,-1,1,16); 1 = same, no negative coords; 2 = {W,H,C,img} order swizzled;
    //           3 = {W,H,img,C} no swizzle; 4 = {W,H,C,img} no swizzle
    bool img_before_c = (variant == 0 || variant == 1 || variant == 3 || variant >= 5);
    if (img_before_c) {
        dims[0] = W; dims[1] = H; dims[2] = IM; dims[3] = C;
        strides[0] = W * 4; strides[1] = (cuuint64_t)C * H * W * 4; strides[2] = (cuuint64_t)H * W * 4;
        box[0] = 32; box[1] = 1; box[2] = 1; box[3] = 16;
        c[0] = variant == 1 ? 0 : -1; c[1] = variant == 1 ? 2 : -1; c[2] = 1; c[3] = 16;
    } else {
        dims[0] = W; dims[1] = H; dims[2] = C; dims[3] = IM;
        strides[0] = W * 4; strides[1] = (cuuint64_t)H * W * 4; strides[2] = (cuuint64_t)C * H * W * 4;
        box[0] = 32; box[1] = 1; box[2] = 16; box[3] = 1;
        c[0] = -1; c[1] = 2; c[2] = 16; c[3] = 1;
    }
    if (variant == 3 || variant == 4) sw = CU_TENSOR_MAP_SWIZZLE_NONE;
    // 5: padded tensor (36 x 34), in-range coords; 6: box hangs over the high edge in w; 7: image index out of range; 8: dst offset 2048
    if (variant >= 5) { c[0] = 2; c[1] = 33; c[2] = 2; c[3] = 112; }
    if (variant == 6) c[0] = 8;
    if (variant == 7) c[2] = 3;
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d encode result %d\n", variant, (int)r);
    if (r != CUDA_SUCCESS) return 0;
    probe<<<1, 128, 4096>>>(map, o, c[0], c[1], c[2], c[3], 512);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d run: %s\n", variant, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        float ho[512]; cudaMemcpy(ho, o, sizeof(ho), cudaMemcpyDeviceToHost);
        printf("first row: %g %g %g ... row1: %g %g\n", ho[0], ho[1], ho[2], ho[32], ho[33]);
    }
    return 0;
}
