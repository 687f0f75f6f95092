// Pinned host -> device: one contiguous copy of a [784 x 60000] float matrix against the same bytes as strided column-chunk
// copies (cudaMemcpy2DAsync), by chunk width.  nvcc -O2 -o h2d_probe h2d_probe.cu ; prints ms per full matrix.
#include <cuda_runtime.h>
#include <cstdio>
#include <algorithm>
int main() {
    const int R = 784, B = 60000;
    float *h, *d;
    cudaMallocHost(&h, (size_t)R * B * 4);
    cudaMalloc(&d, (size_t)R * B * 4);
    for (size_t i = 0; i < (size_t)R * B; i += 1024) h[i] = 1.f;
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto fn) {
        fn(); cudaStreamSynchronize(s);
        cudaEventRecord(e0, s);
        for (int i = 0; i < 5; ++i) fn();
        cudaEventRecord(e1, s); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
    };
    printf("{\"contiguous\": %.4f", time([&] { cudaMemcpyAsync(d, h, (size_t)R * B * 4, cudaMemcpyHostToDevice, s); }));
    for (int cols : {1536, 3072, 6144, 8192, 12288, 15360, 30016}) {
        float ms = time([&] {
            for (int b0 = 0; b0 < B; b0 += cols) {
                int bc = std::min(cols, B - b0);
                cudaMemcpy2DAsync(d + (size_t)R * b0, bc * 4, h + b0, B * 4, bc * 4, R, cudaMemcpyHostToDevice, s);
            }
        });
        printf(", \"cols_%d\": %.4f", cols, ms);
    }
    // row-block chunks (contiguous pieces of the row-major matrix), 10 of them
    printf(", \"row_blocks_10\": %.4f", time([&] {
        for (int r0 = 0; r0 < R; r0 += 79) {
            int rc = std::min(79, R - r0);
            cudaMemcpyAsync(d + (size_t)r0 * B, h + (size_t)r0 * B, (size_t)rc * B * 4, cudaMemcpyHostToDevice, s);
        }
    }));
    printf("}\n");
    return 0;
}
