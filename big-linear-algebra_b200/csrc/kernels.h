// kernels.h -- launchers of the sm_100a kernels.  Every pointer is device-usable (device or
// managed memory); all launches go to the given stream and return immediately.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "../../include/bla.h"

namespace bla {

// ---- elementwise (ew.cu): 128-bit vectorised grid-stride kernels ---------------------------
void k_scale(float* m, size_t n, float f, cudaStream_t s);                       // lib/matrix.c:59
void k_add(float* a, const float* b, size_t n, cudaStream_t s);                  // lib/matrix.c:65
void k_hadamard(float* a, const float* b, size_t n, cudaStream_t s);             // lib/matrix.c:95
void k_copy(float* dst, const float* src, size_t n, cudaStream_t s);             // lib/matrix.c:14
void k_axpy(float* y, const float* x, float alpha, size_t n, cudaStream_t s);    // y += alpha*x (scale+add, mnist_nn.c:303-315)
void k_relu(float* d, size_t n, cudaStream_t s);                                 // lib/util.c:7
void k_relu_ddx(float* d, size_t n, cudaStream_t s);                             // mnist_nn.c:47
void k_relu_backward(const float* src, const float* relu_result, float* dst, size_t n, cudaStream_t s);  // cifar_unet.c:241
void k_add_tile_columns(float* a, int rows, int cols, const float* b, int bcols, cudaStream_t s);        // lib/matrix.c:189
void k_add_tile_rows(float* a, int rows, int cols, const float* b, cudaStream_t s);                      // lib/matrix.c:199
void k_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t s);                      // lib/matrix.c:105
void k_fill_uniform(float* dst, size_t n, unsigned long long seed, float lo, float hi, cudaStream_t s);
void k_u8_to_float(float* dst, const unsigned char* src, size_t n, float scale, cudaStream_t s);
// host_pack.cpp: dst[i] = (unsigned char)src[i]; true when every src[i] is bit for bit one of 0.0f .. 255.0f
bool pack_row_u8(const float* src, unsigned char* dst, int n);

// ---- reductions (reduce.cu): warp-shuffle trees, two deterministic stages ---------------------
// `work` must hold reduce_workspace_bytes() bytes of device scratch.
size_t reduce_workspace_bytes();
void k_row_sum(const float* m, int rows, int cols, float* out, cudaStream_t s);                          // lib/matrix.c:123
void k_col_sum(const float* m, int rows, int cols, float* out, int quirk, void* work, cudaStream_t s);   // lib/matrix.c:138
void k_sum_squares(const float* m, size_t n, double* out1, void* work, cudaStream_t s);                  // lib/matrix.c:150
void k_max(const float* m, size_t n, float* out1, void* work, cudaStream_t s);                           // lib/matrix.c:160
void k_zscore(float* m, size_t n, void* work, cudaStream_t s);                                           // lib/matrix.c:170
void k_softmax_cols(float* d, int rows, int cols, cudaStream_t s);                                       // lib/util.c:15
void k_softmax_rows(float* d, int rows, int cols, cudaStream_t s);                                       // lib/util.c:36
void k_softmax_xent(const float* logits, const float* expected, int classes, int batch, float* probs, float* grad,
                    float grad_scale, double* stats, cudaStream_t s);                                    // mnist_nn.c:234-268

// ---- GEMM (gemm_simt.cu / gemm_tc.cu) ---------------------------------------------------------
// Implicit-GEMM convolution on the tensor path: B is not a matrix but the NCHW input, gathered by TMA.
struct ConvTc {
    int mode;             // 1 forward (B gathered K-major), 2 weight gradient (A = dy, B gathered MN-major)
    const float* in;      // padded NHWC input [imgs][H][W][C]
    float* out;           // [imgs][F][Ho][Wo]
    int imgs, C, H, W, F, k, stride, Ho, Wo, pad_top, pad_left;
    // weight gradient only: when the GEMM is split along K, its reduction writes dW straight in the reference layout [F][C_real][k][k]
    float* dw_final = nullptr; int C_real = 0; bool* wrote_final = nullptr;
    // forward only: fused y += bias[img][filter] and y += addend[img][filter][pixel]
    const float* bias = nullptr; const float* addend = nullptr;
};

struct GemmArgs {
    bool ta, tb;          // A stored [k x m] / B stored [n x k]
    int m, n, k;
    const float* a; int lda;
    const float* b; int ldb;
    float* c; int ldc;
    bla_epilogue epi;     // zero-initialised = plain store
    const ConvTc* conv;   // nullptr for a plain GEMM; else m = F, n = imgs*Ho*Wo, k = k*k*C, a = weights [F][(ki,kj,c)]
    // ReLU masks in bit form (the MLP's relu' gates): 32 columns per word, row i at mask[i * ld .. ), bit j % 32 of word j / 32.
    //   mask_out : bit = (C[i][j] > 0), written next to C by the launches that can (tensor path with TMA stores, its vectorised
    //              split-K reduction); any launch that cannot clears *mask_written, and the caller falls back to the float gate
    //   gate_bits: the same gate as epi.gate, read as one word per row and 32 columns where the kernel supports it (epi.gate must
    //              still be set: it is what every other path reads)
    uint32_t* mask_out; int mask_ld; bool* mask_written;
    const uint32_t* gate_bits; int gate_ld;
    // Split-K partial sums normally come from the stream-ordered pool, which serves ONE stream: a GEMM issued on a second stream that
    // runs beside the library stream brings its own scratch; the pool is then never touched (the tensor path declines a split that
    // does not fit and the FP32 kernel takes fewer slices).
    float* workspace; size_t workspace_floats;
    bool no_tail_split;   // internal: this call already is one half of a main / tail column split (gemm_tc.cu)
    int* plan_main_columns;   // query only: receives the column count of the first launch (n when there is no split); nothing runs
};
// Dispatch on rt().gemm_path and the shape.
void gemm(const GemmArgs& g, cudaStream_t s);
void gemm_simt(const GemmArgs& g, cudaStream_t s);
// returns false when the shape/alignment is not eligible for the tensor path
bool gemm_3xtf32(const GemmArgs& g, cudaStream_t s);
// The padded NHWC copy of a conv input that the tensor path gathers from.  A caller that convolves the SAME tensor twice (forward,
// then the weight gradient in the backward pass) hands the same cache to both calls: the second one skips the transform.  The
// owner must clear `valid` whenever the source tensor changes (the U-Net does at the start of every step).
struct NhwcCache {
    float* xp = nullptr;      // pool block, kept across calls (release with nhwc_cache_release)
    size_t cap = 0;           // floats
    bool valid = false;
    const float* src = nullptr;
    int imgs = 0, C = 0, Cp = 0, H = 0, W = 0, Hp = 0, Wp = 0, pt = 0, pl = 0, dil = 0;
};
void nhwc_cache_release(NhwcCache* c);

// conv2d forward on the tensor path: out = conv(in, w) with w already arranged as [F][(ki, kj, c)] over Cp >= C channels (zero
// weights for the padding channels; Cp a multiple of 16); false if ineligible.
// `in` [imgs][C][Hin][Win] is placed at spacing `dil` inside a logical H x W image (dil = 1, Hin = H for an ordinary conv).
bool conv2d_tc(const float* in, const float* w_taps, float* out, int imgs, int C, int Cp, int Hin, int Win, int dil, int H, int W, int F,
               int k, int stride, int pad_top, int pad_left, NhwcCache* cache, cudaStream_t s, const float* bias = nullptr,
               const float* addend = nullptr);
// weight gradient on the tensor path: into dw_final [F][C][k][k] when the split-K reduction could write it (*wrote_final), else into
// dw_taps [F][(ki, kj, c)] over Cp channels (a multiple of 32) for the caller to un-permute; false if ineligible
bool conv2d_wgrad_tc(const float* x, const float* dy, float* dw_taps, float* dw_final, bool* wrote_final, int imgs, int C, int Cp, int H, int W,
                     int F, int k, int stride, int pad_top, int pad_left, NhwcCache* cache, cudaStream_t s);

// ---- batched device-resident conv2d / group norm (conv_implicit.cu, api_norm.cu) ----------------
// Filters re-laid for the tensor path (taps: forward, flip: dgrad).  A caller whose weights change once per step builds a device
// table of jobs, runs conv_permute_weights_batch once after the update and hands the results to conv2d_forward / conv2d_dgrad.
struct ConvPermuteJob { const float* src; float* dst; int F, C, k2, pad, mode; };
size_t conv_taps_elems(int F, int C, int k);
size_t conv_flip_elems(int F, int C, int k);
ConvPermuteJob conv_taps_job(const float* w, float* dst, int F, int C, int k);
ConvPermuteJob conv_flip_job(const float* w, float* dst, int F, int C, int k);
void conv_permute_weights_batch(const ConvPermuteJob* jobs_device, int njobs, cudaStream_t s);
bool conv_tensor_path_wanted();
// bias [imgs][F] and addend [imgs][F][Ho][Wo] (either may be NULL) are added to the result: in the tensor kernel's epilogue, by
// two elementwise launches after the FP32 kernel.
void conv2d_forward(const float* x, const float* w, float* y, int imgs, int C, int H, int W, int F, int k, int stride, cudaStream_t s,
                    NhwcCache* cache = nullptr, const float* w_taps = nullptr, const float* bias = nullptr, const float* addend = nullptr);
void conv2d_wgrad(const float* x, const float* dy, float* dw, int imgs, int C, int H, int W, int F, int k, int stride, cudaStream_t s,
                  NhwcCache* cache = nullptr);
void conv2d_dgrad(const float* dy, const float* w, float* dx, int imgs, int C, int H, int W, int F, int k, int stride, cudaStream_t s,
                  const float* w_flip = nullptr);
// Neighbours of a group norm fused into its kernels (the U-Net's group_norm -> multi_channel_relu -> _dropout chain,
// cifar_unet.c:1046-1061): forward writes dropout(relu(norm(x))); backward gates the incoming gradient with the same masks
// (ReLU: x > mean; dropout: element i of the tensor is dropped iff uniform_at(drop_seed, i) < drop_rate).
struct GnFuse { int relu; float drop_rate; unsigned long long drop_seed; const float* addend = nullptr; };   // addend: backward only, dx += addend
void k_group_norm_fwd(const float* x, float* y, float* vars, float* means, int images, int C, int HW, int group_size, int quirk,
                      cudaStream_t s, const GnFuse* fuse = nullptr);
void k_group_norm_bwd(const float* dy, float* dx, const float* x, const float* means, const float* stdevs, int images, int C, int HW,
                      int group_size, cudaStream_t s, const GnFuse* fuse = nullptr);

// counter-based generator (splitmix64 of seed + index), identical on host and device
__host__ __device__ inline float uniform_at(unsigned long long seed, unsigned long long i, float lo, float hi) {
    unsigned long long z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    float u = (float)(z >> 40) * (1.0f / 16777216.0f);  // 24 bits -> [0,1)
    return fmaf(hi - lo, u, lo);  // explicit fma: same bits from nvcc and from the host compiler
}

}  // namespace bla
