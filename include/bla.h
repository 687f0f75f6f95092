/*
 * bla.h -- additive C-ABI of libbla.so (big-linear-algebra on B200).
 *
 * The drop-in boundary is the reference's own link line: include/lib/{matrix,layer,conv,norm,util}.h
 * declare exactly the symbols that the reference's lib objects export and its model programs import
 * (SURVEY.md section 8b).  This header adds what a *device-resident* caller needs on top of that,
 * with plain pointers and sizes only (no C++ or torch types):
 *   - device / pinned allocation and explicit transfers, so that a `struct Matrix` can live in HBM
 *     and go through the unchanged matrix.h API without any staging (asynchronously);
 *   - device versions of the loops that the reference keeps *inside its model programs*
 *     (relu', softmax + cross-entropy + argmax, ...: model/mnist_nn.c:38-91,237-257);
 *   - a fused MNIST-MLP trainer, the data-parallel workload of the benchmark
 *     (model/mnist_nn.c:164-394), with its NCCL gradient all-reduce;
 *   - GEMM with transposed operands and fused epilogues (what the reference spells as
 *     matrix_transpose + matrix_multiply + matrix_add_tile_* + activation).
 *
 * Error convention = the reference's (lib/matrix.c:36-39): one line on stdout, then exit(1).
 * There is no CPU fallback anywhere in this library.
 */
#ifndef BLA_H
#define BLA_H

#include <stddef.h>
#include "lib/matrix.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- runtime ------------------------------------------------------------------------------ */

/* Number of usable CUDA devices; 0 when there is none.  Never exits, never initialises a context. */
int bla_device_count(void);
/* Bind the library to a device (default: $LOCAL_RANK if set, else 0).  Must precede other calls. */
void bla_init(int device);
/* Block until everything queued on the library stream has finished. */
void bla_sync(void);
/* The library's cudaStream_t, as an opaque pointer (for CUDA events / graph capture by a host). */
void* bla_stream(void);
/* Adopt a caller's cudaStream_t as the library stream (NULL restores the private stream). */
void bla_set_stream(void* cuda_stream);
const char* bla_version(void);

/* GEMM arithmetic path for every GEMM-shaped call (matrix_multiply*, conv, dense, MLP). */
enum { BLA_GEMM_FP32 = 0,   /* FP32 FMA on the SIMT pipe: <= 1e-5 relative vs the reference */
       BLA_GEMM_3XTF32 = 1, /* three TF32 tcgen05.mma per product, FP32 accumulate in TMEM: <= 1e-3 */
       BLA_GEMM_AUTO = 2 }; /* 3xTF32 for tensor-core sized problems, FP32 otherwise (default; $BLA_PATH) */
void bla_set_gemm_path(int path);
int bla_get_gemm_path(void);
/* 1 when the tcgen05/TMA path is usable on this device+driver; launches of the tensor kernel so far.
 * A GEMM whose operands are not 16-byte aligned (base or row pitch) runs on the FP32 path instead. */
int bla_tc_available(void);
unsigned long long bla_tc_launch_count(void);
/* Columns of a row-major m x n x k product that the FIRST tensor-kernel launch computes: against wave quantisation a product whose
 * last wave would be under half full is issued as a main launch + a split-K tail launch (n when it is not split). */
int bla_tc_main_columns(int m, int n, int k);
/* 1 (default, $BLA_QUIRKS): reproduce reference defects D2 (matrix_col_sum stride) and D5 (group norm
 * divides by the variance); 0: the mathematically intended results. */
void bla_set_quirks(int on);
int bla_get_quirks(void);
/* Counters since load: kernels launched by this library, bytes staged host->device and back. */
unsigned long long bla_launch_count(void);
unsigned long long bla_h2d_bytes(void);
unsigned long long bla_d2h_bytes(void);

/* ---- memory ------------------------------------------------------------------------------- */

void* bla_malloc_device(size_t bytes);
void* bla_malloc_pinned(size_t bytes); /* page-locked host memory for fast transfers */
void* bla_malloc_managed(size_t bytes); /* what the matrix.h API returns to host callers */
void bla_free(void* p);                /* any of the three above */
/* 0 = ordinary host, 1 = managed, 2 = device, 3 = pinned host */
int bla_memory_kind(const void* p);
/* A rows x cols matrix whose data lives in HBM (contents undefined).  Release with free_matrix(). */
struct Matrix* bla_matrix_device(int rows, int cols);
/* Asynchronous copies on the library stream; pair with bla_sync() before touching host memory. */
void bla_copy_h2d(void* dst_device, const void* src_host, size_t bytes);
void bla_copy_d2h(void* dst_host, const void* src_device, size_t bytes);
void bla_copy_d2d(void* dst_device, const void* src_device, size_t bytes);
void bla_memset_zero(void* dst_device, size_t bytes);
/* Deterministic counter-based U[lo,hi) fill (same sequence on any grid; bla_host_uniform is the
 * host twin used by tests and by the CPU baseline). */
void bla_fill_uniform(float* dst_device, size_t n, unsigned long long seed, float lo, float hi);
void bla_host_uniform(float* dst_host, size_t n, unsigned long long seed, float lo, float hi);
/* bytes -> float pixels k (0..255) for MNIST-shaped inputs: dst[i] = (float)src[i] * scale */
void bla_u8_to_float(float* dst_device, const unsigned char* src_device, size_t n, float scale);

/* ---- GEMM --------------------------------------------------------------------------------- */

enum { BLA_ACT_IDENTITY = 0, BLA_ACT_RELU = 1, BLA_ACT_SCALE = 2 };

/* Fused epilogue of bla_gemm_ex.  All pointers may be NULL. */
typedef struct bla_epilogue {
	const float* bias_rows; /* + bias_rows[i]  (matrix_add_tile_columns with a column vector, lib/matrix.c:189) */
	const float* bias_cols; /* + bias_cols[j]  (matrix_add_tile_rows, lib/matrix.c:199) */
	float* pre_activation;  /* optional copy of the value before the activation (ldc layout) */
	const float* gate;      /* result *= (gate[i][j] > 0)   (relu' (.) g, model/mnist_nn.c:276-278) */
	int activation;         /* BLA_ACT_IDENTITY or BLA_ACT_RELU */
	float alpha;            /* result = alpha * (a.b) before bias (0 is treated as 1) */
} bla_epilogue;

/* C[m x n] = op(A)[m x k] . op(B)[k x n], row-major with leading dimensions.
 * trans_a: A is stored [k x m];  trans_b: B is stored [n x k].  Replaces the reference's
 * matrix_transpose/matrix_multiply/matrix_transpose triples (model/mnist_nn.c:266-292). */
void bla_gemm(int trans_a, int trans_b, int m, int n, int k, const float* a, int lda, const float* b, int ldb,
              float* c, int ldc);
void bla_gemm_ex(int trans_a, int trans_b, int m, int n, int k, const float* a, int lda, const float* b, int ldb,
                 float* c, int ldc, const bla_epilogue* epi);

/* ---- device twins of the model-local loops ------------------------------------------------ */

/* model/mnist_nn.c:47-51   d[i] = d[i] > 0 */
void bla_relu_ddx(matrix_float_t* data, int num);
/* model/cifar_unet.c:241-253   dest = relu_result > 0 ? source : 0 */
void bla_relu_backward(const matrix_float_t* source, const matrix_float_t* relu_result, matrix_float_t* dest, size_t n);
/* model/mnist_nn.c:234-268: column softmax of logits[classes x batch], then per column the argmax
 * hit count against one-hot `expected`, the reference's flat-slice cross-entropy (:249-252) summed
 * into loss_sum (double), and grad = (probs - expected) * grad_scale.  probs/grad may alias
 * logits or be NULL.  stats_device = {double loss_sum; double num_correct}, accumulated. */
void bla_softmax_xent(const float* logits, const float* expected, int classes, int batch, float* probs, float* grad,
                      float grad_scale, double* stats_device);

/* ---- batched device-resident group norm (lib/norm.c on [images][C][H*W] tensors) ------------ */

void bla_group_norm(const float* x, float* y, float* vars, float* means, int images, int channels, int hw, int group_size);
void bla_group_norm_ddx(const float* dy, float* dx, const float* x, const float* means, const float* vars, int images,
                        int channels, int hw, int group_size);

/* ---- batched device-resident conv2d as implicit GEMM (lib/conv.c on NCHW tensors) ------------ */

/* x [imgs][C][H][W], w [F][C][k][k], y / dy [imgs][F][Ho][Wo] with Ho = ceil(H/stride); SAME padding
 * as lib/conv.c:12-24.  The im2col matrix is never materialised.  conv (lib/conv.c:205-212): */
void bla_conv2d_forward(const float* x, const float* w, float* y, int imgs, int channels, int height, int width,
                        int filters, int kernel_size, int stride);
/* conv_ddx (lib/conv.c:214-229), the two halves: dw [F][C][k][k] summed over the images ... */
void bla_conv2d_wgrad(const float* x, const float* dy, float* dw, int imgs, int channels, int height, int width,
                      int filters, int kernel_size, int stride);
/* ... and dx [imgs][C][H][W], the exact adjoint for any stride (the reference's is stride-1 only). */
void bla_conv2d_dgrad(const float* dy, const float* w, float* dx, int imgs, int channels, int height, int width,
                      int filters, int kernel_size, int stride);

/* ---- MNIST MLP trainer: model/mnist_nn.c:164-394 as one device-resident step ---------------- */

typedef struct bla_mlp bla_mlp;
/* dims = {inputs, hidden1, hidden2, classes} (784, 256, 128, 10 in the reference, :25-28);
 * max_batch = the largest LOCAL batch (columns per GPU) a step will be given. */
bla_mlp* bla_mlp_create(const int dims[4], int max_batch);
/* the four layer widths the network was created with */
void bla_mlp_dims(const bla_mlp* net, int dims[4]);
void bla_mlp_destroy(bla_mlp* net);
/* Parameters in the reference's layout: W_l [n_l x n_{l-1}] row-major, b_l [n_l] (:165-170).
 * Pointers may be host or device memory. */
void bla_mlp_set_params(bla_mlp* net, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                        const float* b3);
void bla_mlp_get_params(bla_mlp* net, float* w1, float* b1, float* w2, float* b2, float* w3, float* b3);
/* He-uniform weights / zero biases as model/mnist_nn.c:97-144 (`init`), from the counter-based generator. */
void bla_mlp_init_params(bla_mlp* net, unsigned long long seed);
/* The reference's six checkpoint files weights_{1,2,3}.csv / biases_{1,2,3}.csv under `dir` (model/mnist_nn.c:30-35, :344-376). */
void bla_mlp_save_csv(bla_mlp* net, const char* dir);
void bla_mlp_load_csv(bla_mlp* net, const char* dir);
/* One SGD step on x [inputs x batch] RAW pixel values (the 1/255.0F scaling of :218 is fused) and
 * one-hot y [classes x batch]; x/y may be device memory (asynchronous) or host memory (staged).
 * global_batch/col_offset describe the shard (global_batch == batch, col_offset == 0 on one GPU).
 * lr_mult = SGD_LEARN_RATE_MULTIPLIER (0.02).  stats_host, if not NULL, receives {loss_sum, num_correct}
 * of the GLOBAL batch and forces a synchronise; pass NULL to keep the step asynchronous. */
void bla_mlp_train_step(bla_mlp* net, const float* x, const float* y, int batch, int global_batch, int col_offset,
                        float lr_mult, double* stats_host);
/* Same step from byte pixels (MNIST's native storage): x_u8 [inputs x batch]. */
void bla_mlp_train_step_u8(bla_mlp* net, const unsigned char* x_u8, const float* y, int batch, int global_batch,
                           int col_offset, float lr_mult, double* stats_host);
/* Host batches (either step above) are taken in column chunks so that chunk i+1 crosses PCIe while chunk i is trained; the
 * chunks' gradients are summed before the one update, i.e. the same step with a different summation order.  chunk_cols < 0
 * (default, or env BLA_MLP_CHUNK_COLS): automatic -- pinned host batches of >= 16384 columns in ~6144-column chunks (byte
 * batches of >= 32768 columns in two halves: the strided copies need long rows); 0: off; > 0: this chunk width
 * (rounded up to 64 columns, at most 16 chunks) for any host batch. */
void bla_mlp_set_host_chunking(bla_mlp* net, int chunk_cols);
/* Float host batches whose values are the whole numbers 0..255 (pixels as the reference's CSV loader delivers them, mnist_nn.c:
 * 204-209) cross PCIe as bytes: packed on the host, every value checked bit for bit, chunk by chunk behind the chunk before; any
 * other chunk crosses as floats.  Same chunks as bla_mlp_set_host_chunking, same result bit for bit.  mode < 0: automatic (batches of
 * at least 16 MB in a process without a communicator: the ranks of one box share the host's memory bandwidth; default), 0: never, > 0: always.  BLA_MLP_PACK sets the default, BLA_HOST_THREADS the size of the packing pool. */
void bla_mlp_set_host_packing(bla_mlp* net, int mode);
/* The host half of that path on its own (no device needed): x [rows x cols] floats -> bytes, column chunk k (chunk_cols wide, at most
 * 16 chunks) as one contiguous [rows x width_k] block at out + rows * k * chunk_cols; exact[k] (may be NULL) = 1 when every value of
 * chunk k was a whole number 0..255 bit for bit (only then are its bytes meaningful).  Returns 1 when all chunks were exact. */
int bla_pack_pixels(const float* x, int rows, int cols, int chunk_cols, unsigned char* out, int* exact);
/* Forward only (model/mnist_nn.c:446-463, `run`): probs [classes x batch] out. */
void bla_mlp_forward(bla_mlp* net, const float* x, int batch, float* probs);
/* {loss_sum, num_correct} accumulated on the device since the last call (then cleared).  With an active communicator this is a
 * collective (the totals of all ranks are summed): every rank must call it -- or pass stats_host to the step -- at the same point. */
void bla_mlp_read_stats(bla_mlp* net, double* stats_host);

/* ---- data pipeline: lib/mnist_csv2.c and lib/cifar10.c with the dataset resident in HBM ------------------- */

typedef struct bla_mnist bla_mnist;
/* mnist_csv_init (lib/mnist_csv2.c:13-34): rows of `label, 784 pixels,`; parsed by the CSV codec, uploaded once. */
bla_mnist* bla_mnist_from_csv(const char* filepath);
/* The same store from memory: x [examples][features] (sample-major), labels [examples]; host or device pointers. */
bla_mnist* bla_mnist_from_arrays(const float* x_sample_major, const float* labels, int examples, int features);
void bla_mnist_destroy(bla_mnist* data);
int bla_mnist_num_examples(const bla_mnist* data);
/* memset(sampled, 0) + num_sampled = 0 (model/mnist_nn.c:189-190). */
void bla_mnist_reset(bla_mnist* data);
/* get_random_data_take (lib/mnist_csv2.c:41-62) `count` times: the same libc rand() stream and the same index rule (the
 * element after the n-th unsampled one), found with a Fenwick tree in O(log examples) instead of the reference's scan. */
void bla_mnist_sample_take(bla_mnist* data, int count, int* indices_out);
/* The batch matrices of model/mnist_nn.c:199-217 gathered on the device: x_out [features x count] (raw pixel values),
 * y_out [classes x count] one-hot (may be NULL).  Only the indices cross PCIe. */
void bla_mnist_gather(bla_mnist* data, const int* indices_host, int count, float* x_out, float* y_out, int classes);
/* One epoch of model/mnist_nn.c:181-342 (sampler, batch assembly, SGD steps) on the device; stats_host = {average accuracy,
 * average loss} as the reference prints them (:340-341).  The net must have been created with max_batch >= batch_size.
 * Small batches (<= 4096 columns, >= 16 full batches, one GPU) replay one captured CUDA graph of the gather + step per batch,
 * with the same kernels in the same order as the eager loop (env BLA_MLP_GRAPH=0 keeps the eager loop). */
void bla_mlp_train_epoch(bla_mlp* net, bla_mnist* data, int batch_size, float lr_mult, double* stats_host);

/* model/mnist_hinge.c:100-172: ten one-vs-rest hinge classifiers, one full-batch iteration over a device-resident store
 * (two skinny GEMMs over the sample matrix + two small kernels), the reference's partial gradient clear (:126) included. */
typedef struct bla_hinge bla_hinge;
bla_hinge* bla_hinge_create(int features, int classes, int max_examples);
void bla_hinge_destroy(bla_hinge* h);
void bla_hinge_set_weights(bla_hinge* h, const float* w);   /* [classes][features], host or device */
void bla_hinge_get_weights(bla_hinge* h, float* w);
/* norms_host (may be NULL) receives |gradient_p| / N per model (:156). */
void bla_hinge_iteration(bla_hinge* h, bla_mnist* data, float learn_rate, float* norms_host);

typedef struct bla_cifar bla_cifar;
/* One CIFAR-10 binary batch file (10,000 records of 1 label + 3072 pixel bytes, lib/cifar10.c:6-11) uploaded as bytes. */
bla_cifar* bla_cifar_open(const char* filepath);
void bla_cifar_destroy(bla_cifar* data);
int bla_cifar_num_examples(const bla_cifar* data);
/* fill_random_data's draw (lib/cifar10.c:14) `count` times, libc rand(). */
void bla_cifar_sample(bla_cifar* data, int count, int* indices_out);
/* load_example (model/cifar_unet.c:221-233) for a batch: rows flipped bottom-up as lib/cifar10.c:24-31 does, pixels mapped to
 * [-1, 1]; x_out [count][3][32][32] on the device. */
void bla_cifar_gather(bla_cifar* data, const int* indices_host, int count, float* x_out);

/* ---- CSV checkpoint codec (lib/csv.c; include/lib/csv.h has the reference's four functions) ---------------- */

/* Parse a CSV text held in memory with the reference's tokenisation (lib/csv.c:28-54); *values_out is malloc'd. */
size_t bla_csv_parse(const char* text, size_t len, float** values_out);
/* Format rows x cols floats exactly as write_csv_contents would write them; returns the byte count (nothing is stored if
 * it exceeds cap or out is NULL). */
size_t bla_csv_format(const float* data, int cols, size_t rows, char* out, size_t cap);
/* A [rows x cols] tensor in device, managed or host memory <-> the reference's file format, staged through pinned memory. */
void bla_csv_save(const char* filepath, const float* data, int cols, size_t rows);
void bla_csv_load(const char* filepath, float* dst, size_t count);

/* ---- fused self attention: _forward_attention / _backward_attention (cifar_unet.c:999-1022, :1261-1335) ---- */

/* x, out [imgs][C][tokens] (tokens = H*W, key width 16 = SELF_ATTENTION_KEY_DIM); wqkv [C][48] = Q | K | V projection columns,
 * wo [16][C], bo [C].  z [imgs*tokens][C], qkv [imgs*tokens][48], probs [imgs][tokens][tokens] (softmax output) and
 * att [imgs*tokens][16] are written for the backward pass.  Device pointers; asynchronous on the library stream.
 * One transpose, one GEMM for Q|K|V, ONE kernel for scores / softmax / P.V, one GEMM (+bias), one transpose. */
void bla_attention_forward(const float* x, const float* wqkv, const float* wo, const float* bo, float* z, float* qkv,
                           float* probs, float* att, float* out, int imgs, int channels, int tokens);
/* Gradients of the three parameter tensors (summed over images) and, if dx != NULL, of the input.  The softmax Jacobian is
 * taken at the softmax output (the reference feeds it the pre-softmax scores, :1299 -- SURVEY D6). */
void bla_attention_backward(const float* dout, const float* wqkv, const float* wo, const float* z, const float* qkv,
                            const float* probs, const float* att, float* dwqkv, float* dwo, float* dbo, float* dx, int imgs,
                            int channels, int tokens);

/* ---- CIFAR U-Net trainer: model/cifar_unet.c as one batched device-resident step ---------------- */

typedef struct bla_unet bla_unet;
typedef struct bla_unet_config {
	int image_side;     /* IMAGE_HEIGHT = IMAGE_WIDTH = 32 (cifar_unet.c:26-27); a multiple of 8 */
	int dims[4];        /* RESOLUTION_{1..4}_EMBED_DIM = 128, 256, 256, 256 (:29-32) */
	int time_dim;       /* TIME_EMBED_DIM = 512 (:33) */
	int kernel_size;    /* KERNEL_SIZE = 3 (:34) */
	int group_size;     /* GROUP_SIZE = 32 channels per group (:35) */
	int key_dim;        /* SELF_ATTENTION_KEY_DIM = 16 (:36) -- the fused attention kernels are built for 16 */
	float dropout;      /* DROPOUT_RATE = 0.1 (:37); 0 disables the mask */
	int max_imgs;       /* largest batch a step will be given (the reference has no batch axis: 1 image per forward) */
	unsigned long long seed; /* dropout masks: element i of block b at step t is dropped iff
	                          * bla_host_uniform(seed + 7919*t + node_id(b))[i] < dropout */
} bla_unet_config;
/* Builds forward()'s graph (cifar_unet.c:1099-1168): 18 ResNet blocks, 5 attention blocks, 3 stride-2 convs, nearest-
 * neighbour up-sampling, skip concatenations, group norm + ReLU + 3-channel output conv.  Parameters are ONE flat buffer;
 * tensor i is [offset, offset + size) with the reference's per-file layout (conv kernels [F][C][k][k], time_weight
 * [time_dim][C], attention weight [key_dim][C]); Q, K, V projections are packed as the columns of one [C][3*key_dim]. */
bla_unet* bla_unet_create(const bla_unet_config* cfg);
void bla_unet_destroy(bla_unet* net);
size_t bla_unet_num_params(const bla_unet* net);
int bla_unet_num_tensors(const bla_unet* net);
const char* bla_unet_tensor_name(const bla_unet* net, int i);
size_t bla_unet_tensor_offset(const bla_unet* net, int i);
size_t bla_unet_tensor_size(const bla_unet* net, int i);
/* init_parameters (cifar_unet.c:1804-1851): He / Xavier uniform with the reference's fan-ins, zero biases. */
void bla_unet_init_params(bla_unet* net, unsigned long long seed);
void bla_unet_set_params(bla_unet* net, const float* flat);   /* host or device */
void bla_unet_get_params(bla_unet* net, float* flat);
void bla_unet_get_grads(bla_unet* net, float* flat);
/* save_parameters / load_parameters (cifar_unet.c:1484-1802): the reference's directory of CSV files, file for file
 * (data/cifar_unet/<level>/resnet_<i>/conv_1.csv ...; Q, K, V unpacked into query / key / value.csv), lib/csv.c's format.
 * Either program reads what the other wrote.  As in the reference, every ResNet block has a conv_3.csv (zeros here where
 * in == out channels: forward() never reads those kernels, :1062-1066, and this library keeps none), and conv_1.csv /
 * conv_3.csv of down_1/resnet_2 and up_{1..4}/resnet_1 hold only the input channels save_parameters declares (:1557,
 * :1614-1653: 3 of 128, C of 2C); the channels that format drops are written beside them as conv_1_rest.csv /
 * conv_3_rest.csv.  Loading a directory without *_rest.csv (one the reference wrote) leaves those channels unchanged. */
void bla_unet_save_csv(bla_unet* net, const char* dir);
void bla_unet_load_csv(bla_unet* net, const char* dir);
/* forward() on x [imgs][3][H][W] with one time embedding row per image, time_emb [imgs][time_dim] (the reference never
 * initialises its own); out [imgs][3][H][W].  Host or device pointers; dropout off. */
void bla_unet_forward(bla_unet* net, const float* x, const float* time_emb, int imgs, float* out);
/* forward + compute_mse_loss (:1858) + backward (:1351) + SGD: dY = 2 (out - noise) as :1353-1365, gradients summed over
 * the images (and all-reduced over ranks when a communicator is active), params -= lr * grads (lr = 0: gradients only).
 * loss_host, if not NULL, receives the sum over images of the per-image MSE and forces a synchronise. */
void bla_unet_train_step(bla_unet* net, const float* x, const float* time_emb, const float* noise, int imgs, float lr,
                         double* loss_host);

/* ---- NCCL over NVLink: one process per GPU -------------------------------------------------- */

/* Rank 0 fills a 128-byte id (ncclGetUniqueId) that the launcher broadcasts out of band
 * (torch.distributed store, MPI, a file ...); every rank then calls bla_comm_init with it. */
void bla_comm_unique_id(void* id128);
void bla_comm_init(const void* id128, int rank, int world);
int bla_comm_world(void);
/* 1 once this process' small all-reduces run as the library's own kernel over NVLink peer windows (cudaIpc-mapped, csrc/comm.cu;
 * default when every rank can map every other rank's window, BLA_PEER_ALLREDUCE=0 keeps NCCL), 0 while / when they go through NCCL */
int bla_comm_peer_windows(void);
/* collective: 0 = route the all-reduces through NCCL from now on, 1 = back over the peer windows (A/B measurements, fallback) */
void bla_comm_set_peer_windows(int on);
int bla_comm_rank(void);
/* In-place sum all-reduce of device buffers on the library stream. */
void bla_allreduce_sum_f32(float* buf, size_t n);
void bla_allreduce_sum_f64(double* buf, size_t n);
/* Broadcast a device buffer from `root` (row-sharded GEMM: every rank needs all of B). */
void bla_broadcast_f32(float* buf, size_t n, int root);
void bla_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif
