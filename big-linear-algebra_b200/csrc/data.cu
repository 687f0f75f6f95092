// data.cu -- the data pipeline in front of the hot loops: lib/mnist_csv2.c and lib/cifar10.c with the
// dataset resident in HBM and the per-batch gather on the device.
//
// The reference keeps the parsed MNIST CSV in host memory feature-major (mnist_csv2.c:26-31), draws every sample of a batch
// with get_random_data_take -- an O(examples) scan of the `sampled` flags per draw, 3.6e9 steps per epoch at 60,000 examples
// (mnist_csv2.c:41-62) -- and copies 784 strided floats per sample into the batch matrix (mnist_nn.c:204-217).  Here
//   * the file is parsed by the parallel CSV codec (csv_codec.cu) and uploaded ONCE, sample-major (784 contiguous floats per
//     example: a gathered row is 3136 contiguous bytes);
//   * the draw keeps the reference's libc rand() stream and its exact index rule (the element AFTER the n-th unsampled one,
//     quirk included: it can hit an already sampled example) but finds it with a Fenwick tree in O(log examples);
//   * one kernel gathers the drawn rows into the [784 x batch] input matrix (transposed through shared memory, coalesced both
//     ways) and writes the one-hot [10 x batch] expectations; per batch only the 4-byte indices cross PCIe.
// CIFAR-10 (cifar10.c:13-32, cifar_unet.c:221-233): the 30.73 MB batch file is uploaded as bytes; the gather flips the rows
// bottom-up as fill_random_data does and maps pixels to [-1, 1].
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/bla.h"
#include "../../include/lib/csv.h"
#include "kernels.h"
#include "runtime.h"
#include "take_sampler.h"

namespace bla {
bool comm_active();   // comm.cu
}
using namespace bla;

struct bla_mnist {
    int n = 0, features = 0;
    float* x = nullptr;                  // device [n][features]
    float* y = nullptr;                  // device [n] labels
    TakeSampler sampler;                 // get_random_data_take's index rule on a Fenwick tree (take_sampler.h)
    std::vector<char> sampled;           // mnist_csv2.h:11
    int num_sampled = 0;
    int* idx_dev = nullptr;              // staging for a batch's indices
    int* idx_pin = nullptr;
    int idx_cap = 0;
    int* cursor = nullptr;               // device: first index of the batch a replayed step graph gathers (advanced by the graph)
};

struct bla_cifar {
    int n = 0;
    unsigned char* bytes = nullptr;      // device copy of the batch file: n records of 1 label + 3072 pixel bytes
    int* idx_dev = nullptr;
    int* idx_pin = nullptr;
    int idx_cap = 0;
};

namespace {

constexpr int kThreads = 256;

// out_x[f][k] = x[idx[k]][f]  (mnist_nn.c:209-211, the 1/255 scaling stays in the training step);  out_y[c][k] = (label == c)
__global__ void __launch_bounds__(kThreads) mnist_gather_kernel(const float* __restrict__ x, const float* __restrict__ y, const int* __restrict__ idx,
                                                                int count, int features, int classes, float* __restrict__ out_x,
                                                                float* __restrict__ out_y, const int* __restrict__ cursor) {
    __shared__ float tile[32][33];
    if (cursor) idx += *cursor;                                // replayed from a graph: the batch position lives on the device
    const int k0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int kk = ty; kk < 32; kk += 8) {                    // rows of the store: 32 consecutive features = 128 bytes
        const int k = k0 + kk, f = f0 + tx;
        tile[kk][tx] = (k < count && f < features) ? x[(size_t)idx[k] * features + f] : 0.f;
    }
    __syncthreads();
    for (int ff = ty; ff < 32; ff += 8) {
        const int f = f0 + ff, k = k0 + tx;
        if (f < features && k < count) out_x[(size_t)f * count + k] = tile[tx][ff];
    }
    if (blockIdx.y == 0 && out_y) {
        for (int e = threadIdx.x; e < 32 * classes; e += kThreads) {
            const int c = e / 32, k = k0 + (e & 31);
            if (k < count) out_y[(size_t)c * count + k] = ((int)y[idx[k]] == c) ? 1.f : 0.f;   // mnist_nn.c:213-217
        }
    }
}

__global__ void cursor_set_kernel(int* cursor, int value) { *cursor = value; }
__global__ void cursor_advance_kernel(int* cursor, int by) { *cursor += by; }

// out[k][c][i][j] = (byte[idx[k]][1 + c*1024 + (31 - i)*32 + j] - 127.5) / 127.5   (cifar10.c:24-31 flip, cifar_unet.c:226-232)
__global__ void __launch_bounds__(kThreads) cifar_gather_kernel(const unsigned char* __restrict__ bytes, const int* __restrict__ idx, int count,
                                                                float* __restrict__ out) {
    const size_t total = (size_t)count * 3072;
    for (size_t e = (size_t)blockIdx.x * kThreads + threadIdx.x; e < total; e += (size_t)gridDim.x * kThreads) {
        const int k = (int)(e / 3072), r = (int)(e % 3072);
        const int c = r >> 10, i = (r >> 5) & 31, j = r & 31;
        const unsigned char v = bytes[(size_t)idx[k] * 3073 + 1 + c * 1024 + (31 - i) * 32 + j];
        out[e] = ((float)v - 127.5f) / 127.5f;
    }
}

void sampler_reset(bla_mnist* m) {
    m->sampler.reset_all_unsampled(m->n);
    std::fill(m->sampled.begin(), m->sampled.end(), 0);
    m->num_sampled = 0;
}

template <class T>
void ensure_idx(T* s, int count) {
    if (count <= s->idx_cap) return;
    if (s->idx_dev) { pool_free(s->idx_dev); pool_free(s->idx_pin); }
    s->idx_cap = count;
    s->idx_dev = (int*)pool_alloc(kDevice, (size_t)count * sizeof(int));
    s->idx_pin = (int*)pool_alloc(kPinned, (size_t)count * sizeof(int));
}

bla_mnist* mnist_new(const float* x_sample_major, const float* y, int n, int features) {
    rt();
    bla_mnist* m = new bla_mnist();
    m->n = n; m->features = features;
    m->x = (float*)pool_alloc(kDevice, (size_t)(n ? n : 1) * features * sizeof(float));
    m->y = (float*)pool_alloc(kDevice, (size_t)(n ? n : 1) * sizeof(float));
    cudaStream_t s = rt().stream;
    BLA_CUDA(cudaMemcpyAsync(m->x, x_sample_major, (size_t)n * features * sizeof(float), cudaMemcpyDefault, s));
    BLA_CUDA(cudaMemcpyAsync(m->y, y, (size_t)n * sizeof(float), cudaMemcpyDefault, s));
    BLA_CUDA(cudaStreamSynchronize(s));
    rt().h2d_bytes += (size_t)n * (features + 1) * sizeof(float);
    m->sampled.assign(n, 0);
    sampler_reset(m);
    return m;
}

}  // namespace

extern "C" {

// mnist_csv_init (mnist_csv2.c:13-34): every row is `label, 784 pixels,`; num_examples = commas / 785
bla_mnist* bla_mnist_from_csv(const char* filepath) {
    FILE* f = fopen(filepath, "r");
    if (!f) die("bla: cannot open MNIST CSV file %s, exiting", filepath);
    int commas = 0;
    float* v = read_csv_contents_file(f, &commas);
    const int line = 785, n = commas / line;
    std::vector<float> x((size_t)n * (line - 1)), y(n);
    for (int i = 0; i < n; ++i) {
        y[i] = v[(size_t)i * line];
        memcpy(&x[(size_t)i * (line - 1)], v + (size_t)i * line + 1, (line - 1) * sizeof(float));
    }
    free(v);
    return mnist_new(x.data(), y.data(), n, line - 1);
}

bla_mnist* bla_mnist_from_arrays(const float* x_sample_major, const float* labels, int examples, int features) {
    return mnist_new(x_sample_major, labels, examples, features);
}

void bla_mnist_destroy(bla_mnist* m) {
    if (!m) return;
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
    pool_free(m->x); pool_free(m->y);
    if (m->idx_dev) { pool_free(m->idx_dev); pool_free(m->idx_pin); }
    if (m->cursor) pool_free(m->cursor);
    delete m;
}

int bla_mnist_num_examples(const bla_mnist* m) { return m->n; }
}  // extern "C"
// for the other device-resident loops of the library (hinge.cu)
const float* bla_mnist_x_device(const bla_mnist* m) { return m->x; }
const float* bla_mnist_y_device(const bla_mnist* m) { return m->y; }
extern "C" {

// mnist_nn.c:189-190: a new round of SGD
void bla_mnist_reset(bla_mnist* m) { sampler_reset(m); }

// get_random_data_take (mnist_csv2.c:41-62) `count` times: same libc rand() stream, same index rule
void bla_mnist_sample_take(bla_mnist* m, int count, int* indices_out) {
    for (int k = 0; k < count; ++k) indices_out[k] = m->sampler.take(m->sampled.data(), &m->num_sampled);
}

// the batch matrices of mnist_nn.c:199-217 on the device: x_out [features x count] raw pixel values, y_out [classes x count]
static void mnist_gather_launch(bla_mnist* m, const int* idx_dev, int count, float* x_out, float* y_out, int classes, cudaStream_t s,
                                const int* cursor = nullptr) {
    mnist_gather_kernel<<<dim3(ceil_div(count, 32), ceil_div(m->features, 32)), kThreads, 0, s>>>(m->x, m->y, idx_dev, count, m->features, classes,
                                                                                               x_out, y_out, cursor);
    BLA_LAUNCH_CHECK();
    count_launch();
}

void bla_mnist_gather(bla_mnist* m, const int* indices_host, int count, float* x_out, float* y_out, int classes) {
    if (count <= 0) return;
    cudaStream_t s = rt().stream;
    ensure_idx(m, count);
    BLA_CUDA(cudaStreamSynchronize(s));                  // the previous batch's copy out of the pinned staging has finished
    memcpy(m->idx_pin, indices_host, (size_t)count * sizeof(int));
    BLA_CUDA(cudaMemcpyAsync(m->idx_dev, m->idx_pin, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, s));
    rt().h2d_bytes += (size_t)count * sizeof(int);
    mnist_gather_launch(m, m->idx_dev, count, x_out, y_out, classes, s);
}

// One epoch of model/mnist_nn.c:181-342 with the dataset, the batch assembly and the step all on the device.
// stats_host = {epoch average accuracy, epoch average loss} as printed at :340-341.  The draws of the whole epoch are made up
// front (nothing else consumes rand() inside the reference's epoch loop, so the stream is the same) and cross PCIe once.
void bla_mlp_train_epoch(bla_mlp* net, bla_mnist* data, int batch_size, float lr_mult, double* stats_host) {
    const int n = data->n;
    if (n <= 0 || batch_size <= 0) return;
    int dims[4];
    bla_mlp_dims(net, dims);
    if (dims[0] != data->features)
        die("bla: bla_mlp_train_epoch: the network takes %d inputs, the dataset has %d features, exiting", dims[0], data->features);
    const int classes = dims[3];                                                        // 10 in the reference (:16)
    cudaStream_t s = rt().stream;
    const int num_batches = (int)ceil((float)n / (float)batch_size);                    // :187
    bla_mnist_reset(data);                                                              // :189-190
    ensure_idx(data, n);
    BLA_CUDA(cudaStreamSynchronize(s));
    bla_mnist_sample_take(data, n, data->idx_pin);
    BLA_CUDA(cudaMemcpyAsync(data->idx_dev, data->idx_pin, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    rt().h2d_bytes += (size_t)n * sizeof(int);
    float* xb = (float*)pool_alloc(kDevice, (size_t)data->features * batch_size * sizeof(float));
    float* yb = (float*)pool_alloc(kDevice, (size_t)classes * batch_size * sizeof(float));
    double stats[2];
    bla_mlp_read_stats(net, stats);                                                     // clear the accumulators
    auto eager_step = [&](int j) {
        const int remaining = n - j * batch_size;
        const int cnt = remaining > batch_size ? batch_size : remaining;                // :194-195
        mnist_gather_launch(data, data->idx_dev + (size_t)j * batch_size, cnt, xb, yb, classes, s);
        bla_mlp_train_step(net, xb, yb, cnt, cnt, 0, lr_mult, nullptr);
    };
    // Small batches are launch-bound (~20 launches of a few microseconds each): the gather + step of one full batch is captured
    // once into a CUDA graph and replayed for every full batch of the epoch; the batch position is a device counter the graph
    // advances itself, so a replay needs no new parameters.  Step 0 runs eagerly first (pool allocations and one-time
    // attribute calls settle outside the capture); a ragged last batch is another shape and runs eagerly too.  Measured on B200
    // (profiles/r01_epoch_graph.json): 0.118 -> 0.103 s per 60,000-example epoch at batch 64, bit-identical parameters; what is
    // left is the ~100 us of seventeen latency-bound kernels per step.  BLA_MLP_GRAPH=0 launches every step eagerly.
    static int graph_on = -1;
    if (graph_on < 0) { const char* e = getenv("BLA_MLP_GRAPH"); graph_on = e ? atoi(e) : 1; }
    const int full = n / batch_size;
    int j = 0;
    if (graph_on && !comm_active() && full >= 16 && batch_size <= 4096) {
        eager_step(j++);
        if (!data->cursor) data->cursor = (int*)pool_alloc(kDevice, sizeof(int));
        cursor_set_kernel<<<1, 1, 0, s>>>(data->cursor, batch_size);
        BLA_LAUNCH_CHECK();
        count_launch();
        const unsigned long long before = rt().launches;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        // a stream that cannot be captured (the caller handed bla_set_stream the legacy default stream): stay eager
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
            mnist_gather_launch(data, data->idx_dev, batch_size, xb, yb, classes, s, data->cursor);
            bla_mlp_train_step(net, xb, yb, batch_size, batch_size, 0, lr_mult, nullptr);
            cursor_advance_kernel<<<1, 1, 0, s>>>(data->cursor, batch_size);
            BLA_LAUNCH_CHECK();
            count_launch();
            BLA_CUDA(cudaStreamEndCapture(s, &graph));
            const int per_replay = (int)(rt().launches - before);
            rt().launches = before;                                                     // nothing ran while capturing
            BLA_CUDA(cudaGraphInstantiate(&exec, graph, 0));
            for (; j < full; ++j) {
                BLA_CUDA(cudaGraphLaunch(exec, s));
                count_launch(per_replay);
            }
            BLA_CUDA(cudaStreamSynchronize(s));
            BLA_CUDA(cudaGraphExecDestroy(exec));
            BLA_CUDA(cudaGraphDestroy(graph));
        } else {
            cudaGetLastError();
        }
    }
    for (; j < num_batches; ++j) eager_step(j);
    bla_mlp_read_stats(net, stats);
    if (stats_host) { stats_host[0] = stats[1] / (double)(float)n; stats_host[1] = stats[0] / (double)(float)n; }
    pool_free(xb); pool_free(yb);
}

// ---- CIFAR-10 ------------------------------------------------------------------------------------------------------
bla_cifar* bla_cifar_open(const char* filepath) {
    rt();
    FILE* f = fopen(filepath, "rb");
    if (!f) die("bla: cannot open CIFAR-10 batch file %s, exiting", filepath);
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    rewind(f);
    bla_cifar* c = new bla_cifar();
    c->n = (int)(size / 3073);
    unsigned char* pin = (unsigned char*)pool_alloc(kPinned, (size_t)size + 1);
    if (fread(pin, 1, (size_t)size, f) != (size_t)size) die("bla: short read from %s, exiting", filepath);
    fclose(f);
    c->bytes = (unsigned char*)pool_alloc(kDevice, (size_t)size + 1);
    BLA_CUDA(cudaMemcpyAsync(c->bytes, pin, (size_t)size, cudaMemcpyHostToDevice, rt().stream));
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
    rt().h2d_bytes += (size_t)size;
    pool_free(pin);
    return c;
}
void bla_cifar_destroy(bla_cifar* c) {
    if (!c) return;
    BLA_CUDA(cudaStreamSynchronize(rt().stream));
    pool_free(c->bytes);
    if (c->idx_dev) { pool_free(c->idx_dev); pool_free(c->idx_pin); }
    delete c;
}
int bla_cifar_num_examples(const bla_cifar* c) { return c->n; }
// fill_random_data's draw (cifar10.c:14): uniform over the 10,000 records of a batch file, libc rand()
void bla_cifar_sample(bla_cifar* c, int count, int* indices_out) {
    for (int k = 0; k < count; ++k) {
        unsigned int e = (unsigned int)(((float)rand() / ((float)RAND_MAX + 1)) * 10000u);
        if ((int)e >= c->n) e = (unsigned int)(c->n - 1);
        indices_out[k] = (int)e;
    }
}
// load_example (cifar_unet.c:221-233) for a batch: x_out [count][3][32][32] in [-1, 1]
void bla_cifar_gather(bla_cifar* c, const int* indices_host, int count, float* x_out) {
    if (count <= 0) return;
    cudaStream_t s = rt().stream;
    ensure_idx(c, count);
    BLA_CUDA(cudaStreamSynchronize(s));
    memcpy(c->idx_pin, indices_host, (size_t)count * sizeof(int));
    BLA_CUDA(cudaMemcpyAsync(c->idx_dev, c->idx_pin, (size_t)count * sizeof(int), cudaMemcpyHostToDevice, s));
    rt().h2d_bytes += (size_t)count * sizeof(int);
    size_t blocks = ((size_t)count * 3072 + kThreads - 1) / kThreads, cap = (size_t)rt().num_sms * 8;
    if (blocks > cap) blocks = cap;
    cifar_gather_kernel<<<(int)blocks, kThreads, 0, s>>>(c->bytes, c->idx_dev, count, x_out);
    BLA_LAUNCH_CHECK();
    count_launch();
}

}  // extern "C"
