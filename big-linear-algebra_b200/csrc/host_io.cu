// host_io.cu -- the reference's host-side data objects, so that a model program links against libbla.so ALONE:
//   lib/mnist_csv2.h  mnist_csv_init, get_random_data_replace, get_random_data_take, visualize_digit_data   (mnist_nn.c)
//   lib/cifar10.h     the CIFAR10_* constants, fill_random_data                                            (cifar_unet.c)
//   lib/bmp.h         write_bmp_data                                                                         (cifar_unet.c)
// (lib/mnist_csv.h -- mnist_hinge.c's reader -- defines a second, incompatible `struct MnistCSV` and `visualize_digit_data`;
// it lives in its own small library, host_io/mnist_csv.c -> libbla_mnist_csv.so, and is never linked together with this file's
// symbols in one program, as in the reference.)
//
// Host code only: nothing here touches the GPU, none of it needs one.  Behaviour follows the reference files cited at each
// function -- same struct layouts, same text on stdout / stderr, same libc rand() consumption, same bytes in the files -- with
// two differences in kind: the MNIST loader uses the parallel CSV codec and a blocked transpose, and the draw without
// replacement finds its example on a Fenwick tree instead of scanning the flags (take_sampler.h).
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <unistd.h>

#include "../../include/lib/bmp.h"
#include "../../include/lib/cifar10.h"
#include "../../include/lib/csv.h"
#include "../../include/lib/mnist_csv2.h"
#include "take_sampler.h"

namespace {

constexpr int kMnistRow = 785;   // label + 28 x 28 pixels per CSV row (mnist_csv2.c:8)

// The reference keeps all sampler state in the caller's struct; the tree that mirrors `sampled` lives here, keyed by that
// struct.  Model code resets the flags itself (mnist_nn.c:189-191 clears `sampled` and `num_sampled` every epoch), so a draw
// first checks that the struct still looks the way the last draw left it and rebuilds the tree from the flags if not.
struct Mirror {
    bla::TakeSampler tree;
    const char* flags = nullptr;
    int num_sampled = -1;
};
std::unordered_map<const MnistCSV*, Mirror>& mirrors() {
    static std::unordered_map<const MnistCSV*, Mirror> m;
    return m;
}

template <class Fn>
void run_threads(int count, Fn fn) {
    if (count <= 1) { fn(0); return; }
    std::vector<std::thread> th;
    for (int i = 1; i < count; ++i) th.emplace_back(fn, i);
    fn(0);
    for (auto& t : th) t.join();
}

inline void put_le32(uint8_t* p, uint32_t v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

void write_all(int fd, const void* data, size_t bytes, const char* what) {
    if (write(fd, data, bytes) != (ssize_t)bytes) fprintf(stderr, "Error while writing bitmap %s (errno=%d).\n", what, errno);
}

}  // namespace

extern "C" {

// ---- lib/mnist_csv2.h ------------------------------------------------------------------------------------------------------

// mnist_csv2.c:13-34.  csv->file is read to its end and closed (read_csv_contents_file's side effect); every other field is
// overwritten.  X is FEATURE-major, X[example + feature * num_examples] (:29), all three arrays are malloc'd: the caller frees
// them (mnist_nn.c:391-393).
void mnist_csv_init(MnistCSV* csv) {
    int commas = 0;
    float* values = read_csv_contents_file(csv->file, &commas);
    printf("MNIST CSV file contents read!\n");
    const int n = commas / kMnistRow;
    csv->num_examples = n;
    csv->X = (float*)malloc((size_t)n * (kMnistRow - 1) * sizeof(float));
    csv->y = (float*)malloc((size_t)n * sizeof(float));
    csv->num_sampled = 0;
    csv->sampled = (char*)calloc(n > 0 ? (size_t)n : 1, 1);
    // [n x 785] rows -> labels + [784 x n]: tiles of 32 examples, so both sides move whole cache lines
    const unsigned hc = std::thread::hardware_concurrency();
    const int workers = n >= 4096 ? (int)std::min<unsigned>(hc ? hc : 4, 16) : 1;
    const int tiles = (n + 31) / 32;
    float* X = csv->X;
    float* y = csv->y;
    run_threads(workers, [=](int w) {
        for (int t = tiles * w / workers; t < tiles * (w + 1) / workers; ++t) {
            const int i0 = t * 32, i1 = std::min(n, i0 + 32);
            for (int i = i0; i < i1; ++i) y[i] = values[(size_t)i * kMnistRow];
            for (int j = 0; j < kMnistRow - 1; ++j)
                for (int i = i0; i < i1; ++i) X[(size_t)j * n + i] = values[(size_t)i * kMnistRow + 1 + j];
        }
    });
    free(values);
    mirrors().erase(csv);
}

// mnist_csv2.c:36-39: uniform with replacement.  (rand() == RAND_MAX indexes one past the end in the reference; the last
// example is returned here.)
MnistExample get_random_data_replace(MnistCSV* csv) {
    int n = (int)floor((float)csv->num_examples * (float)rand() / (float)RAND_MAX);
    if (n >= csv->num_examples) n = csv->num_examples - 1;
    if (n < 0) n = 0;
    MnistExample ex;
    ex.X = csv->X + n;
    ex.y = csv->y[n];
    ex.num_examples = csv->num_examples;
    return ex;
}

// mnist_csv2.c:41-62: uniform without replacement, the reference's index rule (take_sampler.h)
MnistExample get_random_data_take(MnistCSV* csv) {
    Mirror& m = mirrors()[csv];
    if (m.flags != csv->sampled || m.num_sampled != csv->num_sampled || m.tree.n != csv->num_examples) {
        m.tree.rebuild(csv->sampled, csv->num_examples);
        m.flags = csv->sampled;
    }
    const int i = m.tree.take(csv->sampled, &csv->num_sampled);
    m.num_sampled = csv->num_sampled;
    MnistExample ex;
    ex.X = csv->X + i;
    ex.y = csv->y[i];
    ex.num_examples = csv->num_examples;
    return ex;
}

// mnist_csv2.c:64-79: a 28 x 28 character picture of one example (pixel (i, j) is X[(i * 28 + j) * num_examples])
void visualize_digit_data(MnistExample ex) {
    static const char kRule[] = "============================\n";
    fputs(kRule, stdout);
    printf("Data for digit %f:\n", ex.y);
    char line[30];
    for (int i = 0; i < 28; ++i) {
        for (int j = 0; j < 28; ++j) {
            const float v = ex.X[(size_t)(i * 28 + j) * ex.num_examples];
            line[j] = v < 80 ? ' ' : v < 150 ? ':' : '#';
        }
        line[28] = '\n';
        line[29] = '\0';
        fputs(line, stdout);
    }
    fputs(kRule, stdout);
}

// ---- lib/cifar10.h ---------------------------------------------------------------------------------------------------------
// cifar10.c:6-11: the layout of a CIFAR-10 binary batch file (10,000 records of 1 label byte + 3 x 1024 pixel bytes)
const unsigned int CIFAR10_NUM_EXAMPLES_PER_FILE = 10000;
const unsigned int CIFAR10_LINE_LENGTH = 3073;
const unsigned int CIFAR10_DATA_LENGTH = 3072;
const unsigned int CIFAR10_BATCH_FILE_SIZE = 30730000;
const unsigned int CIFAR10_NUM_PIXELS = 1024;
const unsigned int CIFAR10_EXAMPLE_DIM = 32;

// cifar10.c:13-31: one uniformly drawn record's pixels (the label byte is skipped), each colour plane with its rows in
// bottom-up order.  The file offset is left behind the record, as lseek + read leave it.
void fill_random_data(int fd, uint8_t* arr) {
    const unsigned int example = (unsigned int)(((float)rand() / ((float)RAND_MAX + 1)) * CIFAR10_NUM_EXAMPLES_PER_FILE);
    const off_t at = (off_t)example * CIFAR10_LINE_LENGTH + 1;
    if (lseek(fd, at, SEEK_SET) != at) fprintf(stderr, "Error while seeking to CIFAR10 example %d (errno=%d).\n", example, errno);
    uint8_t record[3072];
    if (read(fd, record, sizeof(record)) != (ssize_t)sizeof(record))
        fprintf(stderr, "Error while reading CIFAR10 example %d (errno=%d).\n", example, errno);
    for (int plane = 0; plane < 3; ++plane)
        for (int row = 0; row < 32; ++row) memcpy(arr + plane * 1024 + row * 32, record + plane * 1024 + (31 - row) * 32, 32);
}

// ---- lib/bmp.h -------------------------------------------------------------------------------------------------------------
// bmp.c:11-100: an uncompressed 24-bit BMP -- 14-byte file header, 40-byte BITMAPINFOHEADER (72 x 72 pixels per metre, 1 in
// the colour-count field, as the reference writes them), then the rows in the order given, blue-green-red, each padded to a
// multiple of four bytes.  Like the reference the file is opened without O_TRUNC, so a longer file that was there keeps its
// tail.  (Byte 33 of the info header is never assigned in the reference -- stack contents; 0 here.)
void write_bmp_data(const char* filepath, BMPData* data) {
    const int fd = open(filepath, O_WRONLY | O_CREAT, 0777);
    const uint32_t w = data->width, h = data->height;
    const uint32_t stride = (24 * w + 31) / 32 * 4;
    uint8_t head[54] = {0};
    head[0] = 'B'; head[1] = 'M';
    put_le32(head + 2, 54 + stride * h);
    put_le32(head + 10, 54);
    put_le32(head + 14, 40);
    put_le32(head + 18, w & 0x7FFFFFFFu);
    put_le32(head + 22, h & 0x7FFFFFFFu);
    head[26] = 1;
    head[28] = 24;
    put_le32(head + 38, 72);
    put_le32(head + 42, 72);
    head[46] = 1;
    write_all(fd, head, 14, "header");
    write_all(fd, head + 14, 40, "info header");
    std::vector<uint8_t> rows((size_t)stride * h, 0);
    for (uint32_t i = 0; i < h; ++i) {
        uint8_t* out = rows.data() + (size_t)i * stride;
        const size_t at = (size_t)i * w;
        for (uint32_t j = 0; j < w; ++j) {
            out[3 * j] = data->blue[at + j];
            out[3 * j + 1] = data->green[at + j];
            out[3 * j + 2] = data->red[at + j];
        }
    }
    write_all(fd, rows.data(), rows.size(), "pixel data");
    close(fd);
}

}  // extern "C"
