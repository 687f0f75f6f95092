// runtime.cu -- see runtime.h.  Host-side only (no kernels).
#include "runtime.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>

#include "../../include/bla.h"

namespace bla {

void die(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    printf("\n");
    fflush(stdout);
    exit(1);
}

void check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        die("bla: CUDA error %d (%s) at %s:%d in `%s`, exiting", (int)e, cudaGetErrorString(e), file, line, what);
    }
}

static Runtime g_rt;
static bool g_init = false;

static int g_pdl_off = 0;
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("BLA_PDL"); on = e ? atoi(e) : 1; }
    return on != 0 && g_pdl_off == 0;
}
PdlOff::PdlOff() { ++g_pdl_off; }
PdlOff::~PdlOff() { --g_pdl_off; }
static int g_pdl_late = 0;
bool pdl_early() { return g_pdl_late == 0; }
PdlLate::PdlLate() { ++g_pdl_late; }
PdlLate::~PdlLate() { --g_pdl_late; }
bool rt_initialised() { return g_init; }

void rt_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        die("bla: no usable CUDA device (%s); libbla has no CPU fallback, exiting",
            e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= n) device = 0;
    if (g_init && g_rt.device == device) return;
    BLA_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    BLA_CUDA(cudaGetDeviceProperties(&p, device));
    g_rt.device = device;
    g_rt.num_sms = p.multiProcessorCount;
    g_rt.cc_major = p.major;
    g_rt.cc_minor = p.minor;
    g_rt.l2_bytes = (size_t)p.l2CacheSize;
    if (p.major != 10) {
        die("bla: device %d is sm_%d%d; this library only carries sm_100a code, exiting", device, p.major, p.minor);
    }
    BLA_CUDA(cudaStreamCreateWithFlags(&g_rt.own_stream, cudaStreamNonBlocking));
    g_rt.stream = g_rt.own_stream;
    const char* path = getenv("BLA_PATH");
    if (path) {
        if (!strcmp(path, "fp32")) g_rt.gemm_path = BLA_GEMM_FP32;
        else if (!strcmp(path, "3xtf32")) g_rt.gemm_path = BLA_GEMM_3XTF32;
        else g_rt.gemm_path = BLA_GEMM_AUTO;
    }
    const char* q = getenv("BLA_QUIRKS");
    if (q) g_rt.quirks = atoi(q) != 0;
    g_init = true;
}

Runtime& rt() {
    if (!g_init) {
        int dev = 0;
        const char* d = getenv("BLA_DEVICE");
        if (!d) d = getenv("LOCAL_RANK");
        if (d) dev = atoi(d);
        rt_init(dev);
    }
    return g_rt;
}

// ---------------------------------------------------------------------------------------------
// Registry + pools
// ---------------------------------------------------------------------------------------------
struct Block {
    size_t size;
    MemKind kind;
    bool in_use;
};
static std::map<uintptr_t, Block> g_blocks;                           // every block we ever created
static std::unordered_map<size_t, std::vector<void*>> g_free[4];     // per kind, by rounded size

static size_t round_size(size_t bytes) {
    if (bytes < 256) return 256;
    size_t p = 256;
    while ((p << 1) <= bytes) p <<= 1;  // largest power of two <= bytes
    size_t step = p >> 3;               // 8 classes per octave: <= 12.5% slack
    if (step < 256) step = 256;
    return (bytes + step - 1) / step * step;
}

void* pool_alloc(MemKind kind, size_t bytes) {
    rt();
    size_t sz = round_size(bytes);
    auto& fl = g_free[kind][sz];
    void* p = nullptr;
    if (!fl.empty()) {
        p = fl.back();
        fl.pop_back();
        g_blocks[(uintptr_t)p].in_use = true;
        return p;
    }
    cudaError_t e;
    switch (kind) {
        case kDevice: e = cudaMalloc(&p, sz); break;
        case kManaged: e = cudaMallocManaged(&p, sz, cudaMemAttachGlobal); break;
        case kPinned: e = cudaMallocHost(&p, sz); break;
        default: die("bla: pool_alloc of host kind");
    }
    if (e != cudaSuccess) {
        // give cached blocks back to the driver once, then retry
        cudaGetLastError();
        BLA_CUDA(cudaStreamSynchronize(g_rt.stream));
        for (int k = 1; k < 4; ++k) {
            for (auto& kv : g_free[k]) {
                for (void* q : kv.second) {
                    g_blocks.erase((uintptr_t)q);
                    if (k == kPinned) cudaFreeHost(q); else cudaFree(q);
                }
                kv.second.clear();
            }
        }
        switch (kind) {
            case kDevice: e = cudaMalloc(&p, sz); break;
            case kManaged: e = cudaMallocManaged(&p, sz, cudaMemAttachGlobal); break;
            default: e = cudaMallocHost(&p, sz); break;
        }
        if (e != cudaSuccess) die("bla: out of memory allocating %zu bytes (kind %d): %s, exiting", sz, (int)kind, cudaGetErrorString(e));
    }
    g_blocks[(uintptr_t)p] = Block{sz, kind, true};
    return p;
}

bool pool_free(void* p) {
    if (!p || g_blocks.empty()) return false;
    auto it = g_blocks.find((uintptr_t)p);
    if (it == g_blocks.end()) return false;
    if (!it->second.in_use) die("bla: double free of library memory %p, exiting", p);
    it->second.in_use = false;
    g_free[it->second.kind][it->second.size].push_back(p);
    return true;
}

// bytes from p to the end of the library block that holds it; 0 for foreign pointers
size_t owned_bytes_from(const void* p) {
    if (g_blocks.empty()) return 0;
    auto it = g_blocks.upper_bound((uintptr_t)p);
    if (it == g_blocks.begin()) return 0;
    --it;
    const uintptr_t end = it->first + it->second.size;
    return (uintptr_t)p < end ? (size_t)(end - (uintptr_t)p) : 0;
}

MemKind classify(const void* p) {
    if (!g_blocks.empty()) {
        auto it = g_blocks.upper_bound((uintptr_t)p);
        if (it != g_blocks.begin()) {
            --it;
            if ((uintptr_t)p < it->first + it->second.size) return it->second.kind;
        }
    }
    rt();
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return kHost;
    }
    switch (a.type) {
        case cudaMemoryTypeDevice: return kDevice;
        case cudaMemoryTypeManaged: return kManaged;
        case cudaMemoryTypeHost: return kPinned;
        default: return kHost;
    }
}

// ---------------------------------------------------------------------------------------------
// CallScope
// ---------------------------------------------------------------------------------------------
CallScope::CallScope() { rt(); }
CallScope::~CallScope() { finish(); }
cudaStream_t CallScope::stream() const { return g_rt.stream; }

void CallScope::note_kind(MemKind k) {
    if (k != kDevice) any_host_visible_ = true;
}

void* CallScope::stage(const void* p, size_t bytes, bool copy_in, bool copy_back) {
    if (bytes == 0) return const_cast<void*>(p);
    MemKind k = classify(p);
    note_kind(k);
    if (k == kDevice || k == kManaged) return const_cast<void*>(p);
    // ordinary or pinned host memory: bounce through HBM
    void* d = pool_alloc(kDevice, bytes);
    temps_.push_back(d);
    if (copy_in) {
        BLA_CUDA(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, g_rt.stream));
        g_rt.h2d_bytes += bytes;
    }
    if (copy_back) backs_.push_back(Back{const_cast<void*>(p), d, bytes});
    return d;
}

const float* CallScope::in(const float* p, size_t n) { return (const float*)stage(p, n * sizeof(float), true, false); }
float* CallScope::out(float* p, size_t n) { return (float*)stage(p, n * sizeof(float), false, true); }
float* CallScope::inout(float* p, size_t n) { return (float*)stage(p, n * sizeof(float), true, true); }

void* CallScope::scratch_bytes(size_t bytes) {
    void* d = pool_alloc(kDevice, bytes ? bytes : 4);
    temps_.push_back(d);
    return d;
}
float* CallScope::scratch(size_t n) { return (float*)scratch_bytes(n * sizeof(float)); }

float* CallScope::new_result(size_t n) { return (float*)pool_alloc(result_kind(), (n ? n : 1) * sizeof(float)); }

void CallScope::finish() {
    if (done_) return;
    done_ = true;
    for (const Back& b : backs_) {
        BLA_CUDA(cudaMemcpyAsync(b.host, b.dev, b.bytes, cudaMemcpyDeviceToHost, g_rt.stream));
        g_rt.d2h_bytes += b.bytes;
    }
    if (any_host_visible_ || force_sync_) BLA_CUDA(cudaStreamSynchronize(g_rt.stream));
    // Single stream: a recycled temp is only ever reused by later work on the same stream.
    for (void* t : temps_) pool_free(t);
    temps_.clear();
    backs_.clear();
}

}  // namespace bla

// ---------------------------------------------------------------------------------------------
// C-ABI: runtime + memory (include/bla.h)
// ---------------------------------------------------------------------------------------------
using namespace bla;

extern "C" {

int bla_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void bla_init(int device) { rt_init(device); }
void bla_sync(void) { BLA_CUDA(cudaStreamSynchronize(rt().stream)); }
void* bla_stream(void) { return (void*)rt().stream; }
// The pool hands a freed block to the next allocation without waiting (stream-ordered reuse: it serves ONE stream), so work in
// flight on the outgoing stream is drained before the library moves to another one.
void bla_set_stream(void* s) {
    cudaStream_t next = s ? (cudaStream_t)s : rt().own_stream;
    if (next != rt().stream && rt().stream) BLA_CUDA(cudaStreamSynchronize(rt().stream));
    rt().stream = next;
}
const char* bla_version(void) { return "bla-b200 0.1 (sm_100a)"; }

void bla_set_gemm_path(int path) {
    if (path < 0 || path > 2) die("bla: unknown GEMM path %d, exiting", path);
    rt().gemm_path = path;
}
int bla_get_gemm_path(void) { return rt().gemm_path; }
void bla_set_quirks(int on) { rt().quirks = on != 0; }
int bla_get_quirks(void) { return rt().quirks; }
unsigned long long bla_launch_count(void) { return g_rt.launches; }
unsigned long long bla_h2d_bytes(void) { return g_rt.h2d_bytes; }
unsigned long long bla_d2h_bytes(void) { return g_rt.d2h_bytes; }

void* bla_malloc_device(size_t bytes) { return pool_alloc(kDevice, bytes); }
void* bla_malloc_pinned(size_t bytes) { return pool_alloc(kPinned, bytes); }
void* bla_malloc_managed(size_t bytes) { return pool_alloc(kManaged, bytes); }
void bla_free(void* p) {
    if (p && !pool_free(p)) die("bla: bla_free of a pointer this library did not allocate (%p), exiting", p);
}
int bla_memory_kind(const void* p) { return (int)classify(p); }

struct Matrix* bla_matrix_device(int rows, int cols) {
    struct Matrix* m = (struct Matrix*)malloc(sizeof(struct Matrix));
    m->rows = rows;
    m->cols = cols;
    m->data = (float*)pool_alloc(kDevice, (size_t)rows * cols * sizeof(float));
    return m;
}

void bla_copy_h2d(void* dst, const void* src, size_t bytes) {
    BLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, rt().stream));
    g_rt.h2d_bytes += bytes;
}
void bla_copy_d2h(void* dst, const void* src, size_t bytes) {
    BLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, rt().stream));
    g_rt.d2h_bytes += bytes;
}
void bla_copy_d2d(void* dst, const void* src, size_t bytes) {
    BLA_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, rt().stream));
}
void bla_memset_zero(void* dst, size_t bytes) { BLA_CUDA(cudaMemsetAsync(dst, 0, bytes, rt().stream)); }

}  // extern "C"
