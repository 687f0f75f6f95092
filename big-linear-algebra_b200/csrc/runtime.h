// runtime.h -- host runtime of libbla.so: device binding, the single library stream, the
// pointer registry + pooled allocators, and the per-call staging scope that lets the reference's
// pointer-only C API (struct Matrix {rows, cols, data}, passed by value, lib/matrix.h:6-11) run on
// HBM-resident data whatever kind of memory the caller hands in.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace bla {

enum MemKind : int { kHost = 0, kManaged = 1, kDevice = 2, kPinned = 3 };

// One line on stdout + exit(1): the reference's error convention (lib/matrix.c:36-39).
[[noreturn]] void die(const char* fmt, ...);
void check(cudaError_t e, const char* what, const char* file, int line);
#define BLA_CUDA(x) ::bla::check((x), #x, __FILE__, __LINE__)
#define BLA_LAUNCH_CHECK() ::bla::check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__)

struct Runtime {
    int device = -1;
    int num_sms = 148;
    int cc_major = 0, cc_minor = 0;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    int gemm_path = 2;  // BLA_GEMM_AUTO
    int quirks = 1;
    unsigned long long launches = 0, h2d_bytes = 0, d2h_bytes = 0;
};

// Lazily binds the device on first use; exits loudly when no CUDA device is usable.
Runtime& rt();
bool rt_initialised();
void rt_init(int device);
inline void count_launch(int n = 1) { rt().launches += n; }

// ---- pooled allocations, all recorded in the registry ----------------------------------------
void* pool_alloc(MemKind kind, size_t bytes);
// Returns true if p was one of ours (and recycles it); false for foreign pointers.
bool pool_free(void* p);
// Kind of an arbitrary pointer: registry hit (interior pointers included), else the driver's view.
MemKind classify(const void* p);
// Bytes readable from p to the end of the library (pool) block that contains it; 0 for pointers the library did not allocate.
size_t owned_bytes_from(const void* p);

// ---- per-call staging scope --------------------------------------------------------------------
// Every reference-API entry point opens a CallScope, asks it for device-usable views of its
// operands, launches on scope.stream(), and closes it.  Host operands are staged through a
// device arena (H2D before, D2H after); managed and device operands are used in place.  The call
// synchronises at the end iff some operand is host-visible -- a pure-device call stays async.
class CallScope {
public:
    CallScope();
    ~CallScope();
    cudaStream_t stream() const;
    // read-only operand
    const float* in(const float* p, size_t n);
    // write-only result in caller storage
    float* out(float* p, size_t n);
    // read-modify-write operand
    float* inout(float* p, size_t n);
    // device scratch that lives until the scope closes
    float* scratch(size_t n);
    void* scratch_bytes(size_t bytes);
    // kind to use for a result the library allocates for this call
    MemKind result_kind() const { return any_host_visible_ ? kManaged : kDevice; }
    void note_kind(MemKind k);
    // allocate a library-owned result buffer of result_kind()
    float* new_result(size_t n);
    // finish explicitly (copy-backs + sync); idempotent.  The destructor calls it.
    void finish();
    // force a synchronise at finish (scalar-returning calls)
    void need_sync() { force_sync_ = true; }

private:
    struct Back { void* host; const void* dev; size_t bytes; };
    std::vector<Back> backs_;
    std::vector<void*> temps_;
    bool any_host_visible_ = false;
    bool force_sync_ = false;
    bool done_ = false;
    void* stage(const void* p, size_t bytes, bool copy_in, bool copy_back);
};

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ---------------------------------------------------------------
// A kernel whose every thread passes pdl_wait() before its first access to global memory may be launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs then become resident -- and run their prologue (barrier
// initialisation, TMEM allocation, descriptor prefetch, index arithmetic) -- while the previous kernel of the stream is still
// draining, and block in pdl_wait() until that kernel has completed and its writes are visible.  pdl_trigger() in the EARLIER kernel
// only says "my dependents may be made resident once all my CTAs are"; it orders nothing.  A kernel launched with the attribute that
// never waits would race with its predecessor, so the attribute is given per launch site (launch_pdl), never globally.
// BLA_PDL=0 launches everything fully serialised (A/B).  PdlOff suspends it for a scope: CTAs that are resident but still waiting hold their
// SM's shared memory, which starves kernels of OTHER streams that could have run there -- measured on the MLP step (bias sums and the
// layer-2 weight gradient run beside the main chain): 338.7 us without, 374.5 us with programmatic launches at 60,000 columns, 97.9 /
// 102.7 us at 7,500; the single-stream U-Net step gains 4 % (12.42 -> 11.91 ms).  profiles/r02_ab.txt
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
struct PdlOff { PdlOff(); ~PdlOff(); };
// Inside a PdlLate scope kernels that take the hint (the tensor GEMM, its split-K fold, the MLP head kernel) do NOT trigger their
// dependents early: the next kernel is only pre-staged and becomes resident when this one's CTAs have exited -- the launch gap is
// hidden, but no waiting CTA holds an SM against the kernels of the step's other streams.
bool pdl_early();
struct PdlLate { PdlLate(); ~PdlLate(); };
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (cluster_x > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster_x;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = (unsigned)n;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace bla
