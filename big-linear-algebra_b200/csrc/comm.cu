// comm.cu -- NCCL over NVLink 5 / NVSwitch, one process per GPU (include/bla.h "NCCL" section).
// The reference has no communication layer at all; the only exchanges of
// the B200 path are the flat gradient all-reduce of the data-parallel MLP step and the broadcast of
// B for the row-sharded GEMM.  NCCL is bound at run time with dlopen so that libbla.so has no
// link-time dependency on it: a process that already carries torch's bundled libnccl.so.2 reuses
// that copy, a plain C program picks up the system one.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/bla.h"
#include "runtime.h"

using namespace bla;

namespace {

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
} g;

void load() {
    if (g.handle) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g.handle) break;
    }
    if (!g.handle) die("bla: cannot load libnccl.so.2 (%s), exiting", dlerror());
#define SYM(field, name)                                              \
    *(void**)(&g.field) = dlsym(g.handle, name);                      \
    if (!g.field) die("bla: libnccl lacks %s, exiting", name)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Broadcast, "ncclBroadcast");
    SYM(AllGather, "ncclAllGather");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

void ok(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) die("bla: NCCL error in %s: %s, exiting", what, g.GetErrorString ? g.GetErrorString(r) : "?");
}

}  // namespace

namespace bla {
bool comm_active() { return g.comm != nullptr && g.world > 1; }
// A second stream for collectives that overlap with compute on the library stream (ordering by events).
cudaStream_t comm_stream() {
    static cudaStream_t s = nullptr;
    if (!s) BLA_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return s;
}
void comm_allreduce_f32_on(float* buf, size_t n, cudaStream_t s) {
    if (comm_active() && n) ok(g.AllReduce(buf, buf, n, ncclFloat32, ncclSum, g.comm, s), "ncclAllReduce");
}
void comm_allreduce_f64_on(double* buf, size_t n, cudaStream_t s) {
    if (comm_active() && n) ok(g.AllReduce(buf, buf, n, ncclFloat64, ncclSum, g.comm, s), "ncclAllReduce");
}
void comm_group_start() { ok(g.GroupStart(), "ncclGroupStart"); }
void comm_group_end() { ok(g.GroupEnd(), "ncclGroupEnd"); }
}  // namespace bla

// ------------------------------------------------------------------------------------------------------------------------
// All-reduce over NVLink peer memory, ONE kernel per call (default when every rank of the communicator can map every other rank's
// window; BLA_PEER_ALLREDUCE=0 keeps NCCL).  The exchanges of this library are small (the MLP's 0.94 MB gradient, the hinge
// classifier's 31 KB): at that size a collective is all latency, and NCCL's is two kernels' worth per call.
//
// Every rank owns a window of device memory (cudaMalloc, exported with cudaIpcGetMemHandle, the handles all-gathered through the
// NCCL communicator once) that every other rank maps.  Window = flag block + 2 (parity) x world (source rank) data slots.
// Kernel, per CTA c -- a CTA only ever talks to CTA c of the other ranks, there is no grid-wide step:
//   1. epoch = my counter[c] + 1, parity = low bit of the window's call counter (both in device memory: the kernel can be replayed
//      from a CUDA graph with unchanged arguments)
//   2. PUSH my slice c of the source buffer into slot [parity][my rank] of every peer's window (posted NVLink writes)
//   3. fence, then store `epoch` into flag [my rank][c] of every peer
//   4. wait until flag [r][c] >= epoch for every peer r in MY window: their slices have landed in my HBM
//   5. sum the world contributions in RANK ORDER (mine from the source buffer, the peers' from my window, L2-coherent loads):
//      every rank computes the same bits; the sum goes to dst, or straight into the parameters (params += alpha * sum)
// Slot reuse needs no second handshake: slot [p] is rewritten two CALLS later, and a peer cannot be two calls behind -- it posted
// the flags of call k - 1 that this rank waited for, after its kernel of call k - 2 (the last reader of the slot) had finished.  The
// parity is per call, not per CTA, so calls of different sizes (different grids) may follow each other.
// Waits are bounded (~4 s): a rank that never arrives raises the window's error word and the CTA leaves dst / the parameters
// untouched; comm_peer_failed() reports it to the host.
// ------------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxPeers = 8;
constexpr int kPeerCtas = 148;               // upper bound of the grid of one call
constexpr int kPeerThreads = 256;
constexpr size_t kFlagBytes = (size_t)kMaxPeers * kPeerCtas * 8 + 256;   // flags [source rank][cta], then the error word

unsigned peer_generation = 0;
struct PeerState {
    bool tried = false, on = false;
    bool enabled = true;                     // bla_comm_set_peer_windows(0): keep the windows, route the all-reduces through NCCL
    size_t floats = 0;                       // capacity of one data slot, in floats
    char* base[kMaxPeers] = {nullptr};       // every rank's window as mapped here (base[rank] is the local allocation)
    unsigned long long* counters = nullptr;  // [kPeerCtas] last flag value of every CTA index, [kPeerCtas] calls so far, [kPeerCtas + 1] end-of-call ticket
} peer;

struct PeerArgs {
    float* slot[kMaxPeers];                  // slot [parity 0][my rank] in every rank's window (parity 1 is `stride` floats further)
    unsigned long long* flag_out[kMaxPeers]; // flag [my rank][0] in every rank's window
    const unsigned long long* flag_in;       // my window's flags: [source rank][cta]
    const float* data_in;                    // my window's slots [parity][source rank]
    unsigned long long* counters;
    unsigned int* error;
    const float* src;                        // this rank's contribution (private memory)
    float* dst;                              // sum, or parameters when `fused`
    float alpha;
    int fused;
    size_t n4;                               // float4 elements of this call
    size_t offset4;                          // where this call's range starts inside a slot
    size_t slot_floats;                      // capacity of one slot
    int world, rank;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// 3. fence + announce `value` in flag [my rank][c] of every peer, 4. wait for flag [r][c] >= value of every peer r in MY window.
// Returns false after the bounded wait (~4 s) gave up on a peer: the window's error word is raised, the caller leaves its
// destination untouched but still advances its counters so that the ranks stay in step.
__device__ __forceinline__ bool peer_handshake(const PeerArgs& a, int c, int t, unsigned long long value, int* failed) {
    if (t == 0) *failed = 0;
    // The CTA barrier orders every thread's pushes before the flag writers' st.release.sys, and a release is cumulative over what
    // happens-before it: ONE system-scope release per peer and CTA.  (Round 2's first version had all 256 threads of all CTAs execute
    // __threadfence_system() -- twice -- before the flags: 28 us per call on 2 GPUs, most of it those ~60,000 MEMBAR.SYS.)
    __syncthreads();
    if (t < a.world && t != a.rank) {
        st_release_sys(a.flag_out[t] + c, value);
        const unsigned long long* f = a.flag_in + (size_t)t * kPeerCtas + c;
        long spins = 0;
        while (ld_acquire_sys(f) < value) {
            if (++spins > 2000) __nanosleep(200);
            if (spins > 20000000L) { atomicExch(a.error, 1u); *failed = 1; break; }
        }
    }
    __syncthreads();
    return *failed == 0;
}
// End of a call, thread 0 of every CTA: the CTA's flag counter, and -- by the last CTA to get here -- the call counter whose low bit
// is the slot parity of the NEXT call (read by every CTA of that call at its start: the kernels of one stream do not overlap).
__device__ __forceinline__ void peer_finish(const PeerArgs& a, int c, unsigned long long flag_value, unsigned long long call) {
    a.counters[c] = flag_value;
    __threadfence();
    unsigned long long* calls = a.counters + kPeerCtas;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(a.counters + kPeerCtas + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        *ticket = 0u;
        *calls = call + 1;
    }
}
__device__ __forceinline__ void peer_apply(const PeerArgs& a, float4* dst4, size_t i, const float4& acc) {
    if (a.fused) {
        float4 p = dst4[i];
        p.x = fmaf(a.alpha, acc.x, p.x); p.y = fmaf(a.alpha, acc.y, p.y); p.z = fmaf(a.alpha, acc.z, p.z); p.w = fmaf(a.alpha, acc.w, p.w);
        dst4[i] = p;
    } else {
        dst4[i] = acc;
    }
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs a) {
    const int c = blockIdx.x, t = threadIdx.x;
    pdl_trigger();
    pdl_wait();   // runtime.h: launched with programmatic serialisation
    const unsigned long long epoch = a.counters[c] + 1;
    const unsigned long long call = *reinterpret_cast<volatile unsigned long long*>(a.counters + kPeerCtas);
    const size_t par_off = (call & 1ull) * (size_t)a.world * a.slot_floats;
    const size_t per = (a.n4 + gridDim.x - 1) / gridDim.x;
    const size_t beg = (size_t)c * per, end = beg + per < a.n4 ? beg + per : a.n4;
    const float4* src4 = reinterpret_cast<const float4*>(a.src);
    // 2. push
    for (size_t i = beg + t; i < end; i += kPeerThreads) {
        const float4 v = src4[i];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < a.world && r != a.rank) reinterpret_cast<float4*>(a.slot[r] + par_off)[a.offset4 + i] = v;
    }
    __shared__ int failed;
    if (peer_handshake(a, c, t, epoch, &failed)) {
        // 5. sum in rank order
        const float4* in4 = reinterpret_cast<const float4*>(a.data_in + par_off);
        const size_t slot4 = a.slot_floats / 4;
        float4* dst4 = reinterpret_cast<float4*>(a.dst);
        for (size_t i = beg + t; i < end; i += kPeerThreads) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {
                if (r < a.world) {
                    const float4 v = r == a.rank ? src4[i] : __ldcg(in4 + (size_t)r * slot4 + a.offset4 + i);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
            peer_apply(a, dst4, i, acc);
        }
    }
    if (t == 0) peer_finish(a, c, epoch, call);   // the flags of this call are out: also after a timeout, so that the ranks stay in step
}

// The same exchange in TWO rounds for larger buffers on 8 ranks: reduce-scatter, then all-gather (at 4 ranks the saved bytes do not pay
// for the second flag round: measured alone 29.9 us in two rounds on 4 GPUs against 18.0 us in one round on 2 and 31.4 / 28.8 us on 8).  The one-round kernel moves
// (world - 1) x n floats out of and into every GPU (6.6 MB for the MLP's gradient on 8 ranks: 34 us measured, most of it transfer);
// here the buffer is cut into `world` owner segments, every rank sends each owner only that owner's segment, the owner adds the
// world contributions in rank order and sends the SUM back to everybody: 2 x (world - 1) / world x n floats per GPU, a quarter of
// the bytes at 8 ranks, for one more flag round.  Every rank applies the owner's bits: identical results on all ranks.
// CTA c handles piece c of EVERY segment, so as before it only ever talks to CTA c of the other ranks.
__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_two_round_kernel(PeerArgs a) {
    const int c = blockIdx.x, t = threadIdx.x, world = a.world, me = a.rank;
    pdl_trigger();
    pdl_wait();
    const unsigned long long f1 = a.counters[c] + 1, f2 = f1 + 1;
    const unsigned long long call = *reinterpret_cast<volatile unsigned long long*>(a.counters + kPeerCtas);
    const size_t par_off = (call & 1ull) * (size_t)world * a.slot_floats;
    const size_t seg4 = (a.n4 + world - 1) / world;                       // float4 per owner segment
    const size_t per = (seg4 + gridDim.x - 1) / gridDim.x;
    const size_t qb = (size_t)c * per, qe = qb + per < seg4 ? qb + per : seg4;
    const float4* src4 = reinterpret_cast<const float4*>(a.src);
    const float4* in4 = reinterpret_cast<const float4*>(a.data_in + par_off);
    const size_t slot4 = a.slot_floats / 4;
    float4* dst4 = reinterpret_cast<float4*>(a.dst);
    // round 1: my values of owner r's segment -> slot [me] of r's window, at the element's own position
    for (size_t q = qb + t; q < qe; q += kPeerThreads) {
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
            const size_t g = (size_t)r * seg4 + q;
            if (r < world && r != me && g < a.n4) reinterpret_cast<float4*>(a.slot[r] + par_off)[a.offset4 + g] = src4[g];
        }
    }
    __shared__ int failed[2];   // one word per round: the second round's reset must not race with late readers of the first
    bool ok = peer_handshake(a, c, t, f1, &failed[0]);
    if (ok) {
        // my segment: rank-ordered sum, applied here and sent to everybody (slot [me] of their windows, my segment's positions)
        for (size_t q = qb + t; q < qe; q += kPeerThreads) {
            const size_t g = (size_t)me * seg4 + q;
            if (g >= a.n4) break;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {
                if (r < world) {
                    const float4 v = r == me ? src4[g] : __ldcg(in4 + (size_t)r * slot4 + a.offset4 + g);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r)
                if (r < world && r != me) reinterpret_cast<float4*>(a.slot[r] + par_off)[a.offset4 + g] = acc;
            peer_apply(a, dst4, g, acc);
        }
    }
    ok = peer_handshake(a, c, t, f2, &failed[1]) && ok;
    if (ok) {
        // round 2: the other owners' sums have landed in my window
        for (size_t q = qb + t; q < qe; q += kPeerThreads) {
#pragma unroll
            for (int r = 0; r < kMaxPeers; ++r) {
                const size_t g = (size_t)r * seg4 + q;
                if (r < world && r != me && g < a.n4) peer_apply(a, dst4, g, __ldcg(in4 + (size_t)r * slot4 + a.offset4 + g));
            }
        }
    }
    if (t == 0) peer_finish(a, c, f2, call);
}

void peer_release() {
    ++peer_generation;
    for (int r = 0; r < kMaxPeers; ++r) {
        if (!peer.base[r]) continue;
        if (r == g.rank) cudaFree(peer.base[r]);
        else cudaIpcCloseMemHandle(peer.base[r]);
        peer.base[r] = nullptr;
    }
    if (peer.counters) { cudaFree(peer.counters); peer.counters = nullptr; }
    cudaGetLastError();
    peer.on = false;
    peer.floats = 0;
}

// collective: every rank calls it with the same `floats`; returns whether ALL ranks mapped all windows
bool peer_setup(size_t floats) {
    peer.tried = true;
    const char* e = getenv("BLA_PEER_ALLREDUCE");
    if ((e && atoi(e) == 0) || !comm_active() || g.world > kMaxPeers) return false;
    cudaStream_t s = rt().stream;
    BLA_CUDA(cudaStreamSynchronize(s));
    peer_release();
    floats = (floats + 3) / 4 * 4;
    const size_t bytes = kFlagBytes + 2 * (size_t)g.world * floats * sizeof(float);
    int good = 1;
    char* mine = nullptr;
    cudaIpcMemHandle_t handle;
    memset(&handle, 0, sizeof(handle));
    if (cudaMalloc((void**)&mine, bytes) != cudaSuccess || cudaMemset(mine, 0, kFlagBytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&handle, mine) != cudaSuccess || cudaMalloc((void**)&peer.counters, (kPeerCtas + 2) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(peer.counters, 0, (kPeerCtas + 2) * sizeof(unsigned long long)) != cudaSuccess) {
        cudaGetLastError();
        good = 0;
    }
    BLA_CUDA(cudaDeviceSynchronize());
    // all-gather the handles (and every rank's verdict so far) through the communicator that is already up
    struct Slot { cudaIpcMemHandle_t handle; int good; int pad[15]; };
    static_assert(sizeof(Slot) % 4 == 0, "slots travel as ints");
    Slot my{};
    my.handle = handle;
    my.good = good;
    Slot* all_dev = (Slot*)pool_alloc(kDevice, sizeof(Slot) * g.world);
    BLA_CUDA(cudaMemcpyAsync(all_dev + g.rank, &my, sizeof(Slot), cudaMemcpyHostToDevice, s));
    ok(g.AllGather(all_dev + g.rank, all_dev, sizeof(Slot) / 4, ncclInt32, g.comm, s), "ncclAllGather");
    std::vector<Slot> all(g.world);
    BLA_CUDA(cudaMemcpyAsync(all.data(), all_dev, sizeof(Slot) * g.world, cudaMemcpyDeviceToHost, s));
    BLA_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < g.world; ++r) good &= all[r].good;
    peer.base[g.rank] = mine;
    if (good) {
        for (int r = 0; r < g.world && good; ++r) {
            if (r == g.rank) continue;
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                good = 0;
            }
            peer.base[r] = (char*)p;
        }
    }
    // a second round so that every rank takes the same path
    int* verdict = (int*)all_dev;
    BLA_CUDA(cudaMemcpyAsync(verdict + g.rank, &good, sizeof(int), cudaMemcpyHostToDevice, s));
    ok(g.AllGather(verdict + g.rank, verdict, 1, ncclInt32, g.comm, s), "ncclAllGather");
    std::vector<int> votes(g.world);
    BLA_CUDA(cudaMemcpyAsync(votes.data(), verdict, sizeof(int) * g.world, cudaMemcpyDeviceToHost, s));
    BLA_CUDA(cudaStreamSynchronize(s));
    pool_free(all_dev);
    for (int v : votes) good &= v;
    if (!good) {
        if (g.rank == 0) fprintf(stderr, "bla: peer windows unavailable (IPC / peer access), the all-reduce stays on NCCL\n");
        peer_release();
        return false;
    }
    peer.floats = floats;
    peer.on = true;
    return true;
}

}  // namespace

namespace bla {
// Can an all-reduce of up to `floats` floats run over the peer windows?  Collective whenever the windows have to be (re)built: every
// rank must ask with the same sizes in the same order.  Must not be called while a stream is being captured.
bool comm_peer_ready(size_t floats) {
    if (!peer.enabled) return false;
    if (!peer.tried || (peer.on && floats > peer.floats)) peer_setup(floats);
    return peer.on && floats <= peer.floats;
}
// dst[0..n) = sum over ranks of src[0..n)            (fused == false)
// dst[0..n) += alpha * sum over ranks of src[0..n)   (fused == true: dst are the parameters)
// `offset` (floats, a multiple of 4) places the call's range inside the slots: concurrent calls (two streams) must use disjoint
// ranges.  n a multiple of 4, src / dst 16-byte aligned.  Every rank makes the same calls in the same order.
void comm_peer_allreduce_f32(const float* src, float* dst, float alpha, bool fused, size_t offset, size_t n, cudaStream_t s) {
    if (!n) return;
    if ((n & 3) || (offset & 3) || offset + n > peer.floats || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15))
        die("bla: peer all-reduce of %zu floats at offset %zu does not fit the window contract, exiting", n, offset);
    PeerArgs a{};
    for (int r = 0; r < g.world; ++r) {
        a.slot[r] = (float*)(peer.base[r] + kFlagBytes) + (size_t)g.rank * peer.floats;
        a.flag_out[r] = (unsigned long long*)peer.base[r] + (size_t)g.rank * kPeerCtas;
    }
    a.flag_in = (const unsigned long long*)peer.base[g.rank];
    a.data_in = (const float*)(peer.base[g.rank] + kFlagBytes);
    a.counters = peer.counters;
    a.error = (unsigned int*)(peer.base[g.rank] + kFlagBytes - 64);
    a.src = src; a.dst = dst; a.alpha = alpha; a.fused = fused ? 1 : 0;
    a.n4 = n / 4; a.offset4 = offset / 4; a.slot_floats = peer.floats;
    a.world = g.world; a.rank = g.rank;
    // kernel and grid depend on (world, n, BLA_PEER_TWO_ROUNDS) only: the same on every rank
    static int two_env = -2;
    if (two_env == -2) { const char* e = getenv("BLA_PEER_TWO_ROUNDS"); two_env = e ? atoi(e) : -1; }
    const bool two = two_env >= 0 ? (two_env != 0 && g.world >= 2) : (g.world >= 8 && n >= 65536);
    if (two) {
        const size_t seg4 = (a.n4 + g.world - 1) / g.world;
        int ctas = (int)std::min<size_t>(kPeerCtas, (seg4 + 127) / 128);
        if (ctas < 1) ctas = 1;
        BLA_CUDA(launch_pdl(peer_allreduce_two_round_kernel, dim3(ctas), dim3(kPeerThreads), 0, s, 1, a));
    } else {
        int ctas = (int)std::min<size_t>(kPeerCtas, (a.n4 + 511) / 512);   // ~512 float4 per CTA
        if (ctas < 1) ctas = 1;
        BLA_CUDA(launch_pdl(peer_allreduce_kernel, dim3(ctas), dim3(kPeerThreads), 0, s, 1, a));
    }
    count_launch();
}
// true if a peer never arrived in some all-reduce since the last check (that call left its destination untouched); clears the word
bool comm_peer_failed() {
    if (!peer.on) return false;
    unsigned int e = 0;
    BLA_CUDA(cudaMemcpy(&e, peer.base[g.rank] + kFlagBytes - 64, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) BLA_CUDA(cudaMemset(peer.base[g.rank] + kFlagBytes - 64, 0, sizeof(e)));
    return e != 0;
}
bool comm_peer_on() { return peer.on && peer.enabled; }
// changes whenever addresses a captured graph may have baked in (the windows) are no longer valid
unsigned comm_generation() { return peer_generation; }
}  // namespace bla

extern "C" {

void bla_comm_unique_id(void* id128) {
    load();
    ncclUniqueId id;
    ok(g.GetUniqueId(&id), "ncclGetUniqueId");
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id128, &id, 128);
}

void bla_comm_init(const void* id128, int rank, int world) {
    load();
    rt();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ok(g.CommInitRank(&g.comm, world, id, rank), "ncclCommInitRank");
    g.rank = rank;
    g.world = world;
}

int bla_comm_world(void) { return g.world; }
// 1 once the gradient all-reduces of this process run over the NVLink peer windows (decided at the first data-parallel step), else 0 (NCCL)
int bla_comm_peer_windows(void) { return peer.on && peer.enabled ? 1 : 0; }
// Collective (every rank, same value, same point of its call sequence): 0 routes the all-reduces through NCCL, 1 back over the windows.
void bla_comm_set_peer_windows(int on) {
    if (rt_initialised()) BLA_CUDA(cudaDeviceSynchronize());
    peer.enabled = on != 0;
    ++peer_generation;   // step graphs captured under the other setting are stale
}
int bla_comm_rank(void) { return g.rank; }

// Small buffers (<= 4 MB, whole float4s) go over the NVLink peer windows as one kernel when the windows are up; everything else,
// and everything when they are not, through NCCL.  Collective either way.
void bla_allreduce_sum_f32(float* buf, size_t n) {
    if (!comm_active() || !n) return;
    if (n <= (1u << 20) && (n & 3) == 0 && ((uintptr_t)buf & 15) == 0 && comm_peer_ready(n)) {
        comm_peer_allreduce_f32(buf, buf, 0.f, false, 0, n, rt().stream);
        return;
    }
    ok(g.AllReduce(buf, buf, n, ncclFloat32, ncclSum, g.comm, rt().stream), "ncclAllReduce");
}
void bla_allreduce_sum_f64(double* buf, size_t n) {
    if (!comm_active() || !n) return;
    ok(g.AllReduce(buf, buf, n, ncclFloat64, ncclSum, g.comm, rt().stream), "ncclAllReduce");
}
void bla_broadcast_f32(float* buf, size_t n, int root) {
    if (!comm_active() || !n) return;
    ok(g.Broadcast(buf, buf, n, ncclFloat32, root, g.comm, rt().stream), "ncclBroadcast");
}
void bla_comm_destroy(void) {
    if (g.comm) {
        BLA_CUDA(cudaDeviceSynchronize());
        peer_release();
        peer.tried = false;
        g.CommDestroy(g.comm);
        g.comm = nullptr;
        g.world = 1;
        g.rank = 0;
    }
}

}  // extern "C"
