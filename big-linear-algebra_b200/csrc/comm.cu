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
// Peer window (opt-in, BLA_PEER_ALLREDUCE=1): the small gradient all-reduce as ONE kernel over NVLink peer memory.
//
// Every rank owns a window of device memory (plain cudaMalloc, exported with cudaIpcGetMemHandle, the handles all-gathered
// through NCCL once) that every other rank of the node maps.  A rank's producers write its contribution into its own window;
// the all-reduce kernel then (1) stores the call's epoch number into its flag slot in every peer's window, (2) waits until
// every peer's epoch has arrived in its own, (3) reads the same slice of all windows over NVLink, adds them in rank order
// (so every rank gets the same bits) and writes the sum to a PRIVATE destination.  One launch, no ring, no staging copy:
// at 1.9 MB (the MLP's gradient) the cost is the NVLink read of (world - 1) x 1.9 MB plus one flag round trip.
// The sources are double-buffered by the caller (a window half is rewritten two steps later, and a peer cannot be two
// all-reduces behind: it would not have posted the epoch this rank waited for in between), so one flag round per call is
// enough.  Waits are bounded: a peer that never arrives sets an error word instead of hanging the GPU.
// Status: written in round 1 after the GPU budget was spent -- not yet run on hardware, hence off by default.
// ------------------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kMaxPeers = 8;
constexpr size_t kFlagBytes = 4096;          // the head of every window: one 64-byte line per writing rank + the error word

struct PeerState {
    bool tried = false, on = false;
    size_t floats = 0;                       // payload capacity of one window, in floats
    char* base[kMaxPeers] = {nullptr};       // every rank's window as mapped here (base[rank] is the local allocation)
    unsigned long long epoch = 0;
} peer;

struct PeerArgs {
    const float* src[kMaxPeers];             // the same slice in every rank's window
    unsigned long long* flag_out[kMaxPeers]; // this rank's slot in every peer's flag block
    const unsigned long long* flag_in;       // the local flag block (slot r is written by rank r)
    unsigned int* error;
    float* dst;
    size_t n;
    unsigned long long epoch;
    int world, rank;
};

__global__ void __launch_bounds__(512) peer_allreduce_kernel(PeerArgs a) {
    // (1) announce: everything this rank's earlier kernels wrote is visible system-wide before the flag is
    if (blockIdx.x == 0 && threadIdx.x < a.world) {
        __threadfence_system();
        *(volatile unsigned long long*)a.flag_out[threadIdx.x] = a.epoch;
    }
    // (2) every CTA waits for every rank's announcement of this (or a later) call
    if (threadIdx.x < a.world) {
        const volatile unsigned long long* f = a.flag_in + threadIdx.x * 8;
        long spins = 0;
        while (*f < a.epoch) {
            __nanosleep(200);
            if (++spins > 10000000L) { atomicExch(a.error, 1u); break; }   // ~2 s: report, do not hang the device
        }
        __threadfence_system();
    }
    __syncthreads();
    // (3) sum the windows in rank order
    const size_t n4 = a.n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r) {
            if (r < a.world) {
                const float4 v = __ldcv(reinterpret_cast<const float4*>(a.src[r]) + i);   // never from a stale cache line
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<float4*>(a.dst)[i] = acc;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < a.world; ++r) acc += __ldcv(a.src[r] + i);
        a.dst[i] = acc;
    }
}

// collective: every rank calls it with the same `floats`; returns whether ALL ranks mapped all windows
bool peer_setup(size_t floats) {
    peer.tried = true;
    const char* e = getenv("BLA_PEER_ALLREDUCE");
    if (!e || atoi(e) == 0 || !comm_active() || g.world > kMaxPeers) return false;
    cudaStream_t s = rt().stream;
    const size_t bytes = kFlagBytes + floats * sizeof(float);
    int good = 1;
    char* mine = nullptr;
    cudaIpcMemHandle_t handle;
    memset(&handle, 0, sizeof(handle));
    if (cudaMalloc((void**)&mine, bytes) != cudaSuccess || cudaMemset(mine, 0, bytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&handle, mine) != cudaSuccess) {
        cudaGetLastError();
        good = 0;
    }
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
    if (good) {
        for (int r = 0; r < g.world && good; ++r) {
            if (r == g.rank) { peer.base[r] = mine; continue; }
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
        return false;
    }
    peer.floats = floats;
    peer.on = true;
    return true;
}

}  // namespace

namespace bla {
// The local window's payload (room for `floats` floats), or nullptr when peer windows are off / unavailable.  Collective on
// the first call; the size of that first call is the capacity.
float* comm_peer_window(size_t floats) {
    if (!peer.tried) peer_setup(floats);
    if (!peer.on || floats > peer.floats) return nullptr;
    return (float*)(peer.base[g.rank] + kFlagBytes);
}
// dst[0..n) = sum over ranks of window[offset .. offset + n) -- `offset` in floats from the payload start, 16-byte aligned.
// Every rank makes the same calls in the same order on stream s.
void comm_peer_allreduce_f32(float* dst, size_t offset, size_t n, cudaStream_t s) {
    if (!n) return;
    PeerArgs a{};
    for (int r = 0; r < g.world; ++r) {
        a.src[r] = (const float*)(peer.base[r] + kFlagBytes) + offset;
        a.flag_out[r] = (unsigned long long*)(peer.base[r] + (size_t)g.rank * 64);
    }
    a.flag_in = (const unsigned long long*)peer.base[g.rank];
    a.error = (unsigned int*)(peer.base[g.rank] + kFlagBytes - 64);
    a.dst = dst;
    a.n = n;
    a.epoch = ++peer.epoch;
    a.world = g.world;
    a.rank = g.rank;
    const int ctas = (int)std::min<size_t>(64, (n / 4 + 511) / 512 + 1);
    peer_allreduce_kernel<<<ctas, 512, 0, s>>>(a);
    BLA_LAUNCH_CHECK();
    count_launch();
}
// true if a peer never arrived in some all-reduce since the last check (the results of that call are garbage)
bool comm_peer_failed() {
    if (!peer.on) return false;
    unsigned int e = 0;
    BLA_CUDA(cudaMemcpy(&e, peer.base[g.rank] + kFlagBytes - 64, sizeof(e), cudaMemcpyDeviceToHost));
    return e != 0;
}
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
int bla_comm_rank(void) { return g.rank; }

void bla_allreduce_sum_f32(float* buf, size_t n) {
    if (!comm_active() || !n) return;
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
        BLA_CUDA(cudaStreamSynchronize(rt().stream));
        g.CommDestroy(g.comm);
        g.comm = nullptr;
        g.world = 1;
        g.rank = 0;
    }
}

}  // extern "C"
