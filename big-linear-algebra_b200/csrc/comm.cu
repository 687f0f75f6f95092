// comm.cu -- NCCL over NVLink 5 / NVSwitch, one process per GPU (include/bla.h "NCCL" section).
// The reference has no communication layer at all; the only exchanges of
// the B200 path are the flat gradient all-reduce of the data-parallel MLP step and the broadcast of
// B for the row-sharded GEMM.  NCCL is bound at run time with dlopen so that libbla.so has no
// link-time dependency on it: a process that already carries torch's bundled libnccl.so.2 reuses
// that copy, a plain C program picks up the system one.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>

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
