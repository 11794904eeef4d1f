// Collectives for the row-sharded multi-GPU path (no counterpart in the reference, which is a
// single-process CPU crate).  One process per GPU; NCCL over NVLink 5 / NVSwitch on the context
// stream.  NCCL is resolved with dlopen at run time so that the library binds to whichever
// libnccl.so.2 the host process already loaded (torch's bundled copy in bench.py) instead of
// dragging in a second one; only the long-stable core entry points are used.
#include <dlfcn.h>
#include <cstring>
#include "rc_internal.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclMax = 2 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& api() {
    static NcclApi a;
    if (!a.lib) {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) RC_THROW(RC_NCCL_ERROR, "libnccl.so.2 not found: %s", dlerror());
#define RC_SYM(field, name)                                                             \
    *(void**)(&a.field) = dlsym(a.lib, name);                                           \
    if (!a.field) RC_THROW(RC_NCCL_ERROR, "symbol %s missing in libnccl", name);
        RC_SYM(GetUniqueId, "ncclGetUniqueId")
        RC_SYM(CommInitRank, "ncclCommInitRank")
        RC_SYM(CommDestroy, "ncclCommDestroy")
        RC_SYM(AllReduce, "ncclAllReduce")
        RC_SYM(AllGather, "ncclAllGather")
        RC_SYM(GetErrorString, "ncclGetErrorString")
#undef RC_SYM
    }
    return a;
}

#define RC_NCCL(expr)                                                                         \
    do {                                                                                      \
        ncclResult_t _r = (expr);                                                             \
        if (_r != 0) RC_THROW(RC_NCCL_ERROR, "%s failed: %s", #expr, api().GetErrorString(_r)); \
    } while (0)

}  // namespace

void comm_get_unique_id(void* out128) {
    ncclUniqueId id;
    RC_NCCL(api().GetUniqueId(&id));
    memcpy(out128, &id, 128);
}

void comm_init(rc_ctx* c, const void* id128, int rank, int nranks) {
    RC_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init: bad rank %d / %d", rank, nranks);
    if (nranks == 1) { c->rank = 0; c->nranks = 1; return; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    RC_CUDA(cudaSetDevice(c->device));
    RC_NCCL(api().CommInitRank(&comm, nranks, id, rank));
    c->comm = comm;
    c->rank = rank;
    c->nranks = nranks;
}

void comm_destroy(rc_ctx* c) {
    if (c->comm) api().CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr;
}

void comm_allreduce_sum(rc_ctx* c, void* buf, size_t count, int dtype) {
    if (c->nranks <= 1 || count == 0) return;
    // complex = pairs of reals
    size_t n = count * ((dtype & 2) ? 2 : 1);
    int nt = (dtype & 1) ? ncclFloat64 : ncclFloat32;
    RC_NCCL(api().AllReduce(buf, buf, n, nt, ncclSum, (ncclComm_t)c->comm, c->stream));
}

void comm_allreduce_max_f64(rc_ctx* c, double* buf, size_t count) {
    if (c->nranks <= 1 || count == 0) return;
    RC_NCCL(api().AllReduce(buf, buf, count, ncclFloat64, ncclMax, (ncclComm_t)c->comm, c->stream));
}

void comm_allgather(rc_ctx* c, const void* send, void* recv, size_t bytes_per_rank) {
    if (c->nranks <= 1) {
        if (send != recv) RC_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream));
        return;
    }
    RC_NCCL(api().AllGather(send, recv, bytes_per_rank, ncclInt8, (ncclComm_t)c->comm, c->stream));
}
