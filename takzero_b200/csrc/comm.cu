// comm.cu -- the multi-GPU exchanges of the self-play path over NCCL (NVLink 5 / NVSwitch inside one node).
//
// The reference scales by running independent processes that share files: every `selfplay` process re-reads
// `model_latest.ot` before each move (selfplay/src/main.rs:107) and `learn` sums the produced positions through
// `buffer_lengths.txt` (learn/src/main.rs:195-209).  Here one process per GPU owns a contiguous range of games and
// those two exchanges are collectives: ncclBroadcast of the ready-to-use 16-bit weight set (nn.cu) and ncclAllReduce
// of the counters.  There is no collective inside a simulation.
//
// libnccl is opened with dlopen (TZ_NCCL_LIB, else "libnccl.so.2"): in a process that already loaded NCCL (PyTorch)
// the same copy is used, and a single-GPU host needs no NCCL at all.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "comm.cuh"

// the slice of nccl.h this file uses (stable NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclUint8 = 1, ncclUint64 = 5 };
enum { ncclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

struct TzComm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

static NcclApi g_nccl;
static thread_local char g_comm_err[256] = "";
const char* comm_last_error() { return g_comm_err; }

static int comm_fail(const char* what, int rc) {
    snprintf(g_comm_err, sizeof(g_comm_err), "%s: %s", what,
             rc != 0 && g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "failed");
    return TZ_ECUDA;
}

static int load_nccl() {
    static std::mutex once;  // handles of different host threads may get here together
    std::lock_guard<std::mutex> lock(once);
    if (g_nccl.lib) return TZ_OK;
    const char* env = getenv("TZ_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* name : names) {
        if (!name || !*name) continue;
        lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        snprintf(g_comm_err, sizeof(g_comm_err), "libnccl.so.2 not found (set TZ_NCCL_LIB): %s", dlerror());
        return TZ_EINVAL;
    }
    NcclApi api;
    api.lib = lib;
#define SYM(field, name)                                                      \
    *(void**)(&api.field) = dlsym(lib, name);                                 \
    if (!api.field) {                                                         \
        snprintf(g_comm_err, sizeof(g_comm_err), "%s missing in libnccl", name); \
        return TZ_EINVAL;                                                     \
    }
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Broadcast, "ncclBroadcast");
    SYM(AllReduce, "ncclAllReduce");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(GetVersion, "ncclGetVersion");
#undef SYM
    g_nccl = api;
    return TZ_OK;
}

int comm_unique_id(void* out128) {
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    const int nrc = g_nccl.GetUniqueId(&id);
    if (nrc != 0) return comm_fail("ncclGetUniqueId", nrc);
    memcpy(out128, id.internal, 128);
    return TZ_OK;
}

int comm_init(tz_handle* h, const void* id128, int nranks, int rank) {
    if (nranks < 1 || rank < 0 || rank >= nranks) {
        snprintf(g_comm_err, sizeof(g_comm_err), "bad rank %d of %d", rank, nranks);
        return TZ_EINVAL;
    }
    comm_destroy(h);
    TzComm* c = new TzComm();
    c->nranks = nranks;
    c->rank = rank;
    if (nranks > 1) {
        int rc = load_nccl();
        if (rc) {
            delete c;
            return rc;
        }
        ncclUniqueId id;
        memcpy(id.internal, id128, 128);
        const int nrc = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
        if (nrc != 0) {
            delete c;
            return comm_fail("ncclCommInitRank", nrc);
        }
    }
    h->comm = c;
    return TZ_OK;
}

void comm_destroy(tz_handle* h) {
    if (!h->comm) return;
    if (h->comm->comm) g_nccl.CommDestroy(h->comm->comm);
    delete h->comm;
    h->comm = nullptr;
}

int comm_nranks(const tz_handle* h) { return h->comm ? h->comm->nranks : 1; }
int comm_rank(const tz_handle* h) { return h->comm ? h->comm->rank : 0; }

int comm_broadcast(tz_handle* h, void* dev_buf, size_t bytes, int root, cudaStream_t st) {
    if (comm_nranks(h) == 1) return TZ_OK;
    const int nrc = g_nccl.Broadcast(dev_buf, dev_buf, bytes, ncclUint8, root, h->comm->comm, st);
    return nrc == 0 ? TZ_OK : comm_fail("ncclBroadcast", nrc);
}

int comm_allreduce_sum_u64(tz_handle* h, unsigned long long* dev_buf, int count, cudaStream_t st) {
    if (comm_nranks(h) == 1) return TZ_OK;
    const int nrc = g_nccl.AllReduce(dev_buf, dev_buf, (size_t)count, ncclUint64, ncclSum, h->comm->comm, st);
    return nrc == 0 ? TZ_OK : comm_fail("ncclAllReduce", nrc);
}
