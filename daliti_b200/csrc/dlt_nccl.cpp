// daliti_b200/csrc/dlt_nccl.cpp -- include/daliti_b200_nccl.h: ncclAllReduce as the reduce callback of a sharded map.
// libnccl.so.2 is resolved with dlopen so that the product library carries no link-time dependency on NCCL; only the five
// entry points used here are declared (by their public signatures).
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>

#include "../../include/daliti_b200_nccl.h"

#if !defined(DLT_EMU)
#include <cuda_runtime.h>
#endif

extern "C" void *dlt_stream(dlt_handle h);  // dlt_api.cu: the stream the handle launches on

namespace {
struct NcclId {
    char internal[DLT_NCCL_ID_BYTES];
};
typedef struct ncclComm *nccl_comm_t;
typedef int (*fn_get_unique_id)(NcclId *);
typedef int (*fn_comm_init_rank)(nccl_comm_t *, int, NcclId, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_all_reduce)(const void *, void *, size_t, int /* ncclDataType_t */, int /* ncclRedOp_t */, nccl_comm_t, void * /* cudaStream_t */);
typedef const char *(*fn_get_error_string)(int);
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;

struct Api {
    void *lib = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_get_error_string get_error_string = nullptr;
    bool ok = false;
};
Api g_api;
std::once_flag g_once;
thread_local std::string g_err;

void load_api() {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        g_api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) return;
    g_api.get_unique_id = (fn_get_unique_id)dlsym(g_api.lib, "ncclGetUniqueId");
    g_api.comm_init_rank = (fn_comm_init_rank)dlsym(g_api.lib, "ncclCommInitRank");
    g_api.comm_destroy = (fn_comm_destroy)dlsym(g_api.lib, "ncclCommDestroy");
    g_api.all_reduce = (fn_all_reduce)dlsym(g_api.lib, "ncclAllReduce");
    g_api.get_error_string = (fn_get_error_string)dlsym(g_api.lib, "ncclGetErrorString");
    g_api.ok = g_api.get_unique_id && g_api.comm_init_rank && g_api.comm_destroy && g_api.all_reduce;
}
bool api() {
    std::call_once(g_once, load_api);
    if (!g_api.ok) g_err = "libnccl.so.2 not found (dlopen): install NCCL or use the peer-mailbox exchange (dlt_lio_peer_attach)";
    return g_api.ok;
}
int fail(int rc, const char *what) {
    g_err = std::string(what) + ": " + ((g_api.get_error_string && rc) ? g_api.get_error_string(rc) : "failed");
    return DLT_E_STATE;
}
}  // namespace

struct dlt_nccl_comm_s {
    nccl_comm_t comm = nullptr;
    void *stream = nullptr;  // cudaStream_t the all-reduce is enqueued on
    dlt_handle handle = nullptr;  // when bound to a handle the stream is looked up per call (dlt_set_stream may change it)
    int rank = 0, world = 1, device = 0;
};

extern "C" {

int dlt_nccl_available(void) { return api() ? 1 : 0; }
const char *dlt_nccl_last_error(void) { return g_err.c_str(); }

int dlt_nccl_unique_id(unsigned char *id) {
    if (!id) return DLT_E_INVALID;
    if (!api()) return DLT_E_STATE;
    NcclId u;
    std::memset(&u, 0, sizeof(u));
    if (int rc = g_api.get_unique_id(&u)) return fail(rc, "ncclGetUniqueId");
    std::memcpy(id, &u, sizeof(u));
    return DLT_OK;
}

int dlt_nccl_create(const unsigned char *id, int rank, int world, int device, dlt_nccl_comm *out) {
    if (!id || !out || world < 1 || rank < 0 || rank >= world) return DLT_E_INVALID;
    *out = nullptr;
    if (!api()) return DLT_E_STATE;
#if !defined(DLT_EMU)
    if (cudaSetDevice(device) != cudaSuccess) {
        g_err = "cudaSetDevice failed";
        return DLT_E_NO_DEVICE;
    }
#endif
    NcclId u;
    std::memcpy(&u, id, sizeof(u));
    dlt_nccl_comm c = new dlt_nccl_comm_s();
    c->rank = rank;
    c->world = world;
    c->device = device;
    if (int rc = g_api.comm_init_rank(&c->comm, world, u, rank)) {
        delete c;
        return fail(rc, "ncclCommInitRank");
    }
    *out = c;
    return DLT_OK;
}

int dlt_nccl_destroy(dlt_nccl_comm c) {
    if (!c) return DLT_E_INVALID;
    if (c->comm && g_api.comm_destroy) g_api.comm_destroy(c->comm);
    delete c;
    return DLT_OK;
}

int dlt_nccl_allreduce(void *ctx, double *buf_dev, int n) {
    dlt_nccl_comm c = static_cast<dlt_nccl_comm>(ctx);
    if (!c || !buf_dev || n < 0) return DLT_E_INVALID;
    if (n == 0 || c->world == 1) return DLT_OK;
    void *stream = c->handle ? dlt_stream(c->handle) : c->stream;
    if (int rc = g_api.all_reduce(buf_dev, buf_dev, (size_t)n, kNcclFloat64, kNcclSum, c->comm, stream)) return fail(rc, "ncclAllReduce");
    return DLT_OK;
}

int dlt_nccl_attach(dlt_nccl_comm c, dlt_lio lio) {
    if (!c || !lio) return DLT_E_INVALID;
    c->handle = dlt_lio_device(lio);
    return dlt_lio_set_reduce(lio, dlt_nccl_allreduce, c, nullptr);
}

int dlt_nccl_attach_handle(dlt_nccl_comm c, dlt_handle h) {
    if (!c || !h) return DLT_E_INVALID;
    c->handle = h;
    return dlt_set_shard_reduce(h, dlt_nccl_allreduce, c);
}

}  // extern "C"
