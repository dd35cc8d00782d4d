// Ghost-row exchange between row slabs: ncclSend/ncclRecv over NVLink 5 / NVSwitch, one group per
// exchange, on the simulation's comm stream.
//
// libnccl is resolved with dlopen on first use so that single-GPU users carry no NCCL dependency and so
// that, inside a process that already loaded a libnccl.so.2 (e.g. torch's), the same copy is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "wsb_internal.h"

namespace wsb {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};

static NcclApi g_api;
static bool g_loaded = false;
static std::string g_load_error;
static std::mutex g_mu;

int nccl_load(const NcclApi **api) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_loaded && g_load_error.empty()) {
        void *h = nullptr;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            g_load_error = std::string("cannot load libnccl.so.2: ") + dlerror();
        } else {
            bool ok = true;
            auto sym = [&](const char *name) -> void * {
                void *p = dlsym(h, name);
                if (!p) {
                    ok = false;
                    g_load_error = std::string("libnccl lacks symbol ") + name;
                }
                return p;
            };
            g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
            g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
            g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
            g_api.Send = (decltype(g_api.Send))sym("ncclSend");
            g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
            g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
            g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
            g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
            g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
            g_loaded = ok;
        }
    }
    if (!g_loaded) return fail(WSB_ERR_NCCL, g_load_error);
    if (api) *api = &g_api;
    return WSB_OK;
}

static int nccl_fail(const NcclApi *api, ncclResult_t r, const char *what) {
    return fail(WSB_ERR_NCCL, std::string("NCCL error: ") + api->GetErrorString(r) + " in " + what);
}

#define WSB_NCCL(api, call)                                      \
    do {                                                         \
        ncclResult_t _r = (call);                                \
        if (_r != ncclSuccess) return nccl_fail(api, _r, #call); \
    } while (0)

int nccl_get_unique_id(void *out128) {
    if (!out128) return fail(WSB_ERR_INVALID_ARGUMENT, "out128 is NULL");
    const NcclApi *api = nullptr;
    WSB_TRY(nccl_load(&api));
    static_assert(sizeof(ncclUniqueId) == WSB_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size changed");
    ncclUniqueId id;
    WSB_NCCL(api, api->GetUniqueId(&id));
    std::memcpy(out128, &id, sizeof(id));
    return WSB_OK;
}

struct HaloComm {
    const NcclApi *api = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    float *scratch = nullptr;  // one float on the device for halo_align
};

int halo_comm_create(int rank, int nranks, const void *unique_id, HaloComm **out) {
    const NcclApi *api = nullptr;
    WSB_TRY(nccl_load(&api));
    ncclUniqueId id;
    std::memcpy(&id, unique_id, sizeof(id));
    HaloComm *c = new HaloComm();
    c->api = api;
    c->rank = rank;
    c->nranks = nranks;
    ncclResult_t r = api->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return nccl_fail(api, r, "ncclCommInitRank");
    }
    if (cudaMalloc(&c->scratch, sizeof(float)) != cudaSuccess || cudaMemset(c->scratch, 0, sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        api->CommDestroy(c->comm);
        delete c;
        return fail(WSB_ERR_CUDA, "cudaMalloc(halo scratch) failed");
    }
    *out = c;
    return WSB_OK;
}

void halo_comm_destroy(HaloComm *c) {
    if (!c) return;
    if (c->comm) c->api->CommDestroy(c->comm);
    if (c->scratch) cudaFree(c->scratch);
    delete c;
}

// A device-side rendezvous of all ranks on `st` (a one-float all-reduce): what follows on the stream starts within
// microseconds on every GPU, whatever the host-side skew between the processes was.
int halo_align(HaloComm *c, cudaStream_t st) {
    const NcclApi *api = c->api;
    WSB_NCCL(api, api->AllReduce(c->scratch, c->scratch, 1, ncclFloat, ncclSum, c->comm, st));
    return WSB_OK;
}

int halo_exchange(HaloComm *c, void *const *planes, int nplanes, size_t elem_size, int pitch, int H, int nrows,
                  int levels, long long level_stride, cudaStream_t st) {
    const NcclApi *api = c->api;
    const size_t row_bytes = (size_t)pitch * elem_size;
    const size_t bytes = (size_t)nrows * row_bytes;
    const int up = c->rank - 1, down = c->rank + 1;
    WSB_NCCL(api, api->GroupStart());
    for (int l = 0; l < levels; ++l)
    for (int k = 0; k < nplanes; ++k) {
        char *o = (char *)planes[k] + (size_t)l * (size_t)level_stride * elem_size;
        if (up >= 0) {
            WSB_NCCL(api, api->Send(o, bytes, ncclInt8, up, c->comm, st));                                // my top rows
            WSB_NCCL(api, api->Recv(o - bytes, bytes, ncclInt8, up, c->comm, st));                        // ghosts above
        }
        if (down < c->nranks) {
            WSB_NCCL(api, api->Send(o + (size_t)(H - nrows) * row_bytes, bytes, ncclInt8, down, c->comm, st));  // my bottom rows
            WSB_NCCL(api, api->Recv(o + (size_t)H * row_bytes, bytes, ncclInt8, down, c->comm, st));     // ghosts below
        }
    }
    WSB_NCCL(api, api->GroupEnd());
    return WSB_OK;
}

}  // namespace wsb
