// Ghost-row exchange between row slabs: ncclSend/ncclRecv over NVLink 5 / NVSwitch, one group per
// exchange, on the simulation's comm stream.
//
// libnccl is resolved with dlopen on first use so that single-GPU users carry no NCCL dependency and so
// that, inside a process that already loaded a libnccl.so.2 (e.g. torch's), the same copy is reused.
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <cstdlib>
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

// ---- peer mappings for the fused ghost exchange (PeerExchange, wsb_internal.h) ---------------------------------
// Each rank exports its six state planes (both buffers of u, v, h) and its flag words with cudaIpcGetMemHandle and
// hands the handles to its two neighbours through the NCCL communicator it already has (one ncclSend/ncclRecv group
// of 592-byte blobs); the neighbours map them with cudaIpcOpenMemHandle. cudaMalloc may carve several small planes
// out of one allocation, and an IPC handle always names the WHOLE allocation, so every pointer travels as
// (handle, offset from the allocation base). All ranks then agree (all-reduce) whether every mapping succeeded:
// the fused exchange is used by all ranks or by none.
namespace {
struct PeerBlob {
    cudaIpcMemHandle_t handle[kPeerPointers];
    unsigned long long offset[kPeerPointers];
    int H;
    int valid;
};

using MemGetAddressRangeFn = int (*)(unsigned long long *, size_t *, unsigned long long);
MemGetAddressRangeFn address_range_fn() {
    static MemGetAddressRangeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<MemGetAddressRangeFn>(p);
    }();
    return fn;
}

bool open_link(const PeerBlob &b, PeerLink *link) {
    if (!b.valid) return false;
    link->H = b.H;
    for (int k = 0; k < kPeerPointers; ++k) {
        // the same allocation may back several pointers: map it once
        void *base = nullptr;
        for (int j = 0; j < k; ++j)
            if (std::memcmp(&b.handle[j], &b.handle[k], sizeof(cudaIpcMemHandle_t)) == 0) base = link->mapped[j];
        if (!base) {
            if (cudaIpcOpenMemHandle(&base, b.handle[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
            link->owned[k] = true;
        }
        link->mapped[k] = base;
        link->ptr[k] = (char *)base + b.offset[k];
    }
    link->present = true;
    return true;
}
}  // namespace

void peer_close(PeerLink *link) {
    for (int k = 0; k < kPeerPointers; ++k) {
        if (link->owned[k] && link->mapped[k]) cudaIpcCloseMemHandle(link->mapped[k]);
        link->owned[k] = false;
        link->mapped[k] = link->ptr[k] = nullptr;
    }
    link->present = false;
}

int peer_setup(HaloComm *c, void *const local[kPeerPointers], int H, PeerLink *up, PeerLink *dn, bool *all_ok,
               cudaStream_t st) {
    const NcclApi *api = c->api;
    *all_ok = false;
    PeerBlob mine;
    std::memset(&mine, 0, sizeof(mine));
    mine.H = H;
    mine.valid = std::getenv("WSB_NO_PEER_EXCHANGE") ? 0 : 1;
    MemGetAddressRangeFn range = address_range_fn();
    if (!range) mine.valid = 0;
    for (int k = 0; k < kPeerPointers && mine.valid; ++k) {
        unsigned long long base = 0;
        size_t size = 0;
        if (range(&base, &size, (unsigned long long)(uintptr_t)local[k]) != 0 ||
            cudaIpcGetMemHandle(&mine.handle[k], (void *)(uintptr_t)base) != cudaSuccess) {
            cudaGetLastError();
            mine.valid = 0;
            break;
        }
        mine.offset[k] = (unsigned long long)(uintptr_t)local[k] - base;
    }
    // blobs travel through device staging buffers: [0] mine, [1] from up, [2] from down
    PeerBlob *dev = nullptr;
    WSB_CUDA(cudaMalloc(&dev, 3 * sizeof(PeerBlob)));
    struct Free {
        void *p;
        ~Free() { cudaFree(p); }
    } free_dev{dev};
    WSB_CUDA(cudaMemsetAsync(dev, 0, 3 * sizeof(PeerBlob), st));
    WSB_CUDA(cudaMemcpyAsync(dev, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    const int r_up = c->rank - 1, r_dn = c->rank + 1;
    WSB_NCCL(api, api->GroupStart());
    if (r_up >= 0) {
        WSB_NCCL(api, api->Send(dev, sizeof(PeerBlob), ncclInt8, r_up, c->comm, st));
        WSB_NCCL(api, api->Recv(dev + 1, sizeof(PeerBlob), ncclInt8, r_up, c->comm, st));
    }
    if (r_dn < c->nranks) {
        WSB_NCCL(api, api->Send(dev, sizeof(PeerBlob), ncclInt8, r_dn, c->comm, st));
        WSB_NCCL(api, api->Recv(dev + 2, sizeof(PeerBlob), ncclInt8, r_dn, c->comm, st));
    }
    WSB_NCCL(api, api->GroupEnd());
    PeerBlob got[2];
    WSB_CUDA(cudaMemcpyAsync(got, dev + 1, 2 * sizeof(PeerBlob), cudaMemcpyDeviceToHost, st));
    WSB_CUDA(cudaStreamSynchronize(st));
    bool ok = mine.valid != 0;
    if (ok && r_up >= 0) ok = open_link(got[0], up);
    if (ok && r_dn < c->nranks) ok = open_link(got[1], dn);
    // every rank uses the fused exchange, or none does
    const float flag = ok ? 1.0f : 0.0f;
    float agreed = 0.0f;
    WSB_CUDA(cudaMemcpyAsync(c->scratch, &flag, sizeof(flag), cudaMemcpyHostToDevice, st));
    WSB_NCCL(api, api->AllReduce(c->scratch, c->scratch, 1, ncclFloat, ncclMin, c->comm, st));
    WSB_CUDA(cudaMemcpyAsync(&agreed, c->scratch, sizeof(agreed), cudaMemcpyDeviceToHost, st));
    WSB_CUDA(cudaMemsetAsync(c->scratch, 0, sizeof(float), st));
    WSB_CUDA(cudaStreamSynchronize(st));
    *all_ok = agreed > 0.5f;
    if (!*all_ok) {
        peer_close(up);
        peer_close(dn);
    }
    return WSB_OK;
}

}  // namespace wsb
