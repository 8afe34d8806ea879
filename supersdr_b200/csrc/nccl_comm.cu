// Multi-GPU plumbing of the C ABI (ssdr_nccl_*): one process per GPU, NCCL over NVLink / NVSwitch, used ONLY to move
// input shards from a root rank to the ranks that own them (and the small pixel rows back) -- channels are independent,
// there is no collective in the math.  libnccl is loaded lazily with dlopen, so
// libssdr_b200.so itself links nothing but the CUDA runtime and loads on hosts without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace ssdr {
int bind_device(int device);
extern std::atomic<int> g_device;

namespace {
struct NcclApi {
    void* so = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void nccl_load() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.so) break;
    }
    if (!g_nccl.so) return;
#define SSDR_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.so, name))
    SSDR_SYM(GetUniqueId, "ncclGetUniqueId"); SSDR_SYM(CommInitRank, "ncclCommInitRank"); SSDR_SYM(CommDestroy, "ncclCommDestroy");
    SSDR_SYM(GetErrorString, "ncclGetErrorString"); SSDR_SYM(Send, "ncclSend"); SSDR_SYM(Recv, "ncclRecv");
    SSDR_SYM(GroupStart, "ncclGroupStart"); SSDR_SYM(GroupEnd, "ncclGroupEnd"); SSDR_SYM(AllReduce, "ncclAllReduce");
    SSDR_SYM(GetVersion, "ncclGetVersion");
#undef SSDR_SYM
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.GetErrorString && g_nccl.Send &&
                g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.AllReduce;
}
bool nccl_ready() {
    std::call_once(g_nccl_once, nccl_load);
    if (!g_nccl.ok) set_error("NCCL is not available (dlopen libnccl.so.2 failed: %s)", g_nccl.so ? "missing symbols" : dlerror());
    return g_nccl.ok;
}
}  // namespace
}  // namespace ssdr

using namespace ssdr;

struct ssdr_comm {
    int device = -1, rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;
    double* d_scalar = nullptr;
};

#define SSDR_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess) { set_error("%s failed: %s", #call, g_nccl.GetErrorString(r__)); return SSDR_E_CUDA; } \
    } while (0)

extern "C" {

int ssdr_nccl_available(void) {
    if (!nccl_ready()) return 0;
    int v = 0;
    if (g_nccl.GetVersion) g_nccl.GetVersion(&v);
    return v > 0 ? v : 1;
}

int ssdr_nccl_unique_id(void* id128) {
    SSDR_ARG(id128 != nullptr, "null id buffer");
    if (!nccl_ready()) return SSDR_E_STATE;
    static_assert(sizeof(ncclUniqueId) == SSDR_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    SSDR_NCCL(g_nccl.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return SSDR_OK;
}

int ssdr_nccl_init(ssdr_comm_t* out, const void* id128, int rank, int world) {
    SSDR_ARG(out && id128, "null argument");
    *out = nullptr;
    SSDR_ARG(world >= 1 && rank >= 0 && rank < world, "rank %d outside world %d", rank, world);
    if (!nccl_ready()) return SSDR_E_STATE;
    int rc = bind_device(g_device.load());
    if (rc) return rc;
    ssdr_comm* c = new ssdr_comm();
    c->rank = rank; c->world = world;
    if (cudaGetDevice(&c->device) != cudaSuccess) c->device = -1;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); delete c; return SSDR_E_CUDA; }
    if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc(&c->d_scalar, sizeof(double)) != cudaSuccess) {
        set_error("communicator stream / scratch allocation failed");
        ssdr_nccl_destroy(c);
        return SSDR_E_CUDA;
    }
    *out = c;
    return SSDR_OK;
}

int ssdr_nccl_destroy(ssdr_comm_t c) {
    if (!c) return SSDR_OK;
    bind_device(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_scalar);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return SSDR_OK;
}

// Root's buffer holds every rank's shard at offsets[r] (bytes, counts[r] bytes each); rank r receives its shard into
// recv_dev.  One grouped send/recv round over NVLink; the root's own shard is a device-to-device copy.
int ssdr_nccl_scatter(ssdr_comm_t c, const void* send_root_dev, const size_t* offsets, const size_t* counts, void* recv_dev, int root) {
    SSDR_ARG(c && offsets && counts, "null argument");
    SSDR_ARG(root >= 0 && root < c->world, "root %d outside world %d", root, c->world);
    SSDR_ARG(c->rank != root || send_root_dev, "the root rank needs a send buffer");
    int rc = bind_device(c->device);
    if (rc) return rc;
    const unsigned char* src = static_cast<const unsigned char*>(send_root_dev);
    SSDR_NCCL(g_nccl.GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->world; ++r)
            if (r != root && counts[r]) SSDR_NCCL(g_nccl.Send(src + offsets[r], counts[r], ncclUint8, r, c->comm, c->st));
    } else if (counts[c->rank]) {
        SSDR_ARG(recv_dev != nullptr, "null receive buffer");
        SSDR_NCCL(g_nccl.Recv(recv_dev, counts[c->rank], ncclUint8, root, c->comm, c->st));
    }
    SSDR_NCCL(g_nccl.GroupEnd());
    if (c->rank == root && counts[root] && recv_dev && recv_dev != src + offsets[root])
        SSDR_CUDA(cudaMemcpyAsync(recv_dev, src + offsets[root], counts[root], cudaMemcpyDeviceToDevice, c->st));
    return SSDR_OK;
}

int ssdr_nccl_gather(ssdr_comm_t c, const void* send_dev, void* recv_root_dev, const size_t* offsets, const size_t* counts, int root) {
    SSDR_ARG(c && offsets && counts, "null argument");
    SSDR_ARG(root >= 0 && root < c->world, "root %d outside world %d", root, c->world);
    SSDR_ARG(c->rank != root || recv_root_dev, "the root rank needs a receive buffer");
    int rc = bind_device(c->device);
    if (rc) return rc;
    unsigned char* dst = static_cast<unsigned char*>(recv_root_dev);
    SSDR_NCCL(g_nccl.GroupStart());
    if (c->rank == root) {
        for (int r = 0; r < c->world; ++r)
            if (r != root && counts[r]) SSDR_NCCL(g_nccl.Recv(dst + offsets[r], counts[r], ncclUint8, r, c->comm, c->st));
    } else if (counts[c->rank]) {
        SSDR_ARG(send_dev != nullptr, "null send buffer");
        SSDR_NCCL(g_nccl.Send(send_dev, counts[c->rank], ncclUint8, root, c->comm, c->st));
    }
    SSDR_NCCL(g_nccl.GroupEnd());
    if (c->rank == root && counts[root] && send_dev && send_dev != dst + offsets[root])
        SSDR_CUDA(cudaMemcpyAsync(dst + offsets[root], send_dev, counts[root], cudaMemcpyDeviceToDevice, c->st));
    return SSDR_OK;
}

int ssdr_nccl_allreduce_max_f64(ssdr_comm_t c, double* value) {
    SSDR_ARG(c && value, "null argument");
    int rc = bind_device(c->device);
    if (rc) return rc;
    SSDR_CUDA(cudaMemcpyAsync(c->d_scalar, value, sizeof(double), cudaMemcpyHostToDevice, c->st));
    SSDR_NCCL(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclFloat64, ncclMax, c->comm, c->st));
    SSDR_CUDA(cudaMemcpyAsync(value, c->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, c->st));
    SSDR_CUDA(cudaStreamSynchronize(c->st));
    return SSDR_OK;
}

int ssdr_nccl_barrier(ssdr_comm_t c) {
    double v = 0.0;
    return ssdr_nccl_allreduce_max_f64(c, &v);
}

int ssdr_nccl_sync(ssdr_comm_t c) {
    SSDR_ARG(c != nullptr, "null communicator");
    int rc = bind_device(c->device);
    if (rc) return rc;
    SSDR_CUDA(cudaStreamSynchronize(c->st));
    return SSDR_OK;
}

}  // extern "C"
