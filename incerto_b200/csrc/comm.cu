// NCCL plumbing (one process per GPU).  The reference has no communication
// backend at all; this exists only for (a) the 1-column halo
// exchange of a row-partitioned forest-fire grid and (b) all-reduce of
// cross-replica / cross-slab aggregate statistics.
//
// NCCL is bound lazily with dlopen so that the library loads (and single-GPU
// simulations run) on hosts without it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

#include "modules.hpp"

namespace incerto
{

struct NcclApi
{
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};

static NcclApi g_nccl;

static int32_t nccl_load()
{
    if (g_nccl.handle) return INCERTO_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names)
    {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h)
    {
        set_error("cannot load NCCL: %s", dlerror());
        return INCERTO_ERR_COMM;
    }
#define INC_SYM(field, name)                                                \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); \
    if (!g_nccl.field)                                                      \
    {                                                                       \
        set_error("NCCL symbol %s missing", name);                          \
        return INCERTO_ERR_COMM;                                            \
    }
    INC_SYM(GetUniqueId, "ncclGetUniqueId")
    INC_SYM(CommInitRank, "ncclCommInitRank")
    INC_SYM(CommDestroy, "ncclCommDestroy")
    INC_SYM(AllReduce, "ncclAllReduce")
    INC_SYM(AllGather, "ncclAllGather")
    INC_SYM(Send, "ncclSend")
    INC_SYM(Recv, "ncclRecv")
    INC_SYM(GroupStart, "ncclGroupStart")
    INC_SYM(GroupEnd, "ncclGroupEnd")
    INC_SYM(GetErrorString, "ncclGetErrorString")
#undef INC_SYM
    g_nccl.handle = h;
    return INCERTO_OK;
}

#define INC_NCCL(call)                                                                   \
    do                                                                                   \
    {                                                                                    \
        ncclResult_t r__ = (call);                                                       \
        if (r__ != ncclSuccess)                                                          \
        {                                                                                \
            set_error("NCCL error %d (%s) in %s", static_cast<int>(r__), g_nccl.GetErrorString(r__), #call); \
            return INCERTO_ERR_COMM;                                                     \
        }                                                                                \
    } while (0)

struct Comm
{
    ncclComm_t comm = nullptr;
    uint32_t rank = 0, world = 1;
};

// Communicators are cached per unique id for the life of the process: creating and destroying an NCCL communicator
// costs ~0.5 s each, which a host that builds one simulation after another (bench.py's end-to-end jobs, replicas
// through fresh handles) must not pay per handle.  Handles that attach with the same id share the communicator; they
// are used from one thread, one collective at a time (the handle contract, incerto_b200.h).
struct CachedComm
{
    uint8_t id[128];
    int device;
    Comm comm;
};
static std::vector<CachedComm*> g_comm_cache;

int32_t comm_unique_id(uint8_t out[128])
{
    INC_TRY(nccl_load());
    ncclUniqueId id;
    INC_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    std::memcpy(out, &id, 128);
    return INCERTO_OK;
}

int32_t comm_attach(Sim& s, const uint8_t idb[128], uint32_t rank, uint32_t world)
{
    INC_TRY(nccl_load());
    if (s.comm) comm_free(s);
    INC_CUDA(cudaSetDevice(s.device));
    for (CachedComm* cc : g_comm_cache)
        if (std::memcmp(cc->id, idb, 128) == 0 && cc->device == s.device)
        {
            if (cc->comm.rank != rank || cc->comm.world != world)
            {
                set_error("attach_comm: this unique id is already bound to rank %u of %u in this process", cc->comm.rank, cc->comm.world);
                return INCERTO_ERR_INVALID_ARGUMENT;
            }
            s.comm = &cc->comm;
            return INCERTO_OK;
        }
    ncclUniqueId id;
    std::memcpy(&id, idb, 128);
    CachedComm* cc = new CachedComm;
    std::memcpy(cc->id, idb, 128);
    cc->device = s.device;
    cc->comm.rank = rank;
    cc->comm.world = world;
    ncclResult_t r = g_nccl.CommInitRank(&cc->comm.comm, static_cast<int>(world), id, static_cast<int>(rank));
    if (r != ncclSuccess)
    {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete cc;
        return INCERTO_ERR_COMM;
    }
    g_comm_cache.push_back(cc);
    s.comm = &cc->comm;
    return INCERTO_OK;
}

// the handle lets go of the (cached, shared) communicator; the communicator itself lives until the process ends
void comm_free(Sim& s) { s.comm = nullptr; }

int32_t comm_allreduce(Sim& s, void* d_buf, uint32_t n, int is_f64, uint32_t op)
{
    if (!s.comm)
    {
        set_error("no communicator attached");
        return INCERTO_ERR_COMM;
    }
    const ncclRedOp_t ops[3] = {ncclSum, ncclMin, ncclMax};
    if (op > 2) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_NCCL(g_nccl.AllReduce(d_buf, d_buf, n, is_f64 ? ncclFloat64 : ncclUint64, ops[op], s.comm->comm, s.stream));
    return INCERTO_OK;
}

// n 64-bit host values (u64 or f64) reduced over the communicator in place
int32_t comm_allreduce_host(Sim& S, void* values, uint32_t n, int is_f64, uint32_t op)
{
    if (!values || !n) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_CUDA(cudaSetDevice(S.device));
    INC_TRY(ensure_scratch(S, static_cast<size_t>(n) * 8));
    INC_CUDA(cudaMemcpyAsync(S.d_scratch, values, static_cast<size_t>(n) * 8, cudaMemcpyHostToDevice, S.stream));
    INC_TRY(comm_allreduce(S, S.d_scratch, n, is_f64, op));
    INC_CUDA(cudaMemcpyAsync(values, S.d_scratch, static_cast<size_t>(n) * 8, cudaMemcpyDeviceToHost, S.stream));
    INC_CUDA(cudaStreamSynchronize(S.stream));
    return INCERTO_OK;
}

// every rank contributes `bytes` bytes; d_recv holds world * bytes
int32_t comm_allgather_bytes(Sim& s, const void* d_send, void* d_recv, size_t bytes)
{
    if (!s.comm)
    {
        set_error("no communicator attached");
        return INCERTO_ERR_COMM;
    }
    INC_NCCL(g_nccl.AllGather(d_send, d_recv, bytes, ncclUint8, s.comm->comm, s.stream));
    return INCERTO_OK;
}

int32_t comm_halo(Sim& s, const void* const* d_send_lo, void* const* d_recv_lo, const void* const* d_send_hi,
                  void* const* d_recv_hi, const size_t* bytes, uint32_t n_msgs)
{
    if (!s.comm)
    {
        set_error("no communicator attached");
        return INCERTO_ERR_COMM;
    }
    Comm& c = *s.comm;
    INC_NCCL(g_nccl.GroupStart());
    // inside the group an error must not return early: the group would stay open and swallow every later NCCL call
    // of this thread.  Remember the first failure, stop queueing, always close the group.
    ncclResult_t first = ncclSuccess;
    const char* where = "";
    auto q = [&](ncclResult_t r, const char* what) {
        if (r != ncclSuccess && first == ncclSuccess)
        {
            first = r;
            where = what;
        }
        return first == ncclSuccess;
    };
    for (uint32_t m = 0; m < n_msgs && first == ncclSuccess; ++m)
    {
        if (c.rank > 0)
        {
            if (!q(g_nccl.Send(d_send_lo[m], bytes[m], ncclUint8, static_cast<int>(c.rank) - 1, c.comm, s.stream), "ncclSend(lower)")) break;
            if (!q(g_nccl.Recv(d_recv_lo[m], bytes[m], ncclUint8, static_cast<int>(c.rank) - 1, c.comm, s.stream), "ncclRecv(lower)")) break;
        }
        if (c.rank + 1 < c.world)
        {
            if (!q(g_nccl.Send(d_send_hi[m], bytes[m], ncclUint8, static_cast<int>(c.rank) + 1, c.comm, s.stream), "ncclSend(upper)")) break;
            if (!q(g_nccl.Recv(d_recv_hi[m], bytes[m], ncclUint8, static_cast<int>(c.rank) + 1, c.comm, s.stream), "ncclRecv(upper)")) break;
        }
    }
    const ncclResult_t end = g_nccl.GroupEnd();
    if (first != ncclSuccess)
    {
        set_error("NCCL error %d (%s) in %s", static_cast<int>(first), g_nccl.GetErrorString(first), where);
        return INCERTO_ERR_COMM;
    }
    INC_NCCL(end);
    return INCERTO_OK;
}

} // namespace incerto

using namespace incerto;

extern "C"
{

int32_t incerto_comm_unique_id(uint8_t out[128])
{
    if (!out) return INCERTO_ERR_INVALID_ARGUMENT;
    return comm_unique_id(out);
}

int32_t incerto_sim_attach_comm(incerto_sim* s, const uint8_t unique_id[128], uint32_t rank, uint32_t world)
{
    Sim* S = reinterpret_cast<Sim*>(s);
    if (!S || !unique_id || world == 0 || rank >= world) return INCERTO_ERR_INVALID_ARGUMENT;
    INC_TRY(comm_attach(*S, unique_id, rank, world));
    // row-partitioned forest: try to map the neighbours' slabs (CUDA IPC over NVLink) so that the halo columns can
    // be stored straight into their memory; NCCL send/recv stays as the fallback transport
    return forest_peer_setup(*S);
}

int32_t incerto_sim_allreduce_u64(incerto_sim* s, uint64_t* values, uint32_t n, uint32_t op)
{
    Sim* S = reinterpret_cast<Sim*>(s);
    if (!S) return INCERTO_ERR_INVALID_ARGUMENT;
    return comm_allreduce_host(*S, values, n, 0, op);
}

int32_t incerto_sim_allreduce_f64(incerto_sim* s, double* values, uint32_t n, uint32_t op)
{
    Sim* S = reinterpret_cast<Sim*>(s);
    if (!S) return INCERTO_ERR_INVALID_ARGUMENT;
    return comm_allreduce_host(*S, values, n, 1, op);
}

} // extern "C"
