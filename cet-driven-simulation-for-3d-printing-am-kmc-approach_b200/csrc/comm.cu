// comm.cu — slab decomposition over the GPUs of one box: NCCL point-to-point halo exchange
// and the small all-reduces of the sweep totals, all enqueued on the context's stream.
//
// Slabs are contiguous runs of planes along axis 0 (the slowest axis, so a halo is one
// contiguous block per field: no pack/unpack kernels).  One process per GPU; the NCCL unique
// id is created by rank 0 (cet_comm_unique_id) and distributed by the host side
// (torch.distributed / any out-of-band channel) before cet_comm_init.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2") so that a process which already
// carries PyTorch's bundled NCCL shares that copy, and single-GPU use needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>
#include "ctx.cuh"

namespace cet {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr;      // optional (NCCL >= 2.18)
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.handle) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    CET_REQUIRE(h != nullptr, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define CET_SYM(field, sym)                                                   \
    *(void **)(&g_nccl.field) = dlsym(h, sym);                                \
    CET_REQUIRE(g_nccl.field != nullptr, "NCCL symbol %s missing", sym)
    CET_SYM(GetUniqueId, "ncclGetUniqueId");
    CET_SYM(CommInitRank, "ncclCommInitRank");
    CET_SYM(CommDestroy, "ncclCommDestroy");
    CET_SYM(Send, "ncclSend");
    CET_SYM(Recv, "ncclRecv");
    CET_SYM(AllReduce, "ncclAllReduce");
    CET_SYM(GroupStart, "ncclGroupStart");
    CET_SYM(GroupEnd, "ncclGroupEnd");
    CET_SYM(GetErrorString, "ncclGetErrorString");
#undef CET_SYM
    *(void **)(&g_nccl.CommSplit) = dlsym(h, "ncclCommSplit");
    g_nccl.handle = h;
    return 0;
}

#define CET_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) {                                                                    \
            cet::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return 2000 + (int)r_;                                                                  \
        }                                                                                           \
    } while (0)

// Exchange `halo` planes with both neighbours.  fields: 1 packed state, 2 theta+phi, 4 T.
int comm_halo_exchange(cet_ctx *c, int fields)
{
    if (c->world <= 1 || c->halo == 0) return 0;
    CET_REQUIRE(c->nccl_comm != nullptr, "halo exchange: cet_comm_init has not been called");
    const int64_t H = c->halo, own = c->i_end - c->i_begin;
    CET_REQUIRE(own >= H, "halo exchange: slab of %lld planes is thinner than the halo %lld", (long long)own,
                (long long)H);
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const size_t n = (size_t)(H * c->plane);
    const int lower = c->rank - 1, upper = c->rank + 1;
    struct Field { char *base; size_t esz; };
    Field fl[4];
    int nf = 0;
    if (fields & 1) fl[nf++] = {(char *)c->vox, 1};
    if (fields & 2) { fl[nf++] = {(char *)c->theta, 8}; fl[nf++] = {(char *)c->phi, 8}; }
    if (fields & 4) fl[nf++] = {(char *)c->T, 8};
    CET_NCCL(g_nccl.GroupStart());
    ncclResult_t first = ncclSuccess;    // an early return would leave the group open: remember the first error, always close
    auto op = [&](ncclResult_t r) { if (first == ncclSuccess) first = r; };
    for (int f = 0; f < nf; ++f) {
        const size_t bytes = n * fl[f].esz;
        char *b = fl[f].base;
        if (lower >= 0) {
            op(g_nccl.Send(b + (size_t)H * c->plane * fl[f].esz, bytes, ncclChar, lower, comm, c->stream));
            op(g_nccl.Recv(b, bytes, ncclChar, lower, comm, c->stream));
        }
        if (upper < c->world) {
            op(g_nccl.Send(b + (size_t)(c->np - 2 * H) * c->plane * fl[f].esz, bytes, ncclChar, upper, comm, c->stream));
            op(g_nccl.Recv(b + (size_t)(c->np - H) * c->plane * fl[f].esz, bytes, ncclChar, upper, comm, c->stream));
        }
    }
    const ncclResult_t end = g_nccl.GroupEnd();
    CET_NCCL(first);
    CET_NCCL(end);
    if (fields & 1) c->nst_valid = false;      // ghost states changed (cet_sweep_run repairs the cache itself)
    if (fields & 2) {          // orientation unit vectors of the refreshed ghost planes
        if (lower >= 0) if (int rc = orient_update(c, 0, H)) return rc;
        if (upper < c->world) if (int rc = orient_update(c, c->np - H, c->np)) return rc;
    }
    return 0;
}

int comm_delta_alloc(cet_ctx *c)
{
    if (c->delta_cap) return 0;
    // 1/32 of the zone sites (~19 % of a plane): at 0.5 % of the sites firing per sweep about 4 % of a plane change
    c->delta_cap = (int64_t)DELTA_ZONE * c->plane / 32 + 1024;
    const size_t bytes = DELTA_HEADER + (size_t)c->delta_cap * sizeof(DeltaEntry);
    for (int f = 0; f < 2; ++f) {
        CET_CUDA(cudaMalloc(&c->delta_send[f], bytes));
        CET_CUDA(cudaMalloc(&c->delta_recv[f], bytes));
        CET_CUDA(cudaMemsetAsync(c->delta_send[f], 0, bytes, c->stream));
        CET_CUDA(cudaMemsetAsync(c->delta_recv[f], 0, bytes, c->stream));
    }
    return 0;
}

// One message per neighbour and direction: header + the full entry capacity (the two sides of an
// NCCL send/recv must agree on the size, and a count round trip would cost another latency).
int comm_delta_exchange(cet_ctx *c)
{
    if (c->world <= 1) return 0;
    CET_REQUIRE(c->nccl_comm != nullptr && c->delta_cap > 0, "delta exchange: communicator / buffers missing");
    ncclComm_t comm = (ncclComm_t)c->nccl_comm;
    const size_t bytes = DELTA_HEADER + (size_t)c->delta_cap * sizeof(DeltaEntry);
    const int lower = c->rank - 1, upper = c->rank + 1;
    CET_NCCL(g_nccl.GroupStart());
    ncclResult_t first = ncclSuccess;
    auto op = [&](ncclResult_t r) { if (first == ncclSuccess) first = r; };
    if (lower >= 0) {
        op(g_nccl.Send(c->delta_send[0], bytes, ncclChar, lower, comm, c->stream));
        op(g_nccl.Recv(c->delta_recv[0], bytes, ncclChar, lower, comm, c->stream));
    }
    if (upper < c->world) {
        op(g_nccl.Send(c->delta_send[1], bytes, ncclChar, upper, comm, c->stream));
        op(g_nccl.Recv(c->delta_recv[1], bytes, ncclChar, upper, comm, c->stream));
    }
    const ncclResult_t end = g_nccl.GroupEnd();
    CET_NCCL(first);
    CET_NCCL(end);
    return 0;
}

// plane_sum[0..n) holds this slab's plane totals (zeros elsewhere, all >= 0) and max_inout the
// local maximum, stored contiguously at plane_sum[n].  Every entry is non-zero on at most one
// rank, so ONE max-all-reduce over n+1 doubles yields both the gathered plane sums (bit-exact,
// independent of the slab count) and the global maximum.
// With a second communicator (cet_comm_init splits one off when NCCL offers ncclCommSplit) the reduction runs on
// the context's side stream: started as soon as the plane sums exist, joined by comm_sweep_reduce_join before the
// kernel that needs the totals — pick, apply, the delta exchange and the refresh run meanwhile, and the rank skew
// the reduction would otherwise wait out on the main stream is absorbed.  Without it: in place, on the main stream.
int comm_sweep_reduce(cet_ctx *c, double *plane_sum, int n, double *max_inout)
{
    if (c->world <= 1) return 0;
    CET_REQUIRE(c->nccl_comm != nullptr, "sweep reduce: cet_comm_init has not been called");
    CET_REQUIRE(max_inout == plane_sum + n, "sweep reduce: max slot must follow the plane sums");
    if (c->nccl_comm2) {
        CET_CUDA(cudaEventRecord(c->ev_reduce_ready, c->stream));
        CET_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_reduce_ready, 0));
        CET_NCCL(g_nccl.AllReduce(plane_sum, plane_sum, (size_t)n + 1, ncclDouble, ncclMax, (ncclComm_t)c->nccl_comm2, c->stream2));
        CET_CUDA(cudaEventRecord(c->ev_reduce_done, c->stream2));
        return 0;
    }
    CET_NCCL(g_nccl.AllReduce(plane_sum, plane_sum, (size_t)n + 1, ncclDouble, ncclMax, (ncclComm_t)c->nccl_comm,
                              c->stream));
    return 0;
}
int comm_sweep_reduce_join(cet_ctx *c)
{
    if (c->world > 1 && c->nccl_comm2) CET_CUDA(cudaStreamWaitEvent(c->stream, c->ev_reduce_done, 0));
    return 0;
}

}  // namespace cet

using namespace cet;

extern "C" {

int cet_comm_unique_id(void *id128)
{
    CET_REQUIRE(id128 != nullptr, "cet_comm_unique_id: NULL");
    if (int rc = nccl_load()) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    CET_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int cet_comm_init(cet_ctx *c, const void *id128, int rank, int world)
{
    CET_REQUIRE(c && id128, "cet_comm_init: NULL argument");
    CET_REQUIRE(world >= 1 && rank >= 0 && rank < world, "cet_comm_init: bad rank %d of %d", rank, world);
    CET_REQUIRE(c->nccl_comm == nullptr, "cet_comm_init: communicator already initialised");
    if (int rc = nccl_load()) return rc;
    cet::DeviceGuard dg(c->device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm;
    CET_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    c->nccl_comm = comm;
    c->rank = rank; c->world = world;
    // a second communicator + stream for the sweeps' totals reduction (comm_sweep_reduce); optional
    if (world > 1 && g_nccl.CommSplit && !c->nccl_comm2) {
        ncclComm_t comm2 = nullptr;
        if (g_nccl.CommSplit(comm, 0, rank, &comm2, nullptr) == ncclSuccess && comm2) {
            if (cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking) == cudaSuccess &&
                cudaEventCreateWithFlags(&c->ev_reduce_ready, cudaEventDisableTiming) == cudaSuccess &&
                cudaEventCreateWithFlags(&c->ev_reduce_done, cudaEventDisableTiming) == cudaSuccess) {
                c->nccl_comm2 = comm2;
            } else {
                g_nccl.CommDestroy(comm2);
                (void)cudaGetLastError();
            }
        }
    }
    return 0;
}

int cet_comm_destroy(cet_ctx *c)
{
    if (!c || !c->nccl_comm) return 0;
    cet::DeviceGuard dg(c->device);
    if (c->nccl_comm2) {
        cudaStreamSynchronize(c->stream2);
        g_nccl.CommDestroy((ncclComm_t)c->nccl_comm2);
        c->nccl_comm2 = nullptr;
    }
    g_nccl.CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
    c->rank = 0; c->world = 1;
    return 0;
}

int cet_halo_exchange(cet_ctx *c, int fields)
{
    CET_REQUIRE(c, "cet_halo_exchange: NULL ctx");
    cet::DeviceGuard dg(c->device);
    lattice_changed(c);                  // the ghost planes of every derived array are stale
    return comm_halo_exchange(c, fields);
}

int cet_allreduce_f64(cet_ctx *c, double *inout_host, int n, int op)
{
    CET_REQUIRE(c && inout_host && n > 0, "cet_allreduce_f64: bad argument");
    cet::DeviceGuard dg(c->device);
    if (c->world <= 1) return 0;
    CET_REQUIRE(c->nccl_comm != nullptr, "cet_allreduce_f64: cet_comm_init has not been called");
    if (int rc = ensure_stage(c, (size_t)n * 8)) return rc;
    CET_CUDA(cudaMemcpyAsync(c->stage, inout_host, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CET_NCCL(g_nccl.AllReduce(c->stage, c->stage, (size_t)n, ncclDouble, op == 1 ? ncclMax : ncclSum,
                              (ncclComm_t)c->nccl_comm, c->stream));
    CET_CUDA(cudaMemcpyAsync(inout_host, c->stage, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CET_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
