// Multi-GPU behind the C ABI: one process per GPU, NCCL over NVLink inside the library.
//
// The reference exchanges ghost patches with MPI_Isend / Irecv once per synchronisation (LIB/MPI/xfer_block_data.f90:10-99), reduces dt with
// MPI_Allreduce(MIN) (LIB/TIME/calculate_time_step.f90:48), moves whole blocks with block_xfer (LIB/MPI/block_xfer_nonblocking.f90:16) and
// keeps light data consistent with MPI_Allgather / Allreduce (LIB/MESH/synchronize_lgt_data.f90).  Here the same four things are entry points
// of libwabbit_gpu.so on an NCCL communicator the library owns, so a Fortran host needs nothing but MPI_Bcast of the 128-byte id:
//
//   wgpu_comm_unique_id / wgpu_comm_init     ncclGetUniqueId on rank 0, MPI_Bcast by the host, ncclCommInitRank on every rank
//   wgpu_comm_set_counts                     per-peer patch / block counts of the exchange declared by wgpu_set_exchange / wgpu_set_halo
//   wgpu_rk_steps                            n Runge-Kutta steps back to back: time, dt and the divergence flag stay on the device, the dt MIN
//                                            is an ncclAllReduce on the device scalar, every stage packs -> grouped ncclSend/ncclRecv on a
//                                            second stream (straight into the patch pool / the halo slots of the stage input: no unpack) ||
//                                            stage kernel on the interior blocks -> stage kernel on the partition-boundary blocks.
//                                            ONE host synchronisation at the end (the reference's N_dt_per_grid loop, performance_test.f90).
//   wgpu_exchange_array                      halo copies (and filtered copies of finer neighbours) of a named array before a wavelet-side call
//   wgpu_ship_blocks                         block_xfer between ranks: gather kernel -> ncclSend / ncclRecv straight into free slots
//   wgpu_comm_allreduce / wgpu_comm_allgatherv_i32   light-data collectives (norms, refinement flags)
//
// NCCL is loaded at run time (dlopen, the copy torch already loaded if there is one): the library itself loads without NCCL.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "wgpu_internal.cuh"

namespace {

typedef struct { char internal[128]; } nccl_id_t;
typedef void *nccl_comm_t;
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };
enum { NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_id_t *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_id_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {   // the copy a host framework (torch) already mapped, else the system's
        api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib)
        for (const char *n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
    if (!api.lib) {
        api.err = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed");
        return nullptr;
    }
    bool ok = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(api.lib, name);
        if (!p) ok = false;
        return p;
    };
    api.GetUniqueId = (int (*)(nccl_id_t *))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(nccl_comm_t *, int, nccl_id_t, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.Send = (int (*)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclRecv");
    api.AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclAllReduce");
    api.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t))sym("ncclAllGather");
    api.GetErrorString = (const char *(*)(int))sym("ncclGetErrorString");
    if (!ok) {
        api.err = "NCCL library lacks a required symbol";
        api.lib = nullptr;
        return nullptr;
    }
    return &api;
}

int32_t mg_fail(wgpu_ctx *ctx, int32_t code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

#define NCCL_CHECK(ctx, api, call)                                                                        \
    do {                                                                                                  \
        int r__ = (call);                                                                                 \
        if (r__ != 0) return mg_fail(ctx, WGPU_ERR_CUDA, std::string(#call) + ": " + (api)->GetErrorString(r__)); \
    } while (0)

// grouped point-to-point exchange: for every peer p, send_counts[p] * unit doubles from send (peer-major) and recv_counts[p] * unit doubles into recv
int32_t exchange(wgpu_ctx *ctx, const double *send, double *recv, const std::vector<int> &send_counts, const std::vector<int> &recv_counts, long long unit,
                 cudaStream_t stream)
{
    NcclApi *api = nccl_api();
    if (!api || !ctx->comm) return mg_fail(ctx, WGPU_ERR_ARG, "no communicator: call wgpu_comm_init first");
    const int W = ctx->comm_world, me = ctx->comm_rank;
    long long so = 0, ro = 0;
    bool any = false;
    for (int p = 0; p < W; ++p) any = any || send_counts[p] || recv_counts[p];
    if (!any) return WGPU_OK;
    NCCL_CHECK(ctx, api, api->GroupStart());
    for (int p = 0; p < W; ++p) {
        if (p != me && send_counts[p]) NCCL_CHECK(ctx, api, api->Send(send + so, (size_t)send_counts[p] * unit, NCCL_FLOAT64, p, ctx->comm, stream));
        if (p != me && recv_counts[p]) NCCL_CHECK(ctx, api, api->Recv(recv + ro, (size_t)recv_counts[p] * unit, NCCL_FLOAT64, p, ctx->comm, stream));
        so += (long long)send_counts[p] * unit;
        ro += (long long)recv_counts[p] * unit;
    }
    NCCL_CHECK(ctx, api, api->GroupEnd());
    return WGPU_OK;
}

int32_t ensure_buf(wgpu_ctx *ctx, double **p, size_t *cap, size_t need)
{
    if (need <= *cap && *p) return WGPU_OK;
    if (*p) {
        cudaFree(*p);
        ctx->dev_bytes -= (int64_t)*cap * 8;
    }
    *p = nullptr;
    *cap = 0;
    const size_t want = need + need / 4 + 1024;
    WGPU_CHECK(ctx, cudaMalloc((void **)p, want * 8));
    ctx->dev_bytes += (int64_t)want * 8;
    *cap = want;
    return WGPU_OK;
}

// ------------------------------------------------------------------------------------------------ peer stores (CUDA IPC over NVLink)
void p2p_teardown(wgpu_ctx *ctx)
{
    for (void *b : ctx->p2p_peer)
        if (b) cudaIpcCloseMemHandle(b);
    ctx->p2p_peer.clear();
    cudaFree(ctx->p2p_mem);
    cudaFree(ctx->d_put_base);
    cudaFree(ctx->d_put_flag);
    cudaFree(ctx->d_send_peer);
    cudaFree(ctx->d_send_idx);
    cudaFree(ctx->d_n_to_peer);
    cudaFree(ctx->d_recv_cnt);
    cudaFree(ctx->d_done);
    ctx->p2p_mem = nullptr;
    ctx->d_put_base = nullptr;
    ctx->d_put_flag = nullptr;
    ctx->d_send_peer = ctx->d_send_idx = ctx->d_n_to_peer = ctx->d_recv_cnt = nullptr;
    ctx->d_done = nullptr;
    ctx->p2p_on = false;
    cudaGetLastError();
}

size_t round256(size_t v) { return (v + 255) & ~(size_t)255; }

// Collective over the communicator (called from wgpu_comm_set_counts on every rank).  Any failure on any rank (no peer access, IPC not
// permitted in this container, ...) leaves every rank on the NCCL path: the outcome is agreed by an all-reduce.
int32_t p2p_setup(wgpu_ctx *ctx)
{
    NcclApi *api = nccl_api();
    const int W = ctx->comm_world, me = ctx->comm_rank;
    int32_t rc;
    if (ctx->p2p_on) {
        // a second exchange on the same communicator: every rank first unmaps its peers' pools, and only when ALL have done so (a collective)
        // are the pools freed -- memory must not be freed while another process still has it mapped
        for (void *&b : ctx->p2p_peer) {
            if (b) cudaIpcCloseMemHandle(b);
            b = nullptr;
        }
        double closed = 1.0;
        if ((rc = wgpu_comm_allreduce(ctx, &closed, 1, 1))) return rc;
    }
    p2p_teardown(ctx);
    ctx->p2p_seq = 0;
    const long long pd = wgpu_patch_doubles(ctx);
    long long n_recv = 0;
    for (int p = 0; p < W; ++p) n_recv += ctx->recv_counts[p];
    int ok = ctx->p2p_want && !(getenv("WGPU_P2P") && atoi(getenv("WGPU_P2P")) == 0) ? 1 : 0;
    const size_t flag_bytes = round256((size_t)W * 4), pool_bytes = round256((size_t)std::max<long long>(n_recv, 1) * pd * 8);
    cudaIpcMemHandle_t handle;
    memset(&handle, 0, sizeof(handle));
    if (ok && cudaMalloc((void **)&ctx->p2p_mem, flag_bytes + 2 * pool_bytes) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(ctx->p2p_mem, 0, flag_bytes) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&handle, ctx->p2p_mem) != cudaSuccess) ok = 0;
    cudaGetLastError();
    // all-gather: [64 B handle][ok][n_recv][recv_counts[W]]
    const size_t rec = (64 + 8 + 4 * (size_t)W + 15) & ~(size_t)15;
    std::vector<char> all(rec * W, 0);
    {
        char *mine = all.data() + rec * me;
        memcpy(mine, &handle, 64);
        int head[2] = {ok, (int)n_recv};
        memcpy(mine + 64, head, 8);
        memcpy(mine + 72, ctx->recv_counts.data(), 4 * (size_t)W);
    }
    if ((rc = ensure_buf(ctx, &ctx->d_xbuf, &ctx->xbuf_cap, rec * W / 8 + 2))) return rc;
    char *d = (char *)ctx->d_xbuf;
    WGPU_CHECK(ctx, cudaMemcpyAsync(d + rec * me, all.data() + rec * me, rec, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, api, api->AllGather(d + rec * me, d, rec, NCCL_INT8, ctx->comm, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(all.data(), d, rec * W, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < W; ++p) ok = ok && *(int *)(all.data() + rec * p + 64);
    ctx->p2p_peer.assign(W, nullptr);
    if (ok)
        for (int p = 0; p < W && ok; ++p) {
            if (p == me || ctx->send_counts[p] == 0) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + rec * p, 64);
            if (cudaIpcOpenMemHandle(&ctx->p2p_peer[p], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                ctx->p2p_peer[p] = nullptr;
                ok = 0;
            }
        }
    cudaGetLastError();
    double agreed = ok;
    if ((rc = wgpu_comm_allreduce(ctx, &agreed, 1, 1))) return rc;   // MIN
    if (agreed < 0.5) {
        p2p_teardown(ctx);
        return WGPU_OK;
    }
    // where my patches go in every peer's pools, my flag word there, and the peer / running index of every send patch (peer-major order)
    std::vector<double *> put_base(2 * (size_t)W, nullptr);
    std::vector<unsigned *> put_flag(W, nullptr);
    std::vector<int> send_peer, send_idx;
    for (int p = 0; p < W; ++p) {
        if (ctx->p2p_peer[p]) {
            const char *rp = all.data() + rec * p;
            const int n_recv_p = *(int *)(rp + 68);
            const int *rc_p = (const int *)(rp + 72);
            long long off = 0;
            for (int r = 0; r < me; ++r) off += rc_p[r];
            const size_t pool_p = round256((size_t)std::max(n_recv_p, 1) * pd * 8);
            char *base = (char *)ctx->p2p_peer[p];
            for (int q = 0; q < 2; ++q) put_base[(size_t)q * W + p] = (double *)(base + flag_bytes + q * pool_p) + off * pd;
            put_flag[p] = (unsigned *)base + me;
        }
        for (int k = 0; k < ctx->send_counts[p]; ++k) {
            send_peer.push_back(p);
            send_idx.push_back(k);
        }
    }
    auto up = [&](auto **dp, const auto &v) -> bool {
        using T = typename std::remove_reference<decltype(v[0])>::type;
        const size_t n = std::max<size_t>(v.size(), 1);
        if (cudaMalloc((void **)dp, n * sizeof(T)) != cudaSuccess) return false;
        return v.empty() || cudaMemcpy(*dp, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
    };
    bool good = up(&ctx->d_put_base, put_base) && up(&ctx->d_put_flag, put_flag) && up(&ctx->d_send_peer, send_peer) && up(&ctx->d_send_idx, send_idx) &&
                up(&ctx->d_n_to_peer, ctx->send_counts) && up(&ctx->d_recv_cnt, ctx->recv_counts);
    good = good && cudaMalloc((void **)&ctx->d_done, 4 * (size_t)W) == cudaSuccess && cudaMemset(ctx->d_done, 0, 4 * (size_t)W) == cudaSuccess;
    if (!good) return mg_fail(ctx, WGPU_ERR_CUDA, "peer-store exchange: out of device memory for the tables");
    ctx->p2p_flag_bytes = flag_bytes;
    ctx->p2p_pool_bytes = pool_bytes;
    ctx->p2p_on = true;
    return WGPU_OK;
}

__global__ void advance_time_kernel(double *t) { t[0] = t[1]; }

}  // namespace

extern "C" {

int32_t wgpu_comm_unique_id(char *id128)
{
    if (!id128) return WGPU_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api) return WGPU_ERR_UNSUPPORTED;
    nccl_id_t id;
    if (api->GetUniqueId(&id) != 0) return WGPU_ERR_CUDA;
    memcpy(id128, id.internal, 128);
    return WGPU_OK;
}

int32_t wgpu_comm_init(wgpu_ctx *ctx, const char *id128, int32_t rank, int32_t world)
{
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return WGPU_ERR_ARG;
    NcclApi *api = nccl_api();
    if (!api) return mg_fail(ctx, WGPU_ERR_UNSUPPORTED, "NCCL is not available in this process");
    if (ctx->comm) return mg_fail(ctx, WGPU_ERR_ARG, "wgpu_comm_init: the context has a communicator already");
    WGPU_CHECK(ctx, cudaSetDevice(ctx->cfg.device));
    nccl_id_t id;
    memcpy(id.internal, id128, 128);
    nccl_comm_t comm = nullptr;
    NCCL_CHECK(ctx, api, api->CommInitRank(&comm, world, id, rank));
    ctx->comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_world = world;
    ctx->send_counts.assign(world, 0);
    ctx->recv_counts.assign(world, 0);
    ctx->rsend_counts.assign(world, 0);
    ctx->rrecv_counts.assign(world, 0);
    {
        // highest priority: the exchange (and the boundary blocks behind it) must not queue behind the interior blocks' CTAs
        int lo = 0, hi = 0;
        WGPU_CHECK(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        WGPU_CHECK(ctx, cudaStreamCreateWithPriority(&ctx->comm_stream, cudaStreamNonBlocking, hi));
    }
    WGPU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming));
    WGPU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_xchg, cudaEventDisableTiming));
    WGPU_CHECK(ctx, cudaMalloc((void **)&ctx->d_comm_scratch, 4096 * 8));
    return WGPU_OK;
}

int32_t wgpu_comm_destroy(wgpu_ctx *ctx)
{
    if (!ctx) return WGPU_ERR_ARG;
    NcclApi *api = nccl_api();
    if (ctx->comm && api) {
        cudaStreamSynchronize(ctx->comm_stream);
        cudaStreamSynchronize(ctx->stream);
        api->CommDestroy(ctx->comm);
    }
    ctx->comm = nullptr;
    if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
    ctx->comm_stream = nullptr;
    if (ctx->ev_pack) cudaEventDestroy(ctx->ev_pack);
    if (ctx->ev_xchg) cudaEventDestroy(ctx->ev_xchg);
    ctx->ev_pack = ctx->ev_xchg = nullptr;
    cudaFree(ctx->d_comm_scratch);
    ctx->d_comm_scratch = nullptr;
    p2p_teardown(ctx);
    cudaFree(ctx->d_xbuf);
    ctx->d_xbuf = nullptr;
    ctx->xbuf_cap = 0;
    ctx->comm_world = 1;
    ctx->comm_rank = 0;
    return WGPU_OK;
}

int32_t wgpu_comm_set_counts(wgpu_ctx *ctx, const int32_t *send_counts, const int32_t *recv_counts, const int32_t *restrict_send_counts,
                             const int32_t *restrict_recv_counts)
{
    if (!ctx || !send_counts || !recv_counts) return WGPU_ERR_ARG;
    if (!ctx->comm) return mg_fail(ctx, WGPU_ERR_ARG, "no communicator: call wgpu_comm_init first");
    const int W = ctx->comm_world;
    long long ns = 0, nr = 0;
    for (int p = 0; p < W; ++p) {
        if (send_counts[p] < 0 || recv_counts[p] < 0) return mg_fail(ctx, WGPU_ERR_ARG, "wgpu_comm_set_counts: negative count");
        ctx->send_counts[p] = send_counts[p];
        ctx->recv_counts[p] = recv_counts[p];
        ctx->rsend_counts[p] = restrict_send_counts ? restrict_send_counts[p] : 0;
        ctx->rrecv_counts[p] = restrict_recv_counts ? restrict_recv_counts[p] : 0;
        ns += send_counts[p];
        nr += recv_counts[p];
    }
    const bool halo = ctx->n_halo_send > 0 || !ctx->h_halo.empty();
    const long long want_s = halo ? ctx->n_halo_send : ctx->n_send, want_r = halo ? (long long)ctx->h_halo.size() : -1;
    if (ns != want_s || (want_r >= 0 && nr != want_r))
        return mg_fail(ctx, WGPU_ERR_ARG, "wgpu_comm_set_counts: the counts do not add up to the lists of wgpu_set_halo / wgpu_set_exchange");
    if (halo || W < 2) {
        p2p_teardown(ctx);
        return WGPU_OK;
    }
    return p2p_setup(ctx);      // face patches: try the peer-store transport (collective; falls back to NCCL send / recv on every rank)
}

int32_t wgpu_comm_set_transport(wgpu_ctx *ctx, int32_t peer_stores)
{
    if (!ctx) return WGPU_ERR_ARG;
    ctx->p2p_want = peer_stores ? 1 : 0;     // takes effect at the next wgpu_comm_set_counts
    return WGPU_OK;
}

int32_t wgpu_comm_transport(const wgpu_ctx *ctx) { return ctx && ctx->p2p_on ? 1 : 0; }

// peer-store variant of stage_exchange: the pack kernel IS the send; it runs on the (high-priority) communication stream, next to the interior
// blocks the main stream starts at once, followed by the wait for the peers' flags and then the partition-boundary blocks
static int32_t stage_exchange_p2p(wgpu_ctx *ctx, int j)
{
    const unsigned seq = ++ctx->p2p_seq;
    const int q = (int)(seq & 1u);
    int32_t rc;
    WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_pack, ctx->stream));                  // the stage input is complete
    WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_pack, 0));
    if ((rc = wgpu_launch_pack_put(ctx, j == 1 ? ctx->U : (((j - 1) & 1) ? ctx->UA : ctx->UB), q, seq, ctx->comm_stream))) return rc;
    if ((rc = wgpu_launch_wait_flags(ctx, seq, ctx->comm_stream))) return rc;
    ctx->d_pool = (double *)(ctx->p2p_mem + ctx->p2p_flag_bytes + (size_t)q * ctx->p2p_pool_bytes);
    WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_xchg, ctx->comm_stream));
    return WGPU_OK;
}

// the exchange of stage j on the communication stream, ordered after the pack kernel and recorded in ev_xchg
static int32_t stage_exchange(wgpu_ctx *ctx, int j)
{
    const bool halo = ctx->n_halo_send > 0 || !ctx->h_halo.empty();
    double *recv = nullptr;
    const double *send = nullptr;
    long long unit;
    if (halo) {
        void *p = nullptr;
        int64_t n = 0;
        int32_t rc = wgpu_rk_stage_halo_pointer(ctx, j, &p, &n);
        if (rc) return rc;
        recv = (double *)p;
        send = ctx->d_halo_send_buf;
        unit = (long long)ctx->nc * ctx->blk_elems;
    } else {
        recv = ctx->d_pool;
        send = ctx->d_send_buf;
        unit = wgpu_patch_doubles(ctx);
    }
    WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_pack, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_pack, 0));
    int32_t rc = exchange(ctx, send, recv, ctx->send_counts, ctx->recv_counts, unit, ctx->comm_stream);
    if (rc) return rc;
    WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_xchg, ctx->comm_stream));
    return WGPU_OK;
}

// WGPU_MG_TRACE=1: timestamps (CUDA events) of the phases of the LAST step of a wgpu_rk_steps call, printed by rank 0 to stderr
struct MgTrace {
    std::vector<cudaEvent_t> ev;
    std::vector<std::string> name;
    bool on = false;
    void mark(const char *n, cudaStream_t s)
    {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.push_back(e);
        name.push_back(n);
    }
    void dump()
    {
        if (!on || ev.empty()) return;
        cudaEventSynchronize(ev.back());
        fprintf(stderr, "wgpu_rk_steps trace (last step), ms since step start:\n");
        for (size_t i = 0; i < ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[0], ev[i]);
            fprintf(stderr, "  %8.3f  %s\n", ms, name[i].c_str());
            }
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        name.clear();
    }
};

int32_t wgpu_rk_steps(wgpu_ctx *ctx, double time, int32_t n_steps, double *time_out, double *dt_last)
{
    if (!ctx || n_steps < 1) return WGPU_ERR_ARG;
    MgTrace tr;
    const bool want_trace = getenv("WGPU_MG_TRACE") && ctx->comm_rank == 0 && n_steps > 3;
    const wgpu_config &c = ctx->cfg;
    const bool multi = ctx->comm && ctx->comm_world > 1;
    NcclApi *api = multi ? nccl_api() : nullptr;
    int32_t rc;
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_time, &time, 8, cudaMemcpyHostToDevice, ctx->stream));
    ctx->time_on_device = true;
    struct Reset {
        wgpu_ctx *c;
        double *pool;
        ~Reset()
        {
            c->time_on_device = false;
            c->d_pool = pool;          // the peer-store transport points the stage kernels at the library's own pools
        }
    } reset{ctx, ctx->d_pool};
    for (int step = 0; step < n_steps; ++step) {
        tr.on = want_trace && step == n_steps - 1;
        tr.mark("step start", ctx->stream);
        if (step > 0) {
            advance_time_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_time);
            ctx->launches++;
        }
        if ((rc = wgpu_rk_begin(ctx, time))) return rc;
        if (multi && !(c.dt_fixed > 0.0)) {
            // MPI_Allreduce(MIN) of calculate_time_step.f90:48 on the device scalar: positive doubles order like their bit patterns
            void *p = nullptr;
            if ((rc = wgpu_dtmin_pointer(ctx, &p))) return rc;
            NCCL_CHECK(ctx, api, api->AllReduce(p, p, 1, NCCL_UINT64, NCCL_MIN, ctx->comm, ctx->stream));
        }
        if ((rc = wgpu_rk_dt(ctx, time))) return rc;
        tr.mark("dt all-reduced + finalised", ctx->stream);
        for (int j = 1; j <= c.n_stages; ++j) {
            if (multi && ctx->p2p_on) {
                if ((rc = stage_exchange_p2p(ctx, j))) return rc;
            } else {
                if ((rc = wgpu_pack_halo(ctx, j))) return rc;
                tr.mark("  packed", ctx->stream);
                if (!multi) {
                    if ((rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_ALL))) return rc;
                    continue;
                }
                if ((rc = stage_exchange(ctx, j))) return rc;
            }
            if (ctx->n_bnd && ctx->n_int && ctx->n_jump == 0) {
                // uniform grid: the partition-boundary blocks run on the communication stream right behind the exchange, CONCURRENTLY with
                // the interior blocks on the main stream (the two launches share the SMs: one tail instead of two partial last waves);
                // the next stage's pack waits for both
                tr.mark("  exchange done (comm stream)", ctx->comm_stream);
                if ((rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_INTERIOR))) return rc;
                tr.mark("  interior done", ctx->stream);
                cudaStream_t main_stream = ctx->stream;
                ctx->stream = ctx->comm_stream;
                rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_BOUNDARY);
                ctx->stream = main_stream;
                if (rc) return rc;
                tr.mark("  boundary done (comm stream)", ctx->comm_stream);
                WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_xchg, ctx->comm_stream));
                WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_xchg, 0));
            } else if (ctx->n_bnd && ctx->n_int) {
                if ((rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_INTERIOR))) return rc;   // while the blocks are in flight
                WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_xchg, 0));
                if ((rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_BOUNDARY))) return rc;   // (level-jump patches are refreshed in between: same stream)
            } else {
                WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_xchg, 0));
                if ((rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_ALL))) return rc;
            }
        }
        if ((rc = wgpu_rk_end_nosync(ctx))) return rc;
        tr.mark("step end", ctx->stream);
    }
    tr.dump();
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_dt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 1, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 2, ctx->d_time + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 3, ctx->d_flags + 5, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (dt_last) *dt_last = ctx->h_pinned[0];
    if (time_out) *time_out = ctx->h_pinned[2];
    if (*(int *)(ctx->h_pinned + 3)) {
        cudaMemsetAsync(ctx->d_flags + 5, 0, sizeof(int), ctx->stream);
        return mg_fail(ctx, WGPU_ERR_CUDA, "peer-store exchange: a peer's ghost patches did not arrive within 5 s");
    }
    if (*(int *)(ctx->h_pinned + 1)) {
        cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream);
        return mg_fail(ctx, WGPU_ERR_DIVERGED, "ACM fail: very very large values in state vector.");
    }
    return WGPU_OK;
}

int32_t wgpu_exchange_array(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t filtered)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (!ctx->comm || ctx->comm_world < 2) return WGPU_OK;
    int32_t rc;
    if ((rc = wgpu_pack_blocks(ctx, array_id, slot))) return rc;
    void *p = nullptr;
    int64_t n = 0;
    if ((rc = wgpu_halo_pointer(ctx, array_id, slot, &p, &n))) return rc;
    const long long unit = (long long)ctx->nc * ctx->blk_elems;
    if ((rc = exchange(ctx, ctx->d_halo_send_buf, (double *)p, ctx->send_counts, ctx->recv_counts, unit, ctx->stream))) return rc;
    if (filtered && ctx->wavelet_set && ctx->wavelet.Y != 0 && !ctx->ignore_filter) {
        if ((rc = wgpu_restrict_pack(ctx, array_id, slot))) return rc;
        if ((rc = wgpu_restrict_halo_pointer(ctx, &p, &n))) return rc;
        const long long runit = (long long)ctx->nc * (ctx->blk_elems >> ctx->cfg.dim);
        if ((rc = exchange(ctx, ctx->d_rhalo_send_buf, (double *)p, ctx->rsend_counts, ctx->rrecv_counts, runit, ctx->stream))) return rc;
    }
    return WGPU_OK;
}

int32_t wgpu_ship_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n_items, const int32_t *src_rank, const int32_t *src_slot,
                         const int32_t *dst_rank, int32_t first_free, int32_t *local_slot, int32_t *next_free)
{
    if (!ctx || n_items < 0 || (n_items > 0 && (!src_rank || !src_slot || !dst_rank || !local_slot)) || !next_free || first_free < 1) return WGPU_ERR_ARG;
    const int W = ctx->comm ? ctx->comm_world : 1, me = ctx->comm ? ctx->comm_rank : 0;
    const int N = ctx->cfg.max_blocks;
    std::vector<std::vector<int>> send_ids(W);
    std::vector<int> send_counts(W, 0), recv_counts(W, 0);
    // pass 1: what I send (item order per peer) and how many I receive per peer
    for (int k = 0; k < n_items; ++k) {
        const int s = src_rank[k], d = dst_rank[k];
        if (s < 0 || s >= W || d < 0 || d >= W) return mg_fail(ctx, WGPU_ERR_ARG, "wgpu_ship_blocks: rank out of range");
        if (s == me && d != me) {
            if (src_slot[k] < 1 || src_slot[k] > N) return mg_fail(ctx, WGPU_ERR_ARG, "wgpu_ship_blocks: source slot out of range");
            send_ids[d].push_back(src_slot[k]);
        }
        if (d == me && s != me) recv_counts[s]++;
    }
    // pass 2: local slot of every item that ends up here: its own slot, or the next free slot in (peer, item) order
    std::vector<int> base(W, 0);
    int nxt = first_free;
    for (int p = 0; p < W; ++p) {
        base[p] = nxt;
        nxt += recv_counts[p];
    }
    if (nxt - 1 > N) return mg_fail(ctx, WGPU_ERR_ARG, "wgpu_ship_blocks: the transfer needs more slots than max_blocks");
    std::vector<int> cur(base);
    int m = 0;
    for (int k = 0; k < n_items; ++k) {
        if (dst_rank[k] != me) continue;
        local_slot[m++] = src_rank[k] == me ? src_slot[k] : cur[src_rank[k]]++;
    }
    *next_free = nxt;
    std::vector<int> flat;
    for (int p = 0; p < W; ++p) {
        send_counts[p] = (int)send_ids[p].size();
        flat.insert(flat.end(), send_ids[p].begin(), send_ids[p].end());
    }
    const int n_send = (int)flat.size(), n_recv = nxt - first_free;
    if (W == 1 || (!n_send && !n_recv)) return WGPU_OK;
    int32_t rc;
    void *arr = nullptr;
    int64_t nd = 0;
    if ((rc = wgpu_device_pointer(ctx, array_id, slot, &arr, &nd))) return rc;
    const long long unit = (long long)ctx->nc * ctx->blk_elems;
    if ((rc = ensure_buf(ctx, &ctx->d_xbuf, &ctx->xbuf_cap, (size_t)std::max(n_send, 1) * unit))) return rc;
    if (n_send && (rc = wgpu_gather_blocks(ctx, array_id, slot, n_send, flat.data(), ctx->d_xbuf))) return rc;
    // received blocks land in consecutive free slots of the array itself, peer by peer: no scatter pass
    double *recv = (double *)arr + (long long)(first_free - 1) * unit;
    if ((rc = exchange(ctx, ctx->d_xbuf, recv, send_counts, recv_counts, unit, ctx->stream))) return rc;
    if (n_recv) {
        if (arr == ctx->U) ctx->dtmin_valid = false;
        ctx->det_cached_for = nullptr;
    }
    return WGPU_OK;
}

int32_t wgpu_comm_allreduce(wgpu_ctx *ctx, double *inout, int32_t n, int32_t op)
{
    if (!ctx || n < 0 || (n > 0 && !inout) || n > 4096) return WGPU_ERR_ARG;
    if (!ctx->comm || ctx->comm_world < 2 || n == 0) return WGPU_OK;
    NcclApi *api = nccl_api();
    const int nop = op == 0 ? NCCL_MAX : (op == 1 ? NCCL_MIN : NCCL_SUM);
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_comm_scratch, inout, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, api, api->AllReduce(ctx->d_comm_scratch, ctx->d_comm_scratch, (size_t)n, NCCL_FLOAT64, nop, ctx->comm, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(inout, ctx->d_comm_scratch, 8 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return WGPU_OK;
}

int32_t wgpu_comm_allgatherv_i32(wgpu_ctx *ctx, const int32_t *mine, const int32_t *counts, int32_t *out)
{
    if (!ctx || !counts || !out) return WGPU_ERR_ARG;
    const int W = ctx->comm ? ctx->comm_world : 1, me = ctx->comm ? ctx->comm_rank : 0;
    long long total = 0, off_me = 0;
    for (int p = 0; p < W; ++p) {
        if (p == me) off_me = total;
        total += counts[p];
    }
    if (counts[me] > 0 && !mine) return WGPU_ERR_ARG;
    if (W == 1) {
        memcpy(out, mine, sizeof(int32_t) * (size_t)counts[0]);
        return WGPU_OK;
    }
    NcclApi *api = nccl_api();
    // variable counts: every rank contributes its segment to a zero-initialised array, summed over the ranks
    int32_t rc;
    const size_t words = ((size_t)total + 1) / 2 + 1;
    if ((rc = ensure_buf(ctx, &ctx->d_xbuf, &ctx->xbuf_cap, words))) return rc;
    int *d = (int *)ctx->d_xbuf;
    WGPU_CHECK(ctx, cudaMemsetAsync(d, 0, sizeof(int) * (size_t)total, ctx->stream));
    if (counts[me]) WGPU_CHECK(ctx, cudaMemcpyAsync(d + off_me, mine, sizeof(int) * (size_t)counts[me], cudaMemcpyHostToDevice, ctx->stream));
    NCCL_CHECK(ctx, api, api->AllReduce(d, d, (size_t)total, NCCL_INT32, NCCL_SUM, ctx->comm, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(out, d, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return WGPU_OK;
}

int32_t wgpu_comm_info(const wgpu_ctx *ctx, int32_t *rank, int32_t *world)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (rank) *rank = ctx->comm ? ctx->comm_rank : 0;
    if (world) *world = ctx->comm ? ctx->comm_world : 1;
    return WGPU_OK;
}

}  // extern "C"
