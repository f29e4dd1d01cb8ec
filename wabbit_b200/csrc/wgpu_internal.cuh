// Internal definitions shared by the CUDA translation units of libwabbit_gpu.so.
// Not part of the public C ABI (include/wabbit_gpu.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "wabbit_gpu.h"

#define WGPU_MAX_LEVELS 32

// ------------------------------------------------------------------------------------------------
// Neighbour direction table.  Device-side the 168-slot hvy_neighbor is folded into 26 directions
// (index = (dz+1)*9 + (dy+1)*3 + (dx+1), 13 = self) with one int32 "source code" each:
//    >= 0 : same-level neighbour resident on this GPU: its 0-based block index (gather from its interior)
//    -1   : no neighbour (non-periodic domain boundary)
//    <=-2 : patch pool entry  pid = -2 - code  (values prepared by a pre-pass: restriction,
//           prediction or a remote GPU's pack kernel); pool_off[pid] is the offset in doubles.
// ------------------------------------------------------------------------------------------------
#define WGPU_NDIR 27
#define WGPU_JUMP_PID (1 << 28)   // pool patch ids >= this are level-jump patches (index = pid - WGPU_JUMP_PID) in the jump pool
#define WGPU_FMAX 12   // largest |tap index| of a wavelet filter

// filter banks of a biorthogonal CDF wavelet (setup_wavelet, LIB/WAVELETS/module_wavelets.f90:1031-1290); tap k at [k+WGPU_FMAX]
struct WaveFilters {
    int X, Y;
    int hd_lo, hd_hi, gd_lo, gd_hi, hr_lo, hr_hi, gr_lo, gr_hi;
    double HD[2 * WGPU_FMAX + 1], GD[2 * WGPU_FMAX + 1], HR[2 * WGPU_FMAX + 1], GR[2 * WGPU_FMAX + 1];
};

struct StageArgs {
    // fields
    const double *u_in;        // stage input, compact [blk][nc][Bs^3]
    const double *u0;          // state at start of the step (compact)
    double *k_out;             // RHS output slot (may be nullptr: not stored)
    double *u_out;             // next stage input / final state (may be nullptr)
    const double *k_prev[WGPU_MAX_STAGES];  // earlier slopes entering u_out
    double coef_prev[WGPU_MAX_STAGES];      // Butcher coefficients (dt applied in-kernel: (dt*a)*k)
    int n_prev;
    double coef_self;          // coefficient of the slope computed by this launch
    int use_self;
    // running final combination ("acc mode", tableaus whose stage inputs only use the previous slope):
    //   acc_out = acc_in + (dt*coef_acc)*k_this   (acc_in == u_in: take the stage input itself)
    const double *acc_in;
    double *acc_out;           // may alias acc_in (each thread reads and writes only its own point)
    double coef_acc;
    int use_acc;
    const double *dt_ptr;      // device scalar
    const double *mask;        // compact [blk][n_mask][Bs^3] or nullptr
    int n_mask;
    // topology
    const int *active;         // [n_active] 0-based block indices
    const int *nbr;            // [max_blocks][27]
    const signed char *level;  // [max_blocks]
    const double *pool;
    const long long *pool_off;
    const double *jpool;       // level-jump face patches (restriction / prediction), patch i at i*jpatch doubles
    long long jpatch;
    // physics
    double dx_lvl[WGPU_MAX_LEVELS][3];
    double c0, nu, gamma_p, C_eta_inv, C_sponge_inv, u_mean_set[3];
    int use_sponge;
    // reductions
    unsigned long long *dtmin_bits;  // atomicMin target for the NEXT step's CFL dt (final stage only), or nullptr
    double CFL, CFL_nu;
    int *diverged;                   // set to 1 if any |u_in| > 1e12
    int dim_min_axes;                // number of axes entering minval(dx(1:dim))
    // analytic mask (wgpu_set_mask_sphere): the penalization term of a translating sphere evaluated in the kernel instead of read from hvy_mask
    int geom;                        // 0: mask arrays, 1: sphere
    const int *ixyz;                 // [max_blocks][3] block coordinates on their level
    double g_c0[3], g_v[3], g_R, g_h;
    double t0, t_cj;                 // stage time = t0 + t_cj * dt
    const double *t0_ptr;            // != nullptr: t0 is read from the device (wgpu_rk_steps)
    int skip_plain;                  // leave blocks whose six face neighbours are resident same-level blocks (stage_kernel_tma took them)
    int plain_hint;                  // 1: every block of this launch is such a block, -1: none is, 0: unknown / mixed
};

#define WGPU_NSTAT 23
// closed-form mask geometry of create_mask_kernel (statistics.cu)
struct MaskGeom {
    int penalization, use_sponge;
    double domain[3], c0[3], v[3], R, h, L_sponge, p_sponge;
};
struct StatArgs {
    double domain[3], c0, gamma_p, C_eta_inv, C_sponge_inv, u_mean_set[3];
    int use_sponge;
};
struct VortArgs {            // vort_block_kernel: FD1 / FD2 taps -H..H of the discretization (module_operators.f90:23-33)
    double domain[3], nu, fd1[7], fd2[7];
    int H;
};

struct wgpu_ctx {
    wgpu_config cfg;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    int64_t dev_bytes = 0;

    int nc = 0;                 // n_eqn
    int64_t blk_elems = 0;      // Bs^3 (or Bs^2)
    int64_t gblk_elems = 0;     // (Bs+2g)^3

    // resident arrays (compact interior layout [blk][comp][z][y][x])
    double *U = nullptr;        // hvy_block
    double *UA = nullptr, *UB = nullptr;   // stage inputs (ping-pong)
    double *K[WGPU_MAX_STAGES] = {nullptr};  // hvy_work slots 2..s+1
    double *MASK = nullptr;
    double *TMP = nullptr;

    // topology
    int n_active = 0;
    int *d_active = nullptr;
    int *d_nbr = nullptr;
    signed char *d_level = nullptr;
    std::vector<int> h_active;
    std::vector<int> h_nbr;
    std::vector<signed char> h_level;
    double *d_pool = nullptr;          // receive buffer of remote face patches (owned by the caller)
    long long *d_pool_off = nullptr;
    // multi-GPU exchange
    std::vector<int> remote_faces;     // (block, dir) pairs with a same-level neighbour on another rank, pending wgpu_set_exchange
    int n_int = 0, n_bnd = 0;          // interior / boundary split of the active list
    int *d_active_int = nullptr, *d_active_bnd = nullptr;
    int n_send = 0;
    int *d_send_blk = nullptr, *d_send_dir = nullptr;
    double *d_send_buf = nullptr;
    // halo blocks: copies of blocks owned by other ranks held in local slots (the slots right behind each other, in the order the
    // owners send them); they are known to the neighbour table and the block lookup but never computed
    std::unordered_map<int, int> halo_map;   // lgt id (1-based) -> local 0-based block index
    std::vector<int> h_halo;                 // 0-based block indices of the halo slots, in receive order
    std::vector<signed char> halo_level_of;  // their mesh levels
    std::vector<int> halo_bnd;               // active blocks with a neighbour in a halo slot
    int n_halo_send = 0;
    int *d_halo_send = nullptr, *d_iota = nullptr;
    int halo_send_cap = 0;
    double *d_halo_send_buf = nullptr;       // owned by the caller
    bool halo_fine_neighbor = false;         // a local block has a FINER neighbour that is a halo block (filtered restriction unavailable)
    // block coordinates + device lookup (level, ix, iy, iz) -> block index, for level-jump patches
    std::vector<int> h_ixyz;           // [max_blocks][3], valid where coords_of[b] != 0
    std::vector<char> h_has_coords;
    std::vector<signed char> h_tc_level;   // level given to wgpu_set_treecodes (also for blocks that are sources only)
    int *d_ixyz = nullptr;
    unsigned long long *d_hkeys = nullptr;
    int *d_hvals = nullptr;
    unsigned hmask = 0;
    size_t hcap = 0;
    // level-jump face patches of the stage kernel (refreshed from the stage input before every stage)
    int n_jump = 0, jump_cap = 0;
    int *d_jump_blk = nullptr, *d_jump_dir = nullptr;
    double *d_jpool = nullptr;
    size_t jpool_cap = 0;
    // level-jump ghost patches of the wavelet kernels: all 26 relations, wjump_depth deep (the widest wavelet filter)
    int n_wjump = 0, wjump_cap = 0, wjump_depth = 0;
    int *d_wjump_blk = nullptr, *d_wjump_dir = nullptr, *d_wnbr = nullptr;
    long long *d_woff = nullptr;
    double *d_wpool = nullptr;
    size_t wpool_cap = 0;
    // coarse extension: (block, direction) pairs whose neighbour is coarser
    int n_ce = 0, ce_cap = 0;
    int *d_ce_blk = nullptr, *d_ce_dir = nullptr;
    // HD-filtered restriction of sync_ghosts_tree (lifted wavelets): the blocks that have a coarser neighbour, the directions in which
    // each of them has a coarser or finer neighbour (bit (dz+1)*9+(dy+1)*3+(dx+1)), and their filtered + decimated copies
    int n_rst = 0, rst_cap = 0;
    int *d_rst_blk = nullptr, *d_rmap = nullptr;
    std::vector<int> h_rmap;           // host copy of d_rmap: block -> rpool entry or -1
    unsigned *d_rst_mask = nullptr;
    double *d_rpool = nullptr;
    size_t rpool_cap = 0;
    // filtered copies of FINER neighbours owned by other ranks: received behind the rank's own entries of rpool (wgpu_set_halo_restrict)
    int n_rhalo_recv = 0, n_rhalo_send = 0, rhalo_send_cap = 0;
    int *d_rhalo_send = nullptr;       // rpool indices of the entries other ranks need
    double *d_rhalo_send_buf = nullptr;   // owned by the caller
    bool ignore_filter = false;        // wavelet-side syncs behave like sync_ghosts_tree(ignore_Filter = .true.)
    bool has_jumps = false;            // some active block has a coarser / finer neighbour
    bool lookup_ready = false;         // block lookup + coordinates of the current topology are on the device
    int act_lo = 0, act_hi = 0;        // [act_lo, act_hi): range of block ids of the current active list (rows of nbr / wnbr that are kept up to date)
    int geom = 0;                      // analytic mask geometry (0 none, 1 sphere) and its parameters
    double g_c0[3] = {0, 0, 0}, g_v[3] = {0, 0, 0}, g_R = 0, g_h = 1;
    bool coords_dirty = true;          // wgpu_set_treecodes changed the block positions since the lookup table was last uploaded
    // topology derived on the device (topology.cu: wgpu_set_grid / wgpu_set_active)
    bool topo_on_device = false;       // the current tables come from wgpu_set_grid (block lookup, level, positions are registered on the device)
    bool rmap_on_device = false;       // d_rmap was filled on the device (h_rmap is not kept)
    unsigned char *d_bflag = nullptr;  // [max_blocks] bit0 registered, bit1 halo copy
    unsigned char *d_rel = nullptr;    // [n_active][27] relation of the active blocks
    size_t rel_cap = 0;
    long long *d_cnt = nullptr;        // per-block counts / scanned offsets of the list compaction
    size_t cnt_cap = 0;
    char *d_topo_in = nullptr;         // staging of the (id, level, treecode) lists
    size_t topo_in_cap = 0;
    int *d_halo_ids = nullptr;
    size_t halo_ids_cap = 0;
    double *d_pd_out = nullptr;        // result scratch of wgpu_patch_details
    size_t pd_cap = 0;
    int *d_idbuf[3] = {nullptr, nullptr, nullptr};   // scratch id lists (refine / coarsen)
    size_t idbuf_cap[3] = {0, 0, 0};
    // Runge-Kutta step in flight
    const double *rk_uin = nullptr;
    int rk_next_stage = 0;
    double rk_time = 0.0;              // time at the start of the step in flight (stage times of the analytic mask)
    bool rk_subdiag = false;

    // scalars
    double *d_dt = nullptr;                   // dt of the current step
    double *d_time = nullptr;                 // [2]: time at the start of the step in flight / after it (wgpu_rk_steps keeps the time on the device)
    bool time_on_device = false;
    unsigned long long *d_dtmin = nullptr;    // [2]: CFL dt candidates (bits), ping-pong
    int dtmin_cur = 0;
    bool dtmin_valid = false;
    int *d_flags = nullptr;                   // [8]: [0] diverged, [2..4] topology.cu (duplicate position, jump flags, bad send list)
    double *h_pinned = nullptr;               // [0] dt, [1] flags (as int bits)

    // wavelets
    WaveFilters wavelet;
    bool wavelet_set = false;
    double *d_det_abs = nullptr, *d_det_sq = nullptr;   // [max_blocks][nc]
    const double *det_cached_for = nullptr;             // decomposed array whose Linfty details d_det_* currently hold (fused into the FWT)
    int *d_status = nullptr;                            // [max_blocks]
    double *d_detail_out = nullptr;                     // [max_blocks][nc]
    unsigned long long *d_norm = nullptr;               // [16]

    // multi-GPU inside the library (multigpu.cu): the NCCL communicator, the per-peer counts of the declared exchange, a second stream
    void *comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    std::vector<int> send_counts, recv_counts, rsend_counts, rrecv_counts;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_pack = nullptr, ev_xchg = nullptr;
    double *d_comm_scratch = nullptr;
    double *d_xbuf = nullptr;          // send buffer of wgpu_ship_blocks / scratch of the light-data collectives
    size_t xbuf_cap = 0;

    // face-patch exchange by peer stores over NVLink (multigpu.cu: p2p_setup): the library owns the receive pools (two, alternating per
    // stage) and a flag word per peer in ONE allocation exported by CUDA IPC; the pack kernel writes every patch straight into the
    // receiver's pool and the last CTA per peer releases that peer's flag; the receiver spins on its flags in front of the boundary blocks
    bool p2p_on = false;
    int p2p_want = 1;                  // wgpu_comm_set_transport: 1 peer stores when available, 0 NCCL send / recv
    char *p2p_mem = nullptr;
    size_t p2p_flag_bytes = 0, p2p_pool_bytes = 0;
    std::vector<void *> p2p_peer;      // opened allocations of the peers this rank sends to
    double **d_put_base = nullptr;     // [2][world] where this rank's patches start in peer p's pool q
    unsigned **d_put_flag = nullptr;   // [world] this rank's flag word on peer p
    int *d_send_peer = nullptr, *d_send_idx = nullptr, *d_n_to_peer = nullptr, *d_recv_cnt = nullptr;
    unsigned *d_done = nullptr;        // [world] CTAs of the running pack launch that have finished, per peer
    unsigned p2p_seq = 0;              // stages exchanged so far (flag value of the next one = p2p_seq + 1)
    double *d_pool_user = nullptr;     // the caller's pool of wgpu_set_exchange (used by the NCCL / host-driven paths)

    double *d_stat = nullptr;          // [max_blocks + 1][WGPU_NSTAT]: per-block partial statistics, the result behind them
    std::vector<double *> kry;         // wgpu_krylov_step: M_krylov + 3 registers (Krylov vectors, perturbed state, reference right-hand side)
    double *d_kry_part = nullptr;      // [max_blocks + 1]: per-block partial scalar products, the result behind them
    void *tma_cache = nullptr;         // tensor maps of the resident arrays (kernels.cu: tma_maps)

    // optional event pairs around stage launches
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;   // [2*i], [2*i+1]
    int prof_n = 0;

    // staging for upload/download
    double *d_stage = nullptr;
    int64_t stage_elems = 0;
    double *h_bounce = nullptr;
    int64_t bounce_elems = 0;
    // copy-engine transfers of page-locked host arrays (move_blocks_dma): mode per direction (0: SM-issued zero-copy kernels, 1: DMA of
    // plane spans + a layout kernel), two staging buffers, a stream per direction
    int xfer_dma[2] = {1, 1};          // [0] upload, [1] download
    char *d_span[2] = {nullptr, nullptr};
    size_t span_cap = 0;
    cudaStream_t dma_stream = nullptr;
    cudaEvent_t ev_dma[2] = {nullptr, nullptr}, ev_lay[2] = {nullptr, nullptr};
};

#define WGPU_CHECK(ctx, call)                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                  \
            return WGPU_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

// capi.cu
int32_t wgpu_rk_end_nosync(wgpu_ctx *ctx);
// kernels.cu
int32_t wgpu_launch_stage(wgpu_ctx *ctx, const StageArgs &a, int n_blocks);
void wgpu_tma_release(wgpu_ctx *ctx);
int32_t wgpu_launch_rkc_combine(wgpu_ctx *ctx, double *out, const double *y00, const double *y1, const double *y0, const double *f1, const double *f0,
                                double cA, double cB, double cC, double cD, double cE, int mode);
int32_t wgpu_launch_pack(wgpu_ctx *ctx, const double *src);
int32_t wgpu_launch_pack_put(wgpu_ctx *ctx, const double *src, int parity, unsigned seq, cudaStream_t st);
int32_t wgpu_launch_wait_flags(wgpu_ctx *ctx, unsigned seq, cudaStream_t st);
// jump.cu
int32_t wgpu_launch_jump_fill(wgpu_ctx *ctx, const double *src);
int32_t wgpu_launch_wjump_fill(wgpu_ctx *ctx, const double *src, const double *ce_coarse = nullptr);
int32_t wgpu_launch_restrict_filter(wgpu_ctx *ctx, const double *src, int nc_src, bool *active);
int32_t wgpu_launch_ce(wgpu_ctx *ctx, double *wd, const double *orig, int Nwcl, int Nwcr, int Nscl, int Nscr, int clear_wc, int copy_sc);
int32_t wgpu_launch_export_regions(wgpu_ctx *ctx, const double *src, double *staged, const int *d_ids, int n, int ncomp_src, int ncomp_host,
                                   int g_sync, int by_id);
int32_t wgpu_launch_refine(wgpu_ctx *ctx, const double *src, double *dst, const int *d_mother, const int *d_daughter, int n);
int32_t wgpu_launch_coarsen(wgpu_ctx *ctx, const double *src, double *dst, const int *d_mother, const int *d_daughter, int n);
int32_t wgpu_launch_copy_blocks(wgpu_ctx *ctx, const double *src, double *dst, const int *d_src_ids, const int *d_dst_ids, int n);
int32_t wgpu_launch_copy_entries(wgpu_ctx *ctx, const double *src, double *dst, const int *d_src_idx, int n, long long per_entry);
// topology.cu
int32_t wgpu_topology_halo_restrict(wgpu_ctx *ctx, const std::vector<int> &recv0, const std::vector<int> &send0);
// statistics.cu
int32_t wgpu_launch_create_mask(wgpu_ctx *ctx, const MaskGeom &gm, double time);
int32_t wgpu_launch_stats(wgpu_ctx *ctx, const double *u, const double *rhs, const double *mask, const StatArgs &sa, double *d_part);
int32_t wgpu_launch_vort_stats(wgpu_ctx *ctx, const double *staged, const int *d_ids, int m, int ncomp, const VortArgs &va, double *d_part);
int32_t wgpu_launch_stats_final(wgpu_ctx *ctx, const double *d_part, double *d_out);
int32_t wgpu_launch_kry_dot(wgpu_ctx *ctx, const double *x, const double *y, double *d_part, double *d_out);
int32_t wgpu_launch_kry_axpy(wgpu_ctx *ctx, double *dst, const double *x, const double *y, int op, double a, double b);
// wavelet.cu
int32_t wgpu_launch_wavelet(wgpu_ctx *ctx, const double *src, double *dst, int inverse, const double *ce_coarse);
int32_t wgpu_launch_blockfilter(wgpu_ctx *ctx, const double *src, double *dst, const double *stencil, int half, unsigned comp_mask, int level_mode);
int32_t wgpu_launch_detail(wgpu_ctx *ctx, const double *wd, int eps_norm, int level_ref);
int32_t wgpu_launch_patch_detail(wgpu_ctx *ctx, const double *wd, const int *d_blk, const int *d_dir, int n, int Nl, int Nr, double *d_out, int eps_norm,
                                 int level_ref);
int32_t wgpu_launch_flags(wgpu_ctx *ctx, const int32_t *thresh_comp, const double *eps_use, int *d_status, double *d_detail_out);
int32_t wgpu_launch_linfty(wgpu_ctx *ctx, const double *u, unsigned long long *d_out);
int32_t wgpu_launch_blocksum(wgpu_ctx *ctx, const double *u, int squared, double *d_out);
int32_t wgpu_launch_dtmin(wgpu_ctx *ctx, const double *u, unsigned long long *dtmin_bits);
int32_t wgpu_launch_dt_finalize(wgpu_ctx *ctx, double time, const unsigned long long *dtmin_bits,
                                unsigned long long *dtmin_next);
int32_t wgpu_launch_extract(wgpu_ctx *ctx, const double *staged, double *dst, const int *d_ids, int n, int ncomp_dst,
                            int ncomp_host, int by_id);
int32_t wgpu_launch_export(wgpu_ctx *ctx, const double *src, double *staged, const int *d_ids, int n, int ncomp_src,
                           int ncomp_host, int g_sync, int by_id);
int32_t wgpu_launch_span_unpack(wgpu_ctx *ctx, const double *stg, double *dst, const int *d_ids, int n, int nc, long long pitch, cudaStream_t st);
int32_t wgpu_launch_span_pack(wgpu_ctx *ctx, const double *src, double *stg, const int *d_ids, int n, int nc, long long pitch, cudaStream_t st);
