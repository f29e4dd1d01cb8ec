// C ABI of libwabbit_gpu.so (include/wabbit_gpu.h): context, resident arrays, topology upload,
// host<->device block movement and the Runge-Kutta driver that sequences the stage kernels.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>

#include "resolve.cuh"
#include "wgpu_internal.cuh"

static std::string g_create_err;

namespace {

int32_t fail(wgpu_ctx *ctx, int32_t code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    else g_create_err = msg;
    return code;
}

template <typename T>
int32_t dmalloc(wgpu_ctx *ctx, T **p, size_t n)
{
    *p = nullptr;
    if (n == 0) return WGPU_OK;
    WGPU_CHECK(ctx, cudaMalloc((void **)p, n * sizeof(T)));
    ctx->dev_bytes += (int64_t)(n * sizeof(T));
    return WGPU_OK;
}

// same-level neighbour code of a direction, as assigned by find_neighbor (LIB/MESH/find_neighbors.f90:60-95)
int same_level_code(const int d[3])
{
    const int nzero = (d[0] == 0) + (d[1] == 0) + (d[2] == 0);
    if (nzero == 2) {
        int code = 1;
        for (int i = 0; i < 3; ++i) {
            if (d[i] != 0) code += 8 * i;
            if (d[i] == 1) code += 4;
        }
        return code;
    }
    if (nzero == 1) {
        int code = 25, apply_free = 1;
        for (int i = 0; i < 3; ++i) {
            if (d[i] == 0) code += 8 * (3 - (i + 1));
            else {
                if (d[i] == 1) code += apply_free * 2;
                apply_free++;
            }
        }
        return code;
    }
    int code = 49;
    for (int i = 0; i < 3; ++i)
        if (d[i] == 1) code += 1 << i;
    return code;
}

// hvy_work slots are allocated on first use: the Runge-Kutta driver needs only slot 2 (as the running final combination) for
// tableaus whose stage inputs use the previous slope only (RK4, Heun, midpoint, Euler); general tableaus and explicit
// wgpu_rhs / wgpu_upload / wavelet calls that name a slot allocate it then.  nullptr (+ ctx->err) if the device is out of memory.
double *ensure_K(wgpu_ctx *ctx, int idx)
{
    if (idx < 0 || idx >= ctx->cfg.n_stages) return nullptr;
    if (!ctx->K[idx]) {
        const size_t n = (size_t)ctx->cfg.max_blocks * ctx->nc * ctx->blk_elems;
        if (dmalloc(ctx, &ctx->K[idx], n) != WGPU_OK) return nullptr;
        if (cudaMemsetAsync(ctx->K[idx], 0, n * sizeof(double), ctx->stream) != cudaSuccess) return nullptr;
    }
    return ctx->K[idx];
}

double *array_ptr(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int *ncomp)
{
    const int s = ctx->cfg.n_stages;
    switch (array_id) {
    case WGPU_HVY_BLOCK: *ncomp = ctx->nc; return ctx->U;
    case WGPU_HVY_WORK:
        *ncomp = ctx->nc;
        if (slot == 1) return ctx->U;   // hvy_work(...,1) is the copy of the state (runge_kutta_generic.f90:63-67)
        if (slot >= 2 && slot <= s + 1) return ensure_K(ctx, slot - 2);
        return nullptr;
    case WGPU_HVY_MASK: *ncomp = ctx->cfg.n_mask; return ctx->MASK;
    case WGPU_HVY_TMP: *ncomp = ctx->nc; return ctx->TMP;
    }
    return nullptr;
}

void fill_common_args(wgpu_ctx *ctx, StageArgs &a)
{
    const wgpu_config &c = ctx->cfg;
    memset(&a, 0, sizeof(a));
    a.active = ctx->d_active;
    a.nbr = ctx->d_nbr;
    a.level = ctx->d_level;
    a.pool = ctx->d_pool;
    a.pool_off = ctx->d_pool_off;
    a.jpool = ctx->d_jpool;
    a.jpatch = (long long)ctx->nc * (c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3)) * c.Bs[0] * (c.dim == 3 ? c.Bs[1] : 1);   // = JumpArgs::jpatch (jump.cu)
    for (int l = 0; l < WGPU_MAX_LEVELS; ++l)
        for (int d = 0; d < 3; ++d) a.dx_lvl[l][d] = ldexp(1.0, -l) * c.domain[d] / (double)c.Bs[d];  // module_treelib.f90:93
    a.c0 = c.c0;
    a.nu = c.nu;
    a.gamma_p = c.gamma_p;
    a.C_eta_inv = 1.0 / c.C_eta;
    a.C_sponge_inv = 1.0 / c.C_sponge;
    for (int d = 0; d < 3; ++d) a.u_mean_set[d] = c.u_mean_set[d];
    a.use_sponge = c.use_sponge;
    a.CFL = c.CFL;
    a.CFL_nu = c.CFL_nu;
    a.diverged = ctx->d_flags;
    a.dim_min_axes = c.dim;
    a.dt_ptr = ctx->d_dt;
    if (c.n_mask >= 5 && (c.penalization || c.use_sponge)) {
        a.mask = ctx->MASK;
        a.n_mask = c.n_mask;
    }
    if (ctx->geom && c.penalization) {   // analytic mask: chi and u_s are evaluated in the kernel (the sponge, if any, still comes from hvy_mask)
        a.geom = ctx->geom;
        a.ixyz = ctx->d_ixyz;
        for (int d = 0; d < 3; ++d) {
            a.g_c0[d] = ctx->g_c0[d];
            a.g_v[d] = ctx->g_v[d];
        }
        a.g_R = ctx->g_R;
        a.g_h = ctx->g_h;
    }
}


int32_t ensure_stage(wgpu_ctx *ctx, int64_t elems)
{
    if (ctx->stage_elems >= elems) return WGPU_OK;
    if (ctx->d_stage) {
        cudaFree(ctx->d_stage);
        ctx->dev_bytes -= ctx->stage_elems * 8;
    }
    ctx->stage_elems = 0;
    int32_t rc = dmalloc(ctx, &ctx->d_stage, (size_t)elems);
    if (rc) return rc;
    ctx->stage_elems = elems;
    return WGPU_OK;
}

// the (block, direction) lists of the stage kernel's level-jump face patches and room for the patches
static int32_t upload_jump_lists(wgpu_ctx *ctx, const std::vector<int> &jump_blk, const std::vector<int> &jump_dir)
{
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks, nj = (int)jump_blk.size();
    int32_t rc;
    if (nj > ctx->jump_cap) {
        cudaFree(ctx->d_jump_blk);
        cudaFree(ctx->d_jump_dir);
        ctx->d_jump_blk = ctx->d_jump_dir = nullptr;
        const int want = std::max(nj, std::min(6 * N, 2 * nj));
        if ((rc = dmalloc(ctx, &ctx->d_jump_blk, (size_t)want))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_jump_dir, (size_t)want))) return rc;
        ctx->jump_cap = want;
    }
    const int H = c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3);
    const size_t need = (size_t)std::max(nj, 1) * ctx->nc * H * c.Bs[0] * c.Bs[1];
    if (need > ctx->jpool_cap) {
        cudaFree(ctx->d_jpool);
        ctx->d_jpool = nullptr;
        ctx->dev_bytes -= (int64_t)ctx->jpool_cap * 8;
        const size_t want = need + need / 2;
        if ((rc = dmalloc(ctx, &ctx->d_jpool, want))) return rc;
        ctx->jpool_cap = want;
    }
    if (nj) {
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_jump_blk, jump_blk.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_jump_dir, jump_dir.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors may go out of scope
    }
    return WGPU_OK;
}

// block lookup + level-jump patch lists of the current topology (needs wgpu_set_treecodes and wgpu_set_wavelet)
int32_t upload_jump_tables(wgpu_ctx *ctx, const std::vector<int> &jump_blk, const std::vector<int> &jump_dir)
{
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks;
    ctx->lookup_ready = false;
    bool coords = (int)ctx->h_has_coords.size() == N;
    for (int k = 0; k < ctx->n_active && coords; ++k) coords = ctx->h_has_coords[ctx->h_active[k]] != 0;
    for (size_t k = 0; k < ctx->h_halo.size() && coords; ++k) coords = ctx->h_has_coords[ctx->h_halo[k]] != 0;
    if (ctx->has_jumps) {
        if (!ctx->wavelet_set)
            return fail(ctx, WGPU_ERR_ARG, "grid has level jumps: call wgpu_set_wavelet first (the predictor order is the wavelet's)");
        if (c.Bs[0] != c.Bs[1] || (c.dim == 3 && c.Bs[0] != c.Bs[2])) return fail(ctx, WGPU_ERR_UNSUPPORTED, "level jumps need square / cubic blocks so far");
        if (!coords) return fail(ctx, WGPU_ERR_ARG, "grid has level jumps: call wgpu_set_treecodes for the active blocks first");
    }
    if (!coords) return WGPU_OK;   // uniform grid without block positions: nothing that needs the lookup can be called
    const int nj = (int)jump_blk.size();
    int32_t rc;
    // the lookup table depends on the registered block positions only (wgpu_set_treecodes), not on the active list: the passes of the
    // full-tree adapt_tree change the active list many times per tree state and reuse it
    if (!ctx->coords_dirty && ctx->d_hkeys && ctx->d_ixyz) {
        if ((rc = upload_jump_lists(ctx, jump_blk, jump_dir))) return rc;
        ctx->lookup_ready = true;
        return WGPU_OK;
    }
    // hash table (level, ix, iy, iz) -> block
    size_t cap = 64;
    // every block wgpu_set_treecodes listed is resident in HBM and can be a source: the active blocks, the halo copies, and blocks a
    // caller names only as sources (the coarser leaves next to the blocks of a full-tree pass, wabbit_b200/fulltree.py)
    int n_known = 0;
    for (int b = 0; b < N; ++b) n_known += ctx->h_has_coords[b] != 0;
    while (cap < (size_t)n_known * 2 + 2) cap <<= 1;
    std::vector<unsigned long long> keys(cap, ~0ull);
    std::vector<int> vals(cap, -1);
    for (int b = 0; b < N; ++b) {
        if (!ctx->h_has_coords[b]) continue;
        const unsigned long long key = blk_key(ctx->h_tc_level[b], ctx->h_ixyz[3 * b], ctx->h_ixyz[3 * b + 1], ctx->h_ixyz[3 * b + 2]);
        unsigned h = blk_hash(key) & (unsigned)(cap - 1);
        while (keys[h] != ~0ull) {
            if (keys[h] == key) return fail(ctx, WGPU_ERR_ARG, "two active blocks share a treecode");
            h = (h + 1) & (unsigned)(cap - 1);
        }
        keys[h] = key;
        vals[h] = b;
    }
    if (cap > ctx->hcap) {
        cudaFree(ctx->d_hkeys);
        cudaFree(ctx->d_hvals);
        ctx->d_hkeys = nullptr;
        ctx->d_hvals = nullptr;
        if ((rc = dmalloc(ctx, &ctx->d_hkeys, cap))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_hvals, cap))) return rc;
        ctx->hcap = cap;
    }
    ctx->hmask = (unsigned)(cap - 1);
    if (!ctx->d_ixyz && (rc = dmalloc(ctx, &ctx->d_ixyz, (size_t)N * 3))) return rc;
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_hkeys, keys.data(), cap * 8, cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_hvals, vals.data(), cap * 4, cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_ixyz, ctx->h_ixyz.data(), sizeof(int) * (size_t)N * 3, cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
    if ((rc = upload_jump_lists(ctx, jump_blk, jump_dir))) return rc;
    ctx->coords_dirty = false;
    ctx->lookup_ready = true;
    return WGPU_OK;
}

// level-jump ghost patches of the wavelet kernels: every (block, direction) whose neighbours are finer or coarser, all 26
// directions, as deep as the widest wavelet filter; d_wnbr is d_nbr with those directions pointing into the pool
int32_t upload_wjump_tables(wgpu_ctx *ctx, const std::vector<int> &blk, const std::vector<int> &dir)
{
    ctx->n_wjump = 0;
    if (!ctx->has_jumps) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks, Bs = c.Bs[0];
    const WaveFilters &w = ctx->wavelet;
    const int F = std::max(std::max(-w.hd_lo, w.hd_hi), std::max(-w.hr_lo, w.hr_hi));
    ctx->wjump_depth = F;
    const int nj = (int)blk.size();
    const size_t r0 = (size_t)ctx->act_lo * WGPU_NDIR, r1 = (size_t)ctx->act_hi * WGPU_NDIR;   // rows of the active id range only
    std::vector<int> wnbr(ctx->h_nbr.begin() + r0, ctx->h_nbr.begin() + r1);
    std::vector<long long> off(std::max(nj, 1));
    long long total = 0;
    for (int i = 0; i < nj; ++i) {
        const int d[3] = {dir[i] % 3 - 1, (dir[i] / 3) % 3 - 1, dir[i] / 9 - 1};
        wnbr[(size_t)blk[i] * WGPU_NDIR + dir[i] - r0] = -2 - i;
        off[i] = total;
        total += (long long)ctx->nc * (d[0] ? F : Bs) * (d[1] ? F : Bs) * (c.dim == 3 ? (d[2] ? F : Bs) : 1);
    }
    for (size_t i = 0; i < wnbr.size(); ++i)
        if (wnbr[i] <= -2 - WGPU_JUMP_PID) wnbr[i] = -1;   // stage-kernel patch ids mean nothing here (cannot happen: overwritten above)
    int32_t rc;
    if (!ctx->d_wnbr && (rc = dmalloc(ctx, &ctx->d_wnbr, (size_t)N * WGPU_NDIR))) return rc;
    if (nj > ctx->wjump_cap) {
        cudaFree(ctx->d_wjump_blk);
        cudaFree(ctx->d_wjump_dir);
        cudaFree(ctx->d_woff);
        ctx->d_wjump_blk = ctx->d_wjump_dir = nullptr;
        ctx->d_woff = nullptr;
        const int want = nj + nj / 2 + 64;
        if ((rc = dmalloc(ctx, &ctx->d_wjump_blk, (size_t)want))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_wjump_dir, (size_t)want))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_woff, (size_t)want))) return rc;
        ctx->wjump_cap = want;
    }
    if ((size_t)total > ctx->wpool_cap) {
        cudaFree(ctx->d_wpool);
        ctx->d_wpool = nullptr;
        ctx->dev_bytes -= (int64_t)ctx->wpool_cap * 8;
        const size_t want = (size_t)total + (size_t)total / 4;
        if ((rc = dmalloc(ctx, &ctx->d_wpool, want))) return rc;
        ctx->wpool_cap = want;
    }
    if (!wnbr.empty()) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_wnbr + r0, wnbr.data(), sizeof(int) * wnbr.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (nj) {
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_wjump_blk, blk.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_wjump_dir, dir.data(), sizeof(int) * nj, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_woff, off.data(), sizeof(long long) * nj, cudaMemcpyHostToDevice, ctx->stream));
    }
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_wjump = nj;
    return WGPU_OK;
}

// stage 1 reads U; stage j>1 reads what stage j-1 wrote: UA for even j, UB for odd j
const double *stage_input_of(wgpu_ctx *ctx, int j) { return j == 1 ? ctx->U : (((j - 1) & 1) ? ctx->UA : ctx->UB); }

int32_t upload_ids(wgpu_ctx *ctx, int which, const std::vector<int> &v)
{
    int *&p = ctx->d_idbuf[which];
    if (v.size() > ctx->idbuf_cap[which]) {
        cudaFree(p);
        p = nullptr;
        int32_t rc = dmalloc(ctx, &p, v.size() + v.size() / 2 + 64);
        if (rc) return rc;
        ctx->idbuf_cap[which] = v.size() + v.size() / 2 + 64;
    }
    if (!v.empty()) WGPU_CHECK(ctx, cudaMemcpyAsync(p, v.data(), sizeof(int) * v.size(), cudaMemcpyHostToDevice, ctx->stream));
    return WGPU_OK;
}

}  // namespace

extern "C" {

int32_t wgpu_set_mask_sphere(wgpu_ctx *ctx, int32_t enable, const double *center0, const double *velocity, double radius, double smoothing_width)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (!enable) {
        ctx->geom = 0;
        return WGPU_OK;
    }
    if (!center0 || !velocity || !(radius > 0.0) || !(smoothing_width > 0.0)) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_mask_sphere: bad geometry");
    if (ctx->cfg.dim != 3 || !ctx->cfg.penalization) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_mask_sphere: needs dim = 3 and penalization = 1");
    for (int d = 0; d < 3; ++d) {
        ctx->g_c0[d] = center0[d];
        ctx->g_v[d] = velocity[d];
    }
    ctx->g_R = radius;
    ctx->g_h = smoothing_width;
    ctx->geom = 1;
    return WGPU_OK;
}


int32_t wgpu_set_treecodes(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active, const int32_t *level, const int64_t *treecode)
{
    if (!ctx || n_active < 0 || (n_active > 0 && (!hvy_active || !level || !treecode))) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks;
    ctx->h_ixyz.assign((size_t)N * 3, 0);
    ctx->h_has_coords.assign(N, 0);
    ctx->h_tc_level.assign(N, 0);
    for (int k = 0; k < n_active; ++k) {
        const int hid = hvy_active[k], J = level[k];
        if (hid < 1 || hid > N) return fail(ctx, WGPU_ERR_ARG, "hvy_active entry out of range");
        if (J < 0 || J > c.Jmax) return fail(ctx, WGPU_ERR_ARG, "block level out of range");
        // decoding_b (LIB/TREE/module_treelib.f90:793-831): digit bit0 -> y, bit1 -> x, bit2 -> z; level bit i sits in digit i+Jmax-J
        int p[3] = {0, 0, 0};
        for (int d = 0; d < c.dim; ++d)
            for (int i = 0; i < J; ++i) p[d] |= (int)((treecode[k] >> ((i + c.Jmax - J) * c.dim + d)) & 1) << i;
        ctx->h_ixyz[3 * (size_t)(hid - 1) + 0] = p[1];
        ctx->h_ixyz[3 * (size_t)(hid - 1) + 1] = p[0];
        ctx->h_ixyz[3 * (size_t)(hid - 1) + 2] = p[2];
        ctx->h_has_coords[hid - 1] = 1;
        ctx->h_tc_level[hid - 1] = (signed char)J;
    }
    ctx->coords_dirty = true;
    return WGPU_OK;
}

int32_t wgpu_create(const wgpu_config *cfg, wgpu_ctx **out)
{
    if (!cfg || !out) return fail(nullptr, WGPU_ERR_ARG, "wgpu_create: null argument");
    *out = nullptr;
    if (cfg->dim != 2 && cfg->dim != 3) return fail(nullptr, WGPU_ERR_ARG, "dim must be 2 or 3");
    for (int d = 0; d < cfg->dim; ++d)
        if (cfg->Bs[d] < 2 || (cfg->Bs[d] & 1))   // read_Bs aborts on odd sizes, module_ini_files_parser_mpi.f90:816
            return fail(nullptr, WGPU_ERR_ARG, "number_block_nodes must be even");
    if (cfg->n_stages < 1 || cfg->n_stages > WGPU_MAX_STAGES) return fail(nullptr, WGPU_ERR_ARG, "n_stages out of range");
    if (cfg->max_blocks < 1) return fail(nullptr, WGPU_ERR_ARG, "max_blocks must be positive");
    if (cfg->Jmax < 0 || cfg->Jmax >= WGPU_MAX_LEVELS) return fail(nullptr, WGPU_ERR_ARG, "Jmax out of range");
    // the wavelet side (decomposition, thresholding, refinement, coarsening, ghost synchronisation) works for any number of components --
    // post_compression_unit_test.f90 runs it with number_equations = 1; the ACM time stepper needs dim + 1 (checked where it starts)
    if (cfg->n_eqn < 1 || cfg->n_eqn > 16) return fail(nullptr, WGPU_ERR_UNSUPPORTED, "number_equations must be in 1..16");
    if (cfg->n_mask != 0 && cfg->n_mask != 5 && cfg->n_mask != 6) return fail(nullptr, WGPU_ERR_ARG, "n_mask must be 0, 5 or 6");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, WGPU_ERR_NO_DEVICE, "no CUDA device: the WABBIT GPU hot path has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, WGPU_ERR_ARG, "device ordinal out of range");

    wgpu_ctx *ctx = new (std::nothrow) wgpu_ctx();
    if (!ctx) return fail(nullptr, WGPU_ERR_ARG, "out of host memory");
    ctx->cfg = *cfg;
    if (cudaSetDevice(cfg->device) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, WGPU_ERR_CUDA, "cudaSetDevice failed");
    }
    ctx->nc = cfg->n_eqn;
    const int Bz = cfg->dim == 3 ? cfg->Bs[2] : 1;
    ctx->blk_elems = (int64_t)cfg->Bs[0] * cfg->Bs[1] * Bz;
    ctx->gblk_elems = (int64_t)(cfg->Bs[0] + 2 * cfg->g) * (cfg->Bs[1] + 2 * cfg->g) * (cfg->dim == 3 ? cfg->Bs[2] + 2 * cfg->g : 1);
    const size_t n = (size_t)cfg->max_blocks * ctx->nc * ctx->blk_elems;

    int32_t rc = WGPU_OK;
    auto A = [&](double **p, size_t cnt) {
        if (rc == WGPU_OK) rc = dmalloc(ctx, p, cnt);
        if (rc == WGPU_OK && *p) {
            if (cudaMemset(*p, 0, cnt * sizeof(double)) != cudaSuccess) rc = WGPU_ERR_CUDA;
        }
    };
    A(&ctx->U, n);
    A(&ctx->UA, n);
    A(&ctx->UB, n);
    A(&ctx->K[0], n);   // further hvy_work slots: on first use (ensure_K)
    if (cfg->n_mask > 0) A(&ctx->MASK, (size_t)cfg->max_blocks * cfg->n_mask * ctx->blk_elems);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_active, (size_t)cfg->max_blocks);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_nbr, (size_t)cfg->max_blocks * WGPU_NDIR);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_level, (size_t)cfg->max_blocks);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_dt, 1);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_dtmin, 2);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_time, 2);
    if (rc == WGPU_OK) rc = dmalloc(ctx, &ctx->d_flags, 8);
    if (rc == WGPU_OK && cudaMemset(ctx->d_flags, 0, 8 * sizeof(int)) != cudaSuccess) rc = WGPU_ERR_CUDA;
    if (rc == WGPU_OK && cudaMallocHost((void **)&ctx->h_pinned, 8 * sizeof(double)) != cudaSuccess) rc = WGPU_ERR_CUDA;
    if (rc != WGPU_OK) {
        g_create_err = ctx->err.empty() ? std::string("device allocation failed") : ctx->err;
        wgpu_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return WGPU_OK;
}

int32_t wgpu_destroy(wgpu_ctx *ctx)
{
    if (!ctx) return WGPU_OK;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    wgpu_comm_destroy(ctx);
    cudaFree(ctx->U);
    cudaFree(ctx->UA);
    cudaFree(ctx->UB);
    for (int s = 0; s < WGPU_MAX_STAGES; ++s) cudaFree(ctx->K[s]);
    for (double *p : ctx->kry) cudaFree(p);
    cudaFree(ctx->d_kry_part);
    cudaFree(ctx->MASK);
    cudaFree(ctx->TMP);
    cudaFree(ctx->d_active);
    cudaFree(ctx->d_nbr);
    cudaFree(ctx->d_level);
    cudaFree(ctx->d_det_abs);
    cudaFree(ctx->d_det_sq);
    cudaFree(ctx->d_detail_out);
    cudaFree(ctx->d_status);
    cudaFree(ctx->d_norm);
    cudaFree(ctx->d_pool_off);
    cudaFree(ctx->d_ixyz);
    cudaFree(ctx->d_hkeys);
    cudaFree(ctx->d_hvals);
    cudaFree(ctx->d_jump_blk);
    cudaFree(ctx->d_jump_dir);
    cudaFree(ctx->d_jpool);
    cudaFree(ctx->d_ce_blk);
    cudaFree(ctx->d_ce_dir);
    cudaFree(ctx->d_halo_send);
    cudaFree(ctx->d_rhalo_send);
    cudaFree(ctx->d_iota);
    cudaFree(ctx->d_rst_blk);
    cudaFree(ctx->d_rst_mask);
    cudaFree(ctx->d_rmap);
    cudaFree(ctx->d_rpool);
    cudaFree(ctx->d_wjump_blk);
    cudaFree(ctx->d_wjump_dir);
    cudaFree(ctx->d_wnbr);
    cudaFree(ctx->d_woff);
    cudaFree(ctx->d_wpool);
    cudaFree(ctx->d_pd_out);
    cudaFree(ctx->d_bflag);
    cudaFree(ctx->d_rel);
    cudaFree(ctx->d_cnt);
    cudaFree(ctx->d_topo_in);
    cudaFree(ctx->d_halo_ids);
    cudaFree(ctx->d_idbuf[0]);
    cudaFree(ctx->d_idbuf[1]);
    cudaFree(ctx->d_idbuf[2]);
    cudaFree(ctx->d_active_int);
    cudaFree(ctx->d_active_bnd);
    cudaFree(ctx->d_send_blk);
    cudaFree(ctx->d_send_dir);
    cudaFree(ctx->d_dt);
    cudaFree(ctx->d_dtmin);
    cudaFree(ctx->d_time);
    cudaFree(ctx->d_flags);
    cudaFree(ctx->d_stage);
    cudaFree(ctx->d_stat);
    for (int k = 0; k < 2; ++k) {
        cudaFree(ctx->d_span[k]);
        if (ctx->ev_dma[k]) cudaEventDestroy(ctx->ev_dma[k]);
        if (ctx->ev_lay[k]) cudaEventDestroy(ctx->ev_lay[k]);
    }
    if (ctx->dma_stream) cudaStreamDestroy(ctx->dma_stream);
    wgpu_tma_release(ctx);
    for (auto &e : ctx->prof_ev) cudaEventDestroy(e);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_bounce) cudaFreeHost(ctx->h_bounce);
    delete ctx;
    return WGPU_OK;
}

int32_t wgpu_last_error(const wgpu_ctx *ctx, char *buf, int32_t len)
{
    if (!buf || len <= 0) return WGPU_ERR_ARG;
    const std::string &s = ctx ? ctx->err : g_create_err;
    snprintf(buf, (size_t)len, "%s", s.c_str());
    return WGPU_OK;
}

int32_t wgpu_set_stream(wgpu_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return WGPU_ERR_ARG;
    ctx->stream = (cudaStream_t)cuda_stream;
    return WGPU_OK;
}

int32_t wgpu_synchronize(wgpu_ctx *ctx)
{
    if (!ctx) return WGPU_ERR_ARG;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return WGPU_OK;
}

int32_t wgpu_profile(wgpu_ctx *ctx, int32_t enable)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (enable && ctx->prof_ev.empty()) {
        ctx->prof_ev.resize(2 * 4096);
        for (auto &e : ctx->prof_ev) WGPU_CHECK(ctx, cudaEventCreate(&e));
    }
    ctx->profiling = enable != 0;
    ctx->prof_n = 0;
    return WGPU_OK;
}

int32_t wgpu_profile_read(wgpu_ctx *ctx, int32_t *n_launches, double *total_ms)
{
    if (!ctx || !n_launches || !total_ms) return WGPU_ERR_ARG;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    double tot = 0.0;
    for (int i = 0; i < ctx->prof_n; ++i) {
        float ms = 0.f;
        WGPU_CHECK(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
        tot += ms;
    }
    *n_launches = ctx->prof_n;
    *total_ms = tot;
    ctx->prof_n = 0;
    return WGPU_OK;
}

int64_t wgpu_launch_count(const wgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }
int64_t wgpu_device_bytes(const wgpu_ctx *ctx) { return ctx ? ctx->dev_bytes : 0; }

int32_t wgpu_device_pointer(wgpu_ctx *ctx, int32_t array_id, int32_t slot, void **ptr, int64_t *n_doubles)
{
    if (!ctx || !ptr) return WGPU_ERR_ARG;
    int nc = 0;
    double *p = array_ptr(ctx, array_id, slot, &nc);
    if (!p) return fail(ctx, WGPU_ERR_ARG, "wgpu_device_pointer: no such array/slot");
    *ptr = p;
    if (n_doubles) *n_doubles = (int64_t)ctx->cfg.max_blocks * nc * ctx->blk_elems;
    return WGPU_OK;
}

int32_t wgpu_set_topology(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active, const int32_t *level, const int32_t *hvy_neighbor,
                          int32_t ld, int32_t rank)
{
    if (!ctx || n_active < 0 || (n_active > 0 && (!hvy_active || !level || !hvy_neighbor))) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks;
    if (n_active > N) return fail(ctx, WGPU_ERR_ARG, "more active blocks than max_blocks");
    ctx->h_active.assign(n_active, 0);
    ctx->remote_faces.clear();
    ctx->topo_on_device = false;
    ctx->rmap_on_device = false;
    // only the rows of the active blocks are read by the kernels: keep the range of active ids up to date, not the whole table (the passes
    // of the full-tree adapt_tree name a few thousand blocks out of max_blocks many times per call)
    int act_lo = N, act_hi = 0;
    for (int k = 0; k < n_active; ++k) {
        if (hvy_active[k] < 1 || hvy_active[k] > N) return fail(ctx, WGPU_ERR_ARG, "hvy_active entry out of range");
        act_lo = std::min(act_lo, hvy_active[k] - 1);
        act_hi = std::max(act_hi, hvy_active[k]);
    }
    if (n_active == 0) act_lo = act_hi = 0;
    if (ctx->h_nbr.size() != (size_t)N * WGPU_NDIR) ctx->h_nbr.assign((size_t)N * WGPU_NDIR, -1);
    else std::fill(ctx->h_nbr.begin() + (size_t)act_lo * WGPU_NDIR, ctx->h_nbr.begin() + (size_t)act_hi * WGPU_NDIR, -1);
    ctx->act_lo = act_lo;
    ctx->act_hi = act_hi;
    ctx->h_level.assign(N, 0);
    ctx->has_jumps = false;
    ctx->det_cached_for = nullptr;
    std::vector<int> jump_blk, jump_dir, wjump_blk, wjump_dir, no_same, ce_blk, ce_dir, rst_blk, rmap(N, -1);
    std::vector<unsigned> rst_mask;
    std::vector<int> halo_users;      // active blocks with a neighbour in a halo slot (partition-boundary blocks)
    ctx->halo_fine_neighbor = false;
    ctx->n_rhalo_recv = ctx->n_rhalo_send = 0;
    for (size_t k = 0; k < ctx->h_halo.size(); ++k) ctx->h_level[ctx->h_halo[k]] = ctx->halo_level_of[k];
    for (int k = 0; k < n_active; ++k) {
        const int hid = hvy_active[k];
        if (hid < 1 || hid > N) return fail(ctx, WGPU_ERR_ARG, "hvy_active entry out of range");
        if (level[k] < 0 || level[k] > c.Jmax) return fail(ctx, WGPU_ERR_ARG, "block level out of range");
        ctx->h_active[k] = hid - 1;
        ctx->h_level[hid - 1] = (signed char)level[k];
        unsigned jump_mask = 0;      // directions with a coarser or finer neighbour
        bool has_coarser = false, uses_halo = false;
        for (int dz = -1; dz <= 1; ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy && !dz) continue;
                    if (c.dim == 2 && dz != 0) continue;
                    const int d[3] = {dx, dy, dz};
                    const int code = same_level_code(d);   // 1..56
                    const int nfree = 1 << ((dx == 0) + (dy == 0) + (dz == 0) - (c.dim == 2 ? 1 : 0));
                    const int lgt = hvy_neighbor[(size_t)(code - 1) * ld + (hid - 1)];
                    int entry = -1;
                    if (lgt >= 1) {
                        const int r = (lgt - 1) / N;           // lgt2proc.f90
                        auto hit = r != rank ? ctx->halo_map.find(lgt) : ctx->halo_map.end();
                        if (hit != ctx->halo_map.end()) {
                            entry = hit->second;               // the neighbour's copy lives in a local halo slot
                            uses_halo = true;
                        } else if (r != rank) {
                            // same-level neighbour on another GPU: resolved to a pool patch by wgpu_set_exchange (faces);
                            // edges/corners are not needed by the star stencils of the time step
                            entry = -1;
                            if ((dx != 0) + (dy != 0) + (dz != 0) == 1) {
                                ctx->remote_faces.push_back(hid - 1);
                                ctx->remote_faces.push_back((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                            }
                        } else
                            entry = (lgt - 1) - r * N;
                    } else {
                        // coarser (+56) or finer (+112) neighbours in this direction?
                        bool jump = false;
                        for (int s = 0; s < nfree; ++s) {
                            const int lc = hvy_neighbor[(size_t)(code - 1 + s + 56) * ld + (hid - 1)];
                            const int lf = hvy_neighbor[(size_t)(code - 1 + s + 112) * ld + (hid - 1)];
                            if (lc >= 1 || lf >= 1) jump = true;
                            if (lc >= 1) has_coarser = true;
                            if (lc >= 1 && (ce_blk.empty() || ce_blk.back() != hid - 1 || ce_dir.back() != (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1))) {
                                ce_blk.push_back(hid - 1);
                                ce_dir.push_back((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                            }
                            for (int pass = 0; pass < 2; ++pass) {
                                const int l = pass ? lf : lc;
                                if (l < 1 || (l - 1) / N == rank) continue;
                                if (!ctx->halo_map.count(l))
                                    return fail(ctx, WGPU_ERR_UNSUPPORTED, "level-jump neighbour on another rank without a halo copy: call wgpu_set_halo first");
                                uses_halo = true;
                                if (pass) ctx->halo_fine_neighbor = true;
                            }
                        }
                        // every direction without a same-level neighbour is a candidate for a wavelet ghost patch: besides the
                        // coarser / finer relations of the table these are the edges and corners that the reference fills through
                        // the extension of a coarser face neighbour's patch (get_indices_of_ghost_patch, neighborhood.f90:158-331)
                        no_same.push_back(hid - 1);
                        no_same.push_back((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                        if (jump) {
                            ctx->has_jumps = true;
                            jump_mask |= 1u << ((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                            // faces become restriction / prediction patches in the jump pool; the star stencils of the time
                            // step do not read edges and corners
                            if ((dx != 0) + (dy != 0) + (dz != 0) == 1) {
                                entry = -2 - (WGPU_JUMP_PID + (int)jump_blk.size());
                                jump_blk.push_back(hid - 1);
                                jump_dir.push_back((dz + 1) * 9 + (dy + 1) * 3 + (dx + 1));
                            }
                        }
                    }
                    ctx->h_nbr[(size_t)(hid - 1) * WGPU_NDIR + (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)] = entry;
                }
        if (uses_halo) halo_users.push_back(hid - 1);
        if (has_coarser) {   // this leaf sends restricted data: restrict_copy_at_CE needs its filtered copy
            rmap[hid - 1] = (int)rst_blk.size();
            rst_blk.push_back(hid - 1);
            rst_mask.push_back(jump_mask);
        }
    }
    ctx->n_active = n_active;
    ctx->n_int = n_active;
    ctx->n_bnd = 0;
    ctx->halo_bnd.assign(halo_users.begin(), halo_users.end());
    ctx->n_jump = (int)jump_blk.size();
    if (ctx->has_jumps && (int)ctx->h_has_coords.size() == N) {
        for (size_t i = 0; i < no_same.size(); i += 2) {
            const int b = no_same[i], d = no_same[i + 1];
            if (!ctx->h_has_coords[b]) continue;
            const int dd[3] = {d % 3 - 1, (d / 3) % 3 - 1, d / 9 - 1};
            const int nb = 1 << ctx->h_level[b];
            bool inside = true;
            for (int a = 0; a < c.dim; ++a) {
                const int q = ctx->h_ixyz[3 * (size_t)b + a] + dd[a];
                if ((q < 0 || q >= nb) && !c.periodic[a]) inside = false;
            }
            if (inside) {
                wjump_blk.push_back(b);
                wjump_dir.push_back(d);
            }
        }
    }
    {
        int32_t rcj = upload_jump_tables(ctx, jump_blk, jump_dir);
        if (rcj) return rcj;
        if ((rcj = upload_wjump_tables(ctx, wjump_blk, wjump_dir))) return rcj;
        ctx->n_rst = 0;
        ctx->h_rmap.assign(N, -1);
        if (!rst_blk.empty()) {
            const size_t nr = rst_blk.size();
            if ((int)nr > ctx->rst_cap) {
                cudaFree(ctx->d_rst_blk);
                cudaFree(ctx->d_rst_mask);
                ctx->d_rst_blk = nullptr;
                ctx->d_rst_mask = nullptr;
                const size_t want = nr + nr / 2 + 64;
                if ((rcj = dmalloc(ctx, &ctx->d_rst_blk, want)) || (rcj = dmalloc(ctx, &ctx->d_rst_mask, want))) return rcj;
                ctx->rst_cap = (int)want;
            }
            const size_t need = nr * ctx->nc * (size_t)(ctx->blk_elems >> c.dim);
            if (need > ctx->rpool_cap) {
                cudaFree(ctx->d_rpool);
                ctx->d_rpool = nullptr;
                ctx->dev_bytes -= (int64_t)ctx->rpool_cap * 8;
                const size_t want = need + need / 4;
                if ((rcj = dmalloc(ctx, &ctx->d_rpool, want))) return rcj;
                ctx->rpool_cap = want;
            }
            if (!ctx->d_rmap && (rcj = dmalloc(ctx, &ctx->d_rmap, (size_t)N))) return rcj;
            WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rst_blk, rst_blk.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, ctx->stream));
            WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rst_mask, rst_mask.data(), sizeof(unsigned) * nr, cudaMemcpyHostToDevice, ctx->stream));
            WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rmap, rmap.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
            ctx->h_rmap = rmap;
            WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->n_rst = (int)nr;
        }
        ctx->n_ce = 0;
        if (!ce_blk.empty()) {
            if ((int)ce_blk.size() > ctx->ce_cap) {
                cudaFree(ctx->d_ce_blk);
                cudaFree(ctx->d_ce_dir);
                ctx->d_ce_blk = ctx->d_ce_dir = nullptr;
                const size_t want = ce_blk.size() + ce_blk.size() / 2 + 64;
                if ((rcj = dmalloc(ctx, &ctx->d_ce_blk, want)) || (rcj = dmalloc(ctx, &ctx->d_ce_dir, want))) return rcj;
                ctx->ce_cap = (int)want;
            }
            WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_ce_blk, ce_blk.data(), sizeof(int) * ce_blk.size(), cudaMemcpyHostToDevice, ctx->stream));
            WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_ce_dir, ce_dir.data(), sizeof(int) * ce_dir.size(), cudaMemcpyHostToDevice, ctx->stream));
            WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->n_ce = (int)ce_blk.size();
        }
    }
    if (!ctx->d_active_int) {
        int32_t rc2 = dmalloc(ctx, &ctx->d_active_int, (size_t)N);
        if (rc2) return rc2;
        if ((rc2 = dmalloc(ctx, &ctx->d_active_bnd, (size_t)N))) return rc2;
    }
    if (!ctx->halo_bnd.empty()) {   // halo mode: partition-boundary blocks wait for the block exchange, the others do not
        std::vector<char> isb(N, 0);
        for (int b : ctx->halo_bnd) isb[b] = 1;
        std::vector<int> ai, ab;
        for (int k = 0; k < n_active; ++k) (isb[ctx->h_active[k]] ? ab : ai).push_back(ctx->h_active[k]);
        ctx->n_int = (int)ai.size();
        ctx->n_bnd = (int)ab.size();
        if (!ai.empty()) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active_int, ai.data(), sizeof(int) * ai.size(), cudaMemcpyHostToDevice, ctx->stream));
        if (!ab.empty()) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active_bnd, ab.data(), sizeof(int) * ab.size(), cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active, ctx->h_active.data(), sizeof(int) * n_active, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_nbr + (size_t)act_lo * WGPU_NDIR, ctx->h_nbr.data() + (size_t)act_lo * WGPU_NDIR,
                                        sizeof(int) * (size_t)(act_hi - act_lo) * WGPU_NDIR, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_level, ctx->h_level.data(), (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    } else if (n_active) {
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active_int, ctx->h_active.data(), sizeof(int) * n_active, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active, ctx->h_active.data(), sizeof(int) * n_active, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_nbr + (size_t)act_lo * WGPU_NDIR, ctx->h_nbr.data() + (size_t)act_lo * WGPU_NDIR,
                                        sizeof(int) * (size_t)(act_hi - act_lo) * WGPU_NDIR, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_level, ctx->h_level.data(), (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->dtmin_valid = false;
    return WGPU_OK;
}

// Copy-engine path of wgpu_upload / wgpu_download(g_sync = 0) for page-locked host arrays in 3-D (see span_unpack_kernel): chunks of blocks
// are double-buffered so that the DMA of chunk k+1 runs while the layout kernel works on chunk k.  `ids`: 0-based block indices (also in
// d_idbuf[0]).
static int32_t move_blocks_dma(wgpu_ctx *ctx, bool up, double *dev, int nc, const std::vector<int> &ids, double *host)
{
    const wgpu_config &c = ctx->cfg;
    const int n = (int)ids.size();
    const int Bx = c.Bs[0], By = c.Bs[1], Bz = c.Bs[2], g = c.g, nx = Bx + 2 * g, ny = By + 2 * g, nz = Bz + 2 * g;
    const size_t plane_pitch = (size_t)nx * ny * 8, span = ((size_t)(By - 1) * nx + Bx) * 8, pitch = (span + 255) & ~(size_t)255;
    const size_t per_block = pitch * Bz * nc;
    static const size_t chunk_mb = getenv("WGPU_DMA_CHUNK_MB") && atoi(getenv("WGPU_DMA_CHUNK_MB")) > 0 ? (size_t)atoi(getenv("WGPU_DMA_CHUNK_MB")) : 16;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, (chunk_mb << 20) / per_block));
    if (ctx->span_cap < per_block * chunk) {
        for (int s = 0; s < 2; ++s) {
            cudaFree(ctx->d_span[s]);
            ctx->d_span[s] = nullptr;
            WGPU_CHECK(ctx, cudaMalloc((void **)&ctx->d_span[s], per_block * chunk));
        }
        ctx->dev_bytes += 2 * (int64_t)(per_block * chunk) - 2 * (int64_t)ctx->span_cap;
        ctx->span_cap = per_block * chunk;
    }
    if (!ctx->dma_stream) {
        WGPU_CHECK(ctx, cudaStreamCreateWithFlags(&ctx->dma_stream, cudaStreamNonBlocking));
        for (int s = 0; s < 2; ++s) {
            WGPU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_dma[s], cudaEventDisableTiming));
            WGPU_CHECK(ctx, cudaEventCreateWithFlags(&ctx->ev_lay[s], cudaEventDisableTiming));
        }
    }
    const int64_t host_block = ctx->gblk_elems * nc;
    auto dma = [&](int s0, int m, char *stg) -> int32_t {
        int i = 0;
        while (i < m) {      // consecutive hvy ids are contiguous in the host array: one 3-D copy per run
            int j = i + 1;
            while (j < m && ids[s0 + j] == ids[s0 + j - 1] + 1) ++j;
            cudaMemcpy3DParms p = {};
            cudaPitchedPtr hp = make_cudaPitchedPtr(host + (int64_t)ids[s0 + i] * host_block, plane_pitch, plane_pitch, (size_t)nz);
            cudaPitchedPtr dp = make_cudaPitchedPtr(stg + (size_t)i * per_block, pitch, pitch, (size_t)Bz);
            const cudaPos hpos = make_cudaPos(((size_t)g * nx + g) * 8, (size_t)g, 0), dpos = make_cudaPos(0, 0, 0);
            p.srcPtr = up ? hp : dp;
            p.srcPos = up ? hpos : dpos;
            p.dstPtr = up ? dp : hp;
            p.dstPos = up ? dpos : hpos;
            p.extent = make_cudaExtent(span, (size_t)Bz, (size_t)nc * (j - i));
            p.kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
            WGPU_CHECK(ctx, cudaMemcpy3DAsync(&p, ctx->dma_stream));
            i = j;
        }
        return WGPU_OK;
    };
    int32_t rc;
    // One transfer per direction at a time in this process: two contexts (trees) that reach their uploads together would share the H2D
    // rate and then their downloads the D2H rate, and stay in phase for ever; taking turns costs nothing (the link is busy either way)
    // and shifts them so that one tree's download runs against the other's upload -- both directions of the link in use.
    static std::mutex link[2];
    std::lock_guard<std::mutex> turn(link[up ? 0 : 1]);
    // the staging buffers are free (every call ends synchronised); the device array must be complete before the first pack reads it /
    // may be overwritten once earlier work on the context's stream is done: the layout kernels run on that stream
    int k = 0;
    for (int s0 = 0; s0 < n; s0 += chunk, ++k) {
        const int m = std::min(chunk, n - s0), s = k & 1;
        if (up) {
            if (k >= 2) WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->dma_stream, ctx->ev_lay[s], 0));
            if ((rc = dma(s0, m, ctx->d_span[s]))) return rc;
            WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_dma[s], ctx->dma_stream));
            WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_dma[s], 0));
            if ((rc = wgpu_launch_span_unpack(ctx, (const double *)ctx->d_span[s], dev, ctx->d_idbuf[0] + s0, m, nc, (long long)(pitch / 8), ctx->stream))) return rc;
            WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_lay[s], ctx->stream));
        } else {
            if (k >= 2) WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_dma[s], 0));
            if ((rc = wgpu_launch_span_pack(ctx, dev, (double *)ctx->d_span[s], ctx->d_idbuf[0] + s0, m, nc, (long long)(pitch / 8), ctx->stream))) return rc;
            WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_lay[s], ctx->stream));
            WGPU_CHECK(ctx, cudaStreamWaitEvent(ctx->dma_stream, ctx->ev_lay[s], 0));
            if ((rc = dma(s0, m, ctx->d_span[s]))) return rc;
            WGPU_CHECK(ctx, cudaEventRecord(ctx->ev_dma[s], ctx->dma_stream));
        }
    }
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->dma_stream));
    return WGPU_OK;
}

// ------------------------------------------------------------------------------------------------ data movement
static int32_t move_blocks(wgpu_ctx *ctx, bool up, int32_t array_id, int32_t slot, const int32_t *hvy_ids, int32_t n, double *host,
                           int32_t ncomp_host, int32_t g_sync)
{
    if (!ctx || n < 0 || (n > 0 && (!hvy_ids || !host))) return WGPU_ERR_ARG;
    int nc = 0;
    double *dev = array_ptr(ctx, array_id, slot, &nc);
    if (!dev) return fail(ctx, WGPU_ERR_ARG, "no such array/slot on the device");
    if (ncomp_host < 1) return fail(ctx, WGPU_ERR_ARG, "ncomp_host must be >= 1");
    const wgpu_config &c = ctx->cfg;
    if (g_sync < 0 || g_sync > c.g) return fail(ctx, WGPU_ERR_ARG, "g_sync must be in 0..g");
    for (int d = 0; d < c.dim; ++d)
        if (g_sync > c.Bs[d]) return fail(ctx, WGPU_ERR_ARG, "g_sync larger than the block");
    const int64_t per_block = ctx->gblk_elems * ncomp_host;
    {
        // Page-locked host arrays (cudaHostRegister / cudaMallocHost on the whole hvy array): the layout kernels read / write the
        // host array directly over PCIe -- only interiors (+ the g_sync shell on download) cross the bus instead of the whole ghosted
        // box ((Bs+2g)^3 / Bs^3 = 2.6x at Bs=16, g=3), and there is no staging copy.
        // (asked of the first listed block, not of the array's base address: a caller may pass the address block 1 WOULD have while only a
        // window of the array exists in memory)
        cudaPointerAttributes attr;
        const int64_t first_off = n > 0 ? (int64_t)(hvy_ids[0] - 1) * per_block : 0;
        if (n > 0 && hvy_ids[0] >= 1 && cudaPointerGetAttributes(&attr, host + first_off) == cudaSuccess && attr.type == cudaMemoryTypeHost &&
            attr.devicePointer) {
            double *hdev = (double *)attr.devicePointer - first_off;
            std::vector<int> ids(n);
            for (int i = 0; i < n; ++i) {
                if (hvy_ids[i] < 1 || hvy_ids[i] > c.max_blocks) return fail(ctx, WGPU_ERR_ARG, "hvy id out of range");
                ids[i] = hvy_ids[i] - 1;
            }
            int32_t rc = upload_ids(ctx, 0, ids);
            if (rc) return rc;
            if (ctx->xfer_dma[up ? 0 : 1] && c.dim == 3 && g_sync == 0 && ncomp_host == nc && n >= 8) {
                // copy engines: interior plane spans by DMA + a layout kernel (both directions of the link at full rate at once); the x ghost
                // nodes between the interior rows of the host array receive their same-level neighbours' values on a download
                if ((rc = move_blocks_dma(ctx, up, dev, nc, ids, host))) return rc;
                if (up && array_id == WGPU_HVY_BLOCK) ctx->dtmin_valid = false;
                if (up) ctx->det_cached_for = nullptr;
                return WGPU_OK;
            }
            for (int s0 = 0; s0 < n && !rc; s0 += 32768) {   // grid.y limit
                const int m = std::min(32768, n - s0);
                if (up) rc = wgpu_launch_extract(ctx, hdev, dev, ctx->d_idbuf[0] + s0, m, nc, ncomp_host, 1);
                else if (ctx->has_jumps) rc = wgpu_launch_export_regions(ctx, dev, hdev, ctx->d_idbuf[0] + s0, m, nc, ncomp_host, g_sync, 1);
                else rc = wgpu_launch_export(ctx, dev, hdev, ctx->d_idbuf[0] + s0, m, nc, ncomp_host, g_sync, 1);
            }
            if (rc) return rc;
            WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
            if (up && array_id == WGPU_HVY_BLOCK) ctx->dtmin_valid = false;
            if (up) ctx->det_cached_for = nullptr;
            return WGPU_OK;
        }
        cudaGetLastError();   // pageable memory: not an error, take the staged path
    }
    const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(256ll << 20) / (per_block * 8)));
    int32_t rc = ensure_stage(ctx, per_block * chunk + (chunk + 1) / 2 + 8);
    if (rc) return rc;
    int *d_ids = (int *)(ctx->d_stage + per_block * chunk);
    std::vector<int> ids(chunk);
    for (int s0 = 0; s0 < n; s0 += chunk) {
        const int m = std::min(chunk, n - s0);
        for (int i = 0; i < m; ++i) {
            const int hid = hvy_ids[s0 + i];
            if (hid < 1 || hid > c.max_blocks) return fail(ctx, WGPU_ERR_ARG, "hvy id out of range");
            ids[i] = hid - 1;
        }
        WGPU_CHECK(ctx, cudaMemcpyAsync(d_ids, ids.data(), sizeof(int) * m, cudaMemcpyHostToDevice, ctx->stream));
        // consecutive hvy ids are contiguous in the host array: move them as one transfer
        auto xfer = [&](bool h2d) -> int32_t {
            int i = 0;
            while (i < m) {
                int j = i + 1;
                while (j < m && ids[j] == ids[j - 1] + 1) ++j;
                double *h = host + (int64_t)ids[i] * per_block;
                double *d = ctx->d_stage + (int64_t)i * per_block;
                const size_t bytes = (size_t)(j - i) * per_block * 8;
                if (h2d) WGPU_CHECK(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
                else WGPU_CHECK(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
                i = j;
            }
            return WGPU_OK;
        };
        if (up) {
            if ((rc = xfer(true))) return rc;
            if ((rc = wgpu_launch_extract(ctx, ctx->d_stage, dev, d_ids, m, nc, ncomp_host, 0))) return rc;
        } else {
            // ghost layers beyond g_sync (and patches without a same-level source) keep the host's values
            if (g_sync < c.g || ncomp_host > nc)
                if ((rc = xfer(true))) return rc;
            if (ctx->has_jumps) rc = wgpu_launch_export_regions(ctx, dev, ctx->d_stage, d_ids, m, nc, ncomp_host, g_sync, 0);
            else rc = wgpu_launch_export(ctx, dev, ctx->d_stage, d_ids, m, nc, ncomp_host, g_sync, 0);
            if (rc) return rc;
            if ((rc = xfer(false))) return rc;
        }
        // the staging buffer and `ids` are reused by the next chunk
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (up && array_id == WGPU_HVY_BLOCK) ctx->dtmin_valid = false;
    if (up) ctx->det_cached_for = nullptr;
    return WGPU_OK;
}

int32_t wgpu_set_transfer_mode(wgpu_ctx *ctx, int32_t upload_mode, int32_t download_mode)
{
    if (!ctx || upload_mode < 0 || upload_mode > 1 || download_mode < 0 || download_mode > 1) return WGPU_ERR_ARG;
    ctx->xfer_dma[0] = upload_mode;
    ctx->xfer_dma[1] = download_mode;
    return WGPU_OK;
}

int32_t wgpu_upload(wgpu_ctx *ctx, int32_t array_id, int32_t slot, const int32_t *hvy_ids, int32_t n, const double *host, int32_t ncomp_host)
{
    return move_blocks(ctx, true, array_id, slot, hvy_ids, n, const_cast<double *>(host), ncomp_host, 0);
}

int32_t wgpu_download(wgpu_ctx *ctx, int32_t array_id, int32_t slot, const int32_t *hvy_ids, int32_t n, double *host, int32_t ncomp_host,
                      int32_t g_sync)
{
    return move_blocks(ctx, false, array_id, slot, hvy_ids, n, host, ncomp_host, g_sync);
}

// what the launcher may assume about the blocks of a stage launch (StageArgs::plain_hint): 1 all of them have six resident same-level face
// neighbours (periodic domain, no level jump, no face patch of another rank), -1 none has (the partition-boundary list of the face-patch
// exchange), 0 mixed / unknown
static int plain_hint(const wgpu_ctx *ctx, int which)
{
    const wgpu_config &c = ctx->cfg;
    const bool periodic = c.periodic[0] && c.periodic[1] && (c.dim < 3 || c.periodic[2]);
    const bool face_patches = ctx->n_bnd > 0 && ctx->h_halo.empty() && ctx->n_halo_send == 0;
    if (which == WGPU_BLOCKS_BOUNDARY && face_patches) return -1;
    if (!periodic || ctx->has_jumps || ctx->n_jump > 0) return 0;
    if (which == WGPU_BLOCKS_INTERIOR) return 1;
    return ctx->n_bnd > 0 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ compute
int32_t wgpu_sync_ghosts(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t g_minus, int32_t g_plus)
{
    if (!ctx) return WGPU_ERR_ARG;
    int nc = 0;
    if (!array_ptr(ctx, array_id, slot, &nc)) return fail(ctx, WGPU_ERR_ARG, "no such array/slot on the device");
    if (g_minus < 0 || g_plus < 0 || g_minus > ctx->cfg.g || g_plus > ctx->cfg.g) return fail(ctx, WGPU_ERR_ARG, "ghost width out of range");
    // Same-level relations are gathered on the fly by the consumers; there is no level-jump / remote patch in
    // the topologies accepted by wgpu_set_topology yet, hence nothing to refresh.
    return WGPU_OK;
}

int32_t wgpu_set_ghost_filter(wgpu_ctx *ctx, int32_t ignore_filter)
{
    if (!ctx) return WGPU_ERR_ARG;
    ctx->ignore_filter = ignore_filter != 0;
    ctx->det_cached_for = nullptr;
    return WGPU_OK;
}

static int32_t check_flags(wgpu_ctx *ctx)
{
    int flags[4];
    WGPU_CHECK(ctx, cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[0]) {
        cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream);
        return fail(ctx, WGPU_ERR_DIVERGED, "ACM fail: very very large values in state vector.");
    }
    return WGPU_OK;
}

int32_t wgpu_rhs(wgpu_ctx *ctx, double time, int32_t src_slot, int32_t dst_slot)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (ctx->nc != ctx->cfg.dim + 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "ACM needs number_equations = dim+1");
    int nc = 0;
    const double *src = src_slot == 0 ? ctx->U : array_ptr(ctx, WGPU_HVY_WORK, src_slot, &nc);
    double *dst = array_ptr(ctx, WGPU_HVY_WORK, dst_slot, &nc);
    if (!src || !dst || dst_slot < 2) return fail(ctx, WGPU_ERR_ARG, "wgpu_rhs: bad slot");
    ctx->det_cached_for = nullptr;
    StageArgs a;
    fill_common_args(ctx, a);
    a.u_in = src;
    a.u0 = src;
    a.k_out = dst;
    a.t0 = time;
    a.t_cj = 0.0;
    int32_t rc = wgpu_launch_jump_fill(ctx, src);   // sync_ghosts_RHS_tree: restriction / prediction face patches
    if (rc) return rc;
    a.plain_hint = plain_hint(ctx, WGPU_BLOCKS_ALL);
    rc = wgpu_launch_stage(ctx, a, ctx->n_active);
    if (rc) return rc;
    return check_flags(ctx);
}

// ------------------------------------------------------------------------------------------------ mask function and statistics on the device
int32_t wgpu_create_mask(wgpu_ctx *ctx, double time, int32_t geometry, const double *center0, const double *velocity, double radius,
                         double smoothing_width, double L_sponge, double p_sponge)
{
    if (!ctx || !center0) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    if (!ctx->MASK || c.n_mask < 5) return fail(ctx, WGPU_ERR_ARG, "wgpu_create_mask: the context has no hvy_mask (n_mask = 0)");
    if (geometry != WGPU_GEOM_CYLINDER && geometry != WGPU_GEOM_SPHERE) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wgpu_create_mask: geometry must be cylinder (1) or sphere (2)");
    if ((geometry == WGPU_GEOM_CYLINDER) != (c.dim == 2)) return fail(ctx, WGPU_ERR_ARG, "wgpu_create_mask: cylinder is 2-D, sphere is 3-D");
    if (c.penalization && !(smoothing_width > 0.0)) return fail(ctx, WGPU_ERR_ARG, "wgpu_create_mask: smoothing width must be positive");
    if (!ctx->lookup_ready || !ctx->d_ixyz) return fail(ctx, WGPU_ERR_ARG, "wgpu_create_mask: block positions unknown (wgpu_set_treecodes / wgpu_set_grid first)");
    MaskGeom gm;
    memset(&gm, 0, sizeof(gm));
    gm.penalization = c.penalization;
    gm.use_sponge = c.use_sponge && c.n_mask > 5;
    for (int d = 0; d < 3; ++d) {
        gm.domain[d] = c.domain[d];
        gm.c0[d] = d < c.dim ? center0[d] : 0.0;
        gm.v[d] = (velocity && d < c.dim) ? velocity[d] : 0.0;
    }
    gm.R = radius;
    gm.h = smoothing_width;
    gm.L_sponge = L_sponge;
    gm.p_sponge = p_sponge;
    if (gm.use_sponge && !(L_sponge > 0.0 && p_sponge >= 1.0)) return fail(ctx, WGPU_ERR_ARG, "wgpu_create_mask: sponge needs L_sponge > 0 and p_sponge >= 1");
    return wgpu_launch_create_mask(ctx, gm, time);
}

int32_t wgpu_statistics(wgpu_ctx *ctx, double time, int32_t flags, double *out)
{
    if (!ctx || !out || flags < 0 || flags > 3) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    if (ctx->nc != c.dim + 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "ACM needs number_equations = dim+1");
    const bool with_divergence = flags & WGPU_STAT_DIVERGENCE, with_vorticity = flags & WGPU_STAT_VORTICITY;
    if (with_vorticity && ctx->comm && ctx->comm_world > 1)
        return fail(ctx, WGPU_ERR_UNSUPPORTED, "statistics: the vorticity-based entries are computed on one rank only");
    int32_t rc;
    if (!ctx->d_stat && (rc = dmalloc(ctx, &ctx->d_stat, ((size_t)c.max_blocks + 1) * WGPU_NSTAT))) return rc;
    const double *rhs = nullptr;
    if (with_divergence) {
        if ((rc = wgpu_rhs(ctx, time, 0, 2))) return rc;          // hvy_work(:,:,:,:,:,2) <- RHS(hvy_block): its pressure row carries div(u)
        int ncw = 0;
        rhs = array_ptr(ctx, WGPU_HVY_WORK, 2, &ncw);
    }
    StatArgs sa;
    memset(&sa, 0, sizeof(sa));
    for (int d = 0; d < 3; ++d) {
        sa.domain[d] = c.domain[d];
        sa.u_mean_set[d] = c.u_mean_set[d];
    }
    sa.c0 = c.c0;
    sa.gamma_p = c.gamma_p;
    sa.C_eta_inv = 1.0 / c.C_eta;
    sa.C_sponge_inv = 1.0 / c.C_sponge;
    sa.use_sponge = c.use_sponge;
    const double *mask = (c.n_mask >= 5 && (c.penalization || c.use_sponge)) ? ctx->MASK : nullptr;
    double *d_out = ctx->d_stat + (size_t)c.max_blocks * WGPU_NSTAT;
    if ((rc = wgpu_launch_stats(ctx, ctx->U, rhs, mask, sa, ctx->d_stat))) return rc;
    if (with_vorticity && ctx->n_active) {
        // compute_vorticity / compute_dissipation read ghost nodes: ghosted copies of the velocity components (the download path's export
        // kernels: what sync_ghosts_tree leaves, level jumps included), a chunk of blocks at a time through the staging buffer
        VortArgs va;
        memset(&va, 0, sizeof(va));
        for (int d = 0; d < 3; ++d) va.domain[d] = c.domain[d];
        va.nu = c.nu;
        va.H = c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3);
        if (va.H > c.g) return fail(ctx, WGPU_ERR_ARG, "statistics: fewer ghost nodes than the stencil half width");
        {   // FD1_C2 / C4 / C6 / CTW4 and FD2_C2 / C4 / C6 (the optimized scheme takes FD2_C4), module_operators.f90:23-33, 385-389: taps -H..H
            const double c2[3] = {-0.5, 0.0, 0.5}, c4[5] = {1.0 / 12.0, -8.0 / 12.0, 0.0, 8.0 / 12.0, -1.0 / 12.0};
            const double c6[7] = {-1.0 / 60.0, 9.0 / 60.0, -45.0 / 60.0, 0.0, 45.0 / 60.0, -9.0 / 60.0, 1.0 / 60.0};
            const double tw[7] = {-0.02651995, 0.18941314, -0.79926643, 0.0, 0.79926643, -0.18941314, 0.02651995};
            const double s2[3] = {1.0, -2.0, 1.0}, s4[5] = {-1.0 / 12.0, 16.0 / 12.0, -30.0 / 12.0, 16.0 / 12.0, -1.0 / 12.0};
            const double s6[7] = {2.0 / 180.0, -27.0 / 180.0, 270.0 / 180.0, -490.0 / 180.0, 270.0 / 180.0, -27.0 / 180.0, 2.0 / 180.0};
            const double *f1 = c.fd == 2 ? c2 : (c.fd == 4 ? c4 : (c.fd == 6 ? c6 : tw));
            for (int t = 0; t < 2 * va.H + 1; ++t) va.fd1[t] = f1[t];
            if (c.fd == 2) for (int t = 0; t < 3; ++t) va.fd2[t] = s2[t];
            else if (c.fd == 6) for (int t = 0; t < 7; ++t) va.fd2[t] = s6[t];
            else for (int t = 0; t < 5; ++t) va.fd2[t + va.H - 2] = s4[t];          // FD2_C4, centred in the 2H+1 taps
        }
        const int64_t per_block = ctx->gblk_elems * c.dim;
        const int n = ctx->n_active;
        const int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(256ll << 20) / (per_block * 8)));
        if ((rc = ensure_stage(ctx, per_block * chunk))) return rc;
        WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_stage, 0, sizeof(double) * (size_t)per_block * chunk, ctx->stream));   // no neighbour: zeros
        for (int s0 = 0; s0 < n; s0 += chunk) {
            const int m = std::min(chunk, n - s0);
            if (ctx->has_jumps) rc = wgpu_launch_export_regions(ctx, ctx->U, ctx->d_stage, ctx->d_active + s0, m, ctx->nc, c.dim, va.H, 0);
            else rc = wgpu_launch_export(ctx, ctx->U, ctx->d_stage, ctx->d_active + s0, m, ctx->nc, c.dim, va.H, 0);
            if (rc) return rc;
            if ((rc = wgpu_launch_vort_stats(ctx, ctx->d_stage, ctx->d_active + s0, m, c.dim, va, ctx->d_stat + (size_t)s0 * WGPU_NSTAT))) return rc;
        }
    }
    if ((rc = wgpu_launch_stats_final(ctx, ctx->d_stat, d_out))) return rc;
    const int n_out = with_vorticity ? WGPU_NSTAT : 19;
    WGPU_CHECK(ctx, cudaMemcpyAsync(out, d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (!with_divergence) out[14] = out[15] = 0.0;
    if (ctx->comm && ctx->comm_world > 1) {      // the MPI_Allreduce calls of the post_stage (statistics_ACM.f90:396-430): SUM, MAX, MIN
        double sums[16], mx[2], mn[1];
        for (int k = 0; k < 13; ++k) sums[k] = out[k];
        for (int k = 0; k < 3; ++k) sums[13 + k] = out[16 + k];
        mx[0] = out[13];
        mx[1] = out[14];
        mn[0] = out[15];
        if ((rc = wgpu_comm_allreduce(ctx, sums, 16, 2))) return rc;
        if ((rc = wgpu_comm_allreduce(ctx, mx, 2, 0))) return rc;
        if ((rc = wgpu_comm_allreduce(ctx, mn, 1, 1))) return rc;
        for (int k = 0; k < 13; ++k) out[k] = sums[k];
        for (int k = 0; k < 3; ++k) out[16 + k] = sums[13 + k];
        out[13] = mx[0];
        out[14] = mx[1];
        out[15] = mn[0];
    }
    return WGPU_OK;
}

static int32_t compute_dt(wgpu_ctx *ctx, double time)
{
    int32_t rc;
    if (ctx->nc != ctx->cfg.dim + 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "ACM needs number_equations = dim+1");
    unsigned long long *cur = ctx->d_dtmin + ctx->dtmin_cur, *nxt = ctx->d_dtmin + (ctx->dtmin_cur ^ 1);
    if (!ctx->dtmin_valid && !(ctx->cfg.dt_fixed > 0.0)) {
        const unsigned long long inf = 0x7FF0000000000000ULL;
        WGPU_CHECK(ctx, cudaMemcpyAsync(cur, &inf, 8, cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = wgpu_launch_dtmin(ctx, ctx->U, cur))) return rc;
    }
    if ((rc = wgpu_launch_dt_finalize(ctx, time, cur, nxt))) return rc;
    ctx->dtmin_cur ^= 1;
    ctx->dtmin_valid = false;
    return WGPU_OK;
}

int32_t wgpu_calculate_time_step(wgpu_ctx *ctx, double time, double *dt)
{
    if (!ctx || !dt) return WGPU_ERR_ARG;
    int32_t rc = compute_dt(ctx, time);
    if (rc) return rc;
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_dt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    *dt = ctx->h_pinned[0];
    return WGPU_OK;
}

// ------------------------------------------------------------------------------------------------ wavelets
// interpolation stencils (module_wavelets.f90:14-20), centred
static void interp_stencil(int order, double *s)
{
    for (int i = 0; i < 2 * order - 1; ++i) s[i] = 0.0;
    if (order == 2) {
        s[0] = 1.0 / 2.0; s[1] = 2.0 / 2.0; s[2] = 1.0 / 2.0;
    } else if (order == 4) {
        const double v[7] = {-1.0, 0.0, 9.0, 16.0, 9.0, 0.0, -1.0};
        for (int i = 0; i < 7; ++i) s[i] = v[i] / 16.0;
    } else {
        const double v[11] = {3.0, 0.0, -25.0, 0.0, 150.0, 256.0, 150.0, 0.0, -25.0, 0.0, 3.0};
        for (int i = 0; i < 11; ++i) s[i] = v[i] / 256.0;
    }
}

int32_t wgpu_set_wavelet(wgpu_ctx *ctx, const char *name, int32_t *g_default, int32_t *g_rhs_default)
{
    if (!ctx || !name) return WGPU_ERR_ARG;
    if (strlen(name) != 5 || strncmp(name, "CDF", 3) != 0) return fail(ctx, 3006221, "Unkown bi-orthogonal wavelet specified.");
    const int X = name[3] - '0', Y = name[4] - '0';
    if ((X != 2 && X != 4 && X != 6) || (Y != 0 && Y != 2 && Y != 4 && Y != 6) || Y > X)
        return fail(ctx, 3006221, "Unkown bi-orthogonal wavelet specified (supported: CDFXY, X in 2,4,6, Y in 0,2,4,6, Y<=X).");
    WaveFilters &w = ctx->wavelet;
    memset(&w, 0, sizeof(w));
    w.X = X;
    w.Y = Y;
    double hr[11], hn[11];
    interp_stencil(X, hr);                                   // HR = interpolation stencil (module_wavelets.f90:1157-1176)
    w.hr_lo = -(X - 1);
    w.hr_hi = X - 1;
    for (int i = w.hr_lo; i <= w.hr_hi; ++i) w.HR[i + WGPU_FMAX] = hr[i + X - 1];
    if (Y == 0) {
        w.hd_lo = w.hd_hi = 0;
        w.HD[WGPU_FMAX] = 1.0;
    } else {                                                 // HD = delta + 1/2 sum_j (-1)^j HR(j) h~(i-j)   (:1218-1234)
        interp_stencil(Y, hn);
        hn[Y - 1] = 0.0;
        w.hd_lo = w.hr_lo - (Y - 1);
        w.hd_hi = w.hr_hi + (Y - 1);
        for (int i = w.hd_lo; i <= w.hd_hi; ++i) {
            double v = i == 0 ? 1.0 : 0.0;
            for (int j = w.hr_lo; j <= w.hr_hi; ++j) {
                if (i - j < -(Y - 1) || i - j > Y - 1) continue;
                v = v + ((j % 2 == 0) ? 1.0 : -1.0) * w.HR[j + WGPU_FMAX] * hn[i - j + Y - 1] / 2.0;
            }
            w.HD[i + WGPU_FMAX] = v;
        }
    }
    w.gd_lo = w.hr_lo;                                       // GD(i) = (-1)^i HR(i), GR(i) = (-1)^i HD(i)   (:1275-1288)
    w.gd_hi = w.hr_hi;
    for (int i = w.gd_lo; i <= w.gd_hi; ++i) w.GD[i + WGPU_FMAX] = ((i % 2 == 0) ? 1.0 : -1.0) * w.HR[i + WGPU_FMAX];
    w.gr_lo = w.hd_lo;
    w.gr_hi = w.hd_hi;
    for (int i = w.gr_lo; i <= w.gr_hi; ++i) w.GR[i + WGPU_FMAX] = ((i % 2 == 0) ? 1.0 : -1.0) * w.HD[i + WGPU_FMAX];
    if (g_default) *g_default = X - 1 + (Y - 1 > 0 ? Y - 1 : 0);
    if (g_rhs_default) *g_rhs_default = X / 2;
    int32_t rc;
    const size_t n = (size_t)ctx->cfg.max_blocks * ctx->nc;
    if (!ctx->TMP) {
        if ((rc = dmalloc(ctx, &ctx->TMP, n * ctx->blk_elems))) return rc;
        WGPU_CHECK(ctx, cudaMemset(ctx->TMP, 0, n * ctx->blk_elems * sizeof(double)));
    }
    if (!ctx->d_det_abs) {
        if ((rc = dmalloc(ctx, &ctx->d_det_abs, n))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_det_sq, n))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_detail_out, n))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_status, (size_t)ctx->cfg.max_blocks))) return rc;
        if ((rc = dmalloc(ctx, &ctx->d_norm, 16))) return rc;
    }
    ctx->wavelet_set = true;
    return WGPU_OK;
}

static int32_t transform(wgpu_ctx *ctx, int32_t src_id, int32_t src_slot, int32_t dst_id, int32_t dst_slot, int inverse)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    const wgpu_config &c = ctx->cfg;
    if (c.Bs[0] != c.Bs[1] || (c.dim == 3 && c.Bs[0] != c.Bs[2])) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wavelet kernels: square / cubic blocks only so far");
    if (!ctx->remote_faces.empty() || (ctx->n_bnd && ctx->halo_bnd.empty()))
        return fail(ctx, WGPU_ERR_UNSUPPORTED, "wavelet kernels: neighbours on other ranks need halo copies (wgpu_set_halo)");
    if (ctx->halo_fine_neighbor && !ctx->n_rhalo_recv && !ctx->ignore_filter && ctx->wavelet.Y != 0)
        return fail(ctx, WGPU_ERR_UNSUPPORTED, "finer neighbours on another rank: their filtered copies must be exchanged first (wgpu_set_halo_restrict, wgpu_restrict_pack)");
    int n1 = 0, n2 = 0;
    const double *src = array_ptr(ctx, src_id, src_slot, &n1);
    double *dst = array_ptr(ctx, dst_id, dst_slot, &n2);
    if (!src || !dst || n1 != ctx->nc || n2 != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wavelet transform: bad array/slot");
    if (src == dst) return fail(ctx, WGPU_ERR_ARG, "wavelet transform: src and dst must differ (neighbours read src halos)");
    if (dst == ctx->U) ctx->dtmin_valid = false;
    ctx->det_cached_for = nullptr;
    return wgpu_launch_wavelet(ctx, src, dst, inverse, nullptr);
}

int32_t wgpu_fwt(wgpu_ctx *ctx, int32_t src_id, int32_t src_slot, int32_t dst_id, int32_t dst_slot)
{
    return transform(ctx, src_id, src_slot, dst_id, dst_slot, 0);
}

int32_t wgpu_iwt(wgpu_ctx *ctx, int32_t src_id, int32_t src_slot, int32_t dst_id, int32_t dst_slot)
{
    return transform(ctx, src_id, src_slot, dst_id, dst_slot, 1);
}

int32_t wgpu_iwt_ce(wgpu_ctx *ctx, int32_t wd_id, int32_t wd_slot, int32_t coarse_id, int32_t coarse_slot, int32_t dst_id, int32_t dst_slot)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    const wgpu_config &c = ctx->cfg;
    if (c.Bs[0] != c.Bs[1] || (c.dim == 3 && c.Bs[0] != c.Bs[2])) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wavelet kernels: square / cubic blocks only so far");
    int n1 = 0, n2 = 0, n3 = 0;
    const double *src = array_ptr(ctx, wd_id, wd_slot, &n1);
    const double *crs = array_ptr(ctx, coarse_id, coarse_slot, &n2);
    double *dst = array_ptr(ctx, dst_id, dst_slot, &n3);
    if (!src || !crs || !dst || n1 != ctx->nc || n2 != ctx->nc || n3 != ctx->nc || src == dst || src == crs)
        return fail(ctx, WGPU_ERR_ARG, "wgpu_iwt_ce: bad array/slot (the coefficients must not share an array with the coarse values or the result)");
    if (!ctx->lookup_ready && ctx->has_jumps) return fail(ctx, WGPU_ERR_ARG, "wgpu_iwt_ce: call wgpu_set_treecodes + wgpu_set_topology first");
    if (dst == ctx->U) ctx->dtmin_valid = false;
    ctx->det_cached_for = nullptr;
    return wgpu_launch_wavelet(ctx, src, dst, 1, crs);
}

int32_t wgpu_norm(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t norm_id, double *out)
{
    if (!ctx || !out) return WGPU_ERR_ARG;
    if (norm_id < 0 || norm_id > 3) return fail(ctx, WGPU_ERR_ARG, "wgpu_norm: norm_id must be 0 Linfty, 1 L1, 2 L2 or 3 H1");
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    int n1 = 0;
    const double *src = array_ptr(ctx, array_id, slot, &n1);
    if (!src || n1 != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_norm: bad array/slot");
    if (norm_id != 0) {
        // L1: sum |u| dV; L2 / H1: sqrt(sum u^2 dV)   (componentWiseNorm_tree.f90:150-197, 283-290), dV = prod(dx) of the block's level
        int32_t rcs = wgpu_launch_blocksum(ctx, src, norm_id != 1, ctx->d_detail_out);
        if (rcs) return rcs;
        std::vector<double> part((size_t)ctx->n_active * ctx->nc);
        WGPU_CHECK(ctx, cudaMemcpyAsync(part.data(), ctx->d_detail_out, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        const wgpu_config &c = ctx->cfg;
        for (int p = 0; p < ctx->nc; ++p) out[p] = 0.0;
        for (int k = 0; k < ctx->n_active; ++k) {
            double dv = 1.0;
            for (int d = 0; d < c.dim; ++d) dv *= ldexp(1.0, -ctx->h_level[ctx->h_active[k]]) * c.domain[d] / (double)c.Bs[d];
            for (int p = 0; p < ctx->nc; ++p) out[p] = out[p] + dv * part[(size_t)k * ctx->nc + p];
        }
        if (norm_id != 1)
            for (int p = 0; p < ctx->nc; ++p) out[p] = sqrt(out[p]);
        return WGPU_OK;
    }
    WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_norm, 0, 16 * sizeof(unsigned long long), ctx->stream));
    int32_t rc = wgpu_launch_linfty(ctx, src, ctx->d_norm);
    if (rc) return rc;
    double h[16];
    WGPU_CHECK(ctx, cudaMemcpyAsync(h, ctx->d_norm, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ctx->nc; ++i) out[i] = h[i];
    return WGPU_OK;
}

int32_t wgpu_threshold(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t eps_norm_id, int32_t level_ref, const int32_t *thresh_comp,
                       const double *eps, const double *norm, int32_t *refinement_status, double *detail_out)
{
    if (!ctx || !thresh_comp || !eps || !refinement_status) return WGPU_ERR_ARG;
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    if (eps_norm_id < 0 || eps_norm_id > 3) return fail(ctx, 241024, "ERROR:Unknown wavelet normalization!");
    if (ctx->nc > 16) return fail(ctx, WGPU_ERR_UNSUPPORTED, "too many components");
    int n1 = 0;
    const double *wd = array_ptr(ctx, array_id, slot, &n1);
    if (!wd || n1 != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_threshold: bad array/slot");
    int32_t rc;
    if ((rc = wgpu_launch_detail(ctx, wd, eps_norm_id, level_ref))) return rc;
    double eps_use[16];
    for (int i = 0; i < ctx->nc; ++i) eps_use[i] = norm ? eps[i] * norm[i] : eps[i];   // threshold_block.f90:111-113
    if ((rc = wgpu_launch_flags(ctx, thresh_comp, eps_use, ctx->d_status, detail_out ? ctx->d_detail_out : nullptr))) return rc;
    WGPU_CHECK(ctx, cudaMemcpyAsync(refinement_status, ctx->d_status, sizeof(int) * ctx->n_active, cudaMemcpyDeviceToHost, ctx->stream));
    if (detail_out)
        WGPU_CHECK(ctx, cudaMemcpyAsync(detail_out, ctx->d_detail_out, sizeof(double) * (size_t)ctx->n_active * ctx->nc, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return WGPU_OK;
}

int32_t wgpu_patch_details(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, const int32_t *dirs, double *detail_out)
{
    return wgpu_patch_details_norm(ctx, array_id, slot, 0, 0, n, hvy_ids, dirs, detail_out);
}

int32_t wgpu_patch_details_norm(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t eps_norm_id, int32_t level_ref, int32_t n, const int32_t *hvy_ids,
                                const int32_t *dirs, double *detail_out)
{
    if (!ctx || n < 0 || (n > 0 && (!hvy_ids || !dirs || !detail_out))) return WGPU_ERR_ARG;
    if (eps_norm_id < 0 || eps_norm_id > 3) return fail(ctx, WGPU_ERR_ARG, "wgpu_patch_details_norm: eps_norm must be 0 (Linfty), 1 (L1), 2 (L2) or 3 (H1)");
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    int n1 = 0;
    const double *wd = array_ptr(ctx, array_id, slot, &n1);
    if (!wd || n1 != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_patch_details: bad array/slot");
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    std::vector<int> ids(2 * (size_t)n);
    for (int k = 0; k < n; ++k) {
        ids[k] = hvy_ids[k] - 1;
        ids[(size_t)n + k] = dirs[k];
        if (ids[k] < 0 || ids[k] >= c.max_blocks || dirs[k] < 0 || dirs[k] >= WGPU_NDIR || dirs[k] == 13)
            return fail(ctx, WGPU_ERR_ARG, "wgpu_patch_details: bad (block, direction) pair");
    }
    int32_t rc = upload_ids(ctx, 2, ids);
    if (rc) return rc;
    if ((size_t)n * ctx->nc > ctx->pd_cap) {   // persistent scratch: cudaMalloc / cudaFree per call cost more than the kernel
        cudaFree(ctx->d_pd_out);
        ctx->d_pd_out = nullptr;
        ctx->pd_cap = 0;
        const size_t want = (size_t)n * ctx->nc + (size_t)n * ctx->nc / 2 + 1024;
        if ((rc = dmalloc(ctx, &ctx->d_pd_out, want))) return rc;
        ctx->pd_cap = want;
    }
    double *d_out = ctx->d_pd_out;
    // strip depth = Nwcl / Nwcr of setup_wavelet incl. the widening to the FD stencil (securityZone_tree.f90:166-167)
    const WaveFilters &w = ctx->wavelet;
    const int H = c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3);
    const int Nscl = std::max(-w.hd_lo - 1, 0), Nscr = w.hd_hi;
    const int Nwcl = std::max(Nscl - w.gd_lo, 2 * H), Nwcr = std::max(Nscr + w.gd_hi, 2 * H);
    rc = wgpu_launch_patch_detail(ctx, wd, ctx->d_idbuf[2], ctx->d_idbuf[2] + n, n, Nwcl, Nwcr, d_out, eps_norm_id, level_ref);
    if (!rc) {
        cudaError_t e = cudaMemcpyAsync(detail_out, d_out, sizeof(double) * (size_t)n * ctx->nc, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            ctx->err = std::string("wgpu_patch_details: ") + cudaGetErrorString(e);
            rc = WGPU_ERR_CUDA;
        }
    }
    return rc;
}

int32_t wgpu_coarse_extension(wgpu_ctx *ctx, int32_t wd_id, int32_t wd_slot, int32_t orig_id, int32_t orig_slot, int32_t clear_wc, int32_t copy_sc)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    int n1 = 0, n2 = 0;
    double *wd = array_ptr(ctx, wd_id, wd_slot, &n1);
    const double *orig = array_ptr(ctx, orig_id, orig_slot, &n2);
    if (!wd || !orig || n1 != ctx->nc || n2 != ctx->nc || wd == orig) return fail(ctx, WGPU_ERR_ARG, "wgpu_coarse_extension: bad array/slot");
    // coarse-extension sizes (setup_wavelet, module_wavelets.f90:1368-1417): Nsc from HD, Nwc = Nsc + |GD|, widened to 2*FD_max_size
    const WaveFilters &w = ctx->wavelet;
    const int H = ctx->cfg.fd == 2 ? 1 : (ctx->cfg.fd == 4 ? 2 : 3);
    const int Nscl = std::max(-w.hd_lo - 1, 0), Nscr = w.hd_hi;
    const int Nwcl = std::max(Nscl - w.gd_lo, 2 * H), Nwcr = std::max(Nscr + w.gd_hi, 2 * H);
    ctx->det_cached_for = nullptr;
    if (wd == ctx->U) ctx->dtmin_valid = false;
    return wgpu_launch_ce(ctx, wd, orig, Nwcl, Nwcr, Nscl, Nscr, clear_wc != 0, copy_sc != 0);
}

// ------------------------------------------------------------------------------------------------ refinement / coarsening
int32_t wgpu_refine(wgpu_ctx *ctx, int32_t n, const int32_t *mother_hvy, const int32_t *daughter_hvy, int32_t n_keep, const int32_t *keep_src,
                    const int32_t *keep_dst)
{
    if (!ctx || n < 0 || (n > 0 && (!mother_hvy || !daughter_hvy)) || (n_keep > 0 && (!keep_src || !keep_dst))) return WGPU_ERR_ARG;
    if (!ctx->wavelet_set) return fail(ctx, 1213149, "The cat is angry: Wavelet-setup not yet called?");
    if (!ctx->lookup_ready) return fail(ctx, WGPU_ERR_ARG, "wgpu_refine: call wgpu_set_treecodes + wgpu_set_topology first");
    const wgpu_config &c = ctx->cfg;
    if (c.Bs[0] != c.Bs[1] || (c.dim == 3 && c.Bs[0] != c.Bs[2])) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wgpu_refine: cubic blocks only");
    if (!ctx->remote_faces.empty() || (ctx->n_bnd && ctx->halo_bnd.empty()))
        return fail(ctx, WGPU_ERR_UNSUPPORTED, "wgpu_refine: neighbours on other ranks need halo copies (wgpu_set_halo)");
    if (ctx->halo_fine_neighbor && !ctx->n_rhalo_recv && !ctx->ignore_filter && ctx->wavelet.Y != 0)
        return fail(ctx, WGPU_ERR_UNSUPPORTED, "finer neighbours on another rank: their filtered copies must be exchanged first (wgpu_set_halo_restrict, wgpu_restrict_pack)");
    const int N = c.max_blocks, nd = 1 << c.dim;
    std::vector<int> mo(n), da((size_t)n * nd), ksrc, kdst;
    std::vector<char> is_mother(N, 0), is_active(N, 0), taken(N, 0);
    for (int k = 0; k < ctx->n_active; ++k) is_active[ctx->h_active[k]] = 1;
    for (int i = 0; i < n; ++i) {
        mo[i] = mother_hvy[i] - 1;
        if (mo[i] < 0 || mo[i] >= N || !is_active[mo[i]] || is_mother[mo[i]]) return fail(ctx, WGPU_ERR_ARG, "wgpu_refine: bad mother id");
        if (ctx->h_level[mo[i]] >= c.Jmax) return fail(ctx, WGPU_ERR_ARG, "wgpu_refine: mother is on Jmax already");
        is_mother[mo[i]] = 1;
    }
    if (n_keep < 0) {   // every block that is not refined stays where it is
        for (int k = 0; k < ctx->n_active; ++k)
            if (!is_mother[ctx->h_active[k]]) {
                ksrc.push_back(ctx->h_active[k]);
                kdst.push_back(ctx->h_active[k]);
            }
    } else {
        for (int i = 0; i < n_keep; ++i) {
            ksrc.push_back(keep_src[i] - 1);
            kdst.push_back(keep_dst[i] - 1);
            if (ksrc[i] < 0 || ksrc[i] >= N || !is_active[ksrc[i]] || is_mother[ksrc[i]] || kdst[i] < 0 || kdst[i] >= N)
                return fail(ctx, WGPU_ERR_ARG, "wgpu_refine: bad keep list entry");
        }
    }
    // the new grid is assembled in the other array: any slot may be a destination, but only once
    for (int v : kdst) {
        if (taken[v]) return fail(ctx, WGPU_ERR_ARG, "wgpu_refine: destination slot used twice");
        taken[v] = 1;
    }
    for (size_t i = 0; i < da.size(); ++i) {
        da[i] = daughter_hvy[i] - 1;
        if (da[i] < 0 || da[i] >= N || taken[da[i]]) return fail(ctx, WGPU_ERR_ARG, "wgpu_refine: daughter id out of range or destination slot used twice");
        taken[da[i]] = 1;
    }
    int32_t rc;
    std::vector<int> pairs(ksrc);
    pairs.insert(pairs.end(), kdst.begin(), kdst.end());
    if ((rc = upload_ids(ctx, 0, mo)) || (rc = upload_ids(ctx, 1, da)) || (rc = upload_ids(ctx, 2, pairs))) return rc;
    // daughters and kept blocks are written to the other array, which then becomes hvy_block: one read and one write of
    // the new grid's data, and no in-place hazard when a slot is reused (refinementExecute.f90: the last daughter overwrites the mother)
    if ((rc = wgpu_launch_refine(ctx, ctx->U, ctx->TMP, ctx->d_idbuf[0], ctx->d_idbuf[1], n))) return rc;
    if ((rc = wgpu_launch_copy_blocks(ctx, ctx->U, ctx->TMP, ctx->d_idbuf[2], ctx->d_idbuf[2] + ksrc.size(), (int)ksrc.size()))) return rc;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    std::swap(ctx->U, ctx->TMP);
    ctx->dtmin_valid = false;
    ctx->det_cached_for = nullptr;
    return WGPU_OK;
}

int32_t wgpu_move_blocks(wgpu_ctx *ctx, int32_t n, const int32_t *src_hvy, const int32_t *dst_hvy)
{
    if (!ctx || n < 0 || (n > 0 && (!src_hvy || !dst_hvy))) return WGPU_ERR_ARG;
    if (!ctx->TMP) return fail(ctx, 1213149, "wgpu_move_blocks needs hvy_tmp as the second buffer: call wgpu_set_wavelet first");
    const int N = ctx->cfg.max_blocks;
    std::vector<int> pairs((size_t)2 * n);
    std::vector<char> taken(N, 0);
    bool identity = true;
    for (int i = 0; i < n; ++i) {
        const int s = src_hvy[i] - 1, d = dst_hvy[i] - 1;
        if (s < 0 || s >= N || d < 0 || d >= N || taken[d]) return fail(ctx, WGPU_ERR_ARG, "wgpu_move_blocks: id out of range or destination used twice");
        taken[d] = 1;
        pairs[i] = s;
        pairs[(size_t)n + i] = d;
        identity = identity && s == d;
    }
    if (identity) return WGPU_OK;
    int32_t rc;
    if ((rc = upload_ids(ctx, 2, pairs))) return rc;
    if ((rc = wgpu_launch_copy_blocks(ctx, ctx->U, ctx->TMP, ctx->d_idbuf[2], ctx->d_idbuf[2] + n, n))) return rc;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    std::swap(ctx->U, ctx->TMP);
    ctx->det_cached_for = nullptr;
    return WGPU_OK;
}

int32_t wgpu_coarsen(wgpu_ctx *ctx, int32_t n, const int32_t *mother_hvy, const int32_t *daughter_hvy, int32_t src_id, int32_t src_slot)
{
    if (!ctx || n < 0 || (n > 0 && (!mother_hvy || !daughter_hvy))) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    if (c.Bs[0] != c.Bs[1] || (c.dim == 3 && c.Bs[0] != c.Bs[2])) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wgpu_coarsen: cubic blocks only");
    int nc = 0;
    const double *src = array_ptr(ctx, src_id, src_slot, &nc);
    if (!src || nc != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_coarsen: bad array/slot");
    if (src == ctx->U) return fail(ctx, WGPU_ERR_ARG, "wgpu_coarsen: the decomposed source must not be hvy_block (a mother may reuse a daughter's slot)");
    const int N = c.max_blocks, nd = 1 << c.dim;
    std::vector<int> mo(n), da((size_t)n * nd);
    for (int i = 0; i < n; ++i) {
        mo[i] = mother_hvy[i] - 1;
        if (mo[i] < 0 || mo[i] >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_coarsen: mother id out of range");
    }
    for (size_t i = 0; i < da.size(); ++i) {
        da[i] = daughter_hvy[i] - 1;
        if (da[i] < 0 || da[i] >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_coarsen: daughter id out of range");
    }
    int32_t rc;
    if ((rc = upload_ids(ctx, 0, mo)) || (rc = upload_ids(ctx, 1, da))) return rc;
    if ((rc = wgpu_launch_coarsen(ctx, src, ctx->U, ctx->d_idbuf[0], ctx->d_idbuf[1], n))) return rc;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->dtmin_valid = false;
    return WGPU_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU exchange
static int fd_halo(const wgpu_ctx *ctx) { return ctx->cfg.fd == 2 ? 1 : (ctx->cfg.fd == 4 ? 2 : 3); }

int64_t wgpu_patch_doubles(const wgpu_ctx *ctx)
{
    if (!ctx) return 0;
    return (int64_t)ctx->nc * fd_halo(ctx) * ctx->cfg.Bs[0] * ctx->cfg.Bs[1];
}

int32_t wgpu_block_count(const wgpu_ctx *ctx, int32_t which)
{
    if (!ctx) return 0;
    return which == WGPU_BLOCKS_INTERIOR ? ctx->n_int : (which == WGPU_BLOCKS_BOUNDARY ? ctx->n_bnd : ctx->n_active);
}

int32_t wgpu_set_halo(wgpu_ctx *ctx, int32_t n_halo, const int32_t *halo_lgt, const int32_t *halo_hvy, const int32_t *halo_level, int32_t n_send,
                      const int32_t *send_hvy, double *send_buf)
{
    if (!ctx || n_halo < 0 || n_send < 0) return WGPU_ERR_ARG;
    if ((n_halo && (!halo_lgt || !halo_hvy || !halo_level)) || (n_send && (!send_hvy || !send_buf))) return WGPU_ERR_ARG;
    const int N = ctx->cfg.max_blocks;
    ctx->halo_map.clear();
    ctx->h_halo.clear();
    ctx->halo_level_of.clear();
    for (int k = 0; k < n_halo; ++k) {
        const int b = halo_hvy[k] - 1;
        if (b < 0 || b >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo: halo slot out of range (max_blocks must hold the own blocks and the halo copies)");
        if (k && b != ctx->h_halo.back() + 1) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo: halo slots must be consecutive, in receive order");
        if (halo_level[k] < 0 || halo_level[k] > ctx->cfg.Jmax) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo: level out of range");
        ctx->halo_map[halo_lgt[k]] = b;
        ctx->h_halo.push_back(b);
        ctx->halo_level_of.push_back((signed char)halo_level[k]);
    }
    std::vector<int> sb(std::max(n_send, 1), 0);
    for (int k = 0; k < n_send; ++k) {
        sb[k] = send_hvy[k] - 1;
        if (sb[k] < 0 || sb[k] >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo: bad send entry");
    }
    int32_t rc;
    if (n_send > ctx->halo_send_cap) {
        cudaFree(ctx->d_halo_send);
        cudaFree(ctx->d_iota);
        ctx->d_halo_send = ctx->d_iota = nullptr;
        const int want = n_send + n_send / 2 + 64;
        if ((rc = dmalloc(ctx, &ctx->d_halo_send, (size_t)want)) || (rc = dmalloc(ctx, &ctx->d_iota, (size_t)want))) return rc;
        std::vector<int> iota(want);
        for (int k = 0; k < want; ++k) iota[k] = k;
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_iota, iota.data(), sizeof(int) * want, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->halo_send_cap = want;
    }
    if (n_send) {
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_halo_send, sb.data(), sizeof(int) * n_send, cudaMemcpyHostToDevice, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->n_halo_send = n_send;
    ctx->d_halo_send_buf = send_buf;
    ctx->lookup_ready = false;
    return WGPU_OK;
}

int32_t wgpu_set_halo_restrict(wgpu_ctx *ctx, int32_t n_recv, const int32_t *recv_halo_hvy, int32_t n_send, const int32_t *send_hvy, double *send_buf)
{
    if (!ctx || n_recv < 0 || n_send < 0 || (n_recv && !recv_halo_hvy) || (n_send && (!send_hvy || !send_buf))) return WGPU_ERR_ARG;
    const int N = ctx->cfg.max_blocks;
    if ((int)ctx->h_rmap.size() != N) ctx->h_rmap.assign(N, -1);
    const size_t entry = (size_t)ctx->nc * (size_t)(ctx->blk_elems >> ctx->cfg.dim);
    const size_t need = ((size_t)ctx->n_rst + n_recv) * entry;
    int32_t rc;
    if (need > ctx->rpool_cap) {
        cudaFree(ctx->d_rpool);
        ctx->d_rpool = nullptr;
        ctx->dev_bytes -= (int64_t)ctx->rpool_cap * 8;
        if ((rc = dmalloc(ctx, &ctx->d_rpool, need + need / 4))) return rc;
        ctx->rpool_cap = need + need / 4;
    }
    if (!ctx->d_rmap && (rc = dmalloc(ctx, &ctx->d_rmap, (size_t)N))) return rc;
    if (ctx->rmap_on_device) {   // the block -> rpool map was derived on the device (wgpu_set_grid): extend / query it there
        std::vector<int> r0(n_recv), s0(n_send);
        for (int k = 0; k < n_recv; ++k) {
            r0[k] = recv_halo_hvy[k] - 1;
            if (r0[k] < 0 || r0[k] >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo_restrict: halo slot out of range");
        }
        for (int k = 0; k < n_send; ++k) {
            s0[k] = send_hvy[k] - 1;
            if (s0[k] < 0 || s0[k] >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo_restrict: send block out of range");
        }
        if (n_send > ctx->rhalo_send_cap) {
            cudaFree(ctx->d_rhalo_send);
            ctx->d_rhalo_send = nullptr;
            if ((rc = dmalloc(ctx, &ctx->d_rhalo_send, (size_t)n_send + n_send / 2 + 64))) return rc;
            ctx->rhalo_send_cap = n_send + n_send / 2 + 64;
        }
        if ((rc = wgpu_topology_halo_restrict(ctx, r0, s0))) return rc;
        ctx->n_rhalo_recv = n_recv;
        ctx->n_rhalo_send = n_send;
        ctx->d_rhalo_send_buf = send_buf;
        return WGPU_OK;
    }
    for (int k = 0; k < n_recv; ++k) {
        const int b = recv_halo_hvy[k] - 1;
        if (b < 0 || b >= N) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo_restrict: halo slot out of range");
        ctx->h_rmap[b] = ctx->n_rst + k;
    }
    std::vector<int> idx(std::max(n_send, 1), 0);
    for (int k = 0; k < n_send; ++k) {
        const int b = send_hvy[k] - 1;
        if (b < 0 || b >= N || ctx->h_rmap[b] < 0 || ctx->h_rmap[b] >= ctx->n_rst)
            return fail(ctx, WGPU_ERR_ARG, "wgpu_set_halo_restrict: a block of the send list has no coarser neighbour (no filtered copy exists)");
        idx[k] = ctx->h_rmap[b];
    }
    if (n_send > ctx->rhalo_send_cap) {
        cudaFree(ctx->d_rhalo_send);
        ctx->d_rhalo_send = nullptr;
        if ((rc = dmalloc(ctx, &ctx->d_rhalo_send, (size_t)n_send + n_send / 2 + 64))) return rc;
        ctx->rhalo_send_cap = n_send + n_send / 2 + 64;
    }
    if (n_send) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rhalo_send, idx.data(), sizeof(int) * n_send, cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_rmap, ctx->h_rmap.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->n_rhalo_recv = n_recv;
    ctx->n_rhalo_send = n_send;
    ctx->d_rhalo_send_buf = send_buf;
    return WGPU_OK;
}

int32_t wgpu_restrict_pack(wgpu_ctx *ctx, int32_t array_id, int32_t slot)
{
    if (!ctx) return WGPU_ERR_ARG;
    int nc = 0;
    const double *src = array_ptr(ctx, array_id, slot, &nc);
    if (!src || nc != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_restrict_pack: bad array/slot");
    bool active = false;
    int32_t rc = wgpu_launch_restrict_filter(ctx, src, nc, &active);
    if (rc) return rc;
    if (!active || ctx->n_rhalo_send == 0) return WGPU_OK;
    return wgpu_launch_copy_entries(ctx, ctx->d_rpool, ctx->d_rhalo_send_buf, ctx->d_rhalo_send, ctx->n_rhalo_send,
                                    (long long)ctx->nc * (ctx->blk_elems >> ctx->cfg.dim));
}

int32_t wgpu_restrict_halo_pointer(wgpu_ctx *ctx, void **ptr, int64_t *n_doubles)
{
    if (!ctx || !ptr || !n_doubles) return WGPU_ERR_ARG;
    const int64_t entry = (int64_t)ctx->nc * (ctx->blk_elems >> ctx->cfg.dim);
    *n_doubles = (int64_t)ctx->n_rhalo_recv * entry;
    *ptr = ctx->n_rhalo_recv ? ctx->d_rpool + (int64_t)ctx->n_rst * entry : nullptr;
    return WGPU_OK;
}

int32_t wgpu_pack_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot)
{
    if (!ctx) return WGPU_ERR_ARG;
    int nc = 0;
    const double *src = array_ptr(ctx, array_id, slot, &nc);
    if (!src || nc != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_pack_blocks: bad array/slot");
    return wgpu_launch_copy_blocks(ctx, src, ctx->d_halo_send_buf, ctx->d_halo_send, ctx->d_iota, ctx->n_halo_send);
}

// whole blocks of a resident array <-> a contiguous device buffer of the caller (block k of the list at k*n_eqn*Bs^dim doubles)
static int32_t gather_scatter(wgpu_ctx *ctx, bool gather, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, double *buf)
{
    if (!ctx || n < 0 || (n > 0 && (!hvy_ids || !buf))) return WGPU_ERR_ARG;
    int nc = 0;
    double *arr = array_ptr(ctx, array_id, slot, &nc);
    if (!arr || nc != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_gather_blocks / wgpu_scatter_blocks: bad array/slot");
    if (n == 0) return WGPU_OK;
    std::vector<int> ids(2 * (size_t)n);
    for (int k = 0; k < n; ++k) {
        ids[k] = hvy_ids[k] - 1;
        ids[n + k] = k;
        if (ids[k] < 0 || ids[k] >= ctx->cfg.max_blocks) return fail(ctx, WGPU_ERR_ARG, "wgpu_gather_blocks / wgpu_scatter_blocks: hvy id out of range");
    }
    int32_t rc = upload_ids(ctx, 2, ids);
    if (rc) return rc;
    const int *d_blk = ctx->d_idbuf[2], *d_seq = ctx->d_idbuf[2] + n;
    rc = gather ? wgpu_launch_copy_blocks(ctx, arr, buf, d_blk, d_seq, n) : wgpu_launch_copy_blocks(ctx, buf, arr, d_seq, d_blk, n);
    if (rc) return rc;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // the id list is reused by the next call
    if (!gather) {
        if (arr == ctx->U) ctx->dtmin_valid = false;
        ctx->det_cached_for = nullptr;
    }
    return WGPU_OK;
}

int32_t wgpu_gather_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, double *device_buf)
{
    return gather_scatter(ctx, true, array_id, slot, n, hvy_ids, device_buf);
}

int32_t wgpu_scatter_blocks(wgpu_ctx *ctx, int32_t array_id, int32_t slot, int32_t n, const int32_t *hvy_ids, const double *device_buf)
{
    return gather_scatter(ctx, false, array_id, slot, n, hvy_ids, const_cast<double *>(device_buf));
}

int32_t wgpu_halo_pointer(wgpu_ctx *ctx, int32_t array_id, int32_t slot, void **ptr, int64_t *n_doubles)
{
    if (!ctx || !ptr || !n_doubles) return WGPU_ERR_ARG;
    int nc = 0;
    double *base = array_ptr(ctx, array_id, slot, &nc);
    if (!base || nc != ctx->nc) return fail(ctx, WGPU_ERR_ARG, "wgpu_halo_pointer: bad array/slot");
    *n_doubles = (int64_t)ctx->h_halo.size() * ctx->nc * ctx->blk_elems;
    *ptr = ctx->h_halo.empty() ? nullptr : base + (int64_t)ctx->h_halo[0] * ctx->nc * ctx->blk_elems;
    return WGPU_OK;
}

int32_t wgpu_rk_stage_halo_pointer(wgpu_ctx *ctx, int32_t stage, void **ptr, int64_t *n_doubles)
{
    if (!ctx || !ptr || !n_doubles || stage < 1 || stage > ctx->cfg.n_stages) return WGPU_ERR_ARG;
    double *base = const_cast<double *>(stage_input_of(ctx, stage));
    *n_doubles = (int64_t)ctx->h_halo.size() * ctx->nc * ctx->blk_elems;
    *ptr = ctx->h_halo.empty() ? nullptr : base + (int64_t)ctx->h_halo[0] * ctx->nc * ctx->blk_elems;
    return WGPU_OK;
}

int32_t wgpu_set_exchange(wgpu_ctx *ctx, int32_t n_recv, const int32_t *recv_hvy, const int32_t *recv_dir, double *pool, int32_t n_send,
                          const int32_t *send_hvy, const int32_t *send_dir, double *send_buf)
{
    if (!ctx || n_recv < 0 || n_send < 0) return WGPU_ERR_ARG;
    if ((n_recv && (!recv_hvy || !recv_dir || !pool)) || (n_send && (!send_hvy || !send_dir || !send_buf))) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    if (c.dim != 3 || c.Bs[0] != c.Bs[1] || c.Bs[0] != c.Bs[2]) return fail(ctx, WGPU_ERR_UNSUPPORTED, "exchange needs cubic 3-D blocks");
    if ((size_t)n_recv * 2 != ctx->remote_faces.size()) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_exchange: receive list does not cover the remote faces of the topology");
    const int N = c.max_blocks;
    const int64_t pd = wgpu_patch_doubles(ctx);
    std::vector<long long> off(n_recv > 0 ? n_recv : 1);
    std::vector<char> is_bnd(N, 0);
    for (int k = 0; k < n_recv; ++k) {
        const int b = recv_hvy[k] - 1, d = recv_dir[k];
        if (b < 0 || b >= N || d < 0 || d >= WGPU_NDIR) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_exchange: bad receive entry");
        ctx->h_nbr[(size_t)b * WGPU_NDIR + d] = -2 - k;
        off[k] = (long long)k * pd;
        is_bnd[b] = 1;
    }
    std::vector<int> ai, ab;
    for (int k = 0; k < ctx->n_active; ++k) (is_bnd[ctx->h_active[k]] ? ab : ai).push_back(ctx->h_active[k]);
    ctx->n_int = (int)ai.size();
    ctx->n_bnd = (int)ab.size();
    cudaFree(ctx->d_pool_off);
    cudaFree(ctx->d_send_blk);
    cudaFree(ctx->d_send_dir);
    ctx->d_pool_off = nullptr;
    ctx->d_send_blk = ctx->d_send_dir = nullptr;
    int32_t rc;
    if ((rc = dmalloc(ctx, &ctx->d_pool_off, (size_t)std::max(n_recv, 1)))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_send_blk, (size_t)std::max(n_send, 1)))) return rc;
    if ((rc = dmalloc(ctx, &ctx->d_send_dir, (size_t)std::max(n_send, 1)))) return rc;
    std::vector<int> sb(std::max(n_send, 1)), sd(std::max(n_send, 1));
    for (int k = 0; k < n_send; ++k) {
        sb[k] = send_hvy[k] - 1;
        sd[k] = send_dir[k];
        if (sb[k] < 0 || sb[k] >= N || sd[k] < 0 || sd[k] >= WGPU_NDIR) return fail(ctx, WGPU_ERR_ARG, "wgpu_set_exchange: bad send entry");
    }
    ctx->n_send = n_send;
    ctx->d_pool = pool;
    ctx->d_send_buf = send_buf;
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_pool_off, off.data(), sizeof(long long) * off.size(), cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_send_blk, sb.data(), sizeof(int) * sb.size(), cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_send_dir, sd.data(), sizeof(int) * sd.size(), cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_nbr, ctx->h_nbr.data(), sizeof(int) * ctx->h_nbr.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!ai.empty()) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active_int, ai.data(), sizeof(int) * ai.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (!ab.empty()) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_active_bnd, ab.data(), sizeof(int) * ab.size(), cudaMemcpyHostToDevice, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->remote_faces.clear();
    return WGPU_OK;
}

static bool exchange_pending(wgpu_ctx *ctx) { return !ctx->remote_faces.empty(); }

// ------------------------------------------------------------------------------------------------ Runge-Kutta, phase by phase
int32_t wgpu_rk_begin(wgpu_ctx *ctx, double time)
{
    if (!ctx) return WGPU_ERR_ARG;
    if (ctx->nc != ctx->cfg.dim + 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "ACM needs number_equations = dim+1");
    ctx->rk_time = time;
    const wgpu_config &c = ctx->cfg;
    if (exchange_pending(ctx)) return fail(ctx, WGPU_ERR_ARG, "topology has neighbours on other ranks: call wgpu_set_exchange first");
    unsigned long long *cur = ctx->d_dtmin + ctx->dtmin_cur;
    if (!ctx->dtmin_valid && !(c.dt_fixed > 0.0)) {
        const unsigned long long inf = 0x7FF0000000000000ULL;
        WGPU_CHECK(ctx, cudaMemcpyAsync(cur, &inf, 8, cudaMemcpyHostToDevice, ctx->stream));
        int32_t rc = wgpu_launch_dtmin(ctx, ctx->U, cur);
        if (rc) return rc;
    }
    const int s = c.n_stages, ld = s + 1;
    // "acc mode": if every stage input only uses the slope of the stage before (a_{j,l} = 0 for l < j-1; true for
    // RK4 and Euler/Heun/midpoint), no slope has to be stored: the final combination is accumulated stage by stage
    // in the array K[0] in exactly the reference's order  ((u0 + dt b1 k1) + dt b2 k2) + ...
    bool subdiag = s >= 2;
    for (int j = 1; j < s && subdiag; ++j)          // 0-based row j = stage j+1, whose previous slope is column j
        for (int l = 1; l < j; ++l)
            if (fabs(c.butcher[(size_t)j * ld + l]) >= 1.0e-8) subdiag = false;
    ctx->rk_subdiag = subdiag;
    ctx->det_cached_for = nullptr;
    ctx->rk_uin = ctx->U;
    ctx->rk_next_stage = 0;   // becomes 1 after wgpu_rk_dt
    return WGPU_OK;
}

int32_t wgpu_dtmin_pointer(wgpu_ctx *ctx, void **ptr)
{
    if (!ctx || !ptr) return WGPU_ERR_ARG;
    *ptr = ctx->d_dtmin + ctx->dtmin_cur;
    return WGPU_OK;
}

int32_t wgpu_rk_dt(wgpu_ctx *ctx, double time)
{
    if (!ctx) return WGPU_ERR_ARG;
    unsigned long long *cur = ctx->d_dtmin + ctx->dtmin_cur, *nxt = ctx->d_dtmin + (ctx->dtmin_cur ^ 1);
    int32_t rc = wgpu_launch_dt_finalize(ctx, time, cur, nxt);
    if (rc) return rc;
    ctx->dtmin_cur ^= 1;
    ctx->dtmin_valid = false;
    ctx->rk_next_stage = 1;
    return WGPU_OK;
}

static const double *stage_input(wgpu_ctx *ctx, int j) { return stage_input_of(ctx, j); }

int32_t wgpu_pack_halo(wgpu_ctx *ctx, int32_t stage)
{
    if (!ctx || stage < 1 || stage > ctx->cfg.n_stages) return WGPU_ERR_ARG;
    if (ctx->n_halo_send)   // halo mode: whole blocks for the peers' halo slots
        return wgpu_launch_copy_blocks(ctx, stage_input(ctx, stage), ctx->d_halo_send_buf, ctx->d_halo_send, ctx->d_iota, ctx->n_halo_send);
    return wgpu_launch_pack(ctx, stage_input(ctx, stage));
}

int32_t wgpu_rk_stage(wgpu_ctx *ctx, int32_t j, int32_t which)
{
    if (!ctx) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    const int s = c.n_stages, ld = s + 1;
    if (j < 1 || j > s || ctx->rk_next_stage < 1) return fail(ctx, WGPU_ERR_ARG, "wgpu_rk_stage: call wgpu_rk_begin / wgpu_rk_dt first");
    unsigned long long *dtmin_next = ctx->d_dtmin + ctx->dtmin_cur;   // reset to +inf by dt_finalize
    const double *uin = stage_input(ctx, j);
    StageArgs a;
    fill_common_args(ctx, a);
    a.u_in = uin;
    a.u0 = ctx->U;
    a.t0 = ctx->rk_time;
    a.t0_ptr = ctx->time_on_device ? ctx->d_time : nullptr;
    a.t_cj = c.butcher[(size_t)(j - 1) * ld];                  // t = time + dt*rk_coeffs(j,1), runge_kutta_generic.f90:78,122
    if (a.geom && !ctx->lookup_ready) return fail(ctx, WGPU_ERR_ARG, "analytic mask: call wgpu_set_treecodes + wgpu_set_topology first");
    const bool last = (j == s);
    // the final state may overwrite U in place (each thread reads its bases only at its own point, halos come
    // from the stage input) unless the stage input IS U (single-stage schemes): then go through UA and swap
    double *uout = last ? (uin == ctx->U ? ctx->UA : ctx->U) : ((j & 1) ? ctx->UA : ctx->UB);
    a.u_out = uout;
    const double *brow = c.butcher + (size_t)s * ld;          // final weights b_j = butcher(s+1, j+1)
    if (ctx->rk_subdiag) {
        double *ACC = ensure_K(ctx, 0);
        if (!ACC) return fail(ctx, WGPU_ERR_CUDA, "out of device memory for hvy_work");
        if (!last) {
            const double *row = c.butcher + (size_t)j * ld;   // row of stage j+1
            a.use_self = fabs(row[j]) >= 1.0e-8;
            a.coef_self = row[j];
            a.acc_in = (j == 1) ? ctx->U : ACC;
            a.acc_out = ACC;
            a.use_acc = fabs(brow[j]) >= 1.0e-8;
            a.coef_acc = brow[j];
        } else {
            a.u0 = ACC;                                        // u = acc_{s-1} + (dt b_s) k_s
            a.use_self = fabs(brow[j]) >= 1.0e-8;
            a.coef_self = brow[j];
        }
    } else {
        a.k_out = last ? nullptr : ensure_K(ctx, j - 1);       // the last slope only enters the final combination
        if (!last && !a.k_out) return fail(ctx, WGPU_ERR_CUDA, "out of device memory for hvy_work");
        // row of the tableau that forms u_out: stage j+1 input (row j+1) or the final weights (row s+1)
        const double *row = last ? brow : c.butcher + (size_t)j * ld;
        a.n_prev = 0;
        for (int l = 1; l < j; ++l) {
            if (fabs(row[l]) < 1.0e-8) continue;               // runge_kutta_generic.f90:99,144
            a.k_prev[a.n_prev] = ensure_K(ctx, l - 1);
            a.coef_prev[a.n_prev] = row[l];
            a.n_prev++;
        }
        a.use_self = fabs(row[j]) >= 1.0e-8;
        a.coef_self = row[j];
    }
    if (last && !(c.dt_fixed > 0.0)) a.dtmin_bits = dtmin_next;
    int nblk = ctx->n_active;
    // level-jump face patches of this stage input: once per stage; in halo mode again before the partition-boundary blocks, whose
    // patches read the halo copies that arrived in the meantime (the patches of interior blocks are rewritten with the same values)
    if (which != WGPU_BLOCKS_BOUNDARY || !ctx->halo_bnd.empty()) {
        int32_t rcj = wgpu_launch_jump_fill(ctx, uin);
        if (rcj) return rcj;
    }
    if (which == WGPU_BLOCKS_INTERIOR) {
        a.active = ctx->d_active_int;
        nblk = ctx->n_int;
    } else if (which == WGPU_BLOCKS_BOUNDARY) {
        a.active = ctx->d_active_bnd;
        nblk = ctx->n_bnd;
    }
    a.plain_hint = plain_hint(ctx, which);
    return wgpu_launch_stage(ctx, a, nblk);
}

}  // extern "C"

// end-of-step bookkeeping without the host read-back of dt and of the divergence flag (wgpu_rk_steps reads them once after the last step)
int32_t wgpu_rk_end_nosync(wgpu_ctx *ctx)
{
    const wgpu_config &c = ctx->cfg;
    const int s = c.n_stages;
    if (stage_input(ctx, s) == ctx->U) std::swap(ctx->U, ctx->UA);   // single-stage scheme: result was written to UA
    if (!(c.dt_fixed > 0.0)) ctx->dtmin_valid = true;
    ctx->rk_next_stage = 0;
    return WGPU_OK;
}

extern "C" {

// ------------------------------------------------------------------------------------------------ Runge-Kutta-Chebychev
int32_t wgpu_rkc_step(wgpu_ctx *ctx, double time, int32_t iteration, int32_t s, const double *mu, const double *mu_tilde, const double *nu,
                      const double *gamma_tilde, const double *c, double *dt)
{
    (void)iteration;
    if (!ctx || !mu || !mu_tilde || !nu || !gamma_tilde || !c || !dt) return WGPU_ERR_ARG;
    if (s < 4) return fail(ctx, 1715929, "runge-kutta-chebychev: s cannot be less than 4");
    if (ctx->nc != ctx->cfg.dim + 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "ACM needs number_equations = dim+1");
    if (ctx->comm && ctx->comm_world > 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wgpu_rkc_step: one rank only so far");
    if (exchange_pending(ctx)) return fail(ctx, WGPU_ERR_ARG, "topology has neighbours on other ranks");
    int32_t rc;
    if ((rc = compute_dt(ctx, time))) return rc;            // calculate_time_step: dt stays on the device
    // six registers, as the reference (runge_kutta_chebychev.f90:26-27): y00 = hvy_block; y0, y1, y2 rotate; F0, F1 swap
    const size_t n = (size_t)ctx->cfg.max_blocks * ctx->nc * ctx->blk_elems;
    for (int k = 0; k < 3; ++k)
        if (!ctx->K[k]) {
            if ((rc = dmalloc(ctx, &ctx->K[k], n))) return rc;
            WGPU_CHECK(ctx, cudaMemsetAsync(ctx->K[k], 0, n * sizeof(double), ctx->stream));
        }
    double *F0 = ctx->K[0], *F1 = ctx->K[1];
    double *y1 = ctx->UA, *y2 = ctx->UB, *spare = ctx->K[2];
    ctx->det_cached_for = nullptr;
    auto rhs_of = [&](const double *src, double *dst, double t_cj) -> int32_t {
        StageArgs a;
        fill_common_args(ctx, a);
        a.u_in = src;
        a.u0 = src;
        a.k_out = dst;
        a.t0 = time;
        a.t_cj = t_cj;                                       // tau = time + c(i-1) * dt for an explicitly time-dependent mask
        int32_t r = wgpu_launch_jump_fill(ctx, src);         // sync_ghosts_RHS_tree
        if (r) return r;
        a.plain_hint = plain_hint(ctx, WGPU_BLOCKS_ALL);
        return wgpu_launch_stage(ctx, a, ctx->n_active);
    };
    if ((rc = rhs_of(ctx->U, F0, 0.0))) return rc;                                                   // F0 = rhs(y00)
    if ((rc = wgpu_launch_rkc_combine(ctx, y1, ctx->U, ctx->U, ctx->U, F0, F0, 0.0, 0.0, 0.0, mu_tilde[0], 0.0, 0))) return rc;   // y1 = y0 + mu~_1 dt F0
    const double *y0c = ctx->U;                                                                      // y0 = y00 before the first rotation
    for (int i = 1; i < s; ++i) {                                                                    // Fortran i = 2 .. s
        if ((rc = rhs_of(y1, F1, c[i - 1]))) return rc;                                              // F1 = rhs(y1) at tau = time + c(i-1) dt
        if ((rc = wgpu_launch_rkc_combine(ctx, y2, ctx->U, y1, y0c, F1, F0, 1.0 - mu[i] - nu[i], mu[i], nu[i], mu_tilde[i], gamma_tilde[i], 1))) return rc;
        if (i < s - 1) {                                                                             // y0 <- y1 <- y2; the old y0 buffer is the next y2
            double *old0 = (y0c == ctx->U) ? spare : const_cast<double *>(y0c);
            y0c = y1;
            y1 = y2;
            y2 = old0;
        }
    }
    // u = y2: the array that holds it becomes hvy_block (pointer swap, as the single-stage path of the Runge-Kutta driver)
    if (y2 == ctx->UA) std::swap(ctx->U, ctx->UA);
    else if (y2 == ctx->UB) std::swap(ctx->U, ctx->UB);
    else std::swap(ctx->U, ctx->K[2]);
    ctx->dtmin_valid = false;
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_dt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    *dt = ctx->h_pinned[0];
    return check_flags(ctx);
}

// ------------------------------------------------------------------------------------------------ Krylov exponential integrator
int32_t wgpu_krylov_step(wgpu_ctx *ctx, double time, int32_t iteration, int32_t M_max, int32_t dynamic, double err_threshold, double *dt,
                         int32_t *M_used, double *err_out)
{
    (void)iteration;
    if (!ctx || !dt) return WGPU_ERR_ARG;
    if (M_max < 1 || M_max > 62) return fail(ctx, WGPU_ERR_ARG, "krylov: M_krylov must be in 1..62");
    if (ctx->nc != ctx->cfg.dim + 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "ACM needs number_equations = dim+1");
    if (ctx->comm && ctx->comm_world > 1) return fail(ctx, WGPU_ERR_UNSUPPORTED, "wgpu_krylov_step: one rank only so far");
    if (exchange_pending(ctx)) return fail(ctx, WGPU_ERR_ARG, "topology has neighbours on other ranks");
    int32_t rc;
    // M_max + 3 registers, as the reference (allocate_forest.f90:98): the Krylov vectors 1..M_max, slot M_max+1 (the Arnoldi work vector, which
    // becomes vector M_max+1 in place), the perturbed state and the reference right-hand side
    const size_t n = (size_t)ctx->cfg.max_blocks * ctx->nc * ctx->blk_elems;
    while ((int)ctx->kry.size() < M_max + 3) {
        double *p = nullptr;
        if ((rc = dmalloc(ctx, &p, n))) return rc;
        WGPU_CHECK(ctx, cudaMemsetAsync(p, 0, n * sizeof(double), ctx->stream));
        ctx->kry.push_back(p);
    }
    if (!ctx->d_kry_part && (rc = dmalloc(ctx, &ctx->d_kry_part, (size_t)ctx->cfg.max_blocks + 1))) return rc;
    double **V = ctx->kry.data();
    double *W = V[M_max], *P = V[M_max + 1], *R = V[M_max + 2];
    double *d_res = ctx->d_kry_part + ctx->cfg.max_blocks;
    ctx->det_cached_for = nullptr;
    auto rhs_of = [&](const double *src, double *dst) -> int32_t {
        StageArgs a;
        fill_common_args(ctx, a);
        a.u_in = src;
        a.u0 = src;
        a.k_out = dst;
        a.t0 = time;
        a.t_cj = 0.0;
        int32_t r = wgpu_launch_jump_fill(ctx, src);         // sync_ghosts_RHS_tree
        if (r) return r;
        a.plain_hint = plain_hint(ctx, WGPU_BLOCKS_ALL);
        return wgpu_launch_stage(ctx, a, ctx->n_active);
    };
    auto dot = [&](const double *x, const double *y, double *res) -> int32_t {      // scalarproduct + get_sum_all: the host needs the value
        int32_t r = wgpu_launch_kry_dot(ctx, x, y, ctx->d_kry_part, d_res);
        if (r) return r;
        WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, d_res, 8, cudaMemcpyDeviceToHost, ctx->stream));
        WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
        *res = ctx->h_pinned[0];
        return WGPU_OK;
    };
    if ((rc = compute_dt(ctx, time))) return rc;             // calculate_time_step
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_dt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    double dtv = ctx->h_pinned[0];
    const double epsm = 2.220446049250313e-16;               // epsilon(1.0_rk)
    double normv, beta, h;
    if ((rc = dot(ctx->U, ctx->U, &normv))) return rc;
    normv = sqrt(normv);
    if (normv < epsm) normv = 1.0;
    const double eps = normv * sqrt(epsm);
    if ((rc = rhs_of(ctx->U, R))) return rc;
    if ((rc = check_flags(ctx))) return rc;
    if ((rc = dot(R, R, &beta))) return rc;
    beta = sqrt(beta);
    if (beta < epsm) beta = 1.0;
    if ((rc = wgpu_launch_kry_axpy(ctx, V[0], R, R, 0, beta, 0.0))) return rc;
    const int LD = M_max + 2;
    std::vector<double> H((size_t)LD * LD, 0.0), phi((size_t)LD * LD, 0.0), Ht, Ex;
    double err = 0.0;
    int M_iter = 0;
    auto phi_of = [&](int M, double h_klein) -> int32_t {    // expM_pade(dt * H_tmp) of the augmented (M+2)^2 matrix, then the error entry
        const int m = M + 2;
        Ht.assign((size_t)m * m, 0.0);
        Ex.assign((size_t)m * m, 0.0);
        for (int i = 0; i < M; ++i)
            for (int j = 0; j < M; ++j) Ht[(size_t)i * m + j] = dtv * H[(size_t)i * LD + j];
        Ht[(size_t)0 * m + M] = dtv * 1.0;
        Ht[(size_t)M * m + M + 1] = dtv * 1.0;
        int32_t r = wgpu_expm_pade(Ht.data(), m, Ex.data());
        if (r) return fail(ctx, 240917, "error in computing exp(t*H)");
        std::fill(phi.begin(), phi.end(), 0.0);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) phi[(size_t)i * LD + j] = Ex[(size_t)i * m + j];
        phi[(size_t)M * LD + M] = h_klein * phi[(size_t)(M - 1) * LD + M + 1];
        err = fabs(beta * phi[(size_t)M * LD + M]);
        return WGPU_OK;
    };
    for (M_iter = 1; M_iter <= M_max; ++M_iter) {
        if ((rc = wgpu_launch_kry_axpy(ctx, P, ctx->U, V[M_iter - 1], 1, eps, 0.0))) return rc;      // perturbed state
        if ((rc = rhs_of(P, W))) return rc;
        if ((rc = wgpu_launch_kry_axpy(ctx, W, W, R, 2, eps, 0.0))) return rc;                       // linearization
        for (int it = 1; it <= M_iter; ++it) {                                                       // Arnoldi, modified Gram-Schmidt
            if ((rc = dot(V[it - 1], W, &h))) return rc;
            H[(size_t)(it - 1) * LD + (M_iter - 1)] = h;
            if ((rc = wgpu_launch_kry_axpy(ctx, W, W, V[it - 1], 3, h, 0.0))) return rc;
        }
        if ((rc = dot(W, W, &h))) return rc;
        h = sqrt(h);
        H[(size_t)M_iter * LD + (M_iter - 1)] = h;
        if ((rc = wgpu_launch_kry_axpy(ctx, V[M_iter], W, W, 0, h, 0.0))) return rc;                 // M_iter = M_max: in place
        if (dynamic || M_iter == M_max) {
            if ((rc = phi_of(M_iter, h))) return rc;
            if (dynamic && M_iter == M_max && err > err_threshold) {
                // the largest subspace and the error is still too large: decrease the time step
                int guard = 0;
                while (err > err_threshold && guard++ < 2000) {
                    dtv = 0.90 * dtv;
                    if ((rc = phi_of(M_iter, h))) return rc;
                }
            }
            if (err <= err_threshold || M_iter == M_max) break;
        }
    }
    if (M_iter > M_max) M_iter = M_max;
    for (int it = 1; it <= M_iter + 1; ++it)
        if ((rc = wgpu_launch_kry_axpy(ctx, ctx->U, ctx->U, V[it - 1], 4, beta, phi[(size_t)(it - 1) * LD + M_iter]))) return rc;
    ctx->dtmin_valid = false;
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    *dt = dtv;
    if (M_used) *M_used = M_iter;
    if (err_out) *err_out = err;
    return check_flags(ctx);
}

// ------------------------------------------------------------------------------------------------ explicit filter
int32_t wgpu_filter(wgpu_ctx *ctx, const char *filter_type, const int32_t *filter_component, int32_t only_maxlevel, int32_t all_except_maxlevel)
{
    if (!ctx || !filter_type) return WGPU_ERR_ARG;
    if (only_maxlevel && all_except_maxlevel)
        return fail(ctx, 251106, "ERROR: Do you want to filter only on max level or all except max level??? Choose one, not both.");
    // filter_wrapper.f90:28-60: explicit_(2n+1)pt == superviscosity_(2n)th, n = 1..10
    int order = 0;
    int pts = 0, ord = 0;
    if (sscanf(filter_type, "explicit_%dpt", &pts) == 1 && pts >= 3 && pts <= 21 && (pts & 1)) order = pts - 1;
    else if (sscanf(filter_type, "superviscosity_%d", &ord) == 1 && ord >= 2 && ord <= 20 && !(ord & 1)) order = ord;
    if (!order) return fail(ctx, 251107, std::string("ERROR: Filter not known: ") + filter_type);
    const int a = order / 2;
    if (a > ctx->cfg.g) return fail(ctx, 251108, "ERROR: Nice filter you've selected there, but its stencil size exceeds the ghost layer thickness. Increase g.");
    // generate_superviscosity_stencil (filter_wrapper.f90:82-104): (-1)^(k+a) binom(2a, a+k), normalised by the sum of the absolute values; negated
    // for the orders 4, 8, 12, ...; then stencil(0) + 1
    double st[2 * WGPU_FMAX + 1], sum_abs = 0.0;
    for (int k = -a; k <= a; ++k) {
        double binom = 1.0;
        for (int j = 1; j <= a + k; ++j) binom = binom * (double)(2 * a - (a + k) + j) / (double)j;     // exact for these sizes
        binom = floor(binom + 0.5);
        st[k + a] = (((k + a) & 1) ? -1.0 : 1.0) * binom;
        sum_abs += fabs(st[k + a]);
    }
    for (int k = 0; k <= 2 * a; ++k) st[k] = st[k] / sum_abs;
    if ((order / 2) % 2 == 0)
        for (int k = 0; k <= 2 * a; ++k) st[k] = -st[k];
    st[a] = st[a] + 1.0;
    unsigned mask = 0;
    for (int c = 0; c < ctx->nc; ++c)
        if (!filter_component || filter_component[c]) mask |= 1u << c;
    if (!ctx->TMP) {                      // hvy_tmp: normally allocated by wgpu_set_wavelet
        const size_t n = (size_t)ctx->cfg.max_blocks * ctx->nc * ctx->blk_elems;
        int32_t rca = dmalloc(ctx, &ctx->TMP, n);
        if (rca) return rca;
        WGPU_CHECK(ctx, cudaMemsetAsync(ctx->TMP, 0, n * sizeof(double), ctx->stream));
    }
    int32_t rc = wgpu_launch_blockfilter(ctx, ctx->U, ctx->TMP, st, a, mask, only_maxlevel ? 1 : (all_except_maxlevel ? 2 : 0));
    if (rc) return rc;
    std::swap(ctx->U, ctx->TMP);          // the filtered array becomes hvy_block
    ctx->dtmin_valid = false;
    ctx->det_cached_for = nullptr;
    return WGPU_OK;
}

int32_t wgpu_rk_end(wgpu_ctx *ctx, double *dt)
{
    if (!ctx || !dt) return WGPU_ERR_ARG;
    wgpu_rk_end_nosync(ctx);
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned, ctx->d_dt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->h_pinned + 1, ctx->d_flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    *dt = ctx->h_pinned[0];
    if (*(int *)(ctx->h_pinned + 1)) {
        cudaMemsetAsync(ctx->d_flags, 0, sizeof(int), ctx->stream);
        return fail(ctx, WGPU_ERR_DIVERGED, "ACM fail: very very large values in state vector.");
    }
    return WGPU_OK;
}

int32_t wgpu_rk_step(wgpu_ctx *ctx, double time, int32_t iteration, double *dt)
{
    (void)iteration;
    if (!ctx || !dt) return WGPU_ERR_ARG;
    int32_t rc;
    // runge_kutta_generic.f90:52-56: ghost sync (fused into the stage kernels) and the time step
    if ((rc = wgpu_rk_begin(ctx, time))) return rc;
    if ((rc = wgpu_rk_dt(ctx, time))) return rc;
    for (int j = 1; j <= ctx->cfg.n_stages; ++j) {
        if ((rc = wgpu_pack_halo(ctx, j))) return rc;   // no-op without remote neighbours
        if ((rc = wgpu_rk_stage(ctx, j, WGPU_BLOCKS_ALL))) return rc;
    }
    return wgpu_rk_end(ctx, dt);
}

}  // extern "C"
