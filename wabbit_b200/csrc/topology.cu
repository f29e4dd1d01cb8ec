// Topology on the device: the neighbour relations of the active blocks, derived on the GPU from the block positions.
//
// The reference keeps the neighbourhood as light data: find_neighbors (LIB/MESH/find_neighbors.f90:18-180) fills the 168-slot
// hvy_neighbor table on the host every time the grid changes (updateNeighbors_tree.f90, updateMetadata_tree.f90), and the reference
// itself flags that step as "very expensive > 50k blocks" (ini_file_to_params.f90:226-228).  wgpu_set_topology consumes that table
// (the Fortran-facing entry point).  wgpu_set_grid is the device-native alternative: the host names the resident blocks (slot, level,
// treecode) and the list of active ones, nothing else; a hash (level, ix, iy, iz) -> slot is built on the device and one thread per
// (active block, direction) classifies the relation exactly as find_neighbor does -- same level, finer (a virtual daughter's
// neighbour exists one level up) or coarser (the block one level down that covers the neighbour position) -- from which the 27-entry
// gather table of the stencil / wavelet kernels, the level-jump patch lists, the coarse-extension list, the list of blocks that send
// restricted data and the interior / partition-boundary split are compacted in the order of the active list (stable: the lists are
// identical to the ones wgpu_set_topology derives from hvy_neighbor, tests/test_gpu_topology.py).
//
// Full-tree passes of adapt_tree register leaves AND mothers: a leaf next to a refined region then finds the region's mother on its own
// level (what sync_TMP_from_MF provides), so those passes need no table of their own -- only their active list (wgpu_set_active).
#include <algorithm>
#include <vector>

#include "resolve.cuh"
#include "wgpu_internal.cuh"

namespace {

enum { REL_NONE = 0, REL_SAME = 1, REL_COARSER = 2, REL_FINER = 3, REL_NOOWNER = 4, REL_KIND = 7, REL_HALO = 8 };
enum { C_JUMP = 0, C_WJUMP, C_WSIZE, C_CE, C_RST, C_BND, C_INT, C_NROWS };

struct TopoArgs {
    BlockLookup L;
    unsigned long long *keys;
    int *vals;
    const int *active;
    int n_active;
    signed char *level;      // [max_blocks]
    int *ixyz;               // [max_blocks][3]
    unsigned char *bflag;    // [max_blocks] bit0 registered, bit1 halo copy
    int *nbr, *wnbr;         // [max_blocks][27]
    unsigned char *rel;      // [n_active][27]
    long long *cnt;          // [C_NROWS][n_active]
    int dim, Bs, F, nc;
    int periodic[3];
    int has_jumps;           // fill pass: global flag
    // outputs of the fill pass
    int *jump_blk, *jump_dir, *wjump_blk, *wjump_dir, *ce_blk, *ce_dir, *rst_blk, *rmap, *active_int, *active_bnd;
    long long *woff;
    unsigned *rst_mask;
    int *err;                // [0] duplicate position, [1] bit0: some coarser / finer relation, bit1: a finer neighbour is a halo copy, [2] bad send list
};

// registration: decode the numerical treecode (decoding_b, LIB/TREE/module_treelib.f90:793-831: digit bit0 -> y, bit1 -> x, bit2 -> z; bit i of
// a coordinate sits in digit i + Jmax - level), remember level / position per slot and insert the position into the hash
__global__ void __launch_bounds__(256) topo_register_kernel(TopoArgs a, int n, const int *__restrict__ ids, const int *__restrict__ lvl,
                                                            const long long *__restrict__ tc, int Jmax)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = ids[i], J = lvl[i];
    int p[3] = {0, 0, 0};
    for (int d = 0; d < a.dim; ++d)
        for (int k = 0; k < J; ++k) p[d] |= (int)((tc[i] >> ((k + Jmax - J) * a.dim + d)) & 1) << k;
    const int x = p[1], y = p[0], z = p[2];
    if (b >= 0) {
        a.ixyz[3 * b] = x;
        a.ixyz[3 * b + 1] = y;
        a.ixyz[3 * b + 2] = z;
        a.level[b] = (signed char)J;
        a.bflag[b] |= 1;
    }
    const unsigned long long key = blk_key(J, x, y, z);
    unsigned h = blk_hash(key) & a.L.mask;
    for (;;) {
        const unsigned long long old = atomicCAS(a.keys + h, ~0ull, key);
        if (old == ~0ull) {
            a.vals[h] = b >= 0 ? b : -2;   // -2: the block exists in the tree, but its data are not resident here (position only)
            return;
        }
        if (old == key) {   // two resident blocks share a position
            atomicExch(a.err, 1);
            return;
        }
        h = (h + 1) & a.L.mask;
    }
}

__global__ void __launch_bounds__(256) topo_mark_halo_kernel(unsigned char *bflag, const int *__restrict__ halo, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bflag[halo[i]] |= 2;
}

// one thread per (active block, direction): the relation find_neighbor establishes (LIB/MESH/find_neighbors.f90:60-180)
__global__ void __launch_bounds__(256) topo_rel_kernel(TopoArgs a)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)a.n_active * 27) return;
    const int k = (int)(t / 27), dir = (int)(t % 27);
    const int b = a.active[k];
    const int d[3] = {dir % 3 - 1, (dir / 3) % 3 - 1, dir / 9 - 1};
    unsigned char rel = REL_NONE;
    int entry = -1;
    if (dir != 13 && !(a.dim == 2 && d[2] != 0)) {
        const int L = a.level[b], nb = 1 << L;
        const int p[3] = {a.ixyz[3 * b], a.ixyz[3 * b + 1], a.ixyz[3 * b + 2]};
        int q[3] = {0, 0, 0};
        bool outside = false;
        for (int ax = 0; ax < a.dim; ++ax) {
            q[ax] = p[ax] + d[ax];
            if (q[ax] < 0 || q[ax] >= nb) {
                if (a.periodic[ax]) q[ax] = (q[ax] + nb) & (nb - 1);
                else outside = true;
            }
        }
        if (!outside) {
            int j = blk_lookup(a.L, L, q[0], q[1], q[2]);
            if (j >= 0) {
                rel = REL_SAME | ((a.bflag[j] & 2) ? REL_HALO : 0);
                entry = j;
            } else if (j == -2) {
                rel = REL_NOOWNER;   // a same-level block exists, but not here: no relation a kernel could use, and none to other levels
            } else {
                // finer: the neighbours, one level up, of the virtual daughters that touch this side
                bool finer = false, halo = false;
                const int nb2 = nb * 2;
                for (int c = 0; c < (1 << a.dim); ++c) {
                    bool touches = true;
                    int r[3] = {0, 0, 0};
                    for (int ax = 0; ax < a.dim; ++ax) {
                        const int bit = (c >> ax) & 1;
                        if ((d[ax] > 0 && !bit) || (d[ax] < 0 && bit)) touches = false;
                        r[ax] = (2 * p[ax] + bit + d[ax] + nb2) & (nb2 - 1);
                    }
                    if (!touches) continue;
                    const int f = blk_lookup(a.L, L + 1, r[0], r[1], r[2]);
                    if (f >= 0) {
                        finer = true;
                        halo = halo || (a.bflag[f] & 2);
                    }
                }
                if (finer) rel = REL_FINER | (halo ? REL_HALO : 0);
                else {
                    // coarser: the block one level down that covers the neighbour position.  find_neighbor registers it only for the block
                    // in the matching corner of its mother (last treecode digit, find_neighbors.f90:96-125): an edge / corner position that a
                    // coarser FACE neighbour covers is not a relation of its own (its ghost patch is part of the face neighbour's patch)
                    j = L > 0 ? blk_lookup(a.L, L - 1, q[0] >> 1, q[1] >> 1, q[2] >> 1) : -1;
                    bool corner = true;
                    for (int ax = 0; ax < a.dim; ++ax)
                        if (d[ax] != 0 && (p[ax] & 1) != (d[ax] > 0 ? 1 : 0)) corner = false;
                    rel = ((j >= 0 || j == -2) && corner) ? REL_COARSER : REL_NOOWNER;   // (-2: known by position only; the relation stands)
                    if (j >= 0 && (a.bflag[j] & 2)) rel |= REL_HALO;
                }
            }
        }
    }
    a.rel[t] = rel;
    a.nbr[(long long)b * 27 + dir] = entry;
}

__device__ __forceinline__ long long wpatch_size(const TopoArgs &a, int dir)
{
    const int d[3] = {dir % 3 - 1, (dir / 3) % 3 - 1, dir / 9 - 1};
    return (long long)a.nc * (d[0] ? a.F : a.Bs) * (d[1] ? a.F : a.Bs) * (a.dim == 3 ? (d[2] ? a.F : a.Bs) : 1);
}

// pass 0: per-block counts; pass 1 (after the scans): write the lists at the scanned offsets, block by block, directions ascending
template <int PASS>
__global__ void __launch_bounds__(256) topo_lists_kernel(TopoArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_active) return;
    const int b = a.active[k], n = a.n_active;
    long long oj = 0, ow = 0, os = 0, oc = 0;
    if (PASS == 1) {
        oj = a.cnt[(long long)C_JUMP * n + k];
        ow = a.cnt[(long long)C_WJUMP * n + k];
        os = a.cnt[(long long)C_WSIZE * n + k];
        oc = a.cnt[(long long)C_CE * n + k];
    }
    long long cj = 0, cw = 0, cs = 0, cc = 0;
    unsigned jump_mask = 0;
    bool has_coarser = false, uses_halo = false, fine_halo = false;
    for (int dir = 0; dir < 27; ++dir) {
        const unsigned char r = a.rel[(long long)k * 27 + dir];
        const int kind = r & REL_KIND;
        if (kind == REL_NONE) {
            if (PASS == 1) a.wnbr[(long long)b * 27 + dir] = -1;
            continue;
        }
        if (r & REL_HALO) uses_halo = true;
        if (kind == REL_FINER && (r & REL_HALO)) fine_halo = true;
        const int nz = (dir % 3 != 1) + ((dir / 3) % 3 != 1) + (dir / 9 != 1);
        const bool jump = kind == REL_COARSER || kind == REL_FINER;
        int entry = PASS == 1 ? a.nbr[(long long)b * 27 + dir] : -1, wentry = entry;
        if (jump) {
            jump_mask |= 1u << dir;
            if (nz == 1) {   // faces become restriction / prediction patches of the stage kernel's jump pool
                if (PASS == 1) {
                    a.jump_blk[oj + cj] = b;
                    a.jump_dir[oj + cj] = dir;
                    entry = -2 - (WGPU_JUMP_PID + (int)(oj + cj));
                    wentry = -1;
                }
                ++cj;
            }
        }
        if (kind == REL_COARSER) {
            has_coarser = true;
            if (PASS == 1) {
                a.ce_blk[oc + cc] = b;
                a.ce_dir[oc + cc] = dir;
            }
            ++cc;
        }
        if (kind != REL_SAME) {   // every direction without a same-level block (inside the domain) is a ghost patch of the wavelet kernels
            if (PASS == 1 && a.has_jumps) {
                a.wjump_blk[ow + cw] = b;
                a.wjump_dir[ow + cw] = dir;
                a.woff[ow + cw] = os + cs;
                wentry = -2 - (int)(ow + cw);
            }
            ++cw;
            cs += wpatch_size(a, dir);
        }
        if (PASS == 1) {
            a.nbr[(long long)b * 27 + dir] = entry;
            a.wnbr[(long long)b * 27 + dir] = wentry;
        }
    }
    if (PASS == 0) {
        a.cnt[(long long)C_JUMP * n + k] = cj;
        a.cnt[(long long)C_WJUMP * n + k] = cw;
        a.cnt[(long long)C_WSIZE * n + k] = cs;
        a.cnt[(long long)C_CE * n + k] = cc;
        a.cnt[(long long)C_RST * n + k] = has_coarser ? 1 : 0;
        a.cnt[(long long)C_BND * n + k] = uses_halo ? 1 : 0;
        a.cnt[(long long)C_INT * n + k] = uses_halo ? 0 : 1;
        const int v = (jump_mask ? 1 : 0) | (fine_halo ? 2 : 0);
        if (v) atomicOr(a.err + 1, v);
    } else {
        if (has_coarser) {
            const long long o = a.cnt[(long long)C_RST * n + k];
            a.rst_blk[o] = b;
            a.rst_mask[o] = jump_mask;
            a.rmap[b] = (int)o;
        }
        if (uses_halo) a.active_bnd[a.cnt[(long long)C_BND * n + k]] = b;
        else a.active_int[a.cnt[(long long)C_INT * n + k]] = b;
    }
}

// exclusive scan of every row of cnt[C_NROWS][n] in place (one CTA per row), totals[row] = sum
__global__ void __launch_bounds__(1024) topo_scan_kernel(long long *cnt, int n, long long *totals)
{
    __shared__ long long part[1024];
    long long *row = cnt + (long long)blockIdx.x * n;
    const int tid = threadIdx.x, chunk = (n + 1023) / 1024;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    long long s = 0;
    for (int i = lo; i < hi; ++i) s += row[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    long long run = part[tid] - s;
    for (int i = lo; i < hi; ++i) {
        const long long v = row[i];
        row[i] = run;
        run += v;
    }
    if (tid == 1023) totals[blockIdx.x] = part[1023];
}

__global__ void __launch_bounds__(256) topo_rhalo_kernel(int *rmap, const int *__restrict__ recv, int n_recv, int n_rst, int *send_idx,
                                                         const int *__restrict__ send, int n_send, int *err)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_recv) rmap[recv[i]] = n_rst + i;
    if (i < n_send) {
        const int r = rmap[send[i]];
        if (r < 0 || r >= n_rst) atomicExch(err + 2, 1);
        send_idx[i] = r;
    }
}

template <typename T>
int32_t ensure(wgpu_ctx *ctx, T **p, size_t *cap, size_t need, size_t slack)
{
    if (need <= *cap && *p) return WGPU_OK;
    if (*p) {
        cudaFree(*p);
        ctx->dev_bytes -= (int64_t)(*cap * sizeof(T));
    }
    *p = nullptr;
    *cap = 0;
    const size_t want = need + slack;
    WGPU_CHECK(ctx, cudaMalloc((void **)p, want * sizeof(T)));
    ctx->dev_bytes += (int64_t)(want * sizeof(T));
    *cap = want;
    return WGPU_OK;
}

int32_t topo_fail(wgpu_ctx *ctx, int32_t code, const char *msg)
{
    ctx->err = msg;
    return code;
}

TopoArgs make_args(wgpu_ctx *ctx)
{
    TopoArgs a;
    memset(&a, 0, sizeof(a));
    const wgpu_config &c = ctx->cfg;
    a.L.keys = ctx->d_hkeys;
    a.L.vals = ctx->d_hvals;
    a.L.mask = ctx->hmask;
    a.keys = ctx->d_hkeys;
    a.vals = ctx->d_hvals;
    a.active = ctx->d_active;
    a.n_active = ctx->n_active;
    a.level = ctx->d_level;
    a.ixyz = ctx->d_ixyz;
    a.bflag = ctx->d_bflag;
    a.nbr = ctx->d_nbr;
    a.wnbr = ctx->d_wnbr;
    a.rel = ctx->d_rel;
    a.cnt = ctx->d_cnt;
    a.dim = c.dim;
    a.Bs = c.Bs[0];
    a.nc = ctx->nc;
    a.err = ctx->d_flags + 2;
    for (int k = 0; k < 3; ++k) a.periodic[k] = c.periodic[k];
    if (ctx->wavelet_set) {
        const WaveFilters &w = ctx->wavelet;
        a.F = std::max(std::max(-w.hd_lo, w.hd_hi), std::max(-w.hr_lo, w.hr_hi));
    }
    return a;
}

// the relation / list passes for the current active list (d_active, n_active already set)
int32_t derive_lists(wgpu_ctx *ctx)
{
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks, n = ctx->n_active;
    int32_t rc;
    ctx->has_jumps = false;
    ctx->n_jump = ctx->n_wjump = ctx->n_ce = ctx->n_rst = 0;
    ctx->n_int = n;
    ctx->n_bnd = 0;
    ctx->halo_bnd.clear();
    ctx->halo_fine_neighbor = false;
    ctx->n_rhalo_recv = ctx->n_rhalo_send = 0;
    ctx->remote_faces.clear();
    ctx->dtmin_valid = false;
    ctx->det_cached_for = nullptr;
    if (n == 0) return WGPU_OK;
    if ((rc = ensure(ctx, &ctx->d_rel, &ctx->rel_cap, (size_t)n * 27, (size_t)n * 27 / 4 + 1024))) return rc;
    if ((rc = ensure(ctx, &ctx->d_cnt, &ctx->cnt_cap, (size_t)n * C_NROWS + 16, (size_t)n * C_NROWS / 4 + 1024))) return rc;
    if (!ctx->d_wnbr) {
        size_t cap = 0;
        if ((rc = ensure(ctx, &ctx->d_wnbr, &cap, (size_t)N * WGPU_NDIR, 0))) return rc;
    }
    if (!ctx->d_rmap) {
        size_t cap = 0;
        if ((rc = ensure(ctx, &ctx->d_rmap, &cap, (size_t)N, 0))) return rc;
    }
    if (!ctx->d_active_int) {
        size_t cap = 0;
        if ((rc = ensure(ctx, &ctx->d_active_int, &cap, (size_t)N, 0)) || (cap = 0, rc = ensure(ctx, &ctx->d_active_bnd, &cap, (size_t)N, 0))) return rc;
    }
    TopoArgs a = make_args(ctx);
    const long long nt = (long long)n * 27;
    topo_rel_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, ctx->stream>>>(a);
    topo_lists_kernel<0><<<(n + 255) / 256, 256, 0, ctx->stream>>>(a);
    long long *d_tot = ctx->d_cnt + (size_t)n * C_NROWS;
    topo_scan_kernel<<<C_NROWS, 1024, 0, ctx->stream>>>(ctx->d_cnt, n, d_tot);
    ctx->launches += 3;
    long long tot[C_NROWS];
    int flags[8];
    WGPU_CHECK(ctx, cudaMemcpyAsync(tot, d_tot, sizeof(tot), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    WGPU_CHECK(ctx, cudaGetLastError());
    if (flags[2]) {
        cudaMemsetAsync(ctx->d_flags + 2, 0, 3 * sizeof(int), ctx->stream);
        return topo_fail(ctx, WGPU_ERR_ARG, "two resident blocks share a treecode");
    }
    // flags[3] bit0: a relation to a coarser / finer block exists somewhere (the count of wavelet patches alone does not say so: a
    // direction inside the domain may simply have no owner among the resident blocks)
    ctx->n_jump = (int)tot[C_JUMP];
    ctx->n_ce = (int)tot[C_CE];
    ctx->n_rst = (int)tot[C_RST];
    ctx->n_bnd = (int)tot[C_BND];
    ctx->n_int = (int)tot[C_INT];
    ctx->has_jumps = (flags[3] & 1) != 0;
    if (ctx->has_jumps) {
        if (!ctx->wavelet_set) return topo_fail(ctx, WGPU_ERR_ARG, "grid has level jumps: call wgpu_set_wavelet first (the predictor order is the wavelet's)");
        if (c.Bs[0] != c.Bs[1] || (c.dim == 3 && c.Bs[0] != c.Bs[2])) return topo_fail(ctx, WGPU_ERR_UNSUPPORTED, "level jumps need square / cubic blocks so far");
    }
    ctx->n_wjump = ctx->has_jumps ? (int)tot[C_WJUMP] : 0;
    ctx->wjump_depth = a.F;
    // room for the lists and the pools
    const int H = c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3);
    size_t cap;
    if (ctx->n_jump > ctx->jump_cap) {
        cudaFree(ctx->d_jump_blk);
        cudaFree(ctx->d_jump_dir);
        ctx->d_jump_blk = ctx->d_jump_dir = nullptr;
        const size_t want = (size_t)ctx->n_jump + ctx->n_jump / 2 + 64;
        cap = 0;
        if ((rc = ensure(ctx, &ctx->d_jump_blk, &cap, want, 0)) || (cap = 0, rc = ensure(ctx, &ctx->d_jump_dir, &cap, want, 0))) return rc;
        ctx->jump_cap = (int)want;
    }
    {
        const size_t need = (size_t)std::max(ctx->n_jump, 1) * ctx->nc * H * c.Bs[0] * (c.dim == 3 ? c.Bs[1] : 1);
        if (need > ctx->jpool_cap) {
            cudaFree(ctx->d_jpool);
            ctx->d_jpool = nullptr;
            ctx->dev_bytes -= (int64_t)ctx->jpool_cap * 8;
            ctx->jpool_cap = 0;
            cap = 0;
            if ((rc = ensure(ctx, &ctx->d_jpool, &cap, need + need / 2, 0))) return rc;
            ctx->jpool_cap = cap;
        }
    }
    if (ctx->n_wjump > ctx->wjump_cap) {
        cudaFree(ctx->d_wjump_blk);
        cudaFree(ctx->d_wjump_dir);
        cudaFree(ctx->d_woff);
        ctx->d_wjump_blk = ctx->d_wjump_dir = nullptr;
        ctx->d_woff = nullptr;
        const size_t want = (size_t)ctx->n_wjump + ctx->n_wjump / 2 + 64;
        cap = 0;
        if ((rc = ensure(ctx, &ctx->d_wjump_blk, &cap, want, 0)) || (cap = 0, rc = ensure(ctx, &ctx->d_wjump_dir, &cap, want, 0)) ||
            (cap = 0, rc = ensure(ctx, &ctx->d_woff, &cap, want, 0)))
            return rc;
        ctx->wjump_cap = (int)want;
    }
    if (ctx->has_jumps && (size_t)tot[C_WSIZE] > ctx->wpool_cap) {
        cudaFree(ctx->d_wpool);
        ctx->d_wpool = nullptr;
        ctx->dev_bytes -= (int64_t)ctx->wpool_cap * 8;
        ctx->wpool_cap = 0;
        cap = 0;
        if ((rc = ensure(ctx, &ctx->d_wpool, &cap, (size_t)tot[C_WSIZE] + (size_t)tot[C_WSIZE] / 4, 0))) return rc;
        ctx->wpool_cap = cap;
    }
    if (ctx->n_ce > ctx->ce_cap) {
        cudaFree(ctx->d_ce_blk);
        cudaFree(ctx->d_ce_dir);
        ctx->d_ce_blk = ctx->d_ce_dir = nullptr;
        const size_t want = (size_t)ctx->n_ce + ctx->n_ce / 2 + 64;
        cap = 0;
        if ((rc = ensure(ctx, &ctx->d_ce_blk, &cap, want, 0)) || (cap = 0, rc = ensure(ctx, &ctx->d_ce_dir, &cap, want, 0))) return rc;
        ctx->ce_cap = (int)want;
    }
    if (ctx->n_rst > ctx->rst_cap) {
        cudaFree(ctx->d_rst_blk);
        cudaFree(ctx->d_rst_mask);
        ctx->d_rst_blk = nullptr;
        ctx->d_rst_mask = nullptr;
        const size_t want = (size_t)ctx->n_rst + ctx->n_rst / 2 + 64;
        cap = 0;
        if ((rc = ensure(ctx, &ctx->d_rst_blk, &cap, want, 0)) || (cap = 0, rc = ensure(ctx, &ctx->d_rst_mask, &cap, want, 0))) return rc;
        ctx->rst_cap = (int)want;
    }
    {
        const size_t need = (size_t)ctx->n_rst * ctx->nc * (size_t)(ctx->blk_elems >> c.dim);
        if (need > ctx->rpool_cap) {
            cudaFree(ctx->d_rpool);
            ctx->d_rpool = nullptr;
            ctx->dev_bytes -= (int64_t)ctx->rpool_cap * 8;
            ctx->rpool_cap = 0;
            cap = 0;
            if ((rc = ensure(ctx, &ctx->d_rpool, &cap, need + need / 4, 0))) return rc;
            ctx->rpool_cap = cap;
        }
    }
    WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_rmap, 0xFF, sizeof(int) * (size_t)N, ctx->stream));
    a = make_args(ctx);
    a.has_jumps = ctx->has_jumps ? 1 : 0;
    a.jump_blk = ctx->d_jump_blk;
    a.jump_dir = ctx->d_jump_dir;
    a.wjump_blk = ctx->d_wjump_blk;
    a.wjump_dir = ctx->d_wjump_dir;
    a.woff = ctx->d_woff;
    a.ce_blk = ctx->d_ce_blk;
    a.ce_dir = ctx->d_ce_dir;
    a.rst_blk = ctx->d_rst_blk;
    a.rst_mask = ctx->d_rst_mask;
    a.rmap = ctx->d_rmap;
    a.active_int = ctx->d_active_int;
    a.active_bnd = ctx->d_active_bnd;
    topo_lists_kernel<1><<<(n + 255) / 256, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    ctx->halo_bnd.assign(ctx->n_bnd > 0 ? 1 : 0, 0);   // non-empty = halo mode (the members are not used on this path)
    ctx->halo_fine_neighbor = (flags[3] & 2) != 0;
    ctx->rmap_on_device = true;
    ctx->h_rmap.clear();
    return WGPU_OK;
}

}  // namespace

// host staging (pinned) for the id / level / treecode lists
static int32_t stage_upload(wgpu_ctx *ctx, const void *src, void *dst, size_t bytes)
{
    if (bytes == 0) return WGPU_OK;
    WGPU_CHECK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return WGPU_OK;
}

static int32_t set_active_list(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active)
{
    const int N = ctx->cfg.max_blocks;
    if (n_active > N) return topo_fail(ctx, WGPU_ERR_ARG, "more active blocks than max_blocks");
    ctx->h_active.resize(n_active);
    for (int k = 0; k < n_active; ++k) {
        const int b = hvy_active[k] - 1;
        if (b < 0 || b >= N || !ctx->h_has_coords[b]) return topo_fail(ctx, WGPU_ERR_ARG, "wgpu_set_grid / wgpu_set_active: active block is not a registered block");
        ctx->h_active[k] = b;
    }
    ctx->n_active = n_active;
    int32_t rc = stage_upload(ctx, ctx->h_active.data(), ctx->d_active, sizeof(int) * (size_t)n_active);
    if (rc) return rc;
    WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_flags + 3, 0, 2 * sizeof(int), ctx->stream));   // the duplicate flag [2] of wgpu_set_grid stays
    return WGPU_OK;
}

extern "C" {

int32_t wgpu_set_grid(wgpu_ctx *ctx, int32_t n_known, const int32_t *hvy_ids, const int32_t *level, const int64_t *treecode, int32_t n_active,
                      const int32_t *hvy_active)
{
    if (!ctx || n_known < 0 || n_active < 0 || (n_known > 0 && (!hvy_ids || !level || !treecode)) || (n_active > 0 && !hvy_active)) return WGPU_ERR_ARG;
    const wgpu_config &c = ctx->cfg;
    const int N = c.max_blocks;
    // host mirrors (level of every resident block: volume weights of wgpu_norm, checks of wgpu_refine)
    ctx->h_ixyz.assign((size_t)N * 3, 0);
    ctx->h_has_coords.assign(N, 0);
    ctx->h_tc_level.assign(N, 0);
    ctx->h_level.assign(N, 0);
    for (int k = 0; k < n_known; ++k) {
        const int b = hvy_ids[k] - 1, J = level[k];
        if (b >= N) return topo_fail(ctx, WGPU_ERR_ARG, "hvy id out of range");
        if (J < 0 || J > c.Jmax) return topo_fail(ctx, WGPU_ERR_ARG, "block level out of range");
        if (b < 0) continue;   // hvy id <= 0: the block is known by position only (its data live on another rank)
        int p[3] = {0, 0, 0};
        for (int d = 0; d < c.dim; ++d)
            for (int i = 0; i < J; ++i) p[d] |= (int)((treecode[k] >> ((i + c.Jmax - J) * c.dim + d)) & 1) << i;
        ctx->h_ixyz[3 * (size_t)b + 0] = p[1];
        ctx->h_ixyz[3 * (size_t)b + 1] = p[0];
        ctx->h_ixyz[3 * (size_t)b + 2] = p[2];
        ctx->h_has_coords[b] = 1;
        ctx->h_tc_level[b] = (signed char)J;
        ctx->h_level[b] = (signed char)J;
    }
    int32_t rc;
    size_t cap = 64;
    while (cap < (size_t)n_known * 2 + 2) cap <<= 1;
    if (cap > ctx->hcap) {
        cudaFree(ctx->d_hkeys);
        cudaFree(ctx->d_hvals);
        ctx->d_hkeys = nullptr;
        ctx->d_hvals = nullptr;
        ctx->hcap = 0;
        size_t c1 = 0, c2 = 0;
        if ((rc = ensure(ctx, &ctx->d_hkeys, &c1, cap, 0)) || (rc = ensure(ctx, &ctx->d_hvals, &c2, cap, 0))) return rc;
        ctx->hcap = cap;
    }
    ctx->hmask = (unsigned)(cap - 1);
    if (!ctx->d_ixyz) {
        size_t c1 = 0;
        if ((rc = ensure(ctx, &ctx->d_ixyz, &c1, (size_t)N * 3, 0))) return rc;
    }
    if (!ctx->d_bflag) {
        size_t c1 = 0;
        if ((rc = ensure(ctx, &ctx->d_bflag, &c1, (size_t)N, 0))) return rc;
    }
    // staging buffer on the device for (ids, level, treecode)
    const size_t off_tc = ((size_t)n_known * 8 + 15) & ~(size_t)15, need = off_tc + (size_t)n_known * 8 + 16;
    if ((rc = ensure(ctx, &ctx->d_topo_in, &ctx->topo_in_cap, need, need / 4))) return rc;
    int *d_ids = (int *)ctx->d_topo_in, *d_lvl = d_ids + n_known;
    long long *d_tc = (long long *)(ctx->d_topo_in + off_tc);
    std::vector<int> ids0(n_known);
    for (int k = 0; k < n_known; ++k) ids0[k] = hvy_ids[k] >= 1 ? hvy_ids[k] - 1 : -1;
    WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_hkeys, 0xFF, cap * sizeof(unsigned long long), ctx->stream));
    WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_bflag, 0, (size_t)N, ctx->stream));
    WGPU_CHECK(ctx, cudaMemsetAsync(ctx->d_flags + 2, 0, 3 * sizeof(int), ctx->stream));
    if ((rc = stage_upload(ctx, ids0.data(), d_ids, sizeof(int) * (size_t)n_known))) return rc;
    if ((rc = stage_upload(ctx, level, d_lvl, sizeof(int) * (size_t)n_known))) return rc;
    if ((rc = stage_upload(ctx, treecode, d_tc, sizeof(long long) * (size_t)n_known))) return rc;
    if (n_known) {
        TopoArgs a = make_args(ctx);
        topo_register_kernel<<<(n_known + 255) / 256, 256, 0, ctx->stream>>>(a, n_known, d_ids, d_lvl, d_tc, c.Jmax);
        ctx->launches++;
    }
    if (!ctx->h_halo.empty()) {   // the halo slots wgpu_set_halo declared: copies of other ranks' blocks, sources only
        const int nh = (int)ctx->h_halo.size();
        if ((rc = ensure(ctx, &ctx->d_halo_ids, &ctx->halo_ids_cap, (size_t)nh, (size_t)nh / 2 + 64))) return rc;
        if ((rc = stage_upload(ctx, ctx->h_halo.data(), ctx->d_halo_ids, sizeof(int) * (size_t)nh))) return rc;
        topo_mark_halo_kernel<<<(nh + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_bflag, ctx->d_halo_ids, nh);
        ctx->launches++;
        for (size_t k = 0; k < ctx->h_halo.size(); ++k) ctx->h_level[ctx->h_halo[k]] = ctx->halo_level_of[k];
    }
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));   // ids0 goes out of scope; pageable sources were consumed
    ctx->coords_dirty = false;
    ctx->lookup_ready = true;
    ctx->topo_on_device = true;
    return wgpu_set_active(ctx, n_active, hvy_active);
}

int32_t wgpu_set_active(wgpu_ctx *ctx, int32_t n_active, const int32_t *hvy_active)
{
    if (!ctx || n_active < 0 || (n_active > 0 && !hvy_active)) return WGPU_ERR_ARG;
    if (!ctx->topo_on_device || !ctx->lookup_ready) return topo_fail(ctx, WGPU_ERR_ARG, "wgpu_set_active: call wgpu_set_grid first");
    int32_t rc = set_active_list(ctx, n_active, hvy_active);
    if (rc) return rc;
    return derive_lists(ctx);
}

int32_t wgpu_topology_tables(wgpu_ctx *ctx, int32_t *nbr27, int32_t *wnbr27, int32_t *counts)
{
    if (!ctx || !counts) return WGPU_ERR_ARG;
    const int N = ctx->cfg.max_blocks;
    if (nbr27) WGPU_CHECK(ctx, cudaMemcpyAsync(nbr27, ctx->d_nbr, sizeof(int) * (size_t)N * WGPU_NDIR, cudaMemcpyDeviceToHost, ctx->stream));
    if (wnbr27 && ctx->d_wnbr) WGPU_CHECK(ctx, cudaMemcpyAsync(wnbr27, ctx->d_wnbr, sizeof(int) * (size_t)N * WGPU_NDIR, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    counts[0] = ctx->n_active;
    counts[1] = ctx->n_jump;
    counts[2] = ctx->n_wjump;
    counts[3] = ctx->n_ce;
    counts[4] = ctx->n_rst;
    counts[5] = ctx->n_int;
    counts[6] = ctx->n_bnd;
    counts[7] = ctx->has_jumps ? 1 : 0;
    return WGPU_OK;
}

int32_t wgpu_topology_list(wgpu_ctx *ctx, int32_t which, int32_t n, int32_t *out_a, int32_t *out_b)
{
    if (!ctx || n < 0) return WGPU_ERR_ARG;
    const int *a = nullptr, *b = nullptr;
    int have = 0;
    switch (which) {
    case 0: a = ctx->d_jump_blk; b = ctx->d_jump_dir; have = ctx->n_jump; break;
    case 1: a = ctx->d_wjump_blk; b = ctx->d_wjump_dir; have = ctx->n_wjump; break;
    case 2: a = ctx->d_ce_blk; b = ctx->d_ce_dir; have = ctx->n_ce; break;
    case 3: a = ctx->d_rst_blk; b = (const int *)ctx->d_rst_mask; have = ctx->n_rst; break;
    case 4: a = ctx->d_active_int; have = ctx->n_int; break;
    case 5: a = ctx->d_active_bnd; have = ctx->n_bnd; break;
    default: return WGPU_ERR_ARG;
    }
    if (n > have) n = have;
    if (n && out_a && a) WGPU_CHECK(ctx, cudaMemcpyAsync(out_a, a, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    if (n && out_b && b) WGPU_CHECK(ctx, cudaMemcpyAsync(out_b, b, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return WGPU_OK;
}

}  // extern "C"

// filtered copies across ranks when the block -> rpool map lives on the device (wgpu_set_halo_restrict after wgpu_set_grid)
int32_t wgpu_topology_halo_restrict(wgpu_ctx *ctx, const std::vector<int> &recv0, const std::vector<int> &send0)
{
    const int nr = (int)recv0.size(), ns = (int)send0.size();
    int32_t rc;
    if ((rc = ensure(ctx, &ctx->d_halo_ids, &ctx->halo_ids_cap, (size_t)nr + ns + 1, (size_t)(nr + ns) / 2 + 64))) return rc;
    if (nr) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_halo_ids, recv0.data(), sizeof(int) * (size_t)nr, cudaMemcpyHostToDevice, ctx->stream));
    if (ns) WGPU_CHECK(ctx, cudaMemcpyAsync(ctx->d_halo_ids + nr, send0.data(), sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, ctx->stream));
    const int n = std::max(nr, ns);
    if (n) {
        topo_rhalo_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_rmap, ctx->d_halo_ids, nr, ctx->n_rst, ctx->d_rhalo_send, ctx->d_halo_ids + nr, ns,
                                                                   ctx->d_flags + 2);
        ctx->launches++;
    }
    int flags[8];
    WGPU_CHECK(ctx, cudaMemcpyAsync(flags, ctx->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    WGPU_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    if (flags[4]) {
        cudaMemsetAsync(ctx->d_flags + 2, 0, 3 * sizeof(int), ctx->stream);
        ctx->err = "wgpu_set_halo_restrict: a block of the send list has no coarser neighbour (no filtered copy exists)";
        return WGPU_ERR_ARG;
    }
    return WGPU_OK;
}
