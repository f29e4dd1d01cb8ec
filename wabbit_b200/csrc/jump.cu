// Kernels that cross mesh levels (all built on fill_region, fill.cuh):
//
//   jump_fill_kernel     level-jump face patches for the stage kernel: the lvl_diff = +1 (restriction) and lvl_diff = -1
//                        (prediction) parts of sync_ghosts_RHS_tree (LIB/MPI/synchronize_ghosts_generic.f90:155-174;
//                        "full_leaf", ignore_Filter = .true.), written into the jump pool in the layout of the receiver's
//                        ghost strip (restrict_data / predict_data, LIB/MPI/restrict_predict_data.f90:45-202).
//   export_regions_kernel  resident layout -> ghosted host layout with a g_sync deep, fully synchronised ghost shell on grids
//                        with level jumps: all 26 relations, copy / decimation / prediction (what sync_ghosts_tree with
//                        ignore_Filter leaves; xfer_block_data.f90:10-601).
//   refine_kernel        refineBlock (LIB/MESH/refinementExecute.f90:1-120): prediction of the ghosted mother to 2^dim
//                        daughters (daughter digit bit0 -> y, bit1 -> x, bit2 -> z), interiors only -- daughter ghosts are
//                        never stored in HBM.
//   coarsen_kernel       sync_D2M (LIB/MESH/executeCoarsening_tree.f90:125-230): the scaling coefficients of a decomposed
//                        daughter (even spaghetti positions) become octant `digit` of its mother.
//
// One CTA per patch / region / (daughter, component).  All of them are HBM-bound gathers of a few hundred to a few thousand
// points; the arithmetic (prediction) is never contracted so that values are bit-identical to the reference's.
#include <algorithm>

#include "fill.cuh"
#include "wgpu_internal.cuh"

namespace {

FillCtx make_fill_ctx(wgpu_ctx *ctx, const double *u, bool filtered = false)
{
    FillCtx f;
    f.u = u;
    f.rpool = filtered ? ctx->d_rpool : nullptr;
    f.rmap = filtered ? ctx->d_rmap : nullptr;
    f.L.keys = ctx->d_hkeys;
    f.L.vals = ctx->d_hvals;
    f.L.mask = ctx->hmask;
    f.nc = ctx->nc;
    f.Bs = ctx->cfg.Bs[0];
    f.dim = ctx->cfg.dim;
    f.order = ctx->wavelet.X;
    for (int k = 0; k < 3; ++k) f.periodic[k] = ctx->cfg.periodic[k];
    return f;
}

size_t g_jump_fill_smem = 0;   // largest dynamic shared memory jump_fill_kernel has been configured for (two launchers share it)

template <typename K>
int32_t ensure_smem(wgpu_ctx *ctx, K kernel, size_t smem, size_t &configured)
{
    if (smem > configured) {
        if (smem > 227 * 1024) {
            ctx->err = "level-jump kernel: shared memory request exceeds 227 KB (block too large)";
            return WGPU_ERR_UNSUPPORTED;
        }
        WGPU_CHECK(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    return WGPU_OK;
}

// ------------------------------------------------------------------------------------------------ filtered restriction
// restrict_copy_at_CE (LIB/MPI/restrict_predict_data.f90:121-172) for every leaf that has a coarser neighbour: blockFilterXYZ_vct with
// the wavelet's HD filter and do_restriction (LIB/WAVELETS/module_wavelets.f90:307-401: x, then y, then z; only the scaling positions
// are filtered, each sum starts from 0 and adds u(i+shift)*HD(shift) for every shift in increasing order), then the scaling positions
// of the Nscl / Nscr strips that face a coarser or finer neighbour are copied back from the unfiltered block
// (coarseExtensionManipulateSC_block).  The filter reads the leaf's same-level neighbours only: everything it would take from other
// ghost nodes ends up in a copy strip (tests/test_oracle_sync.py::test_filtered_sync_equals_geometric_definition_on_graded_grids).
// One CTA per (leaf, component); result: (Bs/2)^3 values in rpool.  x pass from global memory (rows of the leaf and of its same-level
// neighbours), y and z passes in shared memory.
struct RestrictArgs {
    const double *u;
    double *rpool;
    const int *rst_blk;
    const unsigned *rst_mask;
    const int *nbr;
    int nc, Bs, F, lo, hi, Nscl, Nscr;
    int XC;                  // decimated x positions per CTA (blockIdx.z selects the chunk): bounds the shared memory for large blocks
    double HD[2 * WGPU_FMAX + 1];
};

__global__ void __launch_bounds__(256) restrict_filter_kernel(const RestrictArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ int nb[27];
    const int Bs = a.Bs, F = a.F, half = Bs / 2, n = Bs + 2 * F;
    const int b = a.rst_blk[blockIdx.x], c = blockIdx.y;
    const unsigned mask = a.rst_mask[blockIdx.x];
    const long long CS = (long long)Bs * Bs * Bs;
    if (threadIdx.x < 27) nb[threadIdx.x] = threadIdx.x == 13 ? b : a.nbr[(long long)b * 27 + threadIdx.x];
    __syncthreads();
    const int XC = a.XC, x0 = blockIdx.z * XC, xc = min(XC, half - x0);   // this CTA's decimated x positions [x0, x0 + xc)
    double *t1 = sm;                                // [n z][n y][xc x]
    double *t2 = t1 + (size_t)n * n * XC;           // [n z][half y][xc x]
    for (int i = threadIdx.x; i < n * n * xc; i += blockDim.x) {
        const int xl = i % xc, xo = x0 + xl, y = (i / xc) % n - F, z = i / (xc * n) - F;
        const int sy = y < 0 ? -1 : (y >= Bs ? 1 : 0), sz = z < 0 ? -1 : (z >= Bs ? 1 : 0);
        // the row (y, z) of the ghosted block lives in up to three blocks (x neighbours): their row pointers, shifted so that every one of them
        // is indexed by the block-local x of the ghosted row (the kernel is bound by the index arithmetic of this loop, not by its loads)
        const long long row = ((long long)(z - sz * Bs) * Bs + (y - sy * Bs)) * Bs;
        const int *nbrow = nb + (sz + 1) * 9 + (sy + 1) * 3;
        const int sm1 = nbrow[0], s00 = nbrow[1], sp1 = nbrow[2];
        const double *pm = sm1 >= 0 ? a.u + ((long long)sm1 * a.nc + c) * CS + row + Bs : nullptr;     // x in [-Bs, 0)
        const double *p0 = s00 >= 0 ? a.u + ((long long)s00 * a.nc + c) * CS + row : nullptr;          // x in [0, Bs)
        const double *pp = sp1 >= 0 ? a.u + ((long long)sp1 * a.nc + c) * CS + row - Bs : nullptr;     // x in [Bs, 2 Bs)
        double acc = 0.0;
        for (int k = a.lo; k <= a.hi; ++k) {
            const int x = 2 * xo + k;
            const double *ptr = x < 0 ? pm : (x >= Bs ? pp : p0);
            const double v = ptr ? ptr[x] : 0.0;
            acc = __dadd_rn(acc, __dmul_rn(v, a.HD[k + WGPU_FMAX]));
        }
        t1[((size_t)(z + F) * n + (y + F)) * xc + xl] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n * half * xc; i += blockDim.x) {
        const int xl = i % xc, yo = (i / xc) % half, z = i / (xc * half);
        double acc = 0.0;
        for (int k = a.lo; k <= a.hi; ++k) acc = __dadd_rn(acc, __dmul_rn(t1[((size_t)z * n + (2 * yo + F + k)) * xc + xl], a.HD[k + WGPU_FMAX]));
        t2[((size_t)z * half + yo) * xc + xl] = acc;
    }
    __syncthreads();
    double *out = a.rpool + ((long long)blockIdx.x * a.nc + c) * (CS / 8);
    const double *own = a.u + ((long long)b * a.nc + c) * CS;
    for (int i = threadIdx.x; i < half * half * xc; i += blockDim.x) {
        const int xl = i % xc, xo = x0 + xl, yo = (i / xc) % half, zo = i / (xc * half);
        const int p[3] = {2 * xo, 2 * yo, 2 * zo};
        bool copy = false;
        for (int d = 0; d < 27 && !copy; ++d) {
            if (!((mask >> d) & 1u)) continue;
            const int dd[3] = {d % 3 - 1, (d / 3) % 3 - 1, d / 9 - 1};
            bool in = true;
            for (int k = 0; k < 3; ++k) in = in && (dd[k] == 0 || (dd[k] < 0 ? p[k] < a.Nscl : p[k] >= Bs - a.Nscr));
            copy = in;
        }
        double v;
        if (copy) v = own[((long long)p[2] * Bs + p[1]) * Bs + p[0]];
        else {
            v = 0.0;
            for (int k = a.lo; k <= a.hi; ++k) v = __dadd_rn(v, __dmul_rn(t2[((size_t)(p[2] + F + k) * half + yo) * xc + xl], a.HD[k + WGPU_FMAX]));
        }
        out[((size_t)zo * half + yo) * half + xo] = v;
    }
}

// the same for two-dimensional blocks: x pass from global memory, y pass in shared memory; result (Bs/2)^2 values per (leaf, component)
__global__ void __launch_bounds__(256) restrict_filter2d_kernel(const RestrictArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ int nb[9];
    const int Bs = a.Bs, F = a.F, half = Bs / 2, n = Bs + 2 * F;
    const int b = a.rst_blk[blockIdx.x], c = blockIdx.y;
    const unsigned mask = a.rst_mask[blockIdx.x];
    const long long CS = (long long)Bs * Bs;
    if (threadIdx.x < 9) nb[threadIdx.x] = threadIdx.x == 4 ? b : a.nbr[(long long)b * 27 + 9 + threadIdx.x];
    __syncthreads();
    double *t1 = sm;                                // [n y][half x]
    for (int i = threadIdx.x; i < n * half; i += blockDim.x) {
        const int xo = i % half, y = i / half - F;
        const int sy = y < 0 ? -1 : (y >= Bs ? 1 : 0);
        double acc = 0.0;
        for (int k = a.lo; k <= a.hi; ++k) {
            const int x = 2 * xo + k;
            const int sx = x < 0 ? -1 : (x >= Bs ? 1 : 0);
            const int src = nb[(sy + 1) * 3 + (sx + 1)];
            const double v = src >= 0 ? a.u[((long long)src * a.nc + c) * CS + (long long)(y - sy * Bs) * Bs + (x - sx * Bs)] : 0.0;
            acc = __dadd_rn(acc, __dmul_rn(v, a.HD[k + WGPU_FMAX]));
        }
        t1[i] = acc;
    }
    __syncthreads();
    double *out = a.rpool + ((long long)blockIdx.x * a.nc + c) * (CS / 4);
    const double *own = a.u + ((long long)b * a.nc + c) * CS;
    for (int i = threadIdx.x; i < half * half; i += blockDim.x) {
        const int xo = i % half, yo = i / half;
        const int p[2] = {2 * xo, 2 * yo};
        bool copy = false;
        for (int d = 9; d < 18 && !copy; ++d) {
            if (!((mask >> d) & 1u)) continue;
            const int dd[2] = {d % 3 - 1, (d / 3) % 3 - 1};
            bool in = true;
            for (int k = 0; k < 2; ++k) in = in && (dd[k] == 0 || (dd[k] < 0 ? p[k] < a.Nscl : p[k] >= Bs - a.Nscr));
            copy = in;
        }
        double v;
        if (copy) v = own[(long long)p[1] * Bs + p[0]];
        else {
            v = 0.0;
            for (int k = a.lo; k <= a.hi; ++k) v = __dadd_rn(v, __dmul_rn(t1[(size_t)(p[1] + F + k) * half + xo], a.HD[k + WGPU_FMAX]));
        }
        out[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------ jump patches
struct JumpArgs {
    FillCtx f;
    double *jpool;
    long long jpatch;          // size of a patch when all have the same (joff == nullptr)
    const long long *joff;     // offset of every patch in the pool, or nullptr
    const int *jblk, *jdir;
    const signed char *level;
    const int *ixyz;
    int H;                     // depth of the ghost strip
    const double *ce_coarse;   // != nullptr: scaling positions take the coincident value of the coarser leaf in this array, the rest is 0
};

// Ghost patch of a decomposed block that faces a coarser leaf, as sync_SCWC_from_MC + coarse_extension_modify leave it (adapt_tree.f90:
// 686-987, reconstruction_step.f90:3-100): the predictor returns the coincident coarse value at scaling positions (even coordinates on
// the block's lattice) and every other position is a wavelet coefficient, which the coarse extension sets to zero.
__device__ inline void fill_scwc(const FillCtx &a, const double *uc, SrcTable &T, int lvl, const int lo[3], const int ext[3], double *out,
                                 long long sc, long long sy, long long sz, int ncomp, int tid, int nt)
{
    const int Bs = a.Bs, dim = a.dim;
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    const int npts = ext[0] * ext[1] * ext[2];
    int clo[3], chi[3];
    for (int k = 0; k < 3; ++k) {
        clo[k] = k < dim ? lo[k] >> 1 : 0;
        chi[k] = k < dim ? (lo[k] + ext[k] - 1) >> 1 : 0;
    }
    __syncthreads();
    if (lvl > 0) src_table_build(T, a.L, lvl - 1, clo, chi, Bs, dim, a.periodic, tid, nt);
    __syncthreads();
    for (int i = tid; i < ncomp * npts; i += nt) {
        const int c = i / npts, r = i % npts;
        const int x = r % ext[0], y = (r / ext[0]) % ext[1], z = r / (ext[0] * ext[1]);
        const int P[3] = {lo[0] + x, lo[1] + y, lo[2] + z};
        double v = 0.0;
        if (lvl > 0 && !(P[0] & 1) && !(P[1] & 1) && (dim == 2 || !(P[2] & 1))) {
            const int Pc[3] = {P[0] >> 1, P[1] >> 1, dim == 3 ? P[2] >> 1 : 0};
            int sb, so, ro;
            src_resolve(T, Pc, Bs, dim, sb, so, ro);
            if (sb >= 0 && ro < 0) v = uc[((long long)sb * a.nc + c) * CS + so];   // owned by a block of level lvl-1 itself
        }
        out[c * sc + z * sz + y * sy + x] = v;
    }
}

__global__ void __launch_bounds__(128) jump_fill_kernel(const JumpArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SrcTable T;
    const int Bs = a.f.Bs, H = a.H, dim = a.f.dim;
    const int b = a.jblk[blockIdx.x], dcode = a.jdir[blockIdx.x];
    const int d[3] = {dcode % 3 - 1, (dcode / 3) % 3 - 1, dcode / 9 - 1};
    int lo[3], ext[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = a.ixyz[3 * b + k] * Bs + (d[k] < 0 ? -H : (d[k] > 0 ? Bs : 0));
        ext[k] = d[k] ? H : (k < dim ? Bs : 1);
    }
    const long long npts = (long long)ext[0] * ext[1] * ext[2];
    double *out = a.jpool + (a.joff ? a.joff[blockIdx.x] : (long long)blockIdx.x * a.jpatch);
    if (a.ce_coarse) fill_scwc(a.f, a.ce_coarse, T, a.level[b], lo, ext, out, npts, ext[0], (long long)ext[0] * ext[1], a.f.nc, threadIdx.x, blockDim.x);
    else fill_region(a.f, T, sm, a.level[b], lo, ext, out, npts, ext[0], (long long)ext[0] * ext[1], 0, a.f.nc, threadIdx.x, blockDim.x, true);
}

// ------------------------------------------------------------------------------------------------ export with ghosts
struct ExportArgs {
    FillCtx f;
    double *staged;         // [n][ncomp_host][nz][ny][nx]
    const int *ids;
    const signed char *level;
    const int *ixyz;
    int ncomp, ncomp_host, g, gs, by_id;
};

__global__ void __launch_bounds__(128) export_regions_kernel(const ExportArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SrcTable T;
    const int Bs = a.f.Bs, dim = a.f.dim, g = a.g, gs = a.gs;
    const int r = blockIdx.x;
    const int d[3] = {r % 3 - 1, (r / 3) % 3 - 1, r / 9 - 1};
    if (dim == 2 && d[2] != 0) return;
    if (gs == 0 && r != 13) return;
    const int b = a.ids[blockIdx.y];
    const int i = a.by_id ? b : blockIdx.y;
    const int nx = Bs + 2 * g, ny = Bs + 2 * g, nz = dim == 3 ? Bs + 2 * g : 1, gz = dim == 3 ? g : 0;
    int lo[3], ext[3], org[3];
    for (int k = 0; k < 3; ++k) {
        org[k] = d[k] < 0 ? -gs : (d[k] > 0 ? Bs : 0);
        lo[k] = a.ixyz[3 * b + k] * Bs + org[k];
        ext[k] = d[k] ? gs : (k < dim ? Bs : 1);
    }
    const long long sc = (long long)nx * ny * nz;
    double *out = a.staged + (long long)i * a.ncomp_host * sc + ((long long)(org[2] + gz) * ny + (org[1] + g)) * nx + (org[0] + g);
    fill_region(a.f, T, sm, a.level[b], lo, ext, out, sc, nx, (long long)nx * ny, 0, a.ncomp, threadIdx.x, blockDim.x);
}

// ------------------------------------------------------------------------------------------------ refineBlock
struct RefineArgs {
    FillCtx f;              // f.u = source array (mothers and their neighbours)
    double *dst;            // compact array that receives the daughters
    const int *mother;      // [n] 0-based
    const int *daughter;    // [n][2^dim] 0-based, digit order
    const signed char *level;
    const int *ixyz;
    int nslab, fz;          // the daughter's z range is cut into nslab slabs of fz planes (one CTA each): bounds the shared memory for large blocks
    long long nb_max;       // doubles reserved per component of the box in shared memory
};

// grid (2^dim * nslab, n); one z slab of one daughter, ALL components per CTA: the source tables of the up to 7 parts of the box that lie
// outside the mother (one fill_region each: hash lookups, barriers) are resolved once for all components, and the part inside the mother
// is a plain copy of its interior -- the table work per daughter drops from 8 * nc to 7 (14.8 -> measured in profiles/ for 25 000 daughters)
__global__ void __launch_bounds__(256) refine_kernel(const RefineArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SrcTable T;
    const int Bs = a.f.Bs, dim = a.f.dim, order = a.f.order, A = order / 2 - 1, half = Bs / 2, nc = a.f.nc;
    const int digit = blockIdx.x % (1 << dim), slab = blockIdx.x >> dim, m = a.mother[blockIdx.y];
    const int q[3] = {(digit >> 1) & 1, digit & 1, (digit >> 2) & 1};   // refinementExecute.f90: bit0 -> y, bit1 -> x, bit2 -> z
    const int lvl = a.level[m];
    // box of the mother's (ghosted) lattice this daughter is interpolated from: [q*Bs/2 - A, q*Bs/2 + Bs/2 + A]; along z only the part
    // this slab of the daughter (fine planes [slab*fz, slab*fz + fz)) needs
    const int last = dim == 3 ? 2 : -1, z0 = slab * a.fz, fzn = min(a.fz, Bs - z0);
    int clo[3], n[3], flo[3], fext[3];
    for (int k = 0; k < 3; ++k) {
        if (k < dim) {
            clo[k] = a.ixyz[3 * m + k] * Bs + q[k] * half - A;
            n[k] = half + 2 * A + 1;
            flo[k] = 2 * (a.ixyz[3 * m + k] * Bs + q[k] * half);
            fext[k] = Bs;
            if (k == last && a.nslab > 1) {
                flo[k] += z0;
                fext[k] = fzn;
                clo[k] = (flo[k] >> 1) - A;
                n[k] = ((flo[k] + fzn - 1 + 1) >> 1) + A - clo[k] + 1;
            }
        } else {
            clo[k] = 0;
            n[k] = 1;
            flo[k] = 0;
            fext[k] = 1;
        }
    }
    const long long NB = (long long)n[0] * n[1] * n[2];      // one component of the box
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    double *cb = sm;                                         // [nc][n2][n1][n0]
    double *scratch = cb + (size_t)nc * a.nb_max;            // fill_region's prediction scratch; later the intermediates of predict_from_box
    // the box straddles the mother and up to 2^dim - 1 neighbouring cells: fill it cell by cell
    const int m0[3] = {a.ixyz[3 * m] * Bs, a.ixyz[3 * m + 1] * Bs, a.ixyz[3 * m + 2] * Bs};
    for (int part = 0; part < 8; ++part) {
        int lo[3], ext[3];
        bool empty = false;
        for (int k = 0; k < 3; ++k) {
            const int side = (part >> k) & 1;   // 0: the part inside the mother, 1: the part outside
            if (k >= dim) {
                lo[k] = 0;
                ext[k] = 1;
                if (side) empty = true;
                continue;
            }
            const int b0 = clo[k], b1 = clo[k] + n[k] - 1;                 // box
            const int i0 = m0[k], i1 = m0[k] + Bs - 1;                      // mother interior
            if (!side) {
                lo[k] = b0 > i0 ? b0 : i0;
                ext[k] = (b1 < i1 ? b1 : i1) - lo[k] + 1;
            } else if (b0 < i0) {                                           // below the mother (a box reaches out on one side only)
                lo[k] = b0;
                ext[k] = i0 - b0;
            } else {                                                        // above
                lo[k] = i1 + 1;
                ext[k] = b1 - i1;
            }
            if (ext[k] <= 0) empty = true;
        }
        if (empty) continue;
        double *out = cb + ((size_t)(lo[2] - clo[2]) * n[1] + (lo[1] - clo[1])) * n[0] + (lo[0] - clo[0]);
        if (part == 0) {
            // inside the mother: her own interior values, no lookup
            const int npts = ext[0] * ext[1] * ext[2];
            const double *um = a.f.u + (long long)m * nc * CS;
            for (int i = threadIdx.x; i < nc * npts; i += blockDim.x) {
                const int c = i / npts, r = i % npts;
                const int x = r % ext[0], y = (r / ext[0]) % ext[1], z = r / (ext[0] * ext[1]);
                out[c * NB + ((long long)z * n[1] + y) * n[0] + x] =
                    um[c * CS + ((long long)(lo[2] - m0[2] + z) * Bs + (lo[1] - m0[1] + y)) * Bs + (lo[0] - m0[0] + x)];
            }
        } else
            fill_region(a.f, T, scratch, lvl, lo, ext, out, NB, n[0], (long long)n[0] * n[1], 0, nc, threadIdx.x, blockDim.x);
    }
    __syncthreads();
    // predict_from_box keeps its two intermediates right behind the box it reads: from the last component down they only overwrite boxes
    // that have been consumed (and the scratch area)
    for (int c = nc - 1; c >= 0; --c) {
        double *out = a.dst + ((long long)a.daughter[blockIdx.y * (1 << dim) + digit] * nc + c) * CS + (a.nslab > 1 ? (long long)z0 * Bs * Bs : 0);
        predict_from_box(cb + c * NB, clo, n, flo, fext, order, dim, out, Bs, (long long)Bs * Bs, threadIdx.x, blockDim.x);
    }
}

// ------------------------------------------------------------------------------------------------ sync_D2M / block copies
// mother[octant digit] = daughter values at even (scaling-coefficient) positions of `src`
__global__ void __launch_bounds__(256) coarsen_kernel(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ mother,
                                                      const int *__restrict__ daughter, int nc, int Bs, int dim)
{
    const int nd = 1 << dim, half = Bs / 2;
    const int digit = blockIdx.x, m = mother[blockIdx.y], c = blockIdx.z;
    const int d = daughter[blockIdx.y * nd + digit];
    const int q[3] = {(digit >> 1) & 1, digit & 1, (digit >> 2) & 1};
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    const int hz = dim == 3 ? half : 1;
    const double *s = src + ((long long)d * nc + c) * CS;
    double *o = dst + ((long long)m * nc + c) * CS;
    for (int i = threadIdx.x; i < half * half * hz; i += blockDim.x) {
        const int x = i % half, y = (i / half) % half, z = i / (half * half);
        o[((long long)(z + q[2] * hz * (dim == 3)) * Bs + (y + q[1] * half)) * Bs + (x + q[0] * half)] = s[((long long)(2 * z) * Bs + 2 * y) * Bs + 2 * x];
    }
}

__global__ void __launch_bounds__(256) copy_blocks_kernel(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ src_ids,
                                                          const int *__restrict__ dst_ids, long long per_block)
{
    const double2 *s = reinterpret_cast<const double2 *>(src + (long long)src_ids[blockIdx.y] * per_block);
    double2 *o = reinterpret_cast<double2 *>(dst + (long long)dst_ids[blockIdx.y] * per_block);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_block / 2; i += (long long)gridDim.x * blockDim.x) o[i] = s[i];
}

// coarse extension on the interiors of decomposed blocks (coarseExtensionManipulateWC_block / ...SC_block,
// LIB/WAVELETS/module_wavelets.f90:877-1027): one CTA per (block, direction with a coarser neighbour).  In the strip facing the
// neighbour every coefficient that is not a pure scaling coefficient is zeroed (Nwc deep), and the pure scaling positions are
// copied from the original values (Nsc deep).  The ghost-patch half of the reference routine has no counterpart here: ghost
// values are never stored, consumers produce them on the fly.
__global__ void __launch_bounds__(128) ce_kernel(double *__restrict__ wd, const double *__restrict__ orig, const int *__restrict__ blk,
                                                 const int *__restrict__ dir, int nc, int Bs, int dim, int Nwcl, int Nwcr, int Nscl, int Nscr,
                                                 int clear_wc, int copy_sc)
{
    const int b = blk[blockIdx.x], dc = dir[blockIdx.x];
    const int d[3] = {dc % 3 - 1, (dc / 3) % 3 - 1, dc / 9 - 1};
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 0 ? !clear_wc : !copy_sc) continue;
        const int Nl = pass == 0 ? Nwcl : Nscl, Nr = pass == 0 ? Nwcr : Nscr;
        int lo[3], ext[3];
        for (int k = 0; k < 3; ++k) {
            const int B = k < dim ? Bs : 1;
            lo[k] = d[k] > 0 ? B - Nr : 0;
            ext[k] = d[k] < 0 ? Nl : (d[k] > 0 ? Nr : B);
            if (lo[k] < 0) { ext[k] += lo[k]; lo[k] = 0; }
            if (ext[k] > B) ext[k] = B;
        }
        const int npts = ext[0] * ext[1] * ext[2];
        if (npts <= 0) continue;
        for (int i = threadIdx.x; i < nc * npts; i += blockDim.x) {
            const int c = i / npts, r = i % npts;
            const int x = lo[0] + r % ext[0], y = lo[1] + (r / ext[0]) % ext[1], z = lo[2] + r / (ext[0] * ext[1]);
            const bool pure_sc = !(x & 1) && !(y & 1) && (dim == 2 || !(z & 1));
            const long long o = ((long long)b * nc + c) * CS + ((long long)z * Bs + y) * Bs + x;
            if (pass == 0) {
                if (!pure_sc) wd[o] = 0.0;
            } else if (pure_sc) wd[o] = orig[o];
        }
        __syncthreads();
    }
}

}  // namespace

int32_t wgpu_launch_ce(wgpu_ctx *ctx, double *wd, const double *orig, int Nwcl, int Nwcr, int Nscl, int Nscr, int clear_wc, int copy_sc)
{
    if (ctx->n_ce == 0) return WGPU_OK;
    ce_kernel<<<ctx->n_ce, 128, 0, ctx->stream>>>(wd, orig, ctx->d_ce_blk, ctx->d_ce_dir, ctx->nc, ctx->cfg.Bs[0], ctx->cfg.dim, Nwcl, Nwcr, Nscl,
                                                   Nscr, clear_wc, copy_sc);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// *active = true if the filtered copies of `src` are in ctx->d_rpool afterwards (lifted wavelet, filter not ignored, level jumps present)
int32_t wgpu_launch_restrict_filter(wgpu_ctx *ctx, const double *src, int nc_src, bool *active)
{
    *active = false;
    const WaveFilters &w = ctx->wavelet;
    if (ctx->ignore_filter || !ctx->wavelet_set || w.Y == 0 || (ctx->n_rst == 0 && ctx->n_rhalo_recv == 0) || nc_src != ctx->nc) return WGPU_OK;
    if (ctx->n_rst == 0) {   // no block of this rank sends restricted data, but filtered copies of finer neighbours arrived from other ranks
        *active = true;
        return WGPU_OK;
    }
    const wgpu_config &c = ctx->cfg;
    RestrictArgs a;
    a.u = src;
    a.rpool = ctx->d_rpool;
    a.rst_blk = ctx->d_rst_blk;
    a.rst_mask = ctx->d_rst_mask;
    a.nbr = ctx->d_nbr;
    a.nc = ctx->nc;
    a.Bs = c.Bs[0];
    a.lo = w.hd_lo;
    a.hi = w.hd_hi;
    a.F = std::max(-w.hd_lo, w.hd_hi);
    a.Nscl = std::max(-w.hd_lo - 1, 0);   // setup_wavelet, module_wavelets.f90:1368-1376
    a.Nscr = w.hd_hi;
    a.XC = 0;
    for (int k = 0; k < 2 * WGPU_FMAX + 1; ++k) a.HD[k] = w.HD[k];
    const int n = a.Bs + 2 * a.F, half = a.Bs / 2;
    dim3 grid(ctx->n_rst, ctx->nc);
    if (c.dim == 2) {
        const size_t smem2 = sizeof(double) * ((size_t)n * half);
        static size_t configured2 = 0;
        int32_t rc2 = ensure_smem(ctx, restrict_filter2d_kernel, smem2, configured2);
        if (rc2) return rc2;
        restrict_filter2d_kernel<<<grid, 256, smem2, ctx->stream>>>(a);
        ctx->launches++;
        WGPU_CHECK(ctx, cudaGetLastError());
        *active = true;
        return WGPU_OK;
    }
    // all decimated x positions in one CTA if the two intermediates fit into ~100 KB (two CTAs per SM), else chunks of x (independent)
    a.XC = half;
    while (a.XC > 1 && sizeof(double) * ((size_t)n * n * a.XC + (size_t)n * half * a.XC) > 100 * 1024) a.XC = (a.XC + 1) / 2;
    grid.z = (half + a.XC - 1) / a.XC;
    const size_t smem = sizeof(double) * ((size_t)n * n * a.XC + (size_t)n * half * a.XC);
    static size_t configured = 0;
    int32_t rc = ensure_smem(ctx, restrict_filter_kernel, smem, configured);
    if (rc) return rc;
    restrict_filter_kernel<<<grid, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    *active = true;
    return WGPU_OK;
}

int32_t wgpu_launch_jump_fill(wgpu_ctx *ctx, const double *src)
{
    if (ctx->n_jump == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    JumpArgs a;
    a.f = make_fill_ctx(ctx, src);
    a.jpool = ctx->d_jpool;
    a.joff = nullptr;
    a.jblk = ctx->d_jump_blk;
    a.jdir = ctx->d_jump_dir;
    a.level = ctx->d_level;
    a.ixyz = ctx->d_ixyz;
    a.H = c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3);
    a.ce_coarse = nullptr;
    a.jpatch = (long long)ctx->nc * a.H * c.Bs[0] * (c.dim == 3 ? c.Bs[0] : 1);
    size_t best = 0;
    for (int f = 0; f < c.dim; ++f) {
        const int e[3] = {f == 0 ? a.H : c.Bs[0], f == 1 ? a.H : c.Bs[0], c.dim == 3 ? (f == 2 ? a.H : c.Bs[0]) : 1};
        const size_t s = fill_scratch_doubles(e, a.f.order, c.dim);
        best = s > best ? s : best;
    }
    int32_t rc = ensure_smem(ctx, jump_fill_kernel, best * sizeof(double), g_jump_fill_smem);
    if (rc) return rc;
    jump_fill_kernel<<<ctx->n_jump, 128, best * sizeof(double), ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_export_regions(wgpu_ctx *ctx, const double *src, double *staged, const int *d_ids, int n, int ncomp_src, int ncomp_host,
                                   int g_sync, int by_id)
{
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    bool filtered = false;
    if (g_sync > 0) {   // what sync_ghosts_tree leaves in the ghost nodes: filtered restriction for lifted wavelets
        int32_t rcf = wgpu_launch_restrict_filter(ctx, src, ncomp_src, &filtered);
        if (rcf) return rcf;
    }
    ExportArgs a;
    a.f = make_fill_ctx(ctx, src, filtered);
    a.f.nc = ncomp_src;
    a.staged = staged;
    a.ids = d_ids;
    a.level = ctx->d_level;
    a.ixyz = ctx->d_ixyz;
    a.ncomp = ncomp_src < ncomp_host ? ncomp_src : ncomp_host;
    a.ncomp_host = ncomp_host;
    a.g = c.g;
    a.gs = g_sync;
    a.by_id = by_id;
    size_t best = 0;
    for (int r = 0; r < 27; ++r) {
        const int d[3] = {r % 3 - 1, (r / 3) % 3 - 1, r / 9 - 1};
        if (r == 13 || (c.dim == 2 && d[2])) continue;
        const int e[3] = {d[0] ? g_sync : c.Bs[0], d[1] ? g_sync : c.Bs[0], c.dim == 3 ? (d[2] ? g_sync : c.Bs[0]) : 1};
        const size_t s = fill_scratch_doubles(e, a.f.order, c.dim);
        best = s > best ? s : best;
    }
    static size_t configured = 0;
    int32_t rc = ensure_smem(ctx, export_regions_kernel, best * sizeof(double), configured);
    if (rc) return rc;
    dim3 grid(27, n);
    export_regions_kernel<<<grid, 128, best * sizeof(double), ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_refine(wgpu_ctx *ctx, const double *src, double *dst, const int *d_mother, const int *d_daughter, int n)
{
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    bool filtered = false;   // the sync_ghosts_tree before refine_tree (LIB/MAIN/main.f90:314) applies the restriction filter
    int32_t rcf = wgpu_launch_restrict_filter(ctx, src, ctx->nc, &filtered);
    if (rcf) return rcf;
    RefineArgs a;
    a.f = make_fill_ctx(ctx, src, filtered);
    a.dst = dst;
    a.mother = d_mother;
    a.daughter = d_daughter;
    a.level = ctx->d_level;
    a.ixyz = ctx->d_ixyz;
    const int A = a.f.order / 2 - 1, Bs = c.Bs[0], half = Bs / 2;
    const int nn = half + 2 * A + 1;
    // z slabs of the daughter until the coarse boxes of all components and the two intermediates of the prediction fit into ~110 KB
    a.nslab = 1;
    a.fz = Bs;
    size_t smem = 0;
    for (;;) {
        const int nz = c.dim == 3 ? (a.nslab == 1 ? nn : a.fz / 2 + 2 * A + 2) : 1;
        const int n3[3] = {nn, nn, nz}, fe[3] = {Bs, Bs, c.dim == 3 ? a.fz : 1};
        a.nb_max = (long long)n3[0] * n3[1] * n3[2];
        const size_t inter = (size_t)fe[0] * n3[1] * n3[2] + (size_t)fe[0] * fe[1] * n3[2];
        const size_t sub = fill_scratch_doubles(n3, a.f.order, c.dim);
        smem = ((size_t)ctx->nc * a.nb_max + std::max(inter, sub)) * sizeof(double);
        if (smem <= 110 * 1024 || c.dim == 2 || a.fz <= 4) break;
        a.nslab *= 2;
        a.fz = (Bs / a.nslab + 1) & ~1;          // even slabs: fine plane 2i coincides with coarse plane i
    }
    a.nslab = c.dim == 3 ? (Bs + a.fz - 1) / a.fz : 1;
    static size_t configured = 0;
    int32_t rc = ensure_smem(ctx, refine_kernel, smem, configured);
    if (rc) return rc;
    dim3 grid((1 << c.dim) * a.nslab, n);
    refine_kernel<<<grid, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_coarsen(wgpu_ctx *ctx, const double *src, double *dst, const int *d_mother, const int *d_daughter, int n)
{
    if (n == 0) return WGPU_OK;
    dim3 grid(1 << ctx->cfg.dim, n, ctx->nc);
    coarsen_kernel<<<grid, 256, 0, ctx->stream>>>(src, dst, d_mother, d_daughter, ctx->nc, ctx->cfg.Bs[0], ctx->cfg.dim);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_copy_blocks(wgpu_ctx *ctx, const double *src, double *dst, const int *d_src_ids, const int *d_dst_ids, int n)
{
    if (n == 0) return WGPU_OK;
    const long long per_block = (long long)ctx->nc * ctx->blk_elems;
    for (int s0 = 0; s0 < n; s0 += 32768) {   // grid.y limit
        dim3 grid((unsigned)((per_block / 2 + 255) / 256), std::min(32768, n - s0));
        copy_blocks_kernel<<<grid, 256, 0, ctx->stream>>>(src, dst, d_src_ids + s0, d_dst_ids + s0, per_block);
        ctx->launches++;
        WGPU_CHECK(ctx, cudaGetLastError());
    }
    return WGPU_OK;
}

// dst[k] = src[idx[k]] for entries of per_entry doubles: packs filtered copies for other ranks
__global__ void __launch_bounds__(256) copy_entries_kernel(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ idx,
                                                           long long per_entry)
{
    const double *s = src + (long long)idx[blockIdx.y] * per_entry;
    double *o = dst + (long long)blockIdx.y * per_entry;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_entry; i += (long long)gridDim.x * blockDim.x) o[i] = s[i];
}

int32_t wgpu_launch_copy_entries(wgpu_ctx *ctx, const double *src, double *dst, const int *d_src_idx, int n, long long per_entry)
{
    for (int s0 = 0; s0 < n; s0 += 32768) {
        dim3 grid((unsigned)((per_entry + 255) / 256), std::min(32768, n - s0));
        copy_entries_kernel<<<grid, 256, 0, ctx->stream>>>(src, dst + (long long)s0 * per_entry, d_src_idx + s0, per_entry);
        ctx->launches++;
        WGPU_CHECK(ctx, cudaGetLastError());
    }
    return WGPU_OK;
}

int32_t wgpu_launch_wjump_fill(wgpu_ctx *ctx, const double *src, const double *ce_coarse)
{
    if (ctx->n_wjump == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    bool filtered = false;   // sync_TMP_from_all / sync_ghosts_tree before the decomposition: restriction with the HD filter
    if (!ce_coarse) {
        int32_t rcf = wgpu_launch_restrict_filter(ctx, src, ctx->nc, &filtered);
        if (rcf) return rcf;
    }
    JumpArgs a;
    a.f = make_fill_ctx(ctx, src, filtered);
    a.jpool = ctx->d_wpool;
    a.ce_coarse = ce_coarse;
    a.jpatch = 0;
    a.joff = ctx->d_woff;
    a.jblk = ctx->d_wjump_blk;
    a.jdir = ctx->d_wjump_dir;
    a.level = ctx->d_level;
    a.ixyz = ctx->d_ixyz;
    a.H = ctx->wjump_depth;
    size_t best = 0;
    for (int r = 0; r < 27; ++r) {
        const int d[3] = {r % 3 - 1, (r / 3) % 3 - 1, r / 9 - 1};
        if (r == 13 || (c.dim == 2 && d[2])) continue;
        const int e[3] = {d[0] ? a.H : c.Bs[0], d[1] ? a.H : c.Bs[0], c.dim == 3 ? (d[2] ? a.H : c.Bs[0]) : 1};
        const size_t s = fill_scratch_doubles(e, a.f.order, c.dim);
        best = s > best ? s : best;
    }
    int32_t rc = ensure_smem(ctx, jump_fill_kernel, best * sizeof(double), g_jump_fill_smem);
    if (rc) return rc;
    jump_fill_kernel<<<ctx->n_wjump, 128, best * sizeof(double), ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}
