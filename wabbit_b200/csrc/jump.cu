// Level-jump face patches for the stage kernel: the lvl_diff = +1 (restriction) and lvl_diff = -1 (prediction) parts of
// sync_ghosts_RHS_tree (LIB/MPI/synchronize_ghosts_generic.f90:155-174; "full_leaf", ignore_Filter = .true.,
// g_minus = g_plus = stencil half width), written straight into the patch pool in the layout of the receiver's ghost strip.
//
// Reference: restrict_data / predict_data (LIB/MPI/restrict_predict_data.f90:45-202), prediction
//            (LIB/WAVELETS/module_wavelets.f90:96-284), set_send_bounds (LIB/MPI/calc_data_bounds.f90:99-146),
//            get_indices_of_ghost_patch (LIB/TREE/neighborhood.f90:158-331).
//
// One CTA per (receiver block, face) patch.  The ghost strip is H deep and Bs x Bs wide:
//   receiver coarser than the neighbours: every strip point coincides with an interior point of one of the four fine
//     neighbours (decimation, a copy);
//   receiver finer than the neighbour: the strip is interpolated (x, then y, then z, as the reference) from the box of
//     coarse-lattice points around it, each of which is the interior value of whichever leaf owns it (see resolve.cuh).
// HBM traffic: reads ~ the strip's footprint in the neighbours, writes nc*H*Bs^2 doubles per patch.
#include "resolve.cuh"
#include "wgpu_internal.cuh"

namespace {

struct JumpArgs {
    const double *u;
    double *jpool;
    long long jpatch;
    const int *jblk, *jdir;
    const signed char *level;
    const int *ixyz;
    BlockLookup L;
    int nc, Bs, H, order, dim;
    int periodic[3];
};

__device__ __forceinline__ double interp1(const double *p, int stride, int order, const double *c)
{
    // sum_t c[t] * coarse[start + t], products then sums, left to right (module_wavelets.f90:188-283), never contracted
    double acc = __dmul_rn(c[0], p[0]);
    for (int t = 1; t < order; ++t) acc = __dadd_rn(acc, __dmul_rn(c[t], p[t * stride]));
    return acc;
}

__global__ void __launch_bounds__(128) jump_fill_kernel(const JumpArgs a)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ SrcTable T;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int Bs = a.Bs, H = a.H, dim = a.dim;
    const int b = a.jblk[blockIdx.x], dcode = a.jdir[blockIdx.x];
    const int d[3] = {dcode % 3 - 1, (dcode / 3) % 3 - 1, dcode / 9 - 1};
    const int lvl = a.level[b];
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    int org[3], ext[3], bx[3];
    for (int k = 0; k < 3; ++k) {
        bx[k] = a.ixyz[3 * b + k];
        org[k] = d[k] < 0 ? -H : (d[k] > 0 ? Bs : 0);
        ext[k] = d[k] ? H : (k < dim ? Bs : 1);
    }
    const int npts = ext[0] * ext[1] * ext[2];
    double *out = a.jpool + (long long)blockIdx.x * a.jpatch;

    // which kind of patch? the level of the leaf that owns the strip
    int P0[3], lo[3], hi[3];
    for (int k = 0; k < 3; ++k) {
        lo[k] = bx[k] * Bs + org[k];
        hi[k] = lo[k] + ext[k] - 1;
    }
    src_table_build(T, a.L, lvl, lo, hi, Bs, dim, a.periodic, tid, nt);
    __syncthreads();
    {
        for (int k = 0; k < 3; ++k) P0[k] = lo[k];
        int sb, so;
        src_resolve(T, P0, Bs, dim, sb, so);
        if (sb >= 0) {
            // same level or finer owner: copy / decimation
            for (int i = tid; i < a.nc * npts; i += nt) {
                const int c = i / npts, r = i % npts;
                const int P[3] = {lo[0] + r % ext[0], lo[1] + (r / ext[0]) % ext[1], lo[2] + r / (ext[0] * ext[1])};
                src_resolve(T, P, Bs, dim, sb, so);
                out[i] = sb >= 0 ? a.u[((long long)sb * a.nc + c) * CS + so] : 0.0;
            }
            return;
        }
    }
    __syncthreads();

    // coarser owner: prediction from the level-(lvl-1) lattice
    const int order = a.order, A = order / 2 - 1;
    double cf[6];
    if (order == 2) { cf[0] = 0.5; cf[1] = 0.5; }
    else if (order == 4) { cf[0] = -1.0 / 16.0; cf[1] = 9.0 / 16.0; cf[2] = 9.0 / 16.0; cf[3] = -1.0 / 16.0; }
    else { cf[0] = 3.0 / 256.0; cf[1] = -25.0 / 256.0; cf[2] = 150.0 / 256.0; cf[3] = 150.0 / 256.0; cf[4] = -25.0 / 256.0; cf[5] = 3.0 / 256.0; }
    int clo[3], chi[3], n[3];
    for (int k = 0; k < 3; ++k) {
        if (k < dim) {
            clo[k] = (lo[k] >> 1) - A;
            chi[k] = ((hi[k] + 1) >> 1) + A;
        } else clo[k] = chi[k] = 0;
        n[k] = chi[k] - clo[k] + 1;
    }
    src_table_build(T, a.L, lvl - 1, clo, chi, Bs, dim, a.periodic, tid, nt);
    __syncthreads();
    double *cb = sm;                                   // [n2][n1][n0]
    double *t1 = cb + n[0] * n[1] * n[2];              // [n2][n1][e0]
    double *t2 = t1 + ext[0] * n[1] * n[2];            // [n2][e1][e0]
    for (int c = 0; c < a.nc; ++c) {
        for (int i = tid; i < n[0] * n[1] * n[2]; i += nt) {
            const int P[3] = {clo[0] + i % n[0], clo[1] + (i / n[0]) % n[1], clo[2] + i / (n[0] * n[1])};
            int sb, so;
            src_resolve(T, P, Bs, dim, sb, so);
            cb[i] = sb >= 0 ? a.u[((long long)sb * a.nc + c) * CS + so] : 0.0;
        }
        __syncthreads();
        // x
        for (int i = tid; i < ext[0] * n[1] * n[2]; i += nt) {
            const int x = i % ext[0], r = i / ext[0];
            const int G = lo[0] + x;
            const double *row = cb + r * n[0];
            t1[i] = (G & 1) ? interp1(row + ((G - 1) >> 1) - clo[0] - A, 1, order, cf) : row[(G >> 1) - clo[0]];
        }
        __syncthreads();
        // y
        for (int i = tid; i < ext[0] * ext[1] * n[2]; i += nt) {
            const int x = i % ext[0], y = (i / ext[0]) % ext[1], z = i / (ext[0] * ext[1]);
            const int G = lo[1] + y;
            const double *col = t1 + (z * n[1]) * ext[0] + x;
            t2[i] = (G & 1) ? interp1(col + (((G - 1) >> 1) - clo[1] - A) * ext[0], ext[0], order, cf) : col[((G >> 1) - clo[1]) * ext[0]];
        }
        __syncthreads();
        // z
        for (int i = tid; i < npts; i += nt) {
            const int xy = i % (ext[0] * ext[1]), z = i / (ext[0] * ext[1]);
            double v;
            if (dim == 3) {
                const int G = lo[2] + z;
                const int pl = ext[0] * ext[1];
                const double *col = t2 + xy;
                v = (G & 1) ? interp1(col + (((G - 1) >> 1) - clo[2] - A) * pl, pl, order, cf) : col[((G >> 1) - clo[2]) * pl];
            } else v = t2[xy];
            out[(long long)c * npts + i] = v;
        }
        __syncthreads();
    }
}

}  // namespace

int32_t wgpu_launch_jump_fill(wgpu_ctx *ctx, const double *src)
{
    if (ctx->n_jump == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    JumpArgs a;
    a.u = src;
    a.jpool = ctx->d_jpool;
    a.jblk = ctx->d_jump_blk;
    a.jdir = ctx->d_jump_dir;
    a.level = ctx->d_level;
    a.ixyz = ctx->d_ixyz;
    a.L.keys = ctx->d_hkeys;
    a.L.vals = ctx->d_hvals;
    a.L.mask = ctx->hmask;
    a.nc = ctx->nc;
    a.Bs = c.Bs[0];
    a.H = c.fd == 2 ? 1 : (c.fd == 4 ? 2 : 3);
    a.jpatch = (long long)ctx->nc * a.H * c.Bs[0] * (c.dim == 3 ? c.Bs[0] : 1);
    a.order = ctx->wavelet.X;
    a.dim = c.dim;
    for (int k = 0; k < 3; ++k) a.periodic[k] = c.periodic[k];
    // shared memory of the prediction branch: coarse box + two intermediates, largest over the face orientations
    const int A = a.order / 2 - 1, Bs = a.Bs, H = a.H;
    const int nt = Bs / 2 + 1 + 2 * A + 1, nn = H / 2 + 2 + 2 * A + 1;
    const int e3 = c.dim == 3 ? Bs : 1, n3 = c.dim == 3 ? nt : 1;
    size_t best = 0;
    for (int f = 0; f < c.dim; ++f) {
        const int n[3] = {f == 0 ? nn : nt, f == 1 ? nn : nt, c.dim == 3 ? (f == 2 ? nn : nt) : 1};
        const int e[3] = {f == 0 ? H : Bs, f == 1 ? H : Bs, c.dim == 3 ? (f == 2 ? H : Bs) : 1};
        const size_t s = (size_t)n[0] * n[1] * n[2] + (size_t)e[0] * n[1] * n[2] + (size_t)e[0] * e[1] * n[2];
        best = s > best ? s : best;
    }
    (void)e3;
    (void)n3;
    const size_t smem = best * sizeof(double);
    static size_t configured = 0;
    if (smem > configured) {
        WGPU_CHECK(ctx, cudaFuncSetAttribute(jump_fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    jump_fill_kernel<<<ctx->n_jump, 128, smem, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}
