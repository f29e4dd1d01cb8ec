// fill_region: value of every point of a small box of the level-L lattice on a graded leaf grid, i.e. what the reference's
// ghost synchronisation (sync_ghosts_generic, "full_leaf"; LIB/MPI/synchronize_ghosts_generic.f90:181-343)
// leaves in the ghost nodes of a level-L block:
//   owner leaf on level L     -> copy                      (stage 1; xfer_block_data.f90:395-400)
//   owner leaf on level L+1   -> decimation                (restrict_data, LIB/MPI/restrict_predict_data.f90:45-115) of the leaf itself
//                                (ignore_Filter, unlifted wavelets) or of its HD-filtered copy (FillCtx::rpool)
//   owner leaf on level L-1   -> prediction x, y, z        (predict_data :174-202; prediction, LIB/WAVELETS/module_wavelets.f90:96-284)
// The box must lie inside ONE level-L cell (block-sized region), so that all its points share the kind of owner; the
// coarse lattice points the interpolation touches are resolved the same way one level down (see resolve.cuh).
// tests/test_oracle_sync.py::test_sync_equals_geometric_definition_on_graded_grids pins this definition against the
// oracle's restatement of the reference's staged, table-driven algorithm (bit for bit).
#pragma once

#include "resolve.cuh"

struct FillCtx {
    const double *u;      // compact array [blk][nc][Bs^dim]
    // sync_ghosts_tree with a lifted wavelet (ignore_Filter = .false.): values taken from a finer leaf come from its HD-filtered,
    // decimated copy (restrict_copy_at_CE, LIB/MPI/restrict_predict_data.f90:121-172), prepared by restrict_filter_kernel:
    // rpool[rmap[blk]][nc][(Bs/2)^dim]; nullptr = plain decimation (sync_ghosts_RHS_tree, unlifted wavelets)
    const double *rpool;
    const int *rmap;
    BlockLookup L;
    int nc, Bs, dim, order;
    int periodic[3];
};

__device__ __forceinline__ double interp1(const double *p, int stride, int order, const double *c)
{
    // sum_t c[t] * coarse[start + t], products then sums, left to right (module_wavelets.f90:188-283), never contracted
    double acc = __dmul_rn(c[0], p[0]);
    for (int t = 1; t < order; ++t) acc = __dadd_rn(acc, __dmul_rn(c[t], p[t * stride]));
    return acc;
}

// shared-memory doubles the prediction branch needs for a box of extent e[3] (any orientation)
__host__ __device__ inline size_t fill_scratch_doubles(const int e[3], int order, int dim)
{
    const int A = order / 2 - 1;
    int n[3];
    for (int k = 0; k < 3; ++k) n[k] = k < dim ? e[k] / 2 + 2 + 2 * A : 1;
    return (size_t)n[0] * n[1] * n[2] + (size_t)e[0] * n[1] * n[2] + (size_t)e[0] * e[1] * n[2];
}

// Interpolate the box [lo, lo+ext) of the fine lattice from the coarse box cb[n2][n1][n0] whose origin is clo (coarse
// coordinates; fine point G coincides with coarse point G/2).  x, then y, then z as `prediction`.  The two intermediates live
// right behind cb.  out[z*sz + y*sy + x].  All threads of the CTA; ends with a barrier.
__device__ inline void predict_from_box(double *cb, const int clo[3], const int n[3], const int lo[3], const int ext[3], int order, int dim,
                                        double *out, long long sy, long long sz, int tid, int nt)
{
    const int A = order / 2 - 1;
    double cf[6];
    if (order == 2) { cf[0] = 0.5; cf[1] = 0.5; }
    else if (order == 4) { cf[0] = -1.0 / 16.0; cf[1] = 9.0 / 16.0; cf[2] = 9.0 / 16.0; cf[3] = -1.0 / 16.0; }
    else { cf[0] = 3.0 / 256.0; cf[1] = -25.0 / 256.0; cf[2] = 150.0 / 256.0; cf[3] = 150.0 / 256.0; cf[4] = -25.0 / 256.0; cf[5] = 3.0 / 256.0; }
    double *t1 = cb + n[0] * n[1] * n[2];              // [n2][n1][e0]
    double *t2 = t1 + ext[0] * n[1] * n[2];            // [n2][e1][e0]
    for (int i = tid; i < ext[0] * n[1] * n[2]; i += nt) {          // x
        const int x = i % ext[0], r = i / ext[0];
        const int G = lo[0] + x;
        const double *row = cb + r * n[0];
        t1[i] = (G & 1) ? interp1(row + ((G - 1) >> 1) - clo[0] - A, 1, order, cf) : row[(G >> 1) - clo[0]];
    }
    __syncthreads();
    for (int i = tid; i < ext[0] * ext[1] * n[2]; i += nt) {        // y
        const int x = i % ext[0], y = (i / ext[0]) % ext[1], z = i / (ext[0] * ext[1]);
        const int G = lo[1] + y;
        const double *col = t1 + (z * n[1]) * ext[0] + x;
        t2[i] = (G & 1) ? interp1(col + (((G - 1) >> 1) - clo[1] - A) * ext[0], ext[0], order, cf) : col[((G >> 1) - clo[1]) * ext[0]];
    }
    __syncthreads();
    const int pl = ext[0] * ext[1];
    for (int i = tid; i < pl * ext[2]; i += nt) {                   // z
        const int xy = i % pl, z = i / pl;
        double v;
        if (dim == 3) {
            const int G = lo[2] + z;
            const double *col = t2 + xy;
            v = (G & 1) ? interp1(col + (((G - 1) >> 1) - clo[2] - A) * pl, pl, order, cf) : col[((G >> 1) - clo[2]) * pl];
        } else v = t2[xy];
        out[z * sz + (xy / ext[0]) * sy + xy % ext[0]] = v;
    }
    __syncthreads();
}

// All threads of the CTA call this.  lo[3]: unwrapped global lattice coordinates (level `lvl`) of the box origin, ext[3] its
// extents; out[c*sc + z*sz + y*sy + x] receives component c0 + c for c < ncomp.  scratch: fill_scratch_doubles(ext) doubles
// of shared memory.  T: a SrcTable in shared memory.  Returns 0 copy/decimation, 1 prediction, -1 no owner (zeros written if
// zero_if_no_owner, else nothing).
__device__ inline int fill_region(const FillCtx &a, SrcTable &T, double *scratch, int lvl, const int lo[3], const int ext[3], double *out,
                                  long long sc, long long sy, long long sz, int c0, int ncomp, int tid, int nt, bool zero_if_no_owner = false)
{
    const int Bs = a.Bs, dim = a.dim;
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    const int npts = ext[0] * ext[1] * ext[2];
    int hi[3];
    for (int k = 0; k < 3; ++k) hi[k] = lo[k] + ext[k] - 1;
    __syncthreads();   // T and scratch may still be in use by a previous call
    src_table_build(T, a.L, lvl, lo, hi, Bs, dim, a.periodic, tid, nt);
    __syncthreads();
    int sb, so;
    src_resolve(T, lo, Bs, dim, sb, so);
    if (sb >= 0) {
        const long long RS = CS >> dim;
        for (int i = tid; i < ncomp * npts; i += nt) {
            const int c = i / npts, r = i % npts;
            const int x = r % ext[0], y = (r / ext[0]) % ext[1], z = r / (ext[0] * ext[1]);
            const int P[3] = {lo[0] + x, lo[1] + y, lo[2] + z};
            int ro;
            src_resolve(T, P, Bs, dim, sb, so, ro);
            double v = 0.0;
            if (sb >= 0) {
                const int ri = (ro >= 0 && a.rpool) ? a.rmap[sb] : -1;
                v = ri >= 0 ? a.rpool[((long long)ri * a.nc + c0 + c) * RS + ro] : a.u[((long long)sb * a.nc + c0 + c) * CS + so];
            }
            out[c * sc + z * sz + y * sy + x] = v;
        }
        return 0;
    }
    auto no_owner = [&]() {   // pool patches must not keep stale values; the download leaves the host's ghost nodes alone instead
        if (zero_if_no_owner)
            for (int i = tid; i < ncomp * npts; i += nt) {
                const int c = i / npts, r = i % npts;
                out[c * sc + (r / (ext[0] * ext[1])) * sz + ((r / ext[0]) % ext[1]) * sy + r % ext[0]] = 0.0;
            }
        return -1;
    };
    if (lvl == 0) return no_owner();
    // coarser owner: prediction from the level-(lvl-1) lattice
    const int order = a.order, A = order / 2 - 1;
    int clo[3], chi[3], n[3];
    for (int k = 0; k < 3; ++k) {
        if (k < dim) {
            clo[k] = (lo[k] >> 1) - A;
            chi[k] = ((hi[k] + 1) >> 1) + A;
        } else clo[k] = chi[k] = 0;
        n[k] = chi[k] - clo[k] + 1;
    }
    __syncthreads();
    src_table_build(T, a.L, lvl - 1, clo, chi, Bs, dim, a.periodic, tid, nt);
    __syncthreads();
    {
        const int Pc[3] = {lo[0] >> 1, lo[1] >> 1, lo[2] >> 1};
        src_resolve(T, Pc, Bs, dim, sb, so);
        if (sb < 0) return no_owner();                 // outside a non-periodic domain / no such block
    }
    double *cb = scratch;                              // [n2][n1][n0]
    for (int c = 0; c < ncomp; ++c) {
        for (int i = tid; i < n[0] * n[1] * n[2]; i += nt) {
            const int P[3] = {clo[0] + i % n[0], clo[1] + (i / n[0]) % n[1], clo[2] + i / (n[0] * n[1])};
            src_resolve(T, P, Bs, dim, sb, so);
            cb[i] = sb >= 0 ? a.u[((long long)sb * a.nc + c0 + c) * CS + so] : 0.0;
        }
        __syncthreads();
        predict_from_box(cb, clo, n, lo, ext, order, dim, out + c * sc, sy, sz, tid, nt);
    }
    return 1;
}
