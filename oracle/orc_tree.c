/*
 * oracle/orc_tree.c -- tree-level block loops of the time step in C (TEST INFRASTRUCTURE ONLY; see orc_acm.c).
 *
 * Used (a) as the CPU baseline / `bench.py --impl reference` leg: the reference's algorithm with the reference's
 * decomposition -- every block loop split into contiguous chunks of the (space-filling-curve ordered) block list,
 * one chunk per worker (OpenMP static schedule == one MPI rank per core, balanceLoad_tree.f90:600-715), and
 * (b) by tests to cross-check oracle.py's NumPy loops.
 *
 * Same-level ghost synchronisation only: on one node every neighbour is "internal", which the reference copies
 * patch by patch directly from the sender's interior strip into the receiver's ghost strip
 * (unpack_ghostlayers_internal, LIB/MPI/xfer_block_data.f90:321-440, bounds from calc_data_bounds.f90:99-146 and
 * neighborhood.f90:158-285).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef struct orc_acm_params orc_acm_params;
void orc_rhs_acm_3d(const orc_acm_params *p, int g, const int32_t Bs[3], const double dx[3], const double *phi, double *rhs,
                    const double *mask);
void orc_rhs_acm_2d(const orc_acm_params *p, int g, const int32_t Bs[3], const double dx[3], const double *phi, double *rhs,
                    const double *mask);
double orc_get_dt_block(const orc_acm_params *p, int g, const int32_t Bs[3], const double dx[3], const double *u);
void orc_rk_copy_interior(int dim, int g, const int32_t Bs[3], int nc, double *dst, const double *src);
void orc_rk_axpy_interior(int dim, int g, const int32_t Bs[3], int nc, double *y, double dt, double coef, const double *x);

/* nbr[b*27 + (dz+1)*9 + (dy+1)*3 + (dx+1)] = same-level neighbour block index (0-based) or -1 */
void orc_sync_same_level(int nb, const int32_t *nbr, int dim, int g, const int32_t Bs[3], int nc, double *hvy, int gm, int gp)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = sz * nz, sb = sc * nc;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
        for (int dz = (dim == 3 ? -1 : 0); dz <= (dim == 3 ? 1 : 0); ++dz)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    if (!dx && !dy && !dz) continue;
                    const int n = nbr[b * 27 + (dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
                    if (n < 0) continue;
                    const int d[3] = {dx, dy, dz};
                    int r0[3], r1[3], s0[3];
                    for (int a = 0; a < 3; ++a) {
                        const int B = a < dim ? Bs[a] : 1, ga = a < dim ? g : 0;
                        if (d[a] == 0) { r0[a] = ga; r1[a] = B + ga; s0[a] = ga; }
                        else if (d[a] < 0) { r0[a] = ga - gm; r1[a] = ga; s0[a] = B + ga - gm; }
                        else { r0[a] = B + ga; r1[a] = B + ga + gp; s0[a] = ga; }
                    }
                    for (int c = 0; c < nc; ++c)
                        for (int k = r0[2]; k < r1[2]; ++k)
                            for (int j = r0[1]; j < r1[1]; ++j)
                                memcpy(hvy + b * sb + c * sc + k * sz + j * sy + r0[0],
                                       hvy + n * sb + c * sc + (s0[2] + k - r0[2]) * sz + (s0[1] + j - r0[1]) * sy + s0[0],
                                       sizeof(double) * (size_t)(r1[0] - r0[0]));
                }
    }
}

double orc_dt_tree(const orc_acm_params *p, int nb, const double *dx_blocks, int dim, int g, const int32_t Bs[3], int nc,
                   const double *hvy)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sb = (ptrdiff_t)nx * ny * nz * nc;
    double dt = 9.0e9;
#pragma omp parallel for schedule(static) reduction(min : dt)
    for (int b = 0; b < nb; ++b) {
        const double d = orc_get_dt_block(p, g, Bs, dx_blocks + 3 * b, hvy + b * sb);
        if (d < dt) dt = d;
    }
    return dt;
}

void orc_rhs_tree(const orc_acm_params *p, int nb, const double *dx_blocks, int dim, int g, const int32_t Bs[3], int nc,
                  const double *hvy, double *rhs, const double *mask, int nmask)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sc = (ptrdiff_t)nx * ny * nz, sb = sc * nc;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
        const double *m = mask ? mask + (ptrdiff_t)b * sc * nmask : NULL;
        if (dim == 3) orc_rhs_acm_3d(p, g, Bs, dx_blocks + 3 * b, hvy + b * sb, rhs + b * sb, m);
        else orc_rhs_acm_2d(p, g, Bs, dx_blocks + 3 * b, hvy + b * sb, rhs + b * sb, m);
    }
}

/*
 * RungeKuttaGeneric (runge_kutta_generic.f90:50-154) on a same-level grid, dt given.
 * work: (s+1) arrays laid out like hvy, contiguous: work + slot*nb*sb, slot 0 = copy of the state.
 * butcher: row-major (s+1)x(s+1).
 */
void orc_rk_step_same_level(const orc_acm_params *p, int nb, const int32_t *nbr, const double *dx_blocks, int dim, int g, int g_rhs,
                            const int32_t Bs[3], int nc, double *hvy, double *work, const double *butcher, int s, double dt,
                            const double *mask, int nmask)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sb = (ptrdiff_t)nx * ny * nz * nc, sw = sb * nb;
    const int ld = s + 1;
    /* the caller has synchronised the ghosts and computed dt (runge_kutta_generic.f90:52-56) */
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) orc_rk_copy_interior(dim, g, Bs, nc, work + b * sb, hvy + b * sb);
    orc_rhs_tree(p, nb, dx_blocks, dim, g, Bs, nc, hvy, work + 1 * sw, mask, nmask);
    for (int j = 2; j <= s; ++j) {
#pragma omp parallel for schedule(static)
        for (int b = 0; b < nb; ++b) {
            orc_rk_copy_interior(dim, g, Bs, nc, hvy + b * sb, work + b * sb);
            for (int l = 2; l <= j; ++l) {
                const double coef = butcher[(j - 1) * ld + (l - 1)];
                if (fabs(coef) < 1.0e-8) continue;
                orc_rk_axpy_interior(dim, g, Bs, nc, hvy + b * sb, dt, coef, work + (l - 1) * sw + b * sb);
            }
        }
        orc_sync_same_level(nb, nbr, dim, g, Bs, nc, hvy, g_rhs, g_rhs);
        orc_rhs_tree(p, nb, dx_blocks, dim, g, Bs, nc, hvy, work + j * sw, mask, nmask);
    }
#pragma omp parallel for schedule(static)
    for (int b = 0; b < nb; ++b) {
        orc_rk_copy_interior(dim, g, Bs, nc, hvy + b * sb, work + b * sb);
        for (int j = 2; j <= s + 1; ++j) {
            const double coef = butcher[s * ld + (j - 1)];
            if (fabs(coef) < 1.0e-8) continue;
            orc_rk_axpy_interior(dim, g, Bs, nc, hvy + b * sb, dt, coef, work + (j - 1) * sw + b * sb);
        }
    }
}
