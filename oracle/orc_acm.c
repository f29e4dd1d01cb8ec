/*
 * oracle/orc_acm.c -- CPU restatement of WABBIT's ACM right-hand side, time-step
 * restriction and generic Runge-Kutta stage arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (wabbit_b200/) never links or calls it.
 *
 * Every function follows the reference Fortran statement by statement, with the
 * reference's evaluation order (sums left to right, `(sum)*dx_inv`, centre
 * coefficient of first derivatives omitted).  The parity build uses
 * -O2 -ffp-contract=off (the reference's default build is `mpif90 -O3` for generic
 * x86-64: no FMA, LIB/fortran.mk:72-84); the timing build (cpu_baseline) uses
 * -O3 -march=native.
 *
 * Array layout: Fortran column-major, phi(nx,ny,nz,nc) with nx = Bs+2g, x fastest.
 * Indices below are 0-based: Fortran ix = g+1..Bs+g  <->  C i = g..Bs+g-1.
 *
 * Reference:
 *   RHS_3D_acm            LIB/EQUATION/ACMnew/rhs_ACM.f90:927-1779
 *   RHS_2D_acm            LIB/EQUATION/ACMnew/rhs_ACM.f90:292-922
 *   GET_DT_BLOCK_ACM      LIB/EQUATION/ACMnew/module_ACM.f90:617-691
 *   RungeKuttaGeneric     LIB/TIME/runge_kutta_generic.f90:50-154
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define ORC_NCOLORS 16

typedef struct orc_acm_params {
    int32_t dim;            /* 2 or 3 */
    int32_t fd;             /* 2, 4, 6 = FD_{2nd,4th,6th}_central ; 40 = FD_4th_central_optimized (TW4) */
    int32_t skew;           /* params_acm%skew_symmetry */
    int32_t penalization;   /* params_acm%penalization */
    int32_t use_sponge;     /* params_acm%use_sponge */
    int32_t pad_;
    double c0, nu, gamma_p;
    double C_eta;           /* 1/C_eta applied for every colour >=1 (module_ACM.f90:272) */
    double C_sponge;
    double u_mean_set[3];
    double CFL, CFL_eta, CFL_nu;
} orc_acm_params;

/* stencil tables, rhs_ACM.f90:976-985 */
static const double a_FD4[5] = {1.0 / 12.0, -2.0 / 3.0, 0.0, +2.0 / 3.0, -1.0 / 12.0};
static const double b_FD4[5] = {-1.0 / 12.0, 4.0 / 3.0, -5.0 / 2.0, 4.0 / 3.0, -1.0 / 12.0};
static const double a_TW4[7] = {-0.02651995, +0.18941314, -0.79926643, 0.0, 0.79926643, -0.18941314, 0.02651995};
static const double a_FD6[7] = {-1.0 / 60.0, 3.0 / 20.0, -3.0 / 4.0, 0.0, 3.0 / 4.0, -3.0 / 20.0, 1.0 / 60.0};
static const double b_FD6[7] = {1.0 / 90.0, -3.0 / 20.0, 3.0 / 2.0, -49.0 / 18.0, 3.0 / 2.0, -3.0 / 20.0, 1.0 / 90.0};

/* P(c, off) = phi at the current point shifted by `off` along the active direction (stride s) */
#define P(c, o) ph[(c)][idx + (ptrdiff_t)(o) * s]

/* first derivative, 2nd order: (phi(+1) - phi(-1))*dx_inv*0.5   (rhs_ACM.f90:1024) */
#define D1_2(c) ((P(c, 1) - P(c, -1)) * dinv * 0.5)
#define D1P_2(c, d) ((P(c, 1) * P(d, 1) - P(c, -1) * P(d, -1)) * dinv * 0.5)
/* second derivative, 2nd order (rhs_ACM.f90:1048) */
#define D2_2(c) ((P(c, -1) - 2.0 * P(c, 0) + P(c, 1)) * d2inv)

/* 5-point forms (rhs_ACM.f90:1147, 1162, 1175) */
#define D1_4(c) ((a_FD4[0] * P(c, -2) + a_FD4[1] * P(c, -1) + a_FD4[3] * P(c, 1) + a_FD4[4] * P(c, 2)) * dinv)
#define D1P_4(c, d) ((a_FD4[0] * P(c, -2) * P(d, -2) + a_FD4[1] * P(c, -1) * P(d, -1) + a_FD4[3] * P(c, 1) * P(d, 1) + a_FD4[4] * P(c, 2) * P(d, 2)) * dinv)
#define D2_4(c) ((b_FD4[0] * P(c, -2) + b_FD4[1] * P(c, -1) + b_FD4[2] * P(c, 0) + b_FD4[3] * P(c, 1) + b_FD4[4] * P(c, 2)) * d2inv)

/* 7-point forms with coefficient table A / B (rhs_ACM.f90:1352ff FD6, 1469ff TW4) */
#define D1_7(A, c) ((A[0] * P(c, -3) + A[1] * P(c, -2) + A[2] * P(c, -1) + A[4] * P(c, 1) + A[5] * P(c, 2) + A[6] * P(c, 3)) * dinv)
#define D1P_7(A, c, d) ((A[0] * P(c, -3) * P(d, -3) + A[1] * P(c, -2) * P(d, -2) + A[2] * P(c, -1) * P(d, -1) + A[4] * P(c, 1) * P(d, 1) + A[5] * P(c, 2) * P(d, 2) + A[6] * P(c, 3) * P(d, 3)) * dinv)
#define D2_7(B, c) ((B[0] * P(c, -3) + B[1] * P(c, -2) + B[2] * P(c, -1) + B[3] * P(c, 0) + B[4] * P(c, 1) + B[5] * P(c, 2) + B[6] * P(c, 3)) * d2inv)

static inline __attribute__((always_inline)) double d1(const int fd, const double *const *ph, ptrdiff_t idx, ptrdiff_t s, double dinv, int c)
{
    switch (fd) {
    case 2: return D1_2(c);
    case 4: return D1_4(c);
    case 6: return D1_7(a_FD6, c);
    default: return D1_7(a_TW4, c);
    }
}
static inline __attribute__((always_inline)) double d1p(const int fd, const double *const *ph, ptrdiff_t idx, ptrdiff_t s, double dinv, int c, int d)
{
    switch (fd) {
    case 2: return D1P_2(c, d);
    case 4: return D1P_4(c, d);
    case 6: return D1P_7(a_FD6, c, d);
    default: return D1P_7(a_TW4, c, d);
    }
}
static inline __attribute__((always_inline)) double d2(const int fd, const double *const *ph, ptrdiff_t idx, ptrdiff_t s, double d2inv, int c)
{
    switch (fd) {
    case 2: return D2_2(c);
    case 4: return D2_4(c);
    case 6: return D2_7(b_FD6, c);
    default: return D2_4(c); /* TW4 uses the standard 4th-order second derivative, rhs_ACM.f90:1484 */
    }
}

int orc_fd_halfwidth(int fd) { return fd == 2 ? 1 : (fd == 4 ? 2 : 3); }

/*
 * RHS_3D_acm, p_eqn_model='acm' (rhs_ACM.f90:927-1779).
 * mask may be NULL: equivalent to chi == 0 and sponge mask == 0 everywhere
 * (the reference multiplies by mask(:,:,:,1)=0, which only adds a signed zero).
 */
/* The body is compiled once per stencil family and skew flag (fd, skew are literal constants at every call site below, so the selection
 * inside d1 / d1p / d2 folds away); the arithmetic is the same statement sequence for every instance. */
static inline __attribute__((always_inline)) void rhs_acm_3d_body(const int fd, const int skew, const orc_acm_params *p, int g, const int32_t Bs[3],
                                                                  const double dx[3], const double *phi, double *rhs, const double *mask)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = Bs[2] + 2 * g;
    const ptrdiff_t sx = 1, sy = nx, sz = (ptrdiff_t)nx * ny, sc = (ptrdiff_t)nx * ny * nz;
    const double *ph[4] = {phi, phi + sc, phi + 2 * sc, phi + 3 * sc};
    const double c_0 = p->c0, nu = p->nu, gamma = p->gamma_p;
    const double dx_inv = 1.0 / dx[0], dy_inv = 1.0 / dx[1], dz_inv = 1.0 / dx[2];
    const double dx2_inv = 1.0 / (dx[0] * dx[0]), dy2_inv = 1.0 / (dx[1] * dx[1]), dz2_inv = 1.0 / (dx[2] * dx[2]);
    double C_eta_apply_inv[ORC_NCOLORS + 1];
    for (int c = 0; c <= ORC_NCOLORS; ++c) C_eta_apply_inv[c] = 1.0 / p->C_eta;
    C_eta_apply_inv[0] = 0.0;

    for (int iz = g; iz < Bs[2] + g; ++iz)
        for (int iy = g; iy < Bs[1] + g; ++iy)
            for (int ix = g; ix < Bs[0] + g; ++ix) {
                const ptrdiff_t idx = ix + iy * sy + iz * sz;
                const double u_dx = d1(fd, ph, idx, sx, dx_inv, 0), v_dx = d1(fd, ph, idx, sx, dx_inv, 1);
                const double w_dx = d1(fd, ph, idx, sx, dx_inv, 2), p_dx = d1(fd, ph, idx, sx, dx_inv, 3);
                const double u_dy = d1(fd, ph, idx, sy, dy_inv, 0), v_dy = d1(fd, ph, idx, sy, dy_inv, 1);
                const double w_dy = d1(fd, ph, idx, sy, dy_inv, 2), p_dy = d1(fd, ph, idx, sy, dy_inv, 3);
                const double u_dz = d1(fd, ph, idx, sz, dz_inv, 0), v_dz = d1(fd, ph, idx, sz, dz_inv, 1);
                const double w_dz = d1(fd, ph, idx, sz, dz_inv, 2), p_dz = d1(fd, ph, idx, sz, dz_inv, 3);

                const double u_dxdx = d2(fd, ph, idx, sx, dx2_inv, 0), v_dxdx = d2(fd, ph, idx, sx, dx2_inv, 1), w_dxdx = d2(fd, ph, idx, sx, dx2_inv, 2);
                const double u_dydy = d2(fd, ph, idx, sy, dy2_inv, 0), v_dydy = d2(fd, ph, idx, sy, dy2_inv, 1), w_dydy = d2(fd, ph, idx, sy, dy2_inv, 2);
                const double u_dzdz = d2(fd, ph, idx, sz, dz2_inv, 0), v_dzdz = d2(fd, ph, idx, sz, dz2_inv, 1), w_dzdz = d2(fd, ph, idx, sz, dz2_inv, 2);

                const double u = ph[0][idx], v = ph[1][idx], w = ph[2][idx], pp = ph[3][idx];

                double penalx = 0.0, penaly = 0.0, penalz = 0.0;
                if (mask) {
                    /* chi = mask(1) * C_eta_apply_inv(int(mask(5)))   rhs_ACM.f90:1192 */
                    const double chi = mask[idx] * C_eta_apply_inv[(int)mask[idx + 4 * sc]];
                    penalx = -chi * (u - mask[idx + 1 * sc]);
                    penaly = -chi * (v - mask[idx + 2 * sc]);
                    penalz = -chi * (w - mask[idx + 3 * sc]);
                }

                if (skew) {
                    const double uu_dx = d1p(fd, ph, idx, sx, dx_inv, 0, 0), uv_dy = d1p(fd, ph, idx, sy, dy_inv, 0, 1), uw_dz = d1p(fd, ph, idx, sz, dz_inv, 0, 2);
                    const double vu_dx = d1p(fd, ph, idx, sx, dx_inv, 1, 0), vv_dy = d1p(fd, ph, idx, sy, dy_inv, 1, 1), vw_dz = d1p(fd, ph, idx, sz, dz_inv, 1, 2);
                    const double wu_dx = d1p(fd, ph, idx, sx, dx_inv, 2, 0), wv_dy = d1p(fd, ph, idx, sy, dy_inv, 2, 1), ww_dz = d1p(fd, ph, idx, sz, dz_inv, 2, 2);
                    /* rhs_ACM.f90:1199-1202 */
                    rhs[idx + 0 * sc] = -0.5 * (uu_dx + uv_dy + uw_dz + u * u_dx + v * u_dy + w * u_dz) - p_dx + nu * (u_dxdx + u_dydy + u_dzdz) + penalx;
                    rhs[idx + 1 * sc] = -0.5 * (vu_dx + vv_dy + vw_dz + u * v_dx + v * v_dy + w * v_dz) - p_dy + nu * (v_dxdx + v_dydy + v_dzdz) + penaly;
                    rhs[idx + 2 * sc] = -0.5 * (wu_dx + wv_dy + ww_dz + u * w_dx + v * w_dy + w * w_dz) - p_dz + nu * (w_dxdx + w_dydy + w_dzdz) + penalz;
                } else {
                    /* rhs_ACM.f90:1249-1252 */
                    rhs[idx + 0 * sc] = (-u * u_dx - v * u_dy - w * u_dz) - p_dx + nu * (u_dxdx + u_dydy + u_dzdz) + penalx;
                    rhs[idx + 1 * sc] = (-u * v_dx - v * v_dy - w * v_dz) - p_dy + nu * (v_dxdx + v_dydy + v_dzdz) + penaly;
                    rhs[idx + 2 * sc] = (-u * w_dx - v * w_dy - w * w_dz) - p_dz + nu * (w_dxdx + w_dydy + w_dzdz) + penalz;
                }
                rhs[idx + 3 * sc] = -(c_0 * c_0) * (u_dx + v_dy + w_dz) - gamma * pp;
            }

    /* sponge term, rhs_ACM.f90:1734-1753 */
    if (p->use_sponge && mask) {
        const double C_sponge_inv = 1.0 / p->C_sponge;
        for (int iz = g; iz < Bs[2] + g; ++iz)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                for (int ix = g; ix < Bs[0] + g; ++ix) {
                    const ptrdiff_t idx = ix + iy * sy + iz * sz;
                    const double spo = mask[idx + 5 * sc] * C_sponge_inv;
                    rhs[idx + 0 * sc] = rhs[idx + 0 * sc] - (ph[0][idx] - p->u_mean_set[0]) * spo;
                    rhs[idx + 1 * sc] = rhs[idx + 1 * sc] - (ph[1][idx] - p->u_mean_set[1]) * spo;
                    rhs[idx + 2 * sc] = rhs[idx + 2 * sc] - (ph[2][idx] - p->u_mean_set[2]) * spo;
                    rhs[idx + 3 * sc] = rhs[idx + 3 * sc] - (ph[3][idx]) * spo;
                }
    }
}

#define ORC_FD_DISPATCH(body)                                                                     \
    switch (p->fd) {                                                                             \
    case 2: if (p->skew) body(2, 1, p, g, Bs, dx, phi, rhs, mask); else body(2, 0, p, g, Bs, dx, phi, rhs, mask); break;   \
    case 4: if (p->skew) body(4, 1, p, g, Bs, dx, phi, rhs, mask); else body(4, 0, p, g, Bs, dx, phi, rhs, mask); break;   \
    case 6: if (p->skew) body(6, 1, p, g, Bs, dx, phi, rhs, mask); else body(6, 0, p, g, Bs, dx, phi, rhs, mask); break;   \
    default: if (p->skew) body(40, 1, p, g, Bs, dx, phi, rhs, mask); else body(40, 0, p, g, Bs, dx, phi, rhs, mask); break; \
    }

void orc_rhs_acm_3d(const orc_acm_params *p, int g, const int32_t Bs[3], const double dx[3],
                    const double *phi, double *rhs, const double *mask)
{
    ORC_FD_DISPATCH(rhs_acm_3d_body)
}

/*
 * RHS_2D_acm, p_eqn_model='acm', no lamballais geometry (rhs_ACM.f90:292-922).
 * phi(nx,ny,3) = (ux, uy, p).
 */
static inline __attribute__((always_inline)) void rhs_acm_2d_body(const int fd, const int skew, const orc_acm_params *p, int g, const int32_t Bs[3],
                                                                  const double dx[3], const double *phi, double *rhs, const double *mask)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g;
    const ptrdiff_t sx = 1, sy = nx, sc = (ptrdiff_t)nx * ny;
    const double *ph[4] = {phi, phi + sc, phi + 2 * sc, phi + 2 * sc};
    const double c_0 = p->c0, nu = p->nu, gamma = p->gamma_p;
    const double dx_inv = 1.0 / dx[0], dy_inv = 1.0 / dx[1];
    const double dx2_inv = 1.0 / (dx[0] * dx[0]), dy2_inv = 1.0 / (dx[1] * dx[1]);
    double C_eta_apply_inv[ORC_NCOLORS + 1];
    for (int c = 0; c <= ORC_NCOLORS; ++c) C_eta_apply_inv[c] = 1.0 / p->C_eta;
    C_eta_apply_inv[0] = 0.0;

    for (int iy = g; iy < Bs[1] + g; ++iy)
        for (int ix = g; ix < Bs[0] + g; ++ix) {
            const ptrdiff_t idx = ix + iy * sy;
            const double u_dx = d1(fd, ph, idx, sx, dx_inv, 0), v_dx = d1(fd, ph, idx, sx, dx_inv, 1), p_dx = d1(fd, ph, idx, sx, dx_inv, 2);
            const double u_dy = d1(fd, ph, idx, sy, dy_inv, 0), v_dy = d1(fd, ph, idx, sy, dy_inv, 1), p_dy = d1(fd, ph, idx, sy, dy_inv, 2);
            const double u_dxdx = d2(fd, ph, idx, sx, dx2_inv, 0), v_dxdx = d2(fd, ph, idx, sx, dx2_inv, 1);
            const double u_dydy = d2(fd, ph, idx, sy, dy2_inv, 0), v_dydy = d2(fd, ph, idx, sy, dy2_inv, 1);
            const double div_U = u_dx + v_dy;

            double penalx = 0.0, penaly = 0.0;
            if (mask) {
                /* rhs_ACM.f90:600-602: -mask(1)*C_eta_apply_inv(color)*(phi - mask(2)) */
                const int color = (int)mask[idx + 4 * sc];
                penalx = -mask[idx] * C_eta_apply_inv[color] * (ph[0][idx] - mask[idx + 1 * sc]);
                penaly = -mask[idx] * C_eta_apply_inv[color] * (ph[1][idx] - mask[idx + 2 * sc]);
            }
            if (skew) {
                const double uu_dx = d1p(fd, ph, idx, sx, dx_inv, 0, 0), uv_dy = d1p(fd, ph, idx, sy, dy_inv, 0, 1);
                const double vu_dx = d1p(fd, ph, idx, sx, dx_inv, 1, 0), vv_dy = d1p(fd, ph, idx, sy, dy_inv, 1, 1);
                /* rhs_ACM.f90:574-576 */
                rhs[idx + 0 * sc] = -0.5 * (uu_dx + uv_dy + ph[0][idx] * u_dx + ph[1][idx] * u_dy) - p_dx + nu * (u_dxdx + u_dydy) + penalx;
                rhs[idx + 1 * sc] = -0.5 * (vu_dx + vv_dy + ph[0][idx] * v_dx + ph[1][idx] * v_dy) - p_dy + nu * (v_dxdx + v_dydy) + penaly;
            } else {
                /* rhs_ACM.f90:604-605 */
                rhs[idx + 0 * sc] = -ph[0][idx] * u_dx - ph[1][idx] * u_dy - p_dx + nu * (u_dxdx + u_dydy) + penalx;
                rhs[idx + 1 * sc] = -ph[0][idx] * v_dx - ph[1][idx] * v_dy - p_dy + nu * (v_dxdx + v_dydy) + penaly;
            }
            rhs[idx + 2 * sc] = -(c_0 * c_0) * div_U - gamma * ph[2][idx];
        }

    if (p->use_sponge && mask) { /* rhs_ACM.f90:880-892 */
        const double C_sponge_inv = 1.0 / p->C_sponge;
        for (int iy = g; iy < Bs[1] + g; ++iy)
            for (int ix = g; ix < Bs[0] + g; ++ix) {
                const ptrdiff_t idx = ix + iy * sy;
                const double spo = mask[idx + 5 * sc] * C_sponge_inv;
                rhs[idx + 0 * sc] = rhs[idx + 0 * sc] - (ph[0][idx] - p->u_mean_set[0]) * spo;
                rhs[idx + 1 * sc] = rhs[idx + 1 * sc] - (ph[1][idx] - p->u_mean_set[1]) * spo;
                rhs[idx + 2 * sc] = rhs[idx + 2 * sc] - ph[2][idx] * spo;
            }
    }
}

void orc_rhs_acm_2d(const orc_acm_params *p, int g, const int32_t Bs[3], const double dx[3],
                    const double *phi, double *rhs, const double *mask)
{
    ORC_FD_DISPATCH(rhs_acm_2d_body)
}

/* GET_DT_BLOCK_ACM (module_ACM.f90:617-691), without passive scalars. */
double orc_get_dt_block(const orc_acm_params *p, int g, const int32_t Bs[3], const double dx[3], const double *u)
{
    const int dim = p->dim;
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = (ptrdiff_t)nx * ny * nz;
    double u_mag = -INFINITY;
    const int z0 = dim == 3 ? g : 0, z1 = dim == 3 ? Bs[2] + g : 1;
    for (int iz = z0; iz < z1; ++iz)
        for (int iy = g; iy < Bs[1] + g; ++iy)
            for (int ix = g; ix < Bs[0] + g; ++ix) {
                const ptrdiff_t idx = ix + iy * sy + iz * sz;
                double m = u[idx] * u[idx] + u[idx + sc] * u[idx + sc];
                if (dim == 3) m = m + u[idx + 2 * sc] * u[idx + 2 * sc];
                if (m > u_mag) u_mag = m;
            }
    double dxmin = dx[0];
    for (int d = 1; d < dim; ++d) dxmin = dx[d] < dxmin ? dx[d] : dxmin;
    const double u_eigen = sqrt(u_mag) + sqrt(p->c0 * p->c0 + u_mag);
    double dt;
    if (u_eigen >= 1.0e-6) dt = p->CFL * dxmin / u_eigen;
    else dt = 1.0e-2;
    if (p->nu > 1.0e-13) dt = fmin(dt, p->CFL_nu * (dxmin * dxmin) / p->nu);
    if (p->gamma_p > 0) dt = fmin(dt, p->CFL_eta * p->gamma_p);
    if (p->penalization) dt = fmin(dt, p->CFL_eta * p->C_eta);
    if (p->use_sponge) dt = fmin(dt, p->CFL_eta * p->C_sponge);
    return dt;
}

/*
 * RK stage arithmetic on the interior of one ghosted block (runge_kutta_generic.f90:63-67,90-112,136-154).
 *   orc_rk_copy_interior : dst(interior) = src(interior)
 *   orc_rk_axpy_interior : y(interior)   = y(interior) + dt*coef*x(interior), evaluated as (dt*coef)*x
 */
void orc_rk_copy_interior(int dim, int g, const int32_t Bs[3], int nc, double *dst, const double *src)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = (ptrdiff_t)nx * ny * nz;
    const int z0 = dim == 3 ? g : 0, z1 = dim == 3 ? Bs[2] + g : 1;
    for (int c = 0; c < nc; ++c)
        for (int iz = z0; iz < z1; ++iz)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                memcpy(dst + c * sc + iz * sz + iy * sy + g, src + c * sc + iz * sz + iy * sy + g, sizeof(double) * Bs[0]);
}

void orc_rk_axpy_interior(int dim, int g, const int32_t Bs[3], int nc, double *y, double dt, double coef, const double *x)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = (ptrdiff_t)nx * ny * nz;
    const int z0 = dim == 3 ? g : 0, z1 = dim == 3 ? Bs[2] + g : 1;
    const double a = dt * coef;
    for (int c = 0; c < nc; ++c)
        for (int iz = z0; iz < z1; ++iz)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                for (int ix = g; ix < Bs[0] + g; ++ix) {
                    const ptrdiff_t idx = c * sc + iz * sz + iy * sy + ix;
                    y[idx] = y[idx] + a * x[idx];
                }
}

/* integral_stage divergence guard (rhs_ACM.f90:133-146): max |u| over interior, all components */
double orc_max_abs_interior(int dim, int g, const int32_t Bs[3], int nc, const double *u)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = (ptrdiff_t)nx * ny * nz;
    const int z0 = dim == 3 ? g : 0, z1 = dim == 3 ? Bs[2] + g : 1;
    double m = 0.0;
    for (int c = 0; c < nc; ++c)
        for (int iz = z0; iz < z1; ++iz)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                for (int ix = g; ix < Bs[0] + g; ++ix) {
                    const double a = fabs(u[c * sc + iz * sz + iy * sy + ix]);
                    if (a > m) m = a;
                }
    return m;
}
