/*
 * oracle/orc_wavelet.c -- CPU restatement of WABBIT's per-block wavelet numerics (TEST INFRASTRUCTURE ONLY; see orc_acm.c).
 *
 * Reference:
 *   setup_wavelet                          LIB/WAVELETS/module_wavelets.f90:1031-1417   (filter banks HD/GD/HR/GR, CE sizes)
 *   waveletDecomposition_optimized_block   LIB/WAVELETS/wavelet_decomposition_reconstruction.f90:23-389
 *   waveletReconstruction_optimized_block  LIB/WAVELETS/wavelet_decomposition_reconstruction.f90:426-840
 *   wavelet_renorm_block                   LIB/WAVELETS/module_wavelets.f90:1848-1960
 *   threshold_block                        LIB/INDICATORS/threshold_block.f90:1-130
 *   prediction                             LIB/WAVELETS/module_wavelets.f90:96-284
 *   blockFilterXYZ_vct                     LIB/WAVELETS/module_wavelets.f90:307-401
 *   refineBlock                            LIB/MESH/refinementExecute.f90:1-80
 *   componentWiseNorm_tree (Linfty, L2)    LIB/OPERATORS/componentWiseNorm_tree.f90:63-197
 *
 * Layout: Fortran column-major ghosted blocks u(nx,ny,nz,nc), nx = Bs+2g (nz = 1 in 2-D); 0-based indices here.
 * Every filter is evaluated as the reference writes it: one product per non-zero tap, summed in increasing tap order,
 * no FMA (-ffp-contract=off).  The hard-coded and the generic `sum()` branches of the reference coincide under this
 * rule (zero taps add an exact 0).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_FMAX 12   /* largest |tap index| supported (CDF66: 10) */

typedef struct orc_wavelet {
    int32_t X, Y;
    int32_t hd_lo, hd_hi, gd_lo, gd_hi, hr_lo, hr_hi, gr_lo, gr_hi;
    int32_t g_default, lifted;
    int32_t Nscl, Nscr, Nwcl, Nwcr, Nreconl, Nreconr;
    double HD[2 * ORC_FMAX + 1], GD[2 * ORC_FMAX + 1], HR[2 * ORC_FMAX + 1], GR[2 * ORC_FMAX + 1]; /* index tap+ORC_FMAX */
} orc_wavelet;

/* interpolation stencils, module_wavelets.f90:14-20 */
static void stencil_int(int order, double *s /* centred at index order-1, length 2*order-1 */)
{
    int n = 2 * order - 1;
    for (int i = 0; i < n; ++i) s[i] = 0.0;
    if (order == 2) { s[0] = 1.0 / 2.0; s[1] = 2.0 / 2.0; s[2] = 1.0 / 2.0; }
    else if (order == 4) {
        const double v[7] = {-1.0, 0.0, 9.0, 16.0, 9.0, 0.0, -1.0};
        for (int i = 0; i < 7; ++i) s[i] = v[i] / 16.0;
    } else if (order == 6) {
        const double v[11] = {3.0, 0.0, -25.0, 0.0, 150.0, 256.0, 150.0, 0.0, -25.0, 0.0, 3.0};
        for (int i = 0; i < 11; ++i) s[i] = v[i] / 256.0;
    }
}

/* setup_wavelet for CDFXY with X in {2,4,6}, Y in {0,2,4,6}, X >= Y  (module_wavelets.f90:1155-1290) */
int orc_setup_wavelet(const char *name, orc_wavelet *w)
{
    memset(w, 0, sizeof(*w));
    if (strlen(name) != 5 || strncmp(name, "CDF", 3) != 0) return 1;
    const int X = name[3] - '0', Y = name[4] - '0';
    if ((X != 2 && X != 4 && X != 6) || (Y != 0 && Y != 2 && Y != 4 && Y != 6) || Y > X) return 2;
    w->X = X; w->Y = Y; w->lifted = Y != 0;
    double hr[23], hn[23];
    stencil_int(X, hr);
    w->hr_lo = -(X - 1); w->hr_hi = X - 1;
    for (int i = w->hr_lo; i <= w->hr_hi; ++i) w->HR[i + ORC_FMAX] = hr[i + X - 1];
    if (Y == 0) {
        w->hd_lo = w->hd_hi = 0;
        w->HD[ORC_FMAX] = 1.0;
    } else {
        stencil_int(Y, hn);
        hn[Y - 1] = 0.0;                                    /* h_ntilde(0) = 0 */
        w->hd_lo = w->hr_lo - (Y - 1); w->hd_hi = w->hr_hi + (Y - 1);
        for (int i = w->hd_lo; i <= w->hd_hi; ++i) {
            double v = (i == 0) ? 1.0 : 0.0;
            for (int j = w->hr_lo; j <= w->hr_hi; ++j) {
                if (i - j < -(Y - 1) || i - j > Y - 1) continue;
                const double sgn = (j % 2 == 0) ? 1.0 : -1.0;   /* (-1)**j */
                v = v + sgn * w->HR[j + ORC_FMAX] * hn[i - j + Y - 1] / 2.0;
            }
            w->HD[i + ORC_FMAX] = v;
        }
    }
    w->gd_lo = w->hr_lo; w->gd_hi = w->hr_hi;
    for (int i = w->gd_lo; i <= w->gd_hi; ++i) w->GD[i + ORC_FMAX] = ((i % 2 == 0) ? 1.0 : -1.0) * w->HR[i + ORC_FMAX];
    w->gr_lo = w->hd_lo; w->gr_hi = w->hd_hi;
    for (int i = w->gr_lo; i <= w->gr_hi; ++i) w->GR[i + ORC_FMAX] = ((i % 2 == 0) ? 1.0 : -1.0) * w->HD[i + ORC_FMAX];
    w->g_default = X - 1 + (Y - 1 > 0 ? Y - 1 : 0);       /* ini_file_to_params.f90:467 */
    /* coarse-extension sizes before the FD widening, module_wavelets.f90:1368-1384 */
    w->Nscl = abs(w->hd_lo) - 1 > 0 ? abs(w->hd_lo) - 1 : 0;
    w->Nwcl = w->Nscl + abs(w->gd_lo);
    w->Nreconl = w->Nwcl + abs(w->gr_lo);
    w->Nscr = w->hd_hi;
    w->Nwcr = w->Nscr + w->gd_hi;
    w->Nreconr = w->Nwcr + w->gr_hi;
    return 0;
}

static inline double filt(const double *u, ptrdiff_t stride, const double *F, int lo, int hi)
{
    /* sum over non-zero taps, increasing index, product-then-add */
    double acc = 0.0;
    int first = 1;
    for (int k = lo; k <= hi; ++k) {
        const double c = F[k + ORC_FMAX];
        if (c == 0.0) continue;
        const double p = u[k * stride] * c;
        acc = first ? p : acc + p;
        first = 0;
    }
    return acc;
}

/*
 * blockFilterXYZ_vct (LIB/WAVELETS/module_wavelets.f90:307-401): separable filter of a ghosted block, x then y then z; with
 * do_restriction only every second interior point (Fortran g+1, g+3, ...) is filtered in a direction, and the later passes read
 * only those.  The sum starts from 0 and adds u(i+shift)*c(shift) for EVERY shift fl_l..fl_r in increasing order (zero taps
 * included), as the reference loop does.  u and uf are different arrays [nc][nz][ny][nx]; everything that is not filtered is copied.
 */
void orc_block_filter(int dim, int g, const int32_t Bs[3], int nc, const double *u, double *uf, const double *coef, int fl_l, int fl_r,
                      int do_restriction)
{
    const int n[3] = {Bs[0] + 2 * g, Bs[1] + 2 * g, dim == 3 ? Bs[2] + 2 * g : 1};
    const ptrdiff_t sy = n[0], sz = (ptrdiff_t)n[0] * n[1], sc = sz * n[2];
    memcpy(uf, u, sizeof(double) * (size_t)sc * nc);
    if (fl_l == 0 && fl_r == 0 && fabs(coef[ORC_FMAX] - 1.0) <= 1.0e-10) return;
    int ifs[3], ife[3], ils[3], ile[3];
    for (int d = 0; d < 3; ++d) {
        ifs[d] = g + fl_l; ife[d] = Bs[d] + g - 1 + fl_r;    /* 0-based */
        ils[d] = g; ile[d] = Bs[d] + g - 1;
    }
    if (dim == 2) ifs[2] = ife[2] = ils[2] = ile[2] = 0;
    const int s = do_restriction ? 2 : 1;
    double *tmp = (double *)malloc(sizeof(double) * (size_t)sc);
    for (int c = 0; c < nc; ++c) {
        double *o = uf + c * sc;
        memcpy(tmp, o, sizeof(double) * (size_t)sc);
        for (int iz = ifs[2]; iz <= ife[2]; ++iz)
            for (int iy = ifs[1]; iy <= ife[1]; ++iy)
                for (int ix = ils[0]; ix <= ile[0]; ix += s) {
                    double acc = 0.0;
                    for (int k = fl_l; k <= fl_r; ++k) acc = acc + tmp[iz * sz + iy * sy + ix + k] * coef[k + ORC_FMAX];
                    o[iz * sz + iy * sy + ix] = acc;
                }
        memcpy(tmp, o, sizeof(double) * (size_t)sc);
        for (int iz = ifs[2]; iz <= ife[2]; ++iz)
            for (int iy = ils[1]; iy <= ile[1]; iy += s)
                for (int ix = ils[0]; ix <= ile[0]; ix += s) {
                    double acc = 0.0;
                    for (int k = fl_l; k <= fl_r; ++k) acc = acc + tmp[iz * sz + (iy + k) * sy + ix] * coef[k + ORC_FMAX];
                    o[iz * sz + iy * sy + ix] = acc;
                }
        if (dim == 3) {
            memcpy(tmp, o, sizeof(double) * (size_t)sc);
            for (int iz = ils[2]; iz <= ile[2]; iz += s)
                for (int iy = ils[1]; iy <= ile[1]; iy += s)
                    for (int ix = ils[0]; ix <= ile[0]; ix += s) {
                        double acc = 0.0;
                        for (int k = fl_l; k <= fl_r; ++k) acc = acc + tmp[(iz + k) * sz + iy * sy + ix] * coef[k + ORC_FMAX];
                        o[iz * sz + iy * sy + ix] = acc;
                    }
        }
    }
    free(tmp);
}

/*
 * waveletDecomposition_optimized_block: u (ghosts synchronised to depth >= filter size) -> u_d, spaghetti order:
 * SC at interior offsets 0,2,4,... (Fortran g+1, g+3, ...), WC at 1,3,5,...  u_d must be a different array.
 * Only the interior of u_d is meaningful on return (as in the reference).
 */
void orc_fwt_block(const orc_wavelet *w, int dim, int g, const int32_t Bs[3], int nc, const double *u, double *u_d)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = sz * nz;
    int f = 0;
    if (-w->hd_lo > f) f = -w->hd_lo;
    if (w->hd_hi > f) f = w->hd_hi;
    if (-w->gd_lo > f) f = -w->gd_lo;
    if (w->gd_hi > f) f = w->gd_hi;
    const int fz = dim == 3 ? f : 0, gz = dim == 3 ? g : 0, Bz = dim == 3 ? Bs[2] : 1;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz)));
    for (int c = 0; c < nc; ++c) {
        const double *uc = u + c * sc;
        double *dc = u_d + c * sc;
        memcpy(dc, uc, sizeof(double) * (size_t)sc);
        /* X */
        for (int iz = gz - fz; iz < Bz + gz + fz; ++iz)
            for (int iy = g - f; iy < Bs[1] + g + f; ++iy) {
                const double *row = uc + iz * sz + iy * sy;
                for (int ix = g; ix < Bs[0] + g; ix += 2) {
                    buf[ix] = filt(row + ix, 1, w->HD, w->hd_lo, w->hd_hi);
                    buf[ix + 1] = filt(row + ix + 1, 1, w->GD, w->gd_lo, w->gd_hi);
                }
                memcpy(dc + iz * sz + iy * sy + g, buf + g, sizeof(double) * (size_t)Bs[0]);
            }
        /* Y */
        for (int iz = gz - fz; iz < Bz + gz + fz; ++iz)
            for (int ix = g; ix < Bs[0] + g; ++ix) {
                double *col = dc + iz * sz + ix;
                for (int iy = g; iy < Bs[1] + g; iy += 2) {
                    buf[iy] = filt(col + iy * sy, sy, w->HD, w->hd_lo, w->hd_hi);
                    buf[iy + 1] = filt(col + (iy + 1) * sy, sy, w->GD, w->gd_lo, w->gd_hi);
                }
                for (int iy = g; iy < Bs[1] + g; ++iy) col[iy * sy] = buf[iy];
            }
        /* Z */
        if (dim == 3)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                for (int ix = g; ix < Bs[0] + g; ++ix) {
                    double *col = dc + iy * sy + ix;
                    for (int iz = g; iz < Bs[2] + g; iz += 2) {
                        buf[iz] = filt(col + iz * sz, sz, w->HD, w->hd_lo, w->hd_hi);
                        buf[iz + 1] = filt(col + (iz + 1) * sz, sz, w->GD, w->gd_lo, w->gd_hi);
                    }
                    for (int iz = g; iz < Bs[2] + g; ++iz) col[iz * sz] = buf[iz];
                }
    }
    free(buf);
}

/*
 * waveletReconstruction_optimized_block, generic branch (wavelet_decomposition_reconstruction.f90:765-835):
 *   u_r(i) = sum_k SCstuffed(i+k) HR(k)  +  sum_k WCstuffed(i+k) GR(k)
 * with SC/WC taken from the ghosted spaghetti array (ghosts synchronised).  The reference's hard-coded branches group
 * the same terms differently; results agree to round-off (fields are compared at 1e-12, never bit-wise).
 */
void orc_iwt_block(const orc_wavelet *w, int dim, int g, const int32_t Bs[3], int nc, const double *u, double *u_r)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = sz * nz;
    int f = 0;
    if (-w->hr_lo > f) f = -w->hr_lo;
    if (w->hr_hi > f) f = w->hr_hi;
    if (-w->gr_lo > f) f = -w->gr_lo;
    if (w->gr_hi > f) f = w->gr_hi;
    const int fz = dim == 3 ? f : 0, gz = dim == 3 ? g : 0, Bz = dim == 3 ? Bs[2] : 1;
    const int io = g % 2;   /* parity (0-based) of SC positions in the ghosted array: index g, g+2, ... */
    const int nmax = nx > ny ? (nx > nz ? nx : nz) : (ny > nz ? ny : nz);
    double *bs = (double *)calloc((size_t)nmax + 2 * ORC_FMAX + 2, sizeof(double)) + ORC_FMAX;
    double *bw = (double *)calloc((size_t)nmax + 2 * ORC_FMAX + 2, sizeof(double)) + ORC_FMAX;
    double *bo = (double *)malloc(sizeof(double) * (size_t)nmax);
    for (int c = 0; c < nc; ++c) {
        double *rc = u_r + c * sc;
        memcpy(rc, u + c * sc, sizeof(double) * (size_t)sc);
#define LINE(N, B, GG, PTR, STRIDE)                                                                  \
        do {                                                                                         \
            for (int i = 0; i < (N); ++i) {                                                          \
                const double v = (PTR)[(ptrdiff_t)i * (STRIDE)];                                     \
                const int is_sc = ((i % 2) == io);                                                   \
                bs[i] = is_sc ? v : 0.0;                                                             \
                bw[i] = is_sc ? 0.0 : v;                                                             \
            }                                                                                        \
            for (int i = (GG); i < (B) + (GG); ++i) {                                                \
                double a = 0.0, b2 = 0.0;                                                            \
                for (int k = w->hr_lo; k <= w->hr_hi; ++k) a = a + bs[i + k] * w->HR[k + ORC_FMAX];  \
                for (int k = w->gr_lo; k <= w->gr_hi; ++k) b2 = b2 + bw[i + k] * w->GR[k + ORC_FMAX];\
                bo[i] = a + b2;                                                                      \
            }                                                                                        \
            for (int i = (GG); i < (B) + (GG); ++i) (PTR)[(ptrdiff_t)i * (STRIDE)] = bo[i];          \
        } while (0)
        for (int iz = gz - fz; iz < Bz + gz + fz; ++iz)
            for (int iy = g - f; iy < Bs[1] + g + f; ++iy) LINE(nx, Bs[0], g, rc + iz * sz + iy * sy, 1);
        for (int iz = gz - fz; iz < Bz + gz + fz; ++iz)
            for (int ix = g; ix < Bs[0] + g; ++ix) LINE(ny, Bs[1], g, rc + iz * sz + ix, sy);
        if (dim == 3)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                for (int ix = g; ix < Bs[0] + g; ++ix) LINE(nz, Bs[2], g, rc + iy * sy + ix, sz);
#undef LINE
    }
    free(bs - ORC_FMAX);
    free(bw - ORC_FMAX);
    free(bo);
}

/*
 * wavelet_renorm_block + threshold_block (input_is_WD = true, full interior): detail[c] per component.
 * eps_norm: 0 Linfty, 1 L1, 2 L2, 3 H1.  thresh_comp[c]: 0 ignore, 1 own max-norm, >=2 joint group
 * (maxval(sqrt(x**2)) over the group's components).  Returns refinement status: -1 iff all(detail <= eps*norm).
 */
int orc_threshold_block_box(int dim, int g, const int32_t Bs[3], int nc, const double *u_wd, int level, int level_ref, int eps_norm,
                            const int32_t *thresh_comp, const double *eps, const double *norm, double *detail, const int32_t lo[3],
                            const int32_t hi[3]);

int orc_threshold_block(int dim, int g, const int32_t Bs[3], int nc, const double *u_wd, int level, int level_ref, int eps_norm,
                        const int32_t *thresh_comp, const double *eps, const double *norm, double *detail)
{
    const int32_t lo[3] = {0, 0, 0}, hi[3] = {Bs[0] - 1, Bs[1] - 1, dim == 3 ? Bs[2] - 1 : 0};
    return orc_threshold_block_box(dim, g, Bs, nc, u_wd, level, level_ref, eps_norm, thresh_comp, eps, norm, detail, lo, hi);
}

/* the same restricted to the box lo..hi (inclusive, 0-based interior offsets): threshold_block / wavelet_renorm_block with `indices`
 * (threshold_block.f90:30-44, module_wavelets.f90:1876-1905), as addSecurityZone_CE_tree calls them (securityZone_tree.f90:185-200) */
int orc_threshold_block_box(int dim, int g, const int32_t Bs[3], int nc, const double *u_wd, int level, int level_ref, int eps_norm,
                            const int32_t *thresh_comp, const double *eps, const double *norm, double *detail, const int32_t lo[3],
                            const int32_t hi[3])
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = sz * nz;
    const int gz = dim == 3 ? g : 0, Bz = dim == 3 ? Bs[2] : 1;
    double *own = (double *)malloc(sizeof(double) * (size_t)nc), *sq = (double *)malloc(sizeof(double) * (size_t)nc);
    for (int c = 0; c < nc; ++c) { own[c] = -INFINITY; sq[c] = -INFINITY; }
    double fac = 1.0, fdir = 1.0;
    int fdir_div = 0;
    if (eps_norm == 1) { fac = pow(2.0, (double)((level_ref - level - 1) * dim)); fdir = 4.0; fdir_div = 1; }
    if (eps_norm == 2) { fac = pow(2.0, (double)((level_ref - level - 1) * dim) / 2.0); fdir = 2.0; fdir_div = 1; }
    if (eps_norm == 3 && dim == 3) { fac = pow(2.0, (double)(level_ref - level) * (2.0 - dim) / 2.0); fdir = pow(2.0, 2.0 * (dim - 2.0) / 3.0); }
    (void)Bz;
    for (int c = 0; c < nc; ++c)
        for (int iz = lo[2]; iz <= (dim == 3 ? hi[2] : 0); ++iz)
            for (int iy = lo[1]; iy <= hi[1]; ++iy)
                for (int ix = lo[0]; ix <= hi[0]; ++ix) {
                    const int px = ix % 2 == 0, py = iy % 2 == 0, pz = dim == 3 ? iz % 2 == 0 : 1;
                    double v = u_wd[c * sc + (iz + gz) * sz + (iy + g) * sy + (ix + g)];
                    if (px && py && pz) v = 0.0;                       /* pure scaling coefficients are removed */
                    if (eps_norm != 0 && !(eps_norm == 3 && dim != 3)) {
                        v = v * fac;
                        if (px) v = fdir_div ? v / fdir : v * fdir;
                        if (py) v = fdir_div ? v / fdir : v * fdir;
                        if (dim == 3 && pz) v = fdir_div ? v / fdir : v * fdir;
                    }
                    const double a = fabs(v), s = sqrt(v * v);
                    if (a > own[c]) own[c] = a;
                    if (s > sq[c]) sq[c] = s;
                }
    int maxgrp = 0;
    for (int c = 0; c < nc; ++c) { detail[c] = -1.0; if (thresh_comp[c] > maxgrp) maxgrp = thresh_comp[c]; }
    for (int l = 2; l <= maxgrp; ++l) {
        double m = -INFINITY;
        for (int c = 0; c < nc; ++c) if (thresh_comp[c] == l && sq[c] > m) m = sq[c];
        for (int c = 0; c < nc; ++c) if (thresh_comp[c] == l) detail[c] = m;
    }
    for (int c = 0; c < nc; ++c) {
        if (thresh_comp[c] == 1) detail[c] = own[c];
        if (thresh_comp[c] == 0) detail[c] = 0.0;
    }
    int status = -1;
    for (int c = 0; c < nc; ++c) {
        const double e = norm ? eps[c] * norm[c] : eps[c];
        if (!(detail[c] <= e)) status = 0;
    }
    free(own);
    free(sq);
    return status;
}

/* prediction (module_wavelets.f90:96-284): coarse(n) -> fine(2n-1), order 2/4/6; points without a full stencil are 0 */
void orc_prediction(int order, int ncx, int ncy, int ncz, const double *coarse, double *fine)
{
    const int nfx = 2 * ncx - 1, nfy = 2 * ncy - 1, nfz = 2 * ncz - 1;
    const ptrdiff_t fy = nfx, fz = (ptrdiff_t)nfx * nfy;
    double c[6];
    int n = order;
    if (order == 2) { c[0] = 0.5; c[1] = 0.5; }
    else if (order == 4) { const double v[4] = {-1.0, 9.0, 9.0, -1.0}; for (int i = 0; i < 4; ++i) c[i] = v[i] / 16.0; }
    else { const double v[6] = {3.0, -25.0, 150.0, 150.0, -25.0, 3.0}; for (int i = 0; i < 6; ++i) c[i] = v[i] / 256.0; }
    memset(fine, 0, sizeof(double) * (size_t)nfx * nfy * nfz);
    for (int k = 0; k < ncz; ++k)
        for (int j = 0; j < ncy; ++j)
            for (int i = 0; i < ncx; ++i) fine[2 * k * fz + 2 * j * fy + 2 * i] = coarse[((ptrdiff_t)k * ncy + j) * ncx + i];
#define INTERP(P, S) ({ double a_ = c[0] * (P)[-(n - 1) * (S)]; for (int t_ = 1; t_ < n; ++t_) a_ = a_ + c[t_] * (P)[(-(n - 1) + 2 * t_) * (S)]; a_; })
    for (int k = 0; k < nfz; k += 2) {
        for (int j = 0; j < nfy; j += 2)
            for (int i = n - 1; i <= nfx - n; i += 2) { double *p = fine + k * fz + j * fy + i; *p = INTERP(p, 1); }
        for (int j = n - 1; j <= nfy - n; j += 2)
            for (int i = 0; i < nfx; ++i) { double *p = fine + k * fz + j * fy + i; *p = INTERP(p, fy); }
    }
    for (int k = n - 1; k <= nfz - n; k += 2)
        for (int j = 0; j < nfy; ++j)
            for (int i = 0; i < nfx; ++i) { double *p = fine + k * fz + j * fy + i; *p = INTERP(p, fz); }
#undef INTERP
}

/* refineBlock (refinementExecute.f90:1-80): ghosted mother (ghosts synchronised) -> 2^dim ghosted daughters,
 * daughter k: bit0 -> y, bit1 -> x, bit2 -> z.  daughters: [2^dim][nc][nz][ny][nx]. */
void orc_refine_block(int order, int dim, int g, const int32_t Bs[3], int nc, const double *mother, double *daughters)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const int nfx = 2 * nx - 1, nfy = 2 * ny - 1, nfz = 2 * nz - 1;
    const ptrdiff_t sc = (ptrdiff_t)nx * ny * nz;
    double *fine = (double *)malloc(sizeof(double) * (size_t)nfx * nfy * nfz);
    const int nd = 1 << dim;
    for (int c = 0; c < nc; ++c) {
        orc_prediction(order, nx, ny, nz, mother + c * sc, fine);
        for (int k = 0; k < nd; ++k) {
            const int ox = g + Bs[0] * ((k / 2) % 2), oy = g + Bs[1] * (k % 2), oz = dim == 3 ? g + Bs[2] * (k / 4) : 0;
            double *d = daughters + ((ptrdiff_t)k * nc + c) * sc;
            for (int z = 0; z < nz; ++z)
                for (int y = 0; y < ny; ++y)
                    memcpy(d + ((ptrdiff_t)z * ny + y) * nx, fine + ((ptrdiff_t)(z + oz) * nfy + (y + oy)) * nfx + ox, sizeof(double) * (size_t)nx);
        }
    }
    free(fine);
}

/* componentWiseNorm_tree for one block (componentWiseNorm_tree.f90:63-197): running max |u| (Linfty) per component */
void orc_block_linfty(int dim, int g, const int32_t Bs[3], int nc, const double *u, double *norm_inout)
{
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = sz * nz;
    const int gz = dim == 3 ? g : 0, Bz = dim == 3 ? Bs[2] : 1;
    for (int c = 0; c < nc; ++c)
        for (int iz = gz; iz < Bz + gz; ++iz)
            for (int iy = g; iy < Bs[1] + g; ++iy)
                for (int ix = g; ix < Bs[0] + g; ++ix) {
                    const double a = fabs(u[c * sc + iz * sz + iy * sy + ix]);
                    if (a > norm_inout[c]) norm_inout[c] = a;
                }
}
