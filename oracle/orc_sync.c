/*
 * oracle/orc_sync.c -- CPU restatement of WABBIT's ghost-node synchronisation with level jumps
 * (TEST INFRASTRUCTURE ONLY; see orc_acm.c).
 *
 * Reference:
 *   get_indices_of_modify_patch   LIB/TREE/neighborhood.f90:23-137
 *   get_indices_of_ghost_patch    LIB/TREE/neighborhood.f90:158-331
 *   inverse_relation              LIB/TREE/neighborhood.f90:347-381
 *   set_send_bounds / set_recv_bounds   LIB/MPI/calc_data_bounds.f90:46-189
 *   ghosts_setup_patches          LIB/MPI/module_mpi.f90:318-396
 *   prepare_ghost_synch_metadata  LIB/MPI/synchronize_ghosts_generic.f90:352-694   (sync_case "full_leaf" on a leaf grid)
 *   unpack_ghostlayers_internal   LIB/MPI/xfer_block_data.f90:321-440
 *   restrict_data / predict_data  LIB/MPI/restrict_predict_data.f90:45-202        (ignore_Filter = .true.: plain decimation)
 *
 * All index arithmetic keeps the reference's 1-based, inclusive convention (idx[0]=lower, idx[1]=upper per dimension).
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void orc_prediction(int order, int ncx, int ncy, int ncz, const double *coarse, double *fine);

static int in_list(int v, const int *lst, int n)
{
    for (int i = 0; i < n; ++i)
        if (lst[i] == v) return 1;
    return 0;
}
#define IN(v, ...) in_list((v), (const int[]){__VA_ARGS__}, (int)(sizeof((const int[]){__VA_ARGS__}) / sizeof(int)))

/* idx[0][d] lower, idx[1][d] upper */
void orc_get_indices_of_modify_patch(int g, int dim, int relation, int idx[2][3], const int N_xyz[3], const int N_s[3], const int N_e[3],
                                     const int g_m[3], const int g_p[3], int lvl_diff)
{
    (void)g;
    int r = relation > 56 ? (relation - 1) % 56 + 1 : relation;
    for (int d = 0; d < 3; ++d) { idx[0][d] = 1; idx[1][d] = 1; }
    for (int d = 0; d < dim; ++d) { idx[0][d] = 1 + g_m[d]; idx[1][d] = N_xyz[d] - g_p[d]; }
    if (IN(r, 1, 2, 3, 4, 25, 26, 29, 30, 33, 34, 37, 38, 49, 51, 53, 55)) idx[1][0] = N_s[0] + g_m[0];
    if (IN(r, 5, 6, 7, 8, 27, 28, 31, 32, 35, 36, 39, 40, 50, 52, 54, 56)) idx[0][0] = N_xyz[0] - N_e[0] - g_p[0] + 1;
    if (IN(r, 9, 10, 11, 12, 25, 26, 27, 28, 41, 42, 45, 46, 49, 50, 53, 54)) idx[1][1] = N_s[1] + g_m[1];
    if (IN(r, 13, 14, 15, 16, 29, 30, 31, 32, 43, 44, 47, 48, 51, 52, 55, 56)) idx[0][1] = N_xyz[1] - N_e[1] - g_p[1] + 1;
    if (IN(r, 17, 18, 19, 20, 33, 34, 35, 36, 41, 42, 43, 44, 49, 50, 51, 52)) idx[1][2] = N_s[2] + g_m[2];
    if (IN(r, 21, 22, 23, 24, 37, 38, 39, 40, 45, 46, 47, 48, 53, 54, 55, 56)) idx[0][2] = N_xyz[2] - N_e[2] - g_p[2] + 1;
    if (lvl_diff == -1) {
        if (IN(r, 10, 12, 14, 16, 18, 20, 22, 24, 42, 44, 46, 48)) idx[0][0] = N_xyz[0] / 2 - N_e[0] + 2;
        if (IN(r, 9, 11, 13, 15, 17, 19, 21, 23, 41, 43, 45, 47)) idx[1][0] = N_xyz[0] / 2 + N_s[0];
        if (IN(r, 2, 4, 6, 8, 19, 20, 23, 24, 34, 36, 38, 40)) idx[0][1] = N_xyz[1] / 2 - N_e[1] + 2;
        if (IN(r, 1, 3, 5, 7, 17, 18, 21, 22, 33, 35, 37, 39)) idx[1][1] = N_xyz[1] / 2 + N_s[1];
        if (dim == 3) {
            if (IN(r, 3, 4, 7, 8, 11, 12, 15, 16, 26, 28, 30, 32)) idx[0][2] = N_xyz[2] / 2 - N_e[2] + 2;
            if (IN(r, 1, 2, 5, 6, 9, 10, 13, 14, 25, 27, 29, 31)) idx[1][2] = N_xyz[2] / 2 + N_s[2];
        }
    }
}

void orc_get_indices_of_ghost_patch(const int Bs[3], int g, int dim, int relation, int idx[2][3], int gminus, int gplus, int lvl_diff)
{
    int r = relation > 56 ? (relation - 1) % 56 + 1 : relation;
    for (int d = 0; d < 3; ++d) { idx[0][d] = 1; idx[1][d] = 1; }
    for (int d = 0; d < dim; ++d) { idx[0][d] = g + 1; idx[1][d] = g + Bs[d]; }
    if (IN(r, 1, 2, 3, 4, 25, 26, 29, 30, 33, 34, 37, 38, 49, 51, 53, 55)) { idx[0][0] = g - gminus + 1; idx[1][0] = g; }
    if (IN(r, 5, 6, 7, 8, 27, 28, 31, 32, 35, 36, 39, 40, 50, 52, 54, 56)) { idx[0][0] = Bs[0] + g + 1; idx[1][0] = Bs[0] + g + gplus; }
    if (IN(r, 9, 10, 11, 12, 25, 26, 27, 28, 41, 42, 45, 46, 49, 50, 53, 54)) { idx[0][1] = g - gminus + 1; idx[1][1] = g; }
    if (IN(r, 13, 14, 15, 16, 29, 30, 31, 32, 43, 44, 47, 48, 51, 52, 55, 56)) { idx[0][1] = Bs[1] + g + 1; idx[1][1] = Bs[1] + g + gplus; }
    if (IN(r, 17, 18, 19, 20, 33, 34, 35, 36, 41, 42, 43, 44, 49, 50, 51, 52)) { idx[0][2] = g - gminus + 1; idx[1][2] = g; }
    if (IN(r, 21, 22, 23, 24, 37, 38, 39, 40, 45, 46, 47, 48, 53, 54, 55, 56)) { idx[0][2] = Bs[2] + g + 1; idx[1][2] = Bs[2] + g + gplus; }
    if (IN(r, 10, 12, 14, 16, 18, 20, 22, 24, 42, 44, 46, 48)) {
        if (lvl_diff == +1) idx[0][0] = g - gminus + 1;
        else if (lvl_diff == -1) idx[0][0] = g + Bs[0] / 2 + 1;
        idx[1][0] = Bs[0] + g;
    }
    if (IN(r, 9, 11, 13, 15, 17, 19, 21, 23, 41, 43, 45, 47)) {
        idx[0][0] = g + 1;
        if (lvl_diff == +1) idx[1][0] = Bs[0] + g + gplus;
        else if (lvl_diff == -1) idx[1][0] = g + Bs[0] / 2;
    }
    if (IN(r, 2, 4, 6, 8, 19, 20, 23, 24, 34, 36, 38, 40)) {
        if (lvl_diff == +1) idx[0][1] = g - gminus + 1;
        else if (lvl_diff == -1) idx[0][1] = g + Bs[1] / 2 + 1;
        idx[1][1] = Bs[1] + g;
    }
    if (IN(r, 1, 3, 5, 7, 17, 18, 21, 22, 33, 35, 37, 39)) {
        idx[0][1] = g + 1;
        if (lvl_diff == +1) idx[1][1] = Bs[1] + g + gplus;
        else if (lvl_diff == -1) idx[1][1] = g + Bs[1] / 2;
    }
    if (dim == 3) {
        if (IN(r, 3, 4, 7, 8, 11, 12, 15, 16, 26, 28, 30, 32)) {
            if (lvl_diff == +1) idx[0][2] = g - gminus + 1;
            else if (lvl_diff == -1) idx[0][2] = g + Bs[2] / 2 + 1;
            idx[1][2] = Bs[2] + g;
        }
        if (IN(r, 1, 2, 5, 6, 9, 10, 13, 14, 25, 27, 29, 31)) {
            idx[0][2] = g + 1;
            if (lvl_diff == +1) idx[1][2] = Bs[2] + g + gplus;
            else if (lvl_diff == -1) idx[1][2] = g + Bs[2] / 2;
        }
    }
}

int orc_inverse_relation(int relation)
{
    if (relation == 0) return 0;
    if (relation < 0) return relation < -8 ? relation + 8 : relation - 8;
    int inv = (relation - 1) % 56 + 1;
    if (relation > 56) inv += (3 - (relation - 1) / 56) * 56;
    for (int i_dim = 3; i_dim >= 3 - ((relation - 1) % 56) / 24; --i_dim) {
        if ((relation - 1) % (1 << i_dim) >= (1 << (i_dim - 1))) inv -= 1 << (i_dim - 1);
        else inv += 1 << (i_dim - 1);
    }
    return inv;
}

/* a = (predictor order - 2)/2 */
void orc_set_send_bounds(const int Bs[3], int g, int dim, int a, int relation, int lvl_diff, int gminus, int gplus, int bounds[2][3],
                         int buffer[2][3])
{
    int n[3] = {1, 1, 1}, gv[3] = {g, g, g};
    for (int d = 0; d < dim; ++d) n[d] = Bs[d] + 2 * g;
    for (int d = 0; d < 3; ++d) { bounds[0][d] = bounds[1][d] = 1; buffer[0][d] = buffer[1][d] = 1; }
    if (lvl_diff == 0) {
        const int Ns[3] = {gminus, gminus, gminus}, Ne[3] = {gplus, gplus, gplus};
        orc_get_indices_of_modify_patch(g, dim, relation, bounds, n, Ns, Ne, gv, gv, lvl_diff);
    } else if (lvl_diff == +1) {
        const int Ns[3] = {gminus * 2 - 1, gminus * 2 - 1, gminus * 2 - 1}, Ne[3] = {gplus * 2 - 1, gplus * 2 - 1, gplus * 2 - 1};
        const int gp[3] = {g + 1, g + 1, g + 1};
        orc_get_indices_of_modify_patch(g, dim, relation, bounds, n, Ns, Ne, gv, gp, lvl_diff);
        for (int d = 0; d < dim; ++d) buffer[1][d] = (bounds[1][d] - bounds[0][d] + 1 + 1) / 2;
    } else {
        const int Ns[3] = {gminus / 2 + 1, gminus / 2 + 1, gminus / 2 + 1};
        const int Ne[3] = {(gplus + 1) / 2 + 1, (gplus + 1) / 2 + 1, (gplus + 1) / 2 + 1};
        const int gp[3] = {g - 1, g - 1, g - 1};
        orc_get_indices_of_modify_patch(g, dim, relation, bounds, n, Ns, Ne, gv, gp, lvl_diff);
        for (int d = 0; d < dim; ++d) { bounds[0][d] -= a; bounds[1][d] += a; }
        for (int d = 0; d < dim; ++d) {
            if (bounds[0][d] == g + 1 - a || gminus % 2 == 0) buffer[0][d] = 1 + a * 2;
            else buffer[0][d] = 2 + a * 2;
            buffer[1][d] = buffer[0][d] + gplus - 1;
            if (bounds[1][d] - bounds[0][d] > g + a) buffer[1][d] += Bs[d];
        }
    }
}

/*
 * sync_ghosts_generic, sync_case "full_leaf", ignore_Filter = .true., on a leaf-only grid held by one rank.
 * hvy_neighbor: [168][nb] (Fortran hvy_neighbor(nb,168)), 1-based block ids, -1 none.  level[nb].
 * hvy: [nb][nc][nz][ny][nx] ghosted.  order = predictor order (2,4,6).  lifted: 3 stages, else 2 (restriction in stage 1).
 * Returns the number of patches moved.
 */
void orc_block_filter(int dim, int g, const int32_t Bs[3], int nc, const double *u, double *uf, const double *coef, int fl_l, int fl_r,
                      int do_restriction);
void orc_ce_modify_block(int dim, int g, const int32_t Bs32[3], int nc, double *wd, const double *orig, int relation, int Nwcl, int Nwcr,
                         int Nscl, int Nscr, int clear_wc, int copy_sc);

/*
 * ignore_filter = 0 (sync_ghosts_tree's default) with a lifted wavelet: restrict_data takes the decimated values from the HD-filtered
 * sender block (restrict_copy_at_CE, LIB/MPI/restrict_predict_data.f90:121-172: blockFilterXYZ_vct on the whole block after stage 1,
 * then, for every relation whose neighbour is coarser or finer, the scaling positions of the Nscl / Nscr strip are copied back from
 * the unfiltered block).  coefHD[tap + 12], taps hd_lo..hd_hi.
 */
int orc_sync_ghosts_leaf_ex(int nb, const int32_t *hvy_neighbor, const int32_t *level, int dim, int g, const int32_t Bs32[3], int nc,
                            double *hvy, int gminus, int gplus, int order, int lifted, int ignore_filter, const double *coefHD, int hd_lo,
                            int hd_hi, int Nscl, int Nscr)
{
    const int Bs[3] = {Bs32[0], Bs32[1], dim == 3 ? Bs32[2] : 1};
    const int nx = Bs[0] + 2 * g, ny = Bs[1] + 2 * g, nz = dim == 3 ? Bs[2] + 2 * g : 1;
    const ptrdiff_t sy = nx, sz = (ptrdiff_t)nx * ny, sc = sz * nz, sb = sc * nc;
    const int a = (order - 2) / 2;
    const int isUnlifted = lifted ? 0 : 1, Nstages = lifted ? 3 : 2;
    /* ghosts_setup_patches */
    static int P_send[169][2][3], P_recv[169][2][3], P_buf[169][2][3];
    for (int r = 1; r <= 168; ++r) {
        const int ld = r <= 56 ? 0 : (r <= 112 ? +1 : -1);
        orc_get_indices_of_ghost_patch(Bs, g, dim, r, P_recv[r], gminus, gplus, ld);
        orc_set_send_bounds(Bs, g, dim, a, r, ld, gminus, gplus, P_send[r], P_buf[r]);
    }
    int moved = 0;
    double *res = NULL, *box = NULL;
    size_t res_cap = 0, box_cap = 0;
    const int use_filter = lifted && !ignore_filter;
    double *restricted = use_filter ? (double *)malloc(sizeof(double) * (size_t)sb) : NULL;
    int restricted_id = -1;   /* restricted_hvy_ID of the reference: one filtered block is held at a time */
    for (int istage = 1; istage <= Nstages; ++istage) {
        for (int k = 0; k < nb; ++k)
            for (int i_n = 1; i_n <= 168; ++i_n) {
                const int nbid = hvy_neighbor[(size_t)(i_n - 1) * nb + k];
                if (nbid < 1) continue;
                const int lvl_diff = level[k] - level[nbid - 1];
                /* prepare_ghost_synch_metadata, sync_id 2 on a leaf grid: no valid finer/same neighbour can shadow a slot */
                if (!((istage == 1 && lvl_diff == 0) || (istage == 2 - isUnlifted && lvl_diff == +1) || (istage == 3 - isUnlifted && lvl_diff == -1)))
                    continue;
                const int inv = orc_inverse_relation(i_n);
                int(*S)[3] = P_send[i_n], (*R)[3] = P_recv[inv], (*Bf)[3] = P_buf[i_n];
                double *recv = hvy + (ptrdiff_t)(nbid - 1) * sb;
                const double *send = hvy + (ptrdiff_t)k * sb;
                const int ex = R[1][0] - R[0][0] + 1, ey = R[1][1] - R[0][1] + 1, ez = R[1][2] - R[0][2] + 1;
                if (lvl_diff == 0) {
                    for (int c = 0; c < nc; ++c)
                        for (int z = 0; z < ez; ++z)
                            for (int y = 0; y < ey; ++y)
                                for (int x = 0; x < ex; ++x)
                                    recv[c * sc + (R[0][2] - 1 + z) * sz + (R[0][1] - 1 + y) * sy + (R[0][0] - 1 + x)] =
                                        send[c * sc + (S[0][2] - 1 + z) * sz + (S[0][1] - 1 + y) * sy + (S[0][0] - 1 + x)];
                } else if (lvl_diff == +1) {
                    /* restrict_data: res(1:(n+1)/2) = block(ijk1:ijk2:2) of the (filtered) sender; then recv = res(buffer) */
                    if (use_filter) {
                        if (restricted_id != k) {
                            orc_block_filter(dim, g, Bs32, nc, send, restricted, coefHD, hd_lo, hd_hi, 1);
                            for (int j_n = 1; j_n <= 168; ++j_n) {
                                const int nj = hvy_neighbor[(size_t)(j_n - 1) * nb + k];
                                if (nj < 1) continue;
                                const int ld = level[k] - level[nj - 1];
                                if (ld == -1 || ld == +1) orc_ce_modify_block(dim, g, Bs32, nc, restricted, send, j_n, 0, 0, Nscl, Nscr, 0, 1);
                            }
                            restricted_id = k;
                        }
                        send = restricted;
                    }
                    for (int c = 0; c < nc; ++c)
                        for (int z = 0; z < ez; ++z)
                            for (int y = 0; y < ey; ++y)
                                for (int x = 0; x < ex; ++x) {
                                    const int bx = Bf[0][0] - 1 + x, by = Bf[0][1] - 1 + y, bz = Bf[0][2] - 1 + z;
                                    recv[c * sc + (R[0][2] - 1 + z) * sz + (R[0][1] - 1 + y) * sy + (R[0][0] - 1 + x)] =
                                        send[c * sc + (S[0][2] - 1 + 2 * bz) * sz + (S[0][1] - 1 + 2 * by) * sy + (S[0][0] - 1 + 2 * bx)];
                                }
                } else {
                    /* predict_data: prediction of the sender box to 2n-1 points, then recv = pre(buffer) */
                    const int bxn = S[1][0] - S[0][0] + 1, byn = S[1][1] - S[0][1] + 1, bzn = S[1][2] - S[0][2] + 1;
                    const int fx = 2 * bxn - 1, fy = 2 * byn - 1, fz = 2 * bzn - 1;
                    if ((size_t)bxn * byn * bzn > box_cap) { box_cap = (size_t)bxn * byn * bzn; box = (double *)realloc(box, sizeof(double) * box_cap); }
                    if ((size_t)fx * fy * fz > res_cap) { res_cap = (size_t)fx * fy * fz; res = (double *)realloc(res, sizeof(double) * res_cap); }
                    for (int c = 0; c < nc; ++c) {
                        for (int z = 0; z < bzn; ++z)
                            for (int y = 0; y < byn; ++y)
                                for (int x = 0; x < bxn; ++x)
                                    box[((size_t)z * byn + y) * bxn + x] = send[c * sc + (S[0][2] - 1 + z) * sz + (S[0][1] - 1 + y) * sy + (S[0][0] - 1 + x)];
                        orc_prediction(order, bxn, byn, bzn, box, res);
                        for (int z = 0; z < ez; ++z)
                            for (int y = 0; y < ey; ++y)
                                for (int x = 0; x < ex; ++x)
                                    recv[c * sc + (R[0][2] - 1 + z) * sz + (R[0][1] - 1 + y) * sy + (R[0][0] - 1 + x)] =
                                        res[((size_t)(Bf[0][2] - 1 + z) * fy + (Bf[0][1] - 1 + y)) * fx + (Bf[0][0] - 1 + x)];
                    }
                }
                ++moved;
            }
    }
    free(res);
    free(box);
    free(restricted);
    return moved;
}

/* sync_ghosts_generic("full_leaf", ignore_Filter = .true.): sync_ghosts_RHS_tree */
int orc_sync_ghosts_leaf(int nb, const int32_t *hvy_neighbor, const int32_t *level, int dim, int g, const int32_t Bs32[3], int nc,
                         double *hvy, int gminus, int gplus, int order, int lifted)
{
    return orc_sync_ghosts_leaf_ex(nb, hvy_neighbor, level, dim, g, Bs32, nc, hvy, gminus, gplus, order, lifted, 1, NULL, 0, 0, 0, 0);
}

/*
 * Coarse extension on one decomposed block for one neighbour relation (the neighbour in that relation is coarser):
 * coarseExtensionManipulateWC_block + coarseExtensionManipulateSC_block (LIB/WAVELETS/module_wavelets.f90:877-959, 963-1027).
 * wd: [nc][nz][ny][nx] decomposed block in spaghetti order (modified), orig: the values the scaling coefficients are copied from.
 * WC: every coefficient that is not a pure scaling coefficient is set to 0 in the interior strip (Nwcl / Nwcr deep) AND in the
 * whole ghost patch of the relation; SC: the pure scaling positions of the interior strip (Nscl / Nscr deep) are copied from orig.
 */
void orc_ce_modify_block(int dim, int g, const int32_t Bs32[3], int nc, double *wd, const double *orig, int relation, int Nwcl, int Nwcr,
                         int Nscl, int Nscr, int clear_wc, int copy_sc)
{
    const int Bs[3] = {Bs32[0], Bs32[1], dim == 3 ? Bs32[2] : 1};
    const int n[3] = {Bs[0] + 2 * g, Bs[1] + 2 * g, dim == 3 ? Bs[2] + 2 * g : 1};
    const ptrdiff_t sy = n[0], sz = (ptrdiff_t)n[0] * n[1], sc = sz * n[2];
    const int gv[3] = {g, g, g};
    if (clear_wc) {
        for (int i_set = 1; i_set <= 2; ++i_set) {
            int idx[2][3];
            if (i_set == 1) {
                const int Ns[3] = {Nwcl, Nwcl, Nwcl}, Ne[3] = {Nwcr, Nwcr, Nwcr};
                orc_get_indices_of_modify_patch(g, dim, relation, idx, n, Ns, Ne, gv, gv, +1);
            } else orc_get_indices_of_ghost_patch(Bs, g, dim, relation, idx, g, g, +1);
            if (dim == 2) idx[0][2] = idx[1][2] = 1;
            for (int c = 0; c < nc; ++c)
                for (int k = idx[0][2]; k <= idx[1][2]; ++k)
                    for (int j = idx[0][1]; j <= idx[1][1]; ++j)
                        for (int i = idx[0][0]; i <= idx[1][0]; ++i) {
                            /* scaling positions are g+1, g+3, ... (1-based) in every direction */
                            const int scx = ((i - g) & 1) == 1, scy = ((j - g) & 1) == 1, scz = dim == 3 ? ((k - g) & 1) == 1 : 1;
                            if (!(scx && scy && scz)) wd[c * sc + (k - 1) * sz + (j - 1) * sy + (i - 1)] = 0.0;
                        }
        }
    }
    if (copy_sc) {
        int idx[2][3];
        const int Ns[3] = {Nscl, Nscl, Nscl}, Ne[3] = {Nscr, Nscr, Nscr};
        orc_get_indices_of_modify_patch(g, dim, relation, idx, n, Ns, Ne, gv, gv, +1);
        if (dim == 2) idx[0][2] = idx[1][2] = 1;
        for (int c = 0; c < nc; ++c)
            for (int k = idx[0][2]; k <= idx[1][2]; ++k)
                for (int j = idx[0][1]; j <= idx[1][1]; ++j)
                    for (int i = idx[0][0]; i <= idx[1][0]; ++i) {
                        const int scx = ((i - g) & 1) == 1, scy = ((j - g) & 1) == 1, scz = dim == 3 ? ((k - g) & 1) == 1 : 1;
                        if (scx && scy && scz) {
                            const ptrdiff_t o = c * sc + (k - 1) * sz + (j - 1) * sy + (i - 1);
                            wd[o] = orig[o];
                        }
                    }
    }
}
