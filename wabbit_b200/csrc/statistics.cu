// Device-side mask generation and statistics of the ACM physics module -- the "next" row (f2) of the hot-path scope: both run once per
// saved time step / per stage in the reference and read or write every block.
//
//   create_mask_kernel   create_mask_2D_ACM / create_mask_3D_ACM (LIB/EQUATION/ACMnew/create_mask.f90:6-320) for the closed-form geometries:
//                        geometry = cylinder / circle (2-D; draw_circle, LIB/EQUATION/insects/module_geometry.f90:315-381) with the p-norm sponge
//                        (sponge_2D, LIB/EQUATION/ACMnew/sponge.f90) and sphere (3-D; draw_sphere), cosine smoothing step_cosine4
//                        (LIB/HELPER/module_helpers.f90:456-470): the six components [chi, u_s(3), colour, sponge] of hvy_mask, interiors.
//   stats_block_kernel   the integral_stage of STATISTICS_ACM (LIB/EQUATION/ACMnew/statistics_ACM.f90:138-368), per block: mean flow, kinetic and
//                        ACM energy, max |u|^2, divergence extrema (outside the solid), mask / sponge volume, penalization power, residual
//                        velocity, force on colour 1 -- each block's sums times its dV;
//   vort_block_kernel    enstrophy, max |vorticity|, helicity and dissipation of the same routine (statistics_ACM.f90:371-387): compute_vorticity
//                        (LIB/OPERATORS/compute_vorticity.f90:3-67) and compute_dissipation (compute_dissipation.f90:5-78) with the first- and
//                        second-derivative stencils of the discretization (module_operators.f90:23-33) on ghosted copies of the blocks (the
//                        export kernels' staging layout: ghost nodes as sync_ghosts_tree leaves them, level jumps included);
//   stats_final_kernel   the sum / max / min over the blocks in list order (deterministic), the post_stage's MPI reductions follow in capi.cu.
#include <math.h>

#include "wgpu_internal.cuh"

namespace {

__device__ __forceinline__ double step_cosine(double x_rel, double h)
{
    if (x_rel <= -h) return 1.0;
    if (x_rel >= h) return 0.0;
    return 0.5 * (1.0 + cos((x_rel + h) * 3.14159265358979323846 / (2.0 * h)));
}

__global__ void __launch_bounds__(256) create_mask_kernel(double *__restrict__ mask, const int *__restrict__ active, const signed char *__restrict__ level,
                                                          const int *__restrict__ ixyz, int n_mask, int Bx, int By, int Bz, int dim, MaskGeom gm,
                                                          double time)
{
    const int b = active[blockIdx.x];
    const long long CS = (long long)Bx * By * Bz;
    const int lv = level[b];
    const double sc = ldexp(1.0, -lv);
    const double dx = sc * gm.domain[0] / (double)Bx, dy = sc * gm.domain[1] / (double)By, dz = dim == 3 ? sc * gm.domain[2] / (double)Bz : 0.0;
    const double x0 = (double)(ixyz[3 * b] * Bx) * dx, y0 = (double)(ixyz[3 * b + 1] * By) * dy, z0 = dim == 3 ? (double)(ixyz[3 * b + 2] * Bz) * dz : 0.0;
    const double cx = gm.c0[0] + gm.v[0] * time, cy = gm.c0[1] + gm.v[1] * time, cz = gm.c0[2] + gm.v[2] * time;
    double *m = mask + (long long)b * n_mask * CS;
    for (long long e = threadIdx.x; e < CS; e += blockDim.x) {
        const int ix = (int)(e % Bx), iy = (int)((e / Bx) % By), iz = (int)(e / ((long long)Bx * By));
        const double x = (double)ix * dx + x0, y = (double)iy * dy + y0, z = (double)iz * dz + z0;
        double chi = 0.0;
        if (gm.penalization) {
            double r2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
            if (dim == 3) r2 = r2 + (z - cz) * (z - cz);
            chi = step_cosine(sqrt(r2) - gm.R, gm.h);
        }
        m[e] = chi;
        m[CS + e] = gm.v[0];
        m[2 * CS + e] = gm.v[1];
        m[3 * CS + e] = dim == 3 ? gm.v[2] : 0.0;
        m[4 * CS + e] = 1.0;
        if (n_mask > 5) {
            double sp = 0.0;
            if (gm.use_sponge) {      // p-norm sponge: -( ((x-L/2)^p + (y-L/2)^p)^(1/p) - L/2 ), cosine ramp of width L_sponge
                const double off = 0.5 * gm.domain[0];
                double s = pow(x - off, gm.p_sponge) + pow(y - off, gm.p_sponge);
                if (dim == 3) s = s + pow(z - off, gm.p_sponge);
                const double tmp = -(pow(s, 1.0 / gm.p_sponge) - off);
                sp = step_cosine(tmp - 0.5 * gm.L_sponge, 0.5 * gm.L_sponge);
            }
            m[5 * CS + e] = sp;
        }
    }
}

__device__ __forceinline__ double block_sum(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    return s;
}
__device__ __forceinline__ double block_max(double v, double *red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) s = fmax(s, red[i]);
    return s;
}

// u: hvy_block; rhs: RHS of the same state (its pressure component carries the divergence: rhs_p = -c0^2 div(u) - gamma_p p - p chi_sp/C_sp,
// rhs_ACM.f90:1202,1752) or nullptr; out[blk][WGPU_NSTAT]
__global__ void __launch_bounds__(256) stats_block_kernel(const double *__restrict__ u, const double *__restrict__ rhs, const double *__restrict__ mask,
                                                          const int *__restrict__ active, const signed char *__restrict__ level, int nc, int n_mask,
                                                          int Bx, int By, int Bz, int dim, StatArgs sa, double *__restrict__ out)
{
    __shared__ double red[8];
    const int b = active[blockIdx.x];
    const long long CS = (long long)Bx * By * Bz;
    const double *ub = u + (long long)b * nc * CS;
    const double *rb = rhs ? rhs + (long long)b * nc * CS : nullptr;
    const double *mb = mask ? mask + (long long)b * n_mask * CS : nullptr;
    const double sc = ldexp(1.0, -(int)level[b]);
    double dV = (sc * sa.domain[0] / (double)Bx) * (sc * sa.domain[1] / (double)By);
    if (dim == 3) dV *= sc * sa.domain[2] / (double)Bz;
    double s[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) s[k] = 0.0;
    double umag = 0.0, dmax = -1.0e300, dmin = 1.0e300, res[3] = {0.0, 0.0, 0.0};
    const double c02 = sa.c0 * sa.c0;
    for (long long e = threadIdx.x; e < CS; e += blockDim.x) {
        const double v0 = ub[e], v1 = ub[CS + e], v2 = dim == 3 ? ub[2 * CS + e] : 0.0, p = ub[dim * CS + e];
        s[0] += v0;
        s[1] += v1;
        s[2] += v2;
        const double k2 = v0 * v0 + v1 * v1 + v2 * v2;
        s[3] += 0.5 * k2;
        s[4] += 0.5 * p * p / c02 + 0.5 * k2;
        umag = fmax(umag, k2);
        double chi = 0.0, us0 = 0.0, us1 = 0.0, us2 = 0.0, sp = 0.0;
        if (mb) {
            chi = mb[e];
            us0 = mb[CS + e];
            us1 = mb[2 * CS + e];
            us2 = dim == 3 ? mb[3 * CS + e] : 0.0;
            if (n_mask > 5 && sa.use_sponge) sp = mb[5 * CS + e];
        }
        if (rb) {
            double div = -(rb[dim * CS + e] + sa.gamma_p * p + p * sp * sa.C_sponge_inv) / c02;
            if (chi > 0.0) div = 0.0;                      // "mask divergence inside the solid body"
            dmax = fmax(dmax, div);
            dmin = fmin(dmin, div);
        }
        if (mb) {
            s[5] += chi;
            s[6] += sp;
            s[7] += (us0 * (v0 - us0) + us1 * (v1 - us1) + us2 * (v2 - us2)) * chi * sa.C_eta_inv;
            s[8] += ((v0 - us0) * (v0 - us0) + (v1 - us1) * (v1 - us1) + (v2 - us2) * (v2 - us2)) * chi * sa.C_eta_inv;
            s[9] += (v0 * (v0 - sa.u_mean_set[0]) + v1 * (v1 - sa.u_mean_set[1]) + v2 * (v2 - sa.u_mean_set[2]) + p * p / c02) * sp * sa.C_sponge_inv;
            s[10] += chi * (v0 - us0) * sa.C_eta_inv;      // force = - sum(penal), penal = -chi (u - u_s) / C_eta
            s[11] += chi * (v1 - us1) * sa.C_eta_inv;
            s[12] += chi * (v2 - us2) * sa.C_eta_inv;
            res[0] = fmax(res[0], fabs(v0 - us0) * chi);
            res[1] = fmax(res[1], fabs(v1 - us1) * chi);
            res[2] = fmax(res[2], fabs(v2 - us2) * chi);
        }
    }
    double *o = out + (long long)blockIdx.x * WGPU_NSTAT;
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        const double t = block_sum(s[k], red);
        if (threadIdx.x == 0) o[k] = t * dV;
    }
    const double t13 = block_max(umag, red), t14 = block_max(dmax, red), t15 = -block_max(-dmin, red);
    const double r0 = block_max(res[0], red), r1 = block_max(res[1], red), r2 = block_max(res[2], red);
    if (threadIdx.x == 0) {
        o[13] = t13;
        o[14] = t14;
        o[15] = t15;
        o[16] = r0 * dV;
        o[17] = r1 * dV;
        o[18] = r2 * dV;
        o[19] = o[20] = o[21] = o[22] = 0.0;               // vort_block_kernel's entries
    }
}

// staged: [k][ncomp][nz][ny][nx] ghosted copies (g ghost nodes, the innermost H synchronised) of the blocks ids[0..m); writes entries 19..22 of
// out[k0 + k][WGPU_NSTAT]: 0.5 sum |omega|^2 dV, max |omega|, 0.5 sum omega.u dV (3-D), -nu sum u.lap(u) dV (nu > 0)
__global__ void __launch_bounds__(256) vort_block_kernel(const double *__restrict__ staged, const int *__restrict__ ids, const signed char *__restrict__ level,
                                                         int ncomp, int Bx, int By, int Bz, int g, int dim, VortArgs va, double *__restrict__ out)
{
    __shared__ double red[8];
    const int k = blockIdx.x, b = ids[k];
    const int gz = dim == 3 ? g : 0;
    const int nx = Bx + 2 * g, ny = By + 2 * g, nz = Bz + 2 * gz;
    const long long GS = (long long)nx * ny * nz;
    const double *u0 = staged + (long long)k * ncomp * GS, *u1 = u0 + GS, *u2 = u0 + 2 * GS;
    const double sc = ldexp(1.0, -(int)level[b]);
    const double hx = sc * va.domain[0] / (double)Bx, hy = sc * va.domain[1] / (double)By, hz = dim == 3 ? sc * va.domain[2] / (double)Bz : 1.0;
    const double dV = dim == 3 ? hx * hy * hz : hx * hy;
    const double dx_inv = 1.0 / hx, dy_inv = 1.0 / hy, dz_inv = 1.0 / hz;
    const double dx2_inv = 1.0 / (hx * hx), dy2_inv = 1.0 / (hy * hy), dz2_inv = 1.0 / (hz * hz);
    const int H = va.H;
    // sum(FD(s:e) * u(i+s:i+e)): every tap, from 0 in increasing tap order, no contraction
    auto fd = [&](const double *cf, const double *q, long long st) -> double {
        double s = 0.0;
        for (int t = -H; t <= H; ++t) s = __dadd_rn(s, __dmul_rn(cf[t + H], q[t * st]));
        return s;
    };
    double enst = 0.0, vmax = 0.0, hel = 0.0, dis = 0.0;
    const long long CS = (long long)Bx * By * Bz;
    for (long long e = threadIdx.x; e < CS; e += blockDim.x) {
        const int ix = (int)(e % Bx), iy = (int)((e / Bx) % By), iz = (int)(e / ((long long)Bx * By));
        const long long i = ((long long)(iz + gz) * ny + (iy + g)) * nx + (ix + g);
        const long long sy = nx, sz = (long long)nx * ny;
        const double u_dy = __dmul_rn(fd(va.fd1, u0 + i, sy), dy_inv), v_dx = __dmul_rn(fd(va.fd1, u1 + i, 1), dx_inv);
        if (dim == 3) {
            const double u_dz = __dmul_rn(fd(va.fd1, u0 + i, sz), dz_inv), v_dz = __dmul_rn(fd(va.fd1, u1 + i, sz), dz_inv);
            const double w_dx = __dmul_rn(fd(va.fd1, u2 + i, 1), dx_inv), w_dy = __dmul_rn(fd(va.fd1, u2 + i, sy), dy_inv);
            const double o0 = w_dy - v_dz, o1 = u_dz - w_dx, o2 = v_dx - u_dy;
            const double m2 = o0 * o0 + o1 * o1 + o2 * o2;
            enst += m2;
            vmax = fmax(vmax, sqrt(m2));
            hel += o0 * u0[i] + o1 * u1[i] + o2 * u2[i];
        } else {
            const double o2 = v_dx - u_dy;
            enst += o2 * o2;
            vmax = fmax(vmax, fabs(o2));
        }
        if (va.nu > 0.0) {
            double lu = __dmul_rn(fd(va.fd2, u0 + i, 1), dx2_inv) + __dmul_rn(fd(va.fd2, u0 + i, sy), dy2_inv);
            double lv = __dmul_rn(fd(va.fd2, u1 + i, 1), dx2_inv) + __dmul_rn(fd(va.fd2, u1 + i, sy), dy2_inv);
            if (dim == 3) {
                lu = lu + __dmul_rn(fd(va.fd2, u0 + i, sz), dz2_inv);
                lv = lv + __dmul_rn(fd(va.fd2, u1 + i, sz), dz2_inv);
                const double lw = __dmul_rn(fd(va.fd2, u2 + i, 1), dx2_inv) + __dmul_rn(fd(va.fd2, u2 + i, sy), dy2_inv) +
                                  __dmul_rn(fd(va.fd2, u2 + i, sz), dz2_inv);
                dis += u0[i] * lu + u1[i] * lv + u2[i] * lw;
            } else dis += u0[i] * lu + u1[i] * lv;
        }
    }
    const double t0 = block_sum(enst, red), t1 = block_max(vmax, red), t2 = block_sum(hel, red), t3 = block_sum(dis, red);
    if (threadIdx.x == 0) {
        double *o = out + (long long)k * WGPU_NSTAT;
        o[19] = 0.5 * t0 * dV;
        o[20] = t1;
        o[21] = 0.5 * t2 * dV;
        o[22] = -va.nu * t3 * dV;
    }
}

// entry k of the result: sum over the blocks in list order (k < 13, 16..19, 21, 22), max (13, 14, 20), min (15)
__global__ void stats_final_kernel(const double *__restrict__ part, int nb, double *__restrict__ out)
{
    const int k = threadIdx.x;
    if (k >= WGPU_NSTAT) return;
    double acc = (k == 14) ? -1.0e300 : (k == 15 ? 1.0e300 : 0.0);
    for (int i = 0; i < nb; ++i) {
        const double v = part[(long long)i * WGPU_NSTAT + k];
        if (k == 13 || k == 14 || k == 20) acc = fmax(acc, v);
        else if (k == 15) acc = fmin(acc, v);
        else acc += v;
    }
    out[k] = acc;
}

}  // namespace

int32_t wgpu_launch_create_mask(wgpu_ctx *ctx, const MaskGeom &gm, double time)
{
    if (ctx->n_active == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    create_mask_kernel<<<ctx->n_active, 256, 0, ctx->stream>>>(ctx->MASK, ctx->d_active, ctx->d_level, ctx->d_ixyz, c.n_mask, c.Bs[0], c.Bs[1],
                                                              c.dim == 3 ? c.Bs[2] : 1, c.dim, gm, time);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_stats(wgpu_ctx *ctx, const double *u, const double *rhs, const double *mask, const StatArgs &sa, double *d_part)
{
    const wgpu_config &c = ctx->cfg;
    if (ctx->n_active == 0) return WGPU_OK;
    stats_block_kernel<<<ctx->n_active, 256, 0, ctx->stream>>>(u, rhs, mask, ctx->d_active, ctx->d_level, ctx->nc, c.n_mask, c.Bs[0], c.Bs[1],
                                                              c.dim == 3 ? c.Bs[2] : 1, c.dim, sa, d_part);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// m ghosted blocks in `staged` (ids = d_active + k0): entries 19..22 of d_part[k0 .. k0 + m)
int32_t wgpu_launch_vort_stats(wgpu_ctx *ctx, const double *staged, const int *d_ids, int m, int ncomp, const VortArgs &va, double *d_part)
{
    const wgpu_config &c = ctx->cfg;
    if (m == 0) return WGPU_OK;
    vort_block_kernel<<<m, 256, 0, ctx->stream>>>(staged, d_ids, ctx->d_level, ncomp, c.Bs[0], c.Bs[1], c.dim == 3 ? c.Bs[2] : 1, c.g, c.dim, va, d_part);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_stats_final(wgpu_ctx *ctx, const double *d_part, double *d_out)
{
    stats_final_kernel<<<1, 32, 0, ctx->stream>>>(d_part, ctx->n_active, d_out);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}
