// Hand-written sm_100a kernels of the WABBIT block hot path (3-D ACM time stepping).
//
// Data layout in HBM: every heavy array is stored WITHOUT ghost layers, block-major,
//   A[blk][comp][z][y][x],  x fastest, Bs^3 doubles per component (32 KiB at Bs=16, 128-byte rows).
// Ghost values are never materialised for same-level neighbours: the stage kernel gathers the
// halo of a block directly from the neighbours' interiors (or from the patch pool for level-jump /
// remote patches), which fuses sync_ghosts_RHS_tree into the stencil pass.
//
// stage_kernel = one Runge-Kutta stage of RungeKuttaGeneric (LIB/TIME/runge_kutta_generic.f90:72-154):
//   k_j   = RHS_3D_acm(u_j)                       (LIB/EQUATION/ACMnew/rhs_ACM.f90:927-1779)
//   u_j+1 = u_0 + sum_l (dt*a_{j+1,l}) k_l        (runge_kutta_generic.f90:90-112), or the final
//   u     = u_0 + sum_j (dt*b_j) k_j              (runge_kutta_generic.f90:136-154)
// plus, fused in: the integral_stage divergence guard (rhs_ACM.f90:133-146) and, on the final stage,
// GET_DT_BLOCK_ACM's max(u^2+v^2+w^2) for the next step (module_ACM.f90:648-665).
//
// One CTA per block, Bs x Bs threads, marching in z.  z-planes (4 components, xy-halo included) stream
// through a shared-memory ring filled with cp.async (LDGSTS) PF planes ahead of the compute front, so
// the SM keeps several planes of HBM traffic in flight while the FP64 pipe works on the current one.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <unordered_map>

#include "wgpu_internal.cuh"

namespace {

__device__ __forceinline__ void cp_async16(double *smem, const double *g)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(double *smem, const double *g)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Finite-difference tables (rhs_ACM.f90:976-985).  q points at the centre value, q[o] is offset o.
// Evaluation order follows the reference: left-to-right sums, then *dx_inv.
// ---------------------------------------------------------------------------------------------
template <int FD>
struct St;

template <>
struct St<2> {
    static constexpr int H = 1;
    __device__ static __forceinline__ double d1(const double *q, double dinv) { return (q[1] - q[-1]) * dinv * 0.5; }
    __device__ static __forceinline__ double d1p(const double *q, const double *r, double dinv)
    {
        return (q[1] * r[1] - q[-1] * r[-1]) * dinv * 0.5;
    }
    __device__ static __forceinline__ double d2(const double *q, double d2inv) { return (q[-1] - 2.0 * q[0] + q[1]) * d2inv; }
};

template <>
struct St<4> {
    static constexpr int H = 2;
    static constexpr double a0 = 1.0 / 12.0, a1 = -2.0 / 3.0, a3 = 2.0 / 3.0, a4 = -1.0 / 12.0;
    static constexpr double b0 = -1.0 / 12.0, b1 = 4.0 / 3.0, b2 = -5.0 / 2.0, b3 = 4.0 / 3.0, b4 = -1.0 / 12.0;
    __device__ static __forceinline__ double d1(const double *q, double dinv)
    {
        return (a0 * q[-2] + a1 * q[-1] + a3 * q[1] + a4 * q[2]) * dinv;
    }
    __device__ static __forceinline__ double d1p(const double *q, const double *r, double dinv)
    {
        return (a0 * q[-2] * r[-2] + a1 * q[-1] * r[-1] + a3 * q[1] * r[1] + a4 * q[2] * r[2]) * dinv;
    }
    __device__ static __forceinline__ double d2(const double *q, double d2inv)
    {
        return (b0 * q[-2] + b1 * q[-1] + b2 * q[0] + b3 * q[1] + b4 * q[2]) * d2inv;
    }
};

template <>
struct St<6> {
    static constexpr int H = 3;
    static constexpr double a0 = -1.0 / 60.0, a1 = 3.0 / 20.0, a2 = -3.0 / 4.0, a4 = 3.0 / 4.0, a5 = -3.0 / 20.0, a6 = 1.0 / 60.0;
    static constexpr double b0 = 1.0 / 90.0, b1 = -3.0 / 20.0, b2 = 3.0 / 2.0, b3 = -49.0 / 18.0, b4 = 3.0 / 2.0, b5 = -3.0 / 20.0,
                            b6 = 1.0 / 90.0;
    __device__ static __forceinline__ double d1(const double *q, double dinv)
    {
        return (a0 * q[-3] + a1 * q[-2] + a2 * q[-1] + a4 * q[1] + a5 * q[2] + a6 * q[3]) * dinv;
    }
    __device__ static __forceinline__ double d1p(const double *q, const double *r, double dinv)
    {
        return (a0 * q[-3] * r[-3] + a1 * q[-2] * r[-2] + a2 * q[-1] * r[-1] + a4 * q[1] * r[1] + a5 * q[2] * r[2] +
                a6 * q[3] * r[3]) * dinv;
    }
    __device__ static __forceinline__ double d2(const double *q, double d2inv)
    {
        return (b0 * q[-3] + b1 * q[-2] + b2 * q[-1] + b3 * q[0] + b4 * q[1] + b5 * q[2] + b6 * q[3]) * d2inv;
    }
};

// Tam & Webb optimised 4th-order first derivative, standard 4th-order second derivative (rhs_ACM.f90:978,1461ff)
template <>
struct St<40> {
    static constexpr int H = 3;
    static constexpr double a0 = -0.02651995, a1 = +0.18941314, a2 = -0.79926643, a4 = 0.79926643, a5 = -0.18941314, a6 = 0.02651995;
    __device__ static __forceinline__ double d1(const double *q, double dinv)
    {
        return (a0 * q[-3] + a1 * q[-2] + a2 * q[-1] + a4 * q[1] + a5 * q[2] + a6 * q[3]) * dinv;
    }
    __device__ static __forceinline__ double d1p(const double *q, const double *r, double dinv)
    {
        return (a0 * q[-3] * r[-3] + a1 * q[-2] * r[-2] + a2 * q[-1] * r[-1] + a4 * q[1] * r[1] + a5 * q[2] * r[2] +
                a6 * q[3] * r[3]) * dinv;
    }
    __device__ static __forceinline__ double d2(const double *q, double d2inv) { return St<4>::d2(q, d2inv); }
};

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------
// The stage kernel
// ---------------------------------------------------------------------------------------------
template <int FD, int BS>
struct Tile {
    static constexpr int H = St<FD>::H;
    static constexpr int XO = (H + 1) & ~1;        // x offset of the interior inside a smem row (even => 16 B aligned)
    static constexpr int PITCH = BS + 2 * XO;
    static constexpr int ROWS = BS + 2 * H;
    static constexpr int PLANE = ROWS * PITCH;     // one component
    static constexpr int NC = 4;
    static constexpr int SLOT = NC * PLANE;        // one z-plane, all components
    static constexpr size_t SMEM_MAX = 227 * 1024;
    static constexpr size_t SLOT_BYTES = (size_t)SLOT * sizeof(double);
    // planes in flight ahead of the compute front: 3 if the ring fits in shared memory, else fewer
    static constexpr int PF = (2 * H + 1 + 3) * SLOT_BYTES <= SMEM_MAX ? 3 : ((2 * H + 1 + 2) * SLOT_BYTES <= SMEM_MAX ? 2 : 1);
    static constexpr int RING = 2 * H + 1 + PF;
    static constexpr int NT = BS * BS;
    static constexpr size_t SMEM = (size_t)RING * SLOT * sizeof(double);
    static_assert(SMEM <= SMEM_MAX, "plane ring does not fit in shared memory for this (FD, Bs)");
};

// ---------------------------------------------------------------------------------------------
// Per-block load tables.  Where each row (and each x-halo strip) of a plane comes from does not depend on z, so it
// is resolved once per block into shared memory: an element offset (into the stage input or into the patch
// pool) plus the number of elements to add per plane.  The per-plane loader is then a table lookup + cp.async.
//   variant 0: interior planes  zp = 0..BS-1   (rows y = -H..BS+H-1, plane index zp)
//   variant 1: z- halo planes   zp = -H..-1    (rows y = 0..BS-1,    plane index zp+H)
//   variant 2: z+ halo planes   zp = BS..BS+H-1(rows y = 0..BS-1,    plane index zp-BS)
// ---------------------------------------------------------------------------------------------
#define LT_POOL (1ull << 63)
#define LT_JPOOL (1ull << 62)
#define LT_MASK (~(LT_POOL | LT_JPOOL))
#define LT_SKIP (~0ull)
#define LT_ZERO (~0ull - 1ull)

template <int FD, int BS>
struct LoadTables {
    using T = Tile<FD, BS>;
    static constexpr int NROW = T::NC * T::ROWS;
    static constexpr int NXH = T::NC * BS * 2;
    unsigned long long row_off[3][NROW];
    unsigned long long x_off[NXH];
    int row_stride[3][NROW];
    int x_stride[NXH];
};

// element offset of pool patch `code` (<= -2) with the flag of the buffer it lives in: the exchange pool (remote faces)
// or the jump pool (restriction / prediction patches of jump.cu); offsets inside a patch are added by the caller
__device__ __forceinline__ unsigned long long pool_base(const StageArgs &a, int code)
{
    const int pid = -2 - code;
    if (pid >= WGPU_JUMP_PID) return ((unsigned long long)(pid - WGPU_JUMP_PID) * (unsigned long long)a.jpatch) | LT_JPOOL;
    return (unsigned long long)a.pool_off[pid] | LT_POOL;
}

template <int FD, int BS>
__device__ __forceinline__ void build_tables(const StageArgs &a, LoadTables<FD, BS> &lt, int b, const int *code, int tid)
{
    using T = Tile<FD, BS>;
    constexpr int H = T::H, NC = T::NC, NT = T::NT, ROWS = T::ROWS;
    constexpr long long CS = (long long)BS * BS * BS;
    using LT = LoadTables<FD, BS>;
    for (int i = tid; i < 3 * LT::NROW; i += NT) {
        const int v = i / LT::NROW, rid = i % LT::NROW, c = rid / ROWS, y = rid % ROWS - H;
        unsigned long long off = LT_SKIP;
        int stride = BS * BS;
        if (v == 0) {
            if (y >= 0 && y < BS) off = ((unsigned long long)b * NC + c) * CS + y * BS;
            else {
                const int cd = y < 0 ? code[10] : code[16];       // (0,-1,0) -> 10 ; (0,+1,0) -> 16
                const int ys = y < 0 ? BS + y : y - BS, ky = y < 0 ? y + H : y - BS;
                if (cd >= 0) off = ((unsigned long long)cd * NC + c) * CS + ys * BS;
                else if (cd <= -2) {                              // pool patch (Bs, H, Bs)
                    off = pool_base(a, cd) + ((long long)c * BS * H + ky) * BS;
                    stride = H * BS;
                } else off = LT_ZERO;
            }
        } else if (y >= 0 && y < BS) {
            const int cd = v == 1 ? code[4] : code[22];           // (0,0,-1) -> 4 ; (0,0,+1) -> 22
            if (cd >= 0) off = ((unsigned long long)cd * NC + c) * CS + (v == 1 ? (long long)(BS - H) * BS * BS : 0) + y * BS;
            else if (cd <= -2) off = pool_base(a, cd) + ((long long)c * H * BS + y) * BS;  // (Bs,Bs,H)
            else off = LT_ZERO;
        }
        lt.row_off[v][rid] = off;
        lt.row_stride[v][rid] = stride;
    }
    for (int i = tid; i < LT::NXH; i += NT) {
        const int side = i & 1, cy = i >> 1, c = cy / BS, y = cy % BS;
        const int cd = side ? code[14] : code[12];                // (+1,0,0) -> 14 ; (-1,0,0) -> 12
        unsigned long long off;
        int stride = BS * BS;
        if (cd >= 0) off = ((unsigned long long)cd * NC + c) * CS + y * BS + (side ? 0 : BS - H);
        else if (cd <= -2) {                                      // pool patch (H, Bs, Bs)
            off = pool_base(a, cd) + ((long long)c * BS * BS + y) * H;
            stride = BS * H;
        } else off = LT_ZERO;
        lt.x_off[i] = off;
        lt.x_stride[i] = stride;
    }
}

template <int FD, int BS>
__device__ __forceinline__ void load_plane(const StageArgs &a, const LoadTables<FD, BS> &lt, double *sm, int q, int tid)
{
    using T = Tile<FD, BS>;
    using LT = LoadTables<FD, BS>;
    constexpr int H = T::H, XO = T::XO, PITCH = T::PITCH, PLANE = T::PLANE, NT = T::NT, HB = BS / 2;
    const int zp = q - H;
    const int v = zp < 0 ? 1 : (zp >= BS ? 2 : 0);
    const int pz = zp < 0 ? zp + H : (zp >= BS ? zp - BS : zp);
    double *dst = sm + (q % T::RING) * T::SLOT;

#pragma unroll
    for (int k = 0; k < (LT::NROW * HB + NT - 1) / NT; ++k) {
        const int i = tid + k * NT;
        if (i < LT::NROW * HB) {
            const int rid = i / HB, xc = i % HB;
            const unsigned long long off = lt.row_off[v][rid];
            if (off != LT_SKIP) {
                double *d = dst + rid * PITCH + XO + 2 * xc;       // c*PLANE + row*PITCH == rid*PITCH
                if (off == LT_ZERO) {
                    d[0] = 0.0;
                    d[1] = 0.0;
                } else {
                    const double *base = (off & LT_POOL) ? a.pool : ((off & LT_JPOOL) ? a.jpool : a.u_in);
                    cp_async16(d, base + (long long)(off & LT_MASK) + (long long)pz * lt.row_stride[v][rid] + 2 * xc);
                }
            }
        }
    }
    if (v == 0) {
        // x halos: H elements per (component, row, side); 16-byte chunks when H is even, else 8-byte elements
        constexpr int PER = (H % 2 == 0) ? H / 2 : H;
#pragma unroll
        for (int k = 0; k < (LT::NXH * PER + NT - 1) / NT; ++k) {
            const int i = tid + k * NT;
            if (i < LT::NXH * PER) {
                const int sid = i / PER, e = i % PER, side = sid & 1, cy = sid >> 1, c = cy / BS, y = cy % BS;
                const unsigned long long off = lt.x_off[sid];
                double *d = dst + c * PLANE + (y + H) * PITCH + XO + (side ? BS : -H);
                if (off == LT_ZERO) {
                    if (H % 2 == 0) { d[2 * e] = 0.0; d[2 * e + 1] = 0.0; }
                    else d[e] = 0.0;
                } else {
                    const double *src = ((off & LT_POOL) ? a.pool : ((off & LT_JPOOL) ? a.jpool : a.u_in)) + (long long)(off & LT_MASK) +
                                        (long long)pz * lt.x_stride[sid];
                    if (H % 2 == 0) cp_async16(d + 2 * e, src + 2 * e);
                    else cp_async8(d + e, src + e);
                }
            }
        }
    }
}

template <int FD, bool SKEW, int BS, bool GEOM, bool SKIP_PLAIN = false>
__global__ void __launch_bounds__(BS *BS, (Tile<FD, BS>::SMEM <= 110 * 1024 ? 2 : 1))
    stage_kernel(const __grid_constant__ StageArgs a)
{
    using T = Tile<FD, BS>;
    using S = St<FD>;
    constexpr int H = T::H, XO = T::XO, PITCH = T::PITCH, PLANE = T::PLANE, NC = T::NC, RING = T::RING, SLOT = T::SLOT, PF = T::PF;
    constexpr int NQ = BS + 2 * H;
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_code[WGPU_NDIR];
    __shared__ double s_red[32];
    __shared__ LoadTables<FD, BS> s_lt;

    const int tid = threadIdx.x;
    const int tx = tid % BS, ty = tid / BS;
    const int b = a.active[blockIdx.x];
    if (SKIP_PLAIN) {
        // second launch behind stage_kernel_tma: that kernel has advanced every block whose six face neighbours are resident same-level blocks
        // (a template variant: an early return in the default instance costs it 36 bytes of spills and 9 % of its speed with the in-kernel mask)
        const int *nb = a.nbr + b * WGPU_NDIR;
        if ((nb[12] | nb[14] | nb[10] | nb[16] | nb[4] | nb[22]) >= 0) return;
    }
    if (tid < WGPU_NDIR) s_code[tid] = a.nbr[b * WGPU_NDIR + tid];
    __syncthreads();
    build_tables<FD, BS>(a, s_lt, b, s_code, tid);
    __syncthreads();

    // geometry of this block (module_treelib.f90:93): dx = 2^-J * L / Bs
    const int lvl = a.level[b];
    const double dx = a.dx_lvl[lvl][0], dy = a.dx_lvl[lvl][1], dz = a.dx_lvl[lvl][2];
    const double dinv[3] = {1.0 / dx, 1.0 / dy, 1.0 / dz};
    const double d2inv[3] = {1.0 / (dx * dx), 1.0 / (dy * dy), 1.0 / (dz * dz)};
    const double dt = (a.u_out || a.acc_out) ? *a.dt_ptr : 0.0;
    // bases of the two epilogue outputs; when a base is the stage input itself its value is already in smem
    const bool base_u_global = a.u_out && a.u0 != a.u_in;
    const bool base_acc_global = a.acc_out && a.acc_in != a.u_in;
    const double c02 = a.c0 * a.c0;
    // GEOM: the mask function of a translating sphere at this thread's (x, y) column (draw_sphere, LIB/EQUATION/insects/module_geometry.f90;
    // step_cosine4, LIB/HELPER/module_helpers.f90): coordinates x = i*dx + x0 and the squared distance in the reference's operation order
    double gxy2 = 0.0, gz0 = 0.0, gcz = 0.0;
    if (GEOM) {
        const double ts = __dadd_rn(a.t0_ptr ? *a.t0_ptr : a.t0, __dmul_rn(a.t_cj, *a.dt_ptr));
        const double cx = __dadd_rn(a.g_c0[0], __dmul_rn(a.g_v[0], ts)), cy = __dadd_rn(a.g_c0[1], __dmul_rn(a.g_v[1], ts));
        gcz = __dadd_rn(a.g_c0[2], __dmul_rn(a.g_v[2], ts));
        const double x = __dadd_rn(__dmul_rn((double)tx, dx), __dmul_rn((double)(a.ixyz[3 * b] * BS), dx));
        const double y = __dadd_rn(__dmul_rn((double)ty, dy), __dmul_rn((double)(a.ixyz[3 * b + 1] * BS), dy));
        gz0 = __dmul_rn((double)(a.ixyz[3 * b + 2] * BS), dz);
        const double ex = __dsub_rn(x, cx), ey = __dsub_rn(y, cy);
        gxy2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
    }

    // prologue: planes 0 .. 2H+PF-1 in flight
#pragma unroll 1
    for (int q = 0; q < 2 * H + PF; ++q) {
        if (q < NQ) load_plane<FD, BS>(a, s_lt, sm, q, tid);
        cp_async_commit();
    }

    const int cidx = (ty + H) * PITCH + XO + tx;
    constexpr long long CS = (long long)BS * BS * BS;
    double umag_max = 0.0, uabs_max = 0.0;

#pragma unroll 1
    for (int z = 0; z < BS; ++z) {
        // plane q = z+2H must have landed: all but the newest PF-1 groups... groups are committed one per plane,
        // the newest committed plane is z+2H+PF-1, so waiting for <= PF-1 pending groups completes plane z+2H.
        cp_async_wait<PF - 1>();
        __syncthreads();
        {
            // refill the slot of plane z-1 (last read in iteration z-1, which every thread has left: barrier above)
            const int q = z + 2 * H + PF;
            if (q < NQ) load_plane<FD, BS>(a, s_lt, sm, q, tid);
            cp_async_commit();
        }

        // issue the epilogue's global reads now: their latency hides behind the stencil arithmetic of this plane
        const long long gi = ((long long)b * NC) * CS + (long long)z * BS * BS + ty * BS + tx;
        double pre_u[4], pre_acc[4];
        if (base_u_global) {
#pragma unroll
            for (int c = 0; c < 4; ++c) pre_u[c] = __ldg(a.u0 + gi + c * CS);
        }
        if (base_acc_global) {
#pragma unroll
            for (int c = 0; c < 4; ++c) pre_acc[c] = a.acc_in[gi + c * CS];
        }

        double rhs[4];
        double ctr[4];
        {
            double d1v[3][4];  // [dir][comp]
            double d2v[3][3];
            double cv[3][3];   // skew: cv[dir][comp] = d/d(dir) (comp * vel_dir)
#pragma unroll
            for (int dir = 0; dir < 3; ++dir) {
                double qv[4][2 * H + 1];
#pragma unroll
                for (int o = -H; o <= H; ++o) {
                    int off;
                    if (dir == 0) off = ((z + H) % RING) * SLOT + cidx + o;
                    else if (dir == 1) off = ((z + H) % RING) * SLOT + cidx + o * PITCH;
                    else off = ((z + H + o) % RING) * SLOT + cidx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) qv[c][o + H] = sm[off + c * PLANE];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) d1v[dir][c] = S::d1(&qv[c][H], dinv[dir]);
#pragma unroll
                for (int c = 0; c < 3; ++c) d2v[dir][c] = S::d2(&qv[c][H], d2inv[dir]);
                if (SKEW) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) cv[dir][c] = S::d1p(&qv[c][H], &qv[dir][H], dinv[dir]);
                }
                if (dir == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) ctr[c] = qv[c][H];
                }
            }
            const double u = ctr[0], v = ctr[1], w = ctr[2], p = ctr[3];
            double penal[3] = {0.0, 0.0, 0.0};
            const long long g0 = ((long long)b * a.n_mask) * CS + (long long)z * BS * BS + ty * BS + tx;
            if (GEOM) {
                const double ez = __dsub_rn(__dadd_rn(__dmul_rn((double)z, dz), gz0), gcz);
                const double dist = __dsub_rn(sqrt(__dadd_rn(gxy2, __dmul_rn(ez, ez))), a.g_R);
                double m = 0.0;
                if (dist <= -a.g_h) m = 1.0;
                else if (dist < a.g_h) m = 0.5 * (1.0 + cos((dist + a.g_h) * 3.14159265358979323846 / (2.0 * a.g_h)));
                const double chi = m * a.C_eta_inv;
                penal[0] = -chi * (u - a.g_v[0]);
                penal[1] = -chi * (v - a.g_v[1]);
                penal[2] = -chi * (w - a.g_v[2]);
            } else if (a.mask) {
                // chi = mask(1) * C_eta_apply_inv(int(mask(5)))   rhs_ACM.f90:1192-1195
                const int color = (int)a.mask[g0 + 4 * CS];
                const double chi = a.mask[g0] * (color == 0 ? 0.0 : a.C_eta_inv);
                penal[0] = -chi * (u - a.mask[g0 + 1 * CS]);
                penal[1] = -chi * (v - a.mask[g0 + 2 * CS]);
                penal[2] = -chi * (w - a.mask[g0 + 3 * CS]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double adv;
                if (SKEW)  // rhs_ACM.f90:1199-1201
                    adv = -0.5 * (cv[0][c] + cv[1][c] + cv[2][c] + u * d1v[0][c] + v * d1v[1][c] + w * d1v[2][c]);
                else       // rhs_ACM.f90:1249-1251
                    adv = (-u * d1v[0][c] - v * d1v[1][c] - w * d1v[2][c]);
                rhs[c] = adv - d1v[c][3] + a.nu * (d2v[0][c] + d2v[1][c] + d2v[2][c]) + penal[c];
            }
            rhs[3] = -c02 * (d1v[0][0] + d1v[1][1] + d1v[2][2]) - a.gamma_p * p;   // rhs_ACM.f90:1202
            if (a.use_sponge && a.mask) {  // rhs_ACM.f90:1734-1753
                const double spo = a.mask[g0 + 5 * CS] * a.C_sponge_inv;
                rhs[0] = rhs[0] - (u - a.u_mean_set[0]) * spo;
                rhs[1] = rhs[1] - (v - a.u_mean_set[1]) * spo;
                rhs[2] = rhs[2] - (w - a.u_mean_set[2]) * spo;
                rhs[3] = rhs[3] - p * spo;
            }
        }
        uabs_max = fmax(uabs_max, fmax(fmax(fabs(ctr[0]), fabs(ctr[1])), fmax(fabs(ctr[2]), fabs(ctr[3]))));

        // epilogue: store the slope, form the next stage input / the new state and the running final combination
        double un[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (a.k_out) a.k_out[gi + c * CS] = rhs[c];
            if (a.u_out) {
                double acc = base_u_global ? pre_u[c] : ctr[c];
                for (int l = 0; l < a.n_prev; ++l)   // (dt*a_jl)*k_l, increasing l  (runge_kutta_generic.f90:108-110)
                    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_prev[l]), a.k_prev[l][gi + c * CS]));
                if (a.use_self) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_self), rhs[c]));
                a.u_out[gi + c * CS] = acc;
                un[c] = acc;
            }
            if (a.acc_out) {   // runge_kutta_generic.f90:142-153, one term per stage, increasing j
                double acc = base_acc_global ? pre_acc[c] : ctr[c];
                if (a.use_acc) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_acc), rhs[c]));
                a.acc_out[gi + c * CS] = acc;
            }
        }
        if (a.dtmin_bits && a.u_out) {
            // u_mag = u^2 + v^2 + w^2   (module_ACM.f90:651-652), no contraction
            const double m = __dadd_rn(__dadd_rn(__dmul_rn(un[0], un[0]), __dmul_rn(un[1], un[1])), __dmul_rn(un[2], un[2]));
            umag_max = fmax(umag_max, m);
        }
    }
    cp_async_wait<0>();

    // block reductions: divergence guard, and the CFL time step of this block for the next step
    const int lane = tid & 31, wid = tid >> 5;
    constexpr int NW = (BS * BS + 31) / 32;
    uabs_max = warp_max(uabs_max);
    umag_max = warp_max(umag_max);
    if (lane == 0) s_red[wid] = uabs_max;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < NW; ++i) m = fmax(m, s_red[i]);
        if (m > 1.0e12) atomicExch(a.diverged, 1);   // LIM_DIVERGED, rhs_ACM.f90:134
    }
    if (a.dtmin_bits && a.u_out) {
        __syncthreads();
        if (lane == 0) s_red[wid] = umag_max;
        __syncthreads();
        if (tid == 0) {
            double m = 0.0;
            for (int i = 0; i < NW; ++i) m = fmax(m, s_red[i]);
            // module_ACM.f90:657-665
            const double u_eigen = __dadd_rn(sqrt(m), sqrt(__dadd_rn(c02, m)));
            double dxmin = dx;
            if (a.dim_min_axes > 1) dxmin = fmin(dxmin, dy);
            if (a.dim_min_axes > 2) dxmin = fmin(dxmin, dz);
            double dtb = (u_eigen >= 1.0e-6) ? __ddiv_rn(__dmul_rn(a.CFL, dxmin), u_eigen) : 1.0e-2;
            // explicit diffusion, with THIS block's dx (module_ACM.f90:669-671): folded in before the MIN over blocks and ranks
            if (a.nu > 1.0e-13) dtb = fmin(dtb, __ddiv_rn(__dmul_rn(a.CFL_nu, __dmul_rn(dxmin, dxmin)), a.nu));
            atomicMin(a.dtmin_bits, (unsigned long long)__double_as_longlong(dtb));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// stage_kernel_tma<FD, SKEW, BS, GEOM>: the same stage with the plane ring filled by the TMA unit instead of per-thread cp.async.
// For a block whose six face neighbours are resident same-level blocks ("plain": every block of an equidistant periodic grid, the
// interior blocks of a partition, the blocks away from level jumps on a graded grid) every source of a z-plane is a dense box of the
// resident array A[blk*NC + c][z][y][x], described by three tensor maps (boxes Bs x Bs, Bs x H and XW x Bs).  ONE thread issues the
// bulk tensor copies of a plane -- per component: the block's own plane, the y-/y+ halo rows and the x-/x+ halo columns from the four
// neighbours; for the 2H z-halo planes the neighbour's plane -- onto the plane's mbarrier (expect_tx = the plane's bytes); the other
// 255 threads issue no load instruction at all and wait on the mbarrier's phase before they read the plane.  Shared-memory layout of a
// plane, per component (RS doubles): [ROWS = Bs+2H rows][Bs] (y halo rows in place, pitch Bs: TMA writes dense boxes) followed by the
// x- and x+ halo columns as [Bs][XW] each; the x-direction taps of the threads next to the block's x faces read those through four
// loop-invariant offsets.  Blocks that are not plain leave at once; stage_kernel (above) takes them in a second launch (skip_plain).
// ---------------------------------------------------------------------------------------------
template <int FD, int BS, int SLACK_ = 1>
struct TmaTile {
    static constexpr int H = St<FD>::H;
    static constexpr int XW = (H + 1) & ~1;          // x halo columns moved per side (16-byte multiple)
    static constexpr int ROWS = BS + 2 * H;
    static constexpr int NC = 4;
    static constexpr int RS = ROWS * BS + 2 * BS * XW;   // one component of one plane
    static constexpr int SLOT = NC * RS;
    static constexpr int PF = 3;                     // planes in flight ahead of the compute front
    static constexpr int SLACK = SLACK_;             // extra ring slots: a slot is refilled SLACK iterations after its plane died (0: block barrier per plane)
    static constexpr int RING = 2 * H + 1 + PF + SLACK;
    static constexpr int NT = BS * BS;
    static constexpr int NW = NT / 32;
    static constexpr size_t SMEM = (size_t)RING * SLOT * sizeof(double) + 1024;   // + alignment slack
    static constexpr unsigned BYTES_PLANE = NC * (BS * BS + 2 * H * BS + 2 * XW * BS) * 8;
    static constexpr unsigned BYTES_ZHALO = NC * BS * BS * 8;
    static_assert((BS * 8) % 128 == 0, "TMA destinations must stay 128-byte aligned: rows of a multiple of 16 doubles");
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap *map, unsigned long long *bar, double *dst, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     (unsigned)__cvta_generic_to_shared(dst)),
                 "l"((unsigned long long)map), "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

template <int FD, bool SKEW, int BS, bool GEOM, int SLACK>
__global__ void __launch_bounds__(BS *BS, 2)
    stage_kernel_tma(const __grid_constant__ StageArgs a, const __grid_constant__ CUtensorMap tm_int, const __grid_constant__ CUtensorMap tm_y,
                     const __grid_constant__ CUtensorMap tm_x)
{
    using T = TmaTile<FD, BS, SLACK>;
    using S = St<FD>;
    constexpr int H = T::H, XW = T::XW, ROWS = T::ROWS, RS = T::RS, NC = T::NC, RING = T::RING, SLOT = T::SLOT, PF = T::PF;
    constexpr int NQ = BS + 2 * H, NW = T::NW;
    extern __shared__ double sm_raw[];
    __shared__ __align__(8) unsigned long long s_full[RING], s_empty[RING];
    __shared__ int s_code[WGPU_NDIR];
    __shared__ double s_red[32];
    double *sm = (double *)(((unsigned long long)sm_raw + 1023ull) & ~1023ull);

    const int tid = threadIdx.x;
    const int tx = tid % BS, ty = tid / BS;
    const int lane = tid & 31, wid = tid >> 5;
    const int b = a.active[blockIdx.x];
    if (tid < WGPU_NDIR) s_code[tid] = a.nbr[b * WGPU_NDIR + tid];
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < RING; ++s) {
            mbar_init(&s_full[s], 1);        // the arrive.expect_tx of warp 0; the bytes of all warps' copies complete the phase
            mbar_init(&s_empty[s], NW);      // one arrival per warp when it has read the slot's plane for the last time
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((s_code[12] | s_code[14] | s_code[10] | s_code[16] | s_code[4] | s_code[22]) < 0) return;   // not plain: stage_kernel takes this block

    // lane 0 of EVERY warp issues its share of plane q's copies (20 boxes of an interior plane, 4 of a z-halo plane, dealt round robin over
    // the warps), after all warps have released the slot's previous plane
    auto issue = [&](int q) {
        const int s = q % RING, zp = q - H;
        double *dst = sm + s * SLOT;
        unsigned long long *bar = &s_full[s];
        if (SLACK > 0 && q >= RING) mbar_wait(&s_empty[s], ((q / RING) - 1) & 1);
        if (zp >= 0 && zp < BS) {
            if (wid == 0) mbar_expect_tx(bar, T::BYTES_PLANE);
            for (int i = wid; i < 5 * NC; i += NW) {
                const int c = i / 5, kind = i % 5;
                double *d = dst + c * RS;
                // neighbour codes are read from shared memory here (lane 0 only) instead of living in registers through the plane loop
                if (kind == 0) tma_load_4d(&tm_int, bar, d + H * BS, 0, 0, zp, b * NC + c);
                else if (kind == 1) tma_load_4d(&tm_y, bar, d, 0, BS - H, zp, s_code[10] * NC + c);
                else if (kind == 2) tma_load_4d(&tm_y, bar, d + (H + BS) * BS, 0, 0, zp, s_code[16] * NC + c);
                else if (kind == 3) tma_load_4d(&tm_x, bar, d + ROWS * BS, BS - XW, 0, zp, s_code[12] * NC + c);
                else tma_load_4d(&tm_x, bar, d + ROWS * BS + BS * XW, 0, 0, zp, s_code[14] * NC + c);
            }
        } else {
            const int nb = zp < 0 ? s_code[4] : s_code[22], zz = zp < 0 ? BS + zp : zp - BS;
            if (wid == 0) mbar_expect_tx(bar, T::BYTES_ZHALO);
            if (wid < NC) tma_load_4d(&tm_int, bar, dst + wid * RS + H * BS, 0, 0, zz, nb * NC + wid);
        }
    };

    if (lane == 0) {
#pragma unroll 1
        for (int q = 0; q < 2 * H + PF; ++q)
            if (q < NQ) issue(q);
    }
    __syncwarp();

    const int lvl = a.level[b];
    const double dx = a.dx_lvl[lvl][0], dy = a.dx_lvl[lvl][1], dz = a.dx_lvl[lvl][2];
    const double dinv[3] = {1.0 / dx, 1.0 / dy, 1.0 / dz};
    const double d2inv[3] = {1.0 / (dx * dx), 1.0 / (dy * dy), 1.0 / (dz * dz)};
    const double dt = (a.u_out || a.acc_out) ? *a.dt_ptr : 0.0;
    const bool base_u_global = a.u_out && a.u0 != a.u_in;
    const bool base_acc_global = a.acc_out && a.acc_in != a.u_in;
    const double c02 = a.c0 * a.c0;
    double gxy2 = 0.0, gz0 = 0.0, gcz = 0.0;
    if (GEOM) {
        const double ts = __dadd_rn(a.t0_ptr ? *a.t0_ptr : a.t0, __dmul_rn(a.t_cj, *a.dt_ptr));
        const double cx = __dadd_rn(a.g_c0[0], __dmul_rn(a.g_v[0], ts)), cy = __dadd_rn(a.g_c0[1], __dmul_rn(a.g_v[1], ts));
        gcz = __dadd_rn(a.g_c0[2], __dmul_rn(a.g_v[2], ts));
        const double x = __dadd_rn(__dmul_rn((double)tx, dx), __dmul_rn((double)(a.ixyz[3 * b] * BS), dx));
        const double y = __dadd_rn(__dmul_rn((double)ty, dy), __dmul_rn((double)(a.ixyz[3 * b + 1] * BS), dy));
        gz0 = __dmul_rn((double)(a.ixyz[3 * b + 2] * BS), dz);
        const double ex = __dsub_rn(x, cx), ey = __dsub_rn(y, cy);
        gxy2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
    }

    // offsets inside a component region: centre, and the x-direction taps (the x halo columns live behind the rows)
    const int cidx = (ty + H) * BS + tx;
    int xo[2 * H + 1];
#pragma unroll
    for (int o = -H; o <= H; ++o) {
        const int x = tx + o;
        xo[o + H] = x < 0 ? ROWS * BS + ty * XW + (XW + x) : (x >= BS ? ROWS * BS + BS * XW + ty * XW + (x - BS) : cidx + o);
    }
    constexpr long long CS = (long long)BS * BS * BS;
    double umag_max = 0.0, uabs_max = 0.0;

#pragma unroll
    for (int q = 0; q < 2 * H; ++q) mbar_wait(&s_full[q], 0);      // the first iteration reads planes 0 .. 2H; it waits for plane 2H itself

#pragma unroll 1
    for (int z = 0; z < BS; ++z) {
        if (SLACK == 0) __syncthreads();               // every thread has left iteration z-1: the slot of its oldest plane is free
        if (lane == 0) {
            const int q = z + 2 * H + PF;
            if (q < NQ) issue(q);
        }
        __syncwarp();
        mbar_wait(&s_full[(z + 2 * H) % RING], ((z + 2 * H) / RING) & 1);   // plane z+2H has landed (planes z .. z+2H-1 were waited for earlier)

        const long long gi = ((long long)b * NC) * CS + (long long)z * BS * BS + ty * BS + tx;
        double pre_u[4], pre_acc[4];
        if (base_u_global) {
#pragma unroll
            for (int c = 0; c < 4; ++c) pre_u[c] = __ldg(a.u0 + gi + c * CS);
        }
        if (base_acc_global) {
#pragma unroll
            for (int c = 0; c < 4; ++c) pre_acc[c] = a.acc_in[gi + c * CS];
        }

        double rhs[4];
        double ctr[4];
        {
            double d1v[3][4];
            double d2v[3][3];
            double cv[3][3];
#pragma unroll
            for (int dir = 0; dir < 3; ++dir) {
                double qv[4][2 * H + 1];
#pragma unroll
                for (int o = -H; o <= H; ++o) {
                    int off;
                    if (dir == 0) off = ((z + H) % RING) * SLOT + xo[o + H];
                    else if (dir == 1) off = ((z + H) % RING) * SLOT + cidx + o * BS;
                    else off = ((z + H + o) % RING) * SLOT + cidx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) qv[c][o + H] = sm[off + c * RS];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) d1v[dir][c] = S::d1(&qv[c][H], dinv[dir]);
#pragma unroll
                for (int c = 0; c < 3; ++c) d2v[dir][c] = S::d2(&qv[c][H], d2inv[dir]);
                if (SKEW) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) cv[dir][c] = S::d1p(&qv[c][H], &qv[dir][H], dinv[dir]);
                }
                if (dir == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) ctr[c] = qv[c][H];
                }
            }
            const double u = ctr[0], v = ctr[1], w = ctr[2], p = ctr[3];
            double penal[3] = {0.0, 0.0, 0.0};
            const long long g0 = ((long long)b * a.n_mask) * CS + (long long)z * BS * BS + ty * BS + tx;
            if (GEOM) {
                const double ez = __dsub_rn(__dadd_rn(__dmul_rn((double)z, dz), gz0), gcz);
                const double dist = __dsub_rn(sqrt(__dadd_rn(gxy2, __dmul_rn(ez, ez))), a.g_R);
                double m = 0.0;
                if (dist <= -a.g_h) m = 1.0;
                else if (dist < a.g_h) m = 0.5 * (1.0 + cos((dist + a.g_h) * 3.14159265358979323846 / (2.0 * a.g_h)));
                const double chi = m * a.C_eta_inv;
                penal[0] = -chi * (u - a.g_v[0]);
                penal[1] = -chi * (v - a.g_v[1]);
                penal[2] = -chi * (w - a.g_v[2]);
            } else if (a.mask) {
                const int color = (int)a.mask[g0 + 4 * CS];
                const double chi = a.mask[g0] * (color == 0 ? 0.0 : a.C_eta_inv);
                penal[0] = -chi * (u - a.mask[g0 + 1 * CS]);
                penal[1] = -chi * (v - a.mask[g0 + 2 * CS]);
                penal[2] = -chi * (w - a.mask[g0 + 3 * CS]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double adv;
                if (SKEW) adv = -0.5 * (cv[0][c] + cv[1][c] + cv[2][c] + u * d1v[0][c] + v * d1v[1][c] + w * d1v[2][c]);
                else adv = (-u * d1v[0][c] - v * d1v[1][c] - w * d1v[2][c]);
                rhs[c] = adv - d1v[c][3] + a.nu * (d2v[0][c] + d2v[1][c] + d2v[2][c]) + penal[c];
            }
            rhs[3] = -c02 * (d1v[0][0] + d1v[1][1] + d1v[2][2]) - a.gamma_p * p;
            if (a.use_sponge && a.mask) {
                const double spo = a.mask[g0 + 5 * CS] * a.C_sponge_inv;
                rhs[0] = rhs[0] - (u - a.u_mean_set[0]) * spo;
                rhs[1] = rhs[1] - (v - a.u_mean_set[1]) * spo;
                rhs[2] = rhs[2] - (w - a.u_mean_set[2]) * spo;
                rhs[3] = rhs[3] - p * spo;
            }
        }
        uabs_max = fmax(uabs_max, fmax(fmax(fabs(ctr[0]), fabs(ctr[1])), fmax(fabs(ctr[2]), fabs(ctr[3]))));

        double un[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (a.k_out) a.k_out[gi + c * CS] = rhs[c];
            if (a.u_out) {
                double acc = base_u_global ? pre_u[c] : ctr[c];
                for (int l = 0; l < a.n_prev; ++l)
                    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_prev[l]), a.k_prev[l][gi + c * CS]));
                if (a.use_self) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_self), rhs[c]));
                a.u_out[gi + c * CS] = acc;
                un[c] = acc;
            }
            if (a.acc_out) {
                double acc = base_acc_global ? pre_acc[c] : ctr[c];
                if (a.use_acc) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_acc), rhs[c]));
                a.acc_out[gi + c * CS] = acc;
            }
        }
        if (a.dtmin_bits && a.u_out) {
            const double m = __dadd_rn(__dadd_rn(__dmul_rn(un[0], un[0]), __dmul_rn(un[1], un[1])), __dmul_rn(un[2], un[2]));
            umag_max = fmax(umag_max, m);
        }
        if (SLACK > 0) {
            __syncwarp();                              // every lane has consumed its reads of plane z: this warp releases the slot
            if (lane == 0) mbar_arrive(&s_empty[z % RING]);
        }
    }

    uabs_max = warp_max(uabs_max);
    umag_max = warp_max(umag_max);
    if (lane == 0) s_red[wid] = uabs_max;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < NW; ++i) m = fmax(m, s_red[i]);
        if (m > 1.0e12) atomicExch(a.diverged, 1);
    }
    if (a.dtmin_bits && a.u_out) {
        __syncthreads();
        if (lane == 0) s_red[wid] = umag_max;
        __syncthreads();
        if (tid == 0) {
            double m = 0.0;
            for (int i = 0; i < NW; ++i) m = fmax(m, s_red[i]);
            const double u_eigen = __dadd_rn(sqrt(m), sqrt(__dadd_rn(c02, m)));
            double dxmin = dx;
            if (a.dim_min_axes > 1) dxmin = fmin(dxmin, dy);
            if (a.dim_min_axes > 2) dxmin = fmin(dxmin, dz);
            double dtb = (u_eigen >= 1.0e-6) ? __ddiv_rn(__dmul_rn(a.CFL, dxmin), u_eigen) : 1.0e-2;
            if (a.nu > 1.0e-13) dtb = fmin(dtb, __ddiv_rn(__dmul_rn(a.CFL_nu, __dmul_rn(dxmin, dxmin)), a.nu));
            atomicMin(a.dtmin_bits, (unsigned long long)__double_as_longlong(dtb));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// stage_kernel_any<FD, SKEW>: the same stage for ANY even block size (the reference accepts every even Bs, read_Bs in
// LIB/PARAMS/module_ini_files_parser_mpi.f90:816; its own 3-D penalized fixture uses Bs = 26).  The fast kernels above fix Bs at compile time
// (16 / 18 / 20: the CTA shape, the shared-memory ring and the cp.async tables depend on it); this one takes Bs at run time: a CTA of 256
// threads owns an xy tile of T x T points (T = ceil(Bs / ceil(Bs / 16)) <= 16, so Bs = 26 runs as 2 x 2 tiles of 13) and marches in z
// through a ring of 2H + 1 planes (4 components, tile + xy halo).  Every ring entry is resolved per element: own block, face neighbour's
// interior, exchange-pool / jump-pool patch (the layouts of build_tables above) or zero.  Same arithmetic, same epilogue, same fused
// reductions (the CFL candidate of a block is the MIN over its tiles' candidates: dt is monotone in max|u|).  HBM-bound like the fast
// kernels, but with plain loads and two barriers per plane: the fall-back, not the headline.
// ---------------------------------------------------------------------------------------------
template <int FD, bool SKEW>
__global__ void __launch_bounds__(256, 2) stage_kernel_any(const __grid_constant__ StageArgs a, int BS, int T, int NTILE)
{
    using S = St<FD>;
    constexpr int H = S::H, NC = 4;
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_code[WGPU_NDIR];
    __shared__ double s_red[8];
    const int PW = T + 2 * H, PLANE = PW * PW, SLOT = NC * PLANE;
    constexpr int RING = 2 * H + 1;
    const int tid = threadIdx.x;
    const int tile = blockIdx.x % (NTILE * NTILE), bi = blockIdx.x / (NTILE * NTILE);
    const int tx0 = (tile % NTILE) * T, ty0 = (tile / NTILE) * T;
    const int b = a.active[bi];
    if (tid < WGPU_NDIR) s_code[tid] = a.nbr[b * WGPU_NDIR + tid];
    __syncthreads();
    const long long CS = (long long)BS * BS * BS;

    // value of component c at block-local lattice point (gx, gy, gz), gz in [-H, BS+H), at most one coordinate outside [0, BS)
    auto fetch = [&](int c, int gx, int gy, int gz) -> double {
        int cd, dirx = 0, diry = 0, dirz = 0;
        if (gx < 0) dirx = -1; else if (gx >= BS) dirx = 1;
        if (gy < 0) diry = -1; else if (gy >= BS) diry = 1;
        if (gz < 0) dirz = -1; else if (gz >= BS) dirz = 1;
        if (!(dirx | diry | dirz)) return a.u_in[((long long)b * NC + c) * CS + ((long long)gz * BS + gy) * BS + gx];
        cd = s_code[(dirz + 1) * 9 + (diry + 1) * 3 + (dirx + 1)];
        if (cd >= 0) {
            const int x = gx - dirx * BS, y = gy - diry * BS, z = gz - dirz * BS;
            return a.u_in[((long long)cd * NC + c) * CS + ((long long)z * BS + y) * BS + x];
        }
        if (cd == -1) return 0.0;
        const unsigned long long base = pool_base(a, cd);
        const double *pp = (base & LT_POOL) ? a.pool : a.jpool;
        long long o;
        if (dirx) o = (((long long)c * BS + gz) * BS + gy) * H + (dirx < 0 ? gx + H : gx - BS);          // (H, Bs, Bs)
        else if (diry) o = (((long long)c * BS + gz) * H + (diry < 0 ? gy + H : gy - BS)) * BS + gx;     // (Bs, H, Bs)
        else o = (((long long)c * H + (dirz < 0 ? gz + H : gz - BS)) * BS + gy) * BS + gx;               // (Bs, Bs, H)
        return pp[(long long)(base & LT_MASK) + o];
    };
    auto load_plane = [&](int q) {   // q = gz + H
        const int gz = q - H;
        double *dst = sm + (q % RING) * SLOT;
        const bool zin = gz >= 0 && gz < BS;
        for (int i = tid; i < SLOT; i += 256) {
            const int c = i / PLANE, r = i % PLANE, yy = r / PW, xx = r % PW;
            const int gx = tx0 + xx - H, gy = ty0 + yy - H;
            const bool xin = gx >= 0 && gx < BS, yin = gy >= 0 && gy < BS;
            const bool xt = xx >= H && xx < H + T, yt = yy >= H && yy < H + T;     // inside the tile proper
            // needed: tile points (any z), and on interior planes the x / y halos of the tile (star stencil: no corners)
            bool need = xt && yt && xin && yin;
            // x halo of the tile: gy must lie inside the block (and vice versa); the last tile may overhang the block: stay within BS + H
            if (zin && xt != yt) need = (xt ? xin : yin) && gx < BS + H && gy < BS + H;
            dst[i] = need ? fetch(c, gx, gy, gz) : 0.0;
        }
    };

    const int lx = tid % 16, ly = tid / 16;
    const bool act = lx < T && ly < T && tx0 + lx < BS && ty0 + ly < BS;
    const int tx = tx0 + lx, ty = ty0 + ly;
    const int lvl = a.level[b];
    const double dx = a.dx_lvl[lvl][0], dy = a.dx_lvl[lvl][1], dz = a.dx_lvl[lvl][2];
    const double dinv[3] = {1.0 / dx, 1.0 / dy, 1.0 / dz};
    const double d2inv[3] = {1.0 / (dx * dx), 1.0 / (dy * dy), 1.0 / (dz * dz)};
    const double dt = (a.u_out || a.acc_out) ? *a.dt_ptr : 0.0;
    const bool base_u_global = a.u_out && a.u0 != a.u_in;
    const bool base_acc_global = a.acc_out && a.acc_in != a.u_in;
    const double c02 = a.c0 * a.c0;
    double gxy2 = 0.0, gz0 = 0.0, gcz = 0.0;
    if (a.geom) {   // translating sphere, as in the fast kernel (draw_sphere's operation order)
        const double ts = __dadd_rn(a.t0_ptr ? *a.t0_ptr : a.t0, __dmul_rn(a.t_cj, *a.dt_ptr));
        const double cx = __dadd_rn(a.g_c0[0], __dmul_rn(a.g_v[0], ts)), cy = __dadd_rn(a.g_c0[1], __dmul_rn(a.g_v[1], ts));
        gcz = __dadd_rn(a.g_c0[2], __dmul_rn(a.g_v[2], ts));
        const double x = __dadd_rn(__dmul_rn((double)tx, dx), __dmul_rn((double)(a.ixyz[3 * b] * BS), dx));
        const double y = __dadd_rn(__dmul_rn((double)ty, dy), __dmul_rn((double)(a.ixyz[3 * b + 1] * BS), dy));
        gz0 = __dmul_rn((double)(a.ixyz[3 * b + 2] * BS), dz);
        const double ex = __dsub_rn(x, cx), ey = __dsub_rn(y, cy);
        gxy2 = __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
    }
    for (int q = 0; q < 2 * H; ++q) load_plane(q);
    const int cidx = (ly + H) * PW + lx + H;
    double umag_max = 0.0, uabs_max = 0.0;
#pragma unroll 1
    for (int z = 0; z < BS; ++z) {
        load_plane(z + 2 * H);
        __syncthreads();
        if (act) {
            const long long gi = ((long long)b * NC) * CS + (long long)z * BS * BS + ty * BS + tx;
            double rhs[4], ctr[4];
            double d1v[3][4], d2v[3][3], cv[3][3];
#pragma unroll
            for (int dir = 0; dir < 3; ++dir) {
                double qv[4][2 * H + 1];
#pragma unroll
                for (int o = -H; o <= H; ++o) {
                    int off;
                    if (dir == 0) off = ((z + H) % RING) * SLOT + cidx + o;
                    else if (dir == 1) off = ((z + H) % RING) * SLOT + cidx + o * PW;
                    else off = ((z + H + o) % RING) * SLOT + cidx;
#pragma unroll
                    for (int c = 0; c < 4; ++c) qv[c][o + H] = sm[off + c * PLANE];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) d1v[dir][c] = S::d1(&qv[c][H], dinv[dir]);
#pragma unroll
                for (int c = 0; c < 3; ++c) d2v[dir][c] = S::d2(&qv[c][H], d2inv[dir]);
                if (SKEW) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) cv[dir][c] = S::d1p(&qv[c][H], &qv[dir][H], dinv[dir]);
                }
                if (dir == 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) ctr[c] = qv[c][H];
                }
            }
            const double u = ctr[0], v = ctr[1], w = ctr[2], p = ctr[3];
            double penal[3] = {0.0, 0.0, 0.0};
            const long long g0 = ((long long)b * a.n_mask) * CS + (long long)z * BS * BS + ty * BS + tx;
            if (a.geom) {
                const double ez = __dsub_rn(__dadd_rn(__dmul_rn((double)z, dz), gz0), gcz);
                const double dist = __dsub_rn(sqrt(__dadd_rn(gxy2, __dmul_rn(ez, ez))), a.g_R);
                double m = 0.0;
                if (dist <= -a.g_h) m = 1.0;
                else if (dist < a.g_h) m = 0.5 * (1.0 + cos((dist + a.g_h) * 3.14159265358979323846 / (2.0 * a.g_h)));
                const double chi = m * a.C_eta_inv;
                penal[0] = -chi * (u - a.g_v[0]);
                penal[1] = -chi * (v - a.g_v[1]);
                penal[2] = -chi * (w - a.g_v[2]);
            } else if (a.mask) {
                const int color = (int)a.mask[g0 + 4 * CS];
                const double chi = a.mask[g0] * (color == 0 ? 0.0 : a.C_eta_inv);
                penal[0] = -chi * (u - a.mask[g0 + 1 * CS]);
                penal[1] = -chi * (v - a.mask[g0 + 2 * CS]);
                penal[2] = -chi * (w - a.mask[g0 + 3 * CS]);
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double adv;
                if (SKEW) adv = -0.5 * (cv[0][c] + cv[1][c] + cv[2][c] + u * d1v[0][c] + v * d1v[1][c] + w * d1v[2][c]);
                else adv = (-u * d1v[0][c] - v * d1v[1][c] - w * d1v[2][c]);
                rhs[c] = adv - d1v[c][3] + a.nu * (d2v[0][c] + d2v[1][c] + d2v[2][c]) + penal[c];
            }
            rhs[3] = -c02 * (d1v[0][0] + d1v[1][1] + d1v[2][2]) - a.gamma_p * p;
            if (a.use_sponge && a.mask) {
                const double spo = a.mask[g0 + 5 * CS] * a.C_sponge_inv;
                rhs[0] = rhs[0] - (u - a.u_mean_set[0]) * spo;
                rhs[1] = rhs[1] - (v - a.u_mean_set[1]) * spo;
                rhs[2] = rhs[2] - (w - a.u_mean_set[2]) * spo;
                rhs[3] = rhs[3] - p * spo;
            }
            uabs_max = fmax(uabs_max, fmax(fmax(fabs(ctr[0]), fabs(ctr[1])), fmax(fabs(ctr[2]), fabs(ctr[3]))));
            double un[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (a.k_out) a.k_out[gi + c * CS] = rhs[c];
                if (a.u_out) {
                    double acc = base_u_global ? a.u0[gi + c * CS] : ctr[c];
                    for (int l = 0; l < a.n_prev; ++l) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_prev[l]), a.k_prev[l][gi + c * CS]));
                    if (a.use_self) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_self), rhs[c]));
                    a.u_out[gi + c * CS] = acc;
                    un[c] = acc;
                }
                if (a.acc_out) {
                    double acc = base_acc_global ? a.acc_in[gi + c * CS] : ctr[c];
                    if (a.use_acc) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_acc), rhs[c]));
                    a.acc_out[gi + c * CS] = acc;
                }
            }
            if (a.dtmin_bits && a.u_out)
                umag_max = fmax(umag_max, __dadd_rn(__dadd_rn(__dmul_rn(un[0], un[0]), __dmul_rn(un[1], un[1])), __dmul_rn(un[2], un[2])));
        }
        __syncthreads();   // plane z leaves the ring in the next iteration
    }
    const int lane = tid & 31, wid = tid >> 5;
    uabs_max = warp_max(uabs_max);
    umag_max = warp_max(umag_max);
    if (lane == 0) s_red[wid] = uabs_max;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < 8; ++i) m = fmax(m, s_red[i]);
        if (m > 1.0e12) atomicExch(a.diverged, 1);
    }
    if (a.dtmin_bits && a.u_out) {
        __syncthreads();
        if (lane == 0) s_red[wid] = umag_max;
        __syncthreads();
        if (tid == 0) {
            double m = 0.0;
            for (int i = 0; i < 8; ++i) m = fmax(m, s_red[i]);
            const double u_eigen = __dadd_rn(sqrt(m), sqrt(__dadd_rn(c02, m)));
            double dxmin = dx;
            if (a.dim_min_axes > 1) dxmin = fmin(dxmin, dy);
            if (a.dim_min_axes > 2) dxmin = fmin(dxmin, dz);
            double dtb = (u_eigen >= 1.0e-6) ? __ddiv_rn(__dmul_rn(a.CFL, dxmin), u_eigen) : 1.0e-2;
            if (a.nu > 1.0e-13) dtb = fmin(dtb, __ddiv_rn(__dmul_rn(a.CFL_nu, __dmul_rn(dxmin, dxmin)), a.nu));
            atomicMin(a.dtmin_bits, (unsigned long long)__double_as_longlong(dtb));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// 2-D stage kernel: RHS_2D_acm (LIB/EQUATION/ACMnew/rhs_ACM.f90:292-922) + the same Runge-Kutta epilogue.
// One CTA per block; the whole ghosted tile (3 components, Bs+2H squared, at most ~40 KB at Bs=32) is staged in shared memory
// with the four face halos gathered from the neighbours' interiors (star stencils do not read corners).  Bs is a run-time
// value (the reference's 2-D cases use 26 and 32).  HBM-bound: reads 3*(Bs^2 + 4*H*Bs) + bases, writes 3*Bs^2 per output.
// ---------------------------------------------------------------------------------------------
template <int FD, bool SKEW>
__global__ void __launch_bounds__(256) stage_kernel_2d(const __grid_constant__ StageArgs a, int BS)
{
    using S = St<FD>;
    constexpr int H = S::H, NC = 3;
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_red[8];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = a.active[blockIdx.x];
    const int PITCH = BS + 2 * H, PLANE = PITCH * PITCH;
    const long long CS = (long long)BS * BS;
    const int cW = a.nbr[b * WGPU_NDIR + 12], cE = a.nbr[b * WGPU_NDIR + 14], cS = a.nbr[b * WGPU_NDIR + 10], cN = a.nbr[b * WGPU_NDIR + 16];

    // interior + four face strips; corners of the tile are never read
    for (int i = tid; i < NC * BS * BS; i += nt) {
        const int c = i / (BS * BS), r = i % (BS * BS), y = r / BS, x = r % BS;
        sm[c * PLANE + (y + H) * PITCH + x + H] = a.u_in[((long long)b * NC + c) * CS + r];
    }
    for (int i = tid; i < NC * 4 * H * BS; i += nt) {
        const int c = i / (4 * H * BS), r = i % (4 * H * BS), side = r / (H * BS), q = r % (H * BS);
        int sx, sy, tx, ty, nb;
        if (side == 0) { const int h = q % H, y = q / H; nb = cW; sx = BS - H + h; sy = y; tx = h; ty = y + H; }
        else if (side == 1) { const int h = q % H, y = q / H; nb = cE; sx = h; sy = y; tx = BS + H + h; ty = y + H; }
        else if (side == 2) { const int x = q % BS, h = q / BS; nb = cS; sx = x; sy = BS - H + h; tx = x + H; ty = h; }
        else { const int x = q % BS, h = q / BS; nb = cN; sx = x; sy = h; tx = x + H; ty = BS + H + h; }
        double v = 0.0;
        if (nb >= 0) v = a.u_in[((long long)nb * NC + c) * CS + sy * BS + sx];
        else if (nb <= -2) {   // face patch of the jump pool (restriction / prediction across a level jump): x faces (H, Bs), y faces (Bs, H)
            const unsigned long long base = pool_base(a, nb);
            const double *pp = (base & LT_POOL) ? a.pool : a.jpool;
            const int hh = side < 2 ? q % H : q / BS, tt = side < 2 ? q / H : q % BS;
            v = pp[(long long)(base & LT_MASK) + (side < 2 ? ((long long)c * BS + tt) * H + hh : ((long long)c * H + hh) * BS + tt)];
        }
        sm[c * PLANE + ty * PITCH + tx] = v;
    }
    __syncthreads();

    const int lvl = a.level[b];
    const double dx = a.dx_lvl[lvl][0], dy = a.dx_lvl[lvl][1];
    const double dinv[2] = {1.0 / dx, 1.0 / dy};
    const double d2inv[2] = {1.0 / (dx * dx), 1.0 / (dy * dy)};
    const double dt = (a.u_out || a.acc_out) ? *a.dt_ptr : 0.0;
    const bool base_u_global = a.u_out && a.u0 != a.u_in;
    const bool base_acc_global = a.acc_out && a.acc_in != a.u_in;
    const double c02 = a.c0 * a.c0;
    double umag_max = 0.0, uabs_max = 0.0;

    for (int i = tid; i < BS * BS; i += nt) {
        const int ty = i / BS, tx = i % BS;
        const int cidx = (ty + H) * PITCH + tx + H;
        const long long gi = ((long long)b * NC) * CS + i;
        double d1v[2][3], d2v[2][2], cv[2][2], ctr[3];
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            double qv[3][2 * H + 1];
#pragma unroll
            for (int o = -H; o <= H; ++o) {
                const int off = cidx + (dir == 0 ? o : o * PITCH);
#pragma unroll
                for (int c = 0; c < 3; ++c) qv[c][o + H] = sm[off + c * PLANE];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) d1v[dir][c] = S::d1(&qv[c][H], dinv[dir]);
#pragma unroll
            for (int c = 0; c < 2; ++c) d2v[dir][c] = S::d2(&qv[c][H], d2inv[dir]);
            if (SKEW) {
#pragma unroll
                for (int c = 0; c < 2; ++c) cv[dir][c] = S::d1p(&qv[c][H], &qv[dir][H], dinv[dir]);
            }
            if (dir == 0) {
#pragma unroll
                for (int c = 0; c < 3; ++c) ctr[c] = qv[c][H];
            }
        }
        const double u = ctr[0], v = ctr[1], p = ctr[2];
        double penal[2] = {0.0, 0.0};
        const long long g0 = ((long long)b * a.n_mask) * CS + i;
        if (a.mask) {   // rhs_ACM.f90:600-602
            const int color = (int)a.mask[g0 + 4 * CS];
            const double chi = a.mask[g0] * (color == 0 ? 0.0 : a.C_eta_inv);
            penal[0] = -chi * (u - a.mask[g0 + 1 * CS]);
            penal[1] = -chi * (v - a.mask[g0 + 2 * CS]);
        }
        double rhs[3];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double adv;
            if (SKEW) adv = -0.5 * (cv[0][c] + cv[1][c] + u * d1v[0][c] + v * d1v[1][c]);   // rhs_ACM.f90:574-576
            else adv = -u * d1v[0][c] - v * d1v[1][c];                                      // rhs_ACM.f90:604-605
            rhs[c] = adv - d1v[c][2] + a.nu * (d2v[0][c] + d2v[1][c]) + penal[c];
        }
        rhs[2] = -c02 * (d1v[0][0] + d1v[1][1]) - a.gamma_p * p;
        if (a.use_sponge && a.mask) {   // rhs_ACM.f90:880-892
            const double spo = a.mask[g0 + 5 * CS] * a.C_sponge_inv;
            rhs[0] = rhs[0] - (u - a.u_mean_set[0]) * spo;
            rhs[1] = rhs[1] - (v - a.u_mean_set[1]) * spo;
            rhs[2] = rhs[2] - p * spo;
        }
        uabs_max = fmax(uabs_max, fmax(fmax(fabs(u), fabs(v)), fabs(p)));

        double un[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            if (a.k_out) a.k_out[gi + c * CS] = rhs[c];
            if (a.u_out) {
                double acc = base_u_global ? a.u0[gi + c * CS] : ctr[c];
                for (int l = 0; l < a.n_prev; ++l)
                    acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_prev[l]), a.k_prev[l][gi + c * CS]));
                if (a.use_self) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_self), rhs[c]));
                a.u_out[gi + c * CS] = acc;
                un[c] = acc;
            }
            if (a.acc_out) {
                double acc = base_acc_global ? a.acc_in[gi + c * CS] : ctr[c];
                if (a.use_acc) acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dt, a.coef_acc), rhs[c]));
                a.acc_out[gi + c * CS] = acc;
            }
        }
        if (a.dtmin_bits && a.u_out) umag_max = fmax(umag_max, __dadd_rn(__dmul_rn(un[0], un[0]), __dmul_rn(un[1], un[1])));
    }

    const int lane = tid & 31, wid = tid >> 5, NW = nt >> 5;
    uabs_max = warp_max(uabs_max);
    umag_max = warp_max(umag_max);
    if (lane == 0) s_red[wid] = uabs_max;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < NW; ++i) m = fmax(m, s_red[i]);
        if (m > 1.0e12) atomicExch(a.diverged, 1);
    }
    if (a.dtmin_bits && a.u_out) {
        __syncthreads();
        if (lane == 0) s_red[wid] = umag_max;
        __syncthreads();
        if (tid == 0) {
            double m = 0.0;
            for (int i = 0; i < NW; ++i) m = fmax(m, s_red[i]);
            const double u_eigen = __dadd_rn(sqrt(m), sqrt(__dadd_rn(c02, m)));
            const double dxmin = fmin(dx, dy);
            double dtb = (u_eigen >= 1.0e-6) ? __ddiv_rn(__dmul_rn(a.CFL, dxmin), u_eigen) : 1.0e-2;
            if (a.nu > 1.0e-13) dtb = fmin(dtb, __ddiv_rn(__dmul_rn(a.CFL_nu, __dmul_rn(dxmin, dxmin)), a.nu));   // module_ACM.f90:669-671
            atomicMin(a.dtmin_bits, (unsigned long long)__double_as_longlong(dtb));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// GET_DT_BLOCK_ACM's block loop as a standalone reduction (used when no fused value is available,
// e.g. right after an upload):  dtmin = min_b CFL*min(dx_b)/(sqrt(umag_b) + sqrt(c0^2 + umag_b))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dtmin_kernel(const double *__restrict__ u, const int *__restrict__ active,
                                                    const signed char *__restrict__ level, StageArgs a, int nc, long long CS,
                                                    int dim, unsigned long long *dtmin_bits)
{
    __shared__ double s_red[8];
    const int b = active[blockIdx.x];
    const double *ub = u + (long long)b * nc * CS;
    double m = 0.0;
    for (long long i = threadIdx.x; i < CS; i += blockDim.x) {
        double v = __dadd_rn(__dmul_rn(ub[i], ub[i]), __dmul_rn(ub[i + CS], ub[i + CS]));
        if (dim == 3) v = __dadd_rn(v, __dmul_rn(ub[i + 2 * CS], ub[i + 2 * CS]));
        m = fmax(m, v);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, s_red[i]);
        const int lvl = level[b];
        double dxmin = a.dx_lvl[lvl][0];
        for (int d = 1; d < dim; ++d) dxmin = fmin(dxmin, a.dx_lvl[lvl][d]);
        const double c02 = a.c0 * a.c0;
        const double u_eigen = __dadd_rn(sqrt(m), sqrt(__dadd_rn(c02, m)));
        double dtb = (u_eigen >= 1.0e-6) ? __ddiv_rn(__dmul_rn(a.CFL, dxmin), u_eigen) : 1.0e-2;
        if (a.nu > 1.0e-13) dtb = fmin(dtb, __ddiv_rn(__dmul_rn(a.CFL_nu, __dmul_rn(dxmin, dxmin)), a.nu));   // module_ACM.f90:669-671
        atomicMin(dtmin_bits, (unsigned long long)__double_as_longlong(dtb));
    }
}

// calculate_time_step (LIB/TIME/calculate_time_step.f90:19-118) after the global MIN; single thread.
struct DtArgs {
    double time, dt_fixed, dt_max, time_max, write_time, write_time_first, tsave_stats;
    double CFL_eta, gamma_p, C_eta, C_sponge;
    int penalization, use_sponge, write_fixed_time;
    double *time_dev;   // != nullptr: the time lives on the device (wgpu_rk_steps): read it here and advance it by dt
};

__global__ void dt_finalize_kernel(DtArgs p, const unsigned long long *dtmin_bits, unsigned long long *dtmin_next, double *dt_out,
                                   double *dt_host)
{
    double dt = 9.0e9;
    if (p.dt_fixed > 0.0) {
        dt = p.dt_fixed;
    } else {
        dt = fmin(dt, __longlong_as_double((long long)*dtmin_bits));
        // module_ACM.f90:673-689: the block-independent limits.  The diffusion limit CFL_nu dx^2 / nu depends on the block's dx and is part
        // of the per-block candidate (stage_kernel / dtmin_kernel), so that it takes part in the MIN over the blocks of ALL ranks
        if (p.gamma_p > 0) dt = fmin(dt, __dmul_rn(p.CFL_eta, p.gamma_p));
        if (p.penalization) dt = fmin(dt, __dmul_rn(p.CFL_eta, p.C_eta));
        if (p.use_sponge) dt = fmin(dt, __dmul_rn(p.CFL_eta, p.C_sponge));
        if (p.dt_max > 0.0) dt = fmin(p.dt_max, dt);
    }
    const double time = p.time_dev ? *p.time_dev : p.time;
    if (p.write_fixed_time) {
        if (fmod(time + dt, p.write_time) < fmod(time + 1e-12, p.write_time) && !(fabs(fmod(time, p.write_time)) < 1e-12) &&
            time + 1e-12 > p.write_time_first)
            dt = p.write_time - fmod(time, p.write_time);
    }
    if (fabs(p.tsave_stats - 9999999.9) > 1e-3) {
        if (fmod(time + dt, p.tsave_stats) < fmod(time + 1e-12, p.tsave_stats) && !(fabs(fmod(time, p.tsave_stats)) < 1e-12))
            dt = p.tsave_stats - fmod(time, p.tsave_stats);
    }
    if (time + dt > p.time_max && time <= p.time_max) dt = p.time_max - time;
    *dt_out = dt;
    if (p.time_dev) p.time_dev[1] = time + dt;   // [0] stays the time at the start of the step (stage times of the analytic mask)
    if (dt_host) *dt_host = dt;
    if (dtmin_next) *dtmin_next = 0x7FF0000000000000ULL;  // +inf
}

// ---------------------------------------------------------------------------------------------
// host layout <-> resident layout
// ---------------------------------------------------------------------------------------------
// staged: [n][ncomp_host][nz][ny][nx] ghosted (Fortran hvy(:,:,:,:,k)); dst: compact [blk][ncomp_dst][Bs^3]
// Both layout kernels are persistent: a capped number of CTAs strides over the work items (block, component, chunk of 256 points).
// On page-locked host arrays they are PCIe-bound and need few SMs; the cap leaves room for a second transfer kernel in the opposite
// direction and for the stage kernel of another tree to run at the same time (full-duplex link, see bench.py's end-to-end leg).
__global__ void __launch_bounds__(256) extract_kernel(const double *__restrict__ staged, double *__restrict__ dst, const int *__restrict__ ids,
                                                      int n, int nc, int ncomp_dst, int ncomp_host, int Bx, int By, int Bz, int g, int gz, int by_id)
{
    const int nx = Bx + 2 * g, ny = By + 2 * g, nz = Bz + 2 * gz;
    const long long CS = (long long)Bx * By * Bz;
    const long long nchunk = (CS + blockDim.x - 1) / blockDim.x, total = nchunk * nc * n;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const long long e = (w % nchunk) * blockDim.x + threadIdx.x;
        if (e >= CS) continue;
        const int c = (int)((w / nchunk) % nc), k = (int)(w / (nchunk * nc));
        const int b = ids[k];
        const int i = by_id ? b : k;       // which staged block (by_id: `staged` is the whole host array)
        const int x = e % Bx, y = (e / Bx) % By, z = e / ((long long)Bx * By);
        const double *s = staged + ((long long)i * ncomp_host + c) * nx * ny * nz;
        dst[((long long)b * ncomp_dst + c) * CS + e] = s[((long long)(z + gz) * ny + (y + g)) * nx + (x + g)];
    }
}

// compact -> ghosted staging, ghost shell of width gs gathered from same-level neighbours (all 26 relations:
// the copy sync_ghosts_generic stage 1 performs, synchronize_ghosts_generic.f90:266-339)
__global__ void __launch_bounds__(256) export_kernel(const double *__restrict__ src, double *__restrict__ staged, const int *__restrict__ ids,
                                                     const int *__restrict__ nbr, int n, int nc, int ncomp_src, int ncomp_host, int Bx, int By, int Bz,
                                                     int g, int gz, int gs, int gsz, int by_id)
{
    const int nx = Bx + 2 * g, ny = By + 2 * g, nz = Bz + 2 * gz;
    const int ex = Bx + 2 * gs, ey = By + 2 * gs, ez = Bz + 2 * gsz;
    const long long npts = (long long)ex * ey * ez, CS = (long long)Bx * By * Bz;
    const long long nchunk = (npts + blockDim.x - 1) / blockDim.x, total = nchunk * nc * n;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const long long e = (w % nchunk) * blockDim.x + threadIdx.x;
        if (e >= npts) continue;
        const int c = (int)((w / nchunk) % nc), k = (int)(w / (nchunk * nc));
        const int b = ids[k];
        const int i = by_id ? b : k;
        int x = (int)(e % ex) - gs, y = (int)((e / ex) % ey) - gs, z = (int)(e / ((long long)ex * ey)) - gsz;
        const int dxi = x < 0 ? -1 : (x >= Bx ? 1 : 0), dyi = y < 0 ? -1 : (y >= By ? 1 : 0), dzi = z < 0 ? -1 : (z >= Bz ? 1 : 0);
        int sb = b;
        if (dxi | dyi | dzi) {
            sb = nbr[b * WGPU_NDIR + (dzi + 1) * 9 + (dyi + 1) * 3 + (dxi + 1)];
            if (sb < 0) continue;   // no direct same-level source: leave the staged value untouched
        }
        const int xs = x - dxi * Bx, ys = y - dyi * By, zs = z - dzi * Bz;
        double *d = staged + ((long long)i * ncomp_host + c) * nx * ny * nz;
        d[((long long)(z + gz) * ny + (y + g)) * nx + (x + g)] = src[((long long)sb * ncomp_src + c) * CS + ((long long)zs * By + ys) * Bx + xs];
    }
}

// ---------------------------------------------------------------------------------------------
// Copy-engine transfers of page-locked host arrays (3-D): instead of SM-issued loads / stores over PCIe, the DMA engines move, for every
// interior xy plane of a block, the contiguous SPAN from the first to the last interior node of that plane ((By-1)*nx + Bx doubles: the
// interior rows and the x ghost nodes between them; 1.35x the interior at Bs=16, g=3) between the host array and a device staging buffer
// -- one cudaMemcpy3DAsync per run of consecutive blocks (row = span, height = Bz planes, depth = components x blocks).  H2D and D2H use
// different engines and, unlike SM-issued accesses, run at full rate in both directions at once.  These two kernels convert between the
// staging layout [k][c][z][span_pitch] and the resident layout; the download side fills the x ghost nodes inside the span with the
// same-level x neighbours' values (what sync_ghosts leaves there) or 0 where there is no such neighbour on this GPU.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) span_unpack_kernel(const double *__restrict__ stg, double *__restrict__ dst, const int *__restrict__ ids,
                                                          int n, int nc, int Bx, int By, int Bz, int nx, long long pitch)
{
    const int plane = Bx * By;
    const long long nchunk = (plane + blockDim.x - 1) / blockDim.x, total = nchunk * Bz * nc * n;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int e = (int)(w % nchunk) * blockDim.x + threadIdx.x;
        if (e >= plane) continue;
        const long long pz = w / nchunk;                 // (k * nc + c) * Bz + z
        const int k = (int)(pz / ((long long)nc * Bz));
        const long long cz = pz % ((long long)nc * Bz);
        const int x = e % Bx, y = e / Bx;
        dst[((long long)ids[k] * nc * Bz + cz) * plane + e] = stg[pz * pitch + y * nx + x];
    }
}

__global__ void __launch_bounds__(256) span_pack_kernel(const double *__restrict__ src, double *__restrict__ stg, const int *__restrict__ ids,
                                                        const int *__restrict__ nbr, int n, int nc, int Bx, int By, int Bz, int nx, long long pitch)
{
    const int plane = Bx * By, span = (By - 1) * nx + Bx;
    const long long nchunk = (span + blockDim.x - 1) / blockDim.x, total = nchunk * Bz * nc * n;
    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int e = (int)(w % nchunk) * blockDim.x + threadIdx.x;
        if (e >= span) continue;
        const long long pz = w / nchunk;
        const int k = (int)(pz / ((long long)nc * Bz));
        const long long cz = pz % ((long long)nc * Bz);
        const int b = ids[k];
        int x = e % nx, y = e / nx, sb = b;
        if (x >= Bx) {
            if (x < Bx + (nx - Bx) / 2) {       // right ghost nodes of row y
                sb = nbr[b * WGPU_NDIR + 14];
                x -= Bx;
            } else {                            // left ghost nodes of row y + 1
                sb = nbr[b * WGPU_NDIR + 12];
                x += Bx - nx;
                y += 1;
            }
        }
        stg[pz * pitch + e] = sb >= 0 ? src[((long long)sb * nc * Bz + cz) * plane + y * Bx + x] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Pack kernel: the sender side of the inter-GPU ghost exchange (send_prepare_external,
// LIB/MPI/xfer_block_data.f90:106-246, same-level relations).  One CTA per face patch: copies the g_rhs-deep interior
// strip facing the remote neighbour into the send buffer, already in the layout of the receiver's ghost strip, so the
// receiver's stage kernel reads it in place.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kernel(const double *__restrict__ src, double *__restrict__ send, const int *__restrict__ blk,
                                                   const int *__restrict__ dir, int nc, int Bs, int H)
{
    const int i = blockIdx.x;
    const int b = blk[i], d = dir[i];
    const int dx = d % 3 - 1, dy = (d / 3) % 3 - 1, dz = d / 9 - 1;   // direction from the SENDER to the receiver
    // strip extents in the sender's interior
    const int ex = dx ? H : Bs, ey = dy ? H : Bs, ez = dz ? H : Bs;
    const int x0 = dx > 0 ? Bs - H : 0, y0 = dy > 0 ? Bs - H : 0, z0 = dz > 0 ? Bs - H : 0;
    const long long CS = (long long)Bs * Bs * Bs;
    const int n = ex * ey * ez;
    double *out = send + (long long)i * nc * n;
    const double *in = src + (long long)b * nc * CS;
    for (int e = threadIdx.x; e < nc * n; e += blockDim.x) {
        const int c = e / n, r = e % n, x = r % ex, y = (r / ex) % ey, z = r / (ex * ey);
        out[e] = in[c * CS + ((long long)(z0 + z) * Bs + (y0 + y)) * Bs + (x0 + x)];
    }
}

// ---------------------------------------------------------------------------------------------
// Pack + put: the exchange itself, fused into the pack kernel.  One CTA per face patch writes the strip straight into the RECEIVER's patch
// pool over NVLink (peer memory opened through CUDA IPC, multigpu.cu), in the layout of the receiver's ghost strip.  The last CTA to finish
// the patches of a peer releases that peer's flag word (value = stage sequence number); wait_flags_kernel on the receiver acquires it in
// front of the partition-boundary blocks.  No send buffer, no NCCL kernel, no unpack: one store per ghost value crosses the link.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_put_kernel(const double *__restrict__ src, const int *__restrict__ blk, const int *__restrict__ dir,
                                                       const int *__restrict__ peer, const int *__restrict__ idx, double *const *__restrict__ put_base,
                                                       unsigned *const *__restrict__ put_flag, unsigned *__restrict__ done,
                                                       const int *__restrict__ n_to_peer, unsigned seq, int nc, int Bs, int H)
{
    const int i = blockIdx.x;
    const int b = blk[i], d = dir[i], p = peer[i];
    const int dx = d % 3 - 1, dy = (d / 3) % 3 - 1, dz = d / 9 - 1;
    const int ex = dx ? H : Bs, ey = dy ? H : Bs, ez = dz ? H : Bs;
    const int x0 = dx > 0 ? Bs - H : 0, y0 = dy > 0 ? Bs - H : 0, z0 = dz > 0 ? Bs - H : 0;
    const long long CS = (long long)Bs * Bs * Bs;
    const int n = ex * ey * ez;
    double *out = put_base[p] + (long long)idx[i] * nc * n;
    const double *in = src + (long long)b * nc * CS;
    for (int e = threadIdx.x; e < nc * n; e += blockDim.x) {
        const int c = e / n, r = e % n, x = r % ex, y = (r / ex) % ey, z = r / (ex * ey);
        out[e] = in[c * CS + ((long long)(z0 + z) * Bs + (y0 + y)) * Bs + (x0 + x)];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&done[p], 1u);
        if (prev + 1u == (unsigned)n_to_peer[p]) {
            done[p] = 0;               // every CTA of this launch for peer p has passed: ready for the next launch
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(put_flag[p]), "r"(seq) : "memory");
        }
    }
}

// one thread per peer: spin until the peer's patches of stage `seq` have landed (bounded: 5 s, then flags[5] = 1 and the step reports it)
__global__ void wait_flags_kernel(const unsigned *flags, const int *__restrict__ recv_cnt, int world, unsigned seq, int *err)
{
    const int p = threadIdx.x;
    if (p >= world || recv_cnt[p] == 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + p) : "memory");
        if ((int)(v - seq) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 5000000000ull) {
            *err = 1;
            break;
        }
        __nanosleep(200);
    }
}

// ---------------------------------------------------------------------------------------------
// tensor maps of a resident array for stage_kernel_tma (cuTensorMapEncodeTiled through the runtime's driver entry point: libcuda is not
// linked).  Rank 4: x, y, z, (block * NC + component); boxes Bs x Bs (a plane), Bs x H (y halo rows), XW x Bs (x halo columns).
// ---------------------------------------------------------------------------------------------
struct TmaMaps {
    CUtensorMap m[3];
};
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
        cudaGetLastError();
    }
    return fn;
}

const TmaMaps *tma_maps(wgpu_ctx *ctx, const double *base, int BS, int H)
{
    typedef std::unordered_map<const void *, TmaMaps> Cache;
    if (!ctx->tma_cache) ctx->tma_cache = new Cache();
    Cache &cache = *(Cache *)ctx->tma_cache;
    auto it = cache.find(base);
    if (it != cache.end()) return &it->second;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return nullptr;
    const int XW = (H + 1) & ~1;
    const cuuint64_t dims[4] = {(cuuint64_t)BS, (cuuint64_t)BS, (cuuint64_t)BS, (cuuint64_t)ctx->nc * (cuuint64_t)ctx->cfg.max_blocks};
    const cuuint64_t strides[3] = {(cuuint64_t)BS * 8, (cuuint64_t)BS * BS * 8, (cuuint64_t)BS * BS * BS * 8};
    const cuuint32_t boxes[3][4] = {{(cuuint32_t)BS, (cuuint32_t)BS, 1, 1}, {(cuuint32_t)BS, (cuuint32_t)H, 1, 1}, {(cuuint32_t)XW, (cuuint32_t)BS, 1, 1}};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    TmaMaps t;
    for (int k = 0; k < 3; ++k)
        if (enc(&t.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void *)base, dims, strides, boxes[k], estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return nullptr;
    return &cache.emplace(base, t).first->second;
}

template <int FD, bool SKEW, int BS>
int32_t launch_stage_t(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    using T = Tile<FD, BS>;
    static bool configured = false;
    if (!configured) {
        WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel<FD, SKEW, BS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
        if (FD == 4) WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel<FD, SKEW, BS, (FD == 4)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
        configured = true;
    }
    const bool prof = ctx->profiling && ctx->prof_n < (int)ctx->prof_ev.size() / 2;
    if (prof) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream);
    if (a.geom && FD != 4) {
        ctx->err = "analytic mask: the stage kernel is instantiated for FD_4th_central only";
        return WGPU_ERR_UNSUPPORTED;
    }
    StageArgs rest = a;
    bool rest_needed = true;
    if constexpr (FD == 4 && BS == 16) {
        // plain blocks through the TMA kernel; whatever it leaves (blocks at level jumps, partition or domain boundaries) through the
        // cp.async kernel with skip_plain.  plain_hint: 1 every block of this launch is plain, -1 none is (skip the TMA launch)
        // Measured on B200 (32 768 blocks, RK4 step): cp.async kernel 21.1 ms; TMA with one issuing thread and the block barrier 21.6 ms; TMA
        // with the copies dealt over the warps and per-slot empty barriers (no block barrier) 26.9 ms -- the 16-byte-wide x-halo boxes and
        // 20 small copies per plane cost the TMA unit more than the loader instructions cost the SMs.  Opt-in: WGPU_STAGE_TMA=1 (block barrier,
        // shared issue) or 2 (empty barriers).
        static const int mode = getenv("WGPU_STAGE_TMA") ? atoi(getenv("WGPU_STAGE_TMA")) : 0;
        const TmaMaps *tm = (mode > 0 && a.plain_hint >= 0) ? tma_maps(ctx, a.u_in, BS, St<FD>::H) : nullptr;
        if (tm) {
            static bool tconf = false;
            if (!tconf) {
                WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel_tma<FD, SKEW, BS, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaTile<FD, BS, 0>::SMEM));
                WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel_tma<FD, SKEW, BS, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaTile<FD, BS, 0>::SMEM));
                WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel_tma<FD, SKEW, BS, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaTile<FD, BS, 1>::SMEM));
                WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel_tma<FD, SKEW, BS, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmaTile<FD, BS, 1>::SMEM));
                tconf = true;
            }
            constexpr int NT = TmaTile<FD, BS, 0>::NT;
            if (mode == 1) {
                if (a.geom) stage_kernel_tma<FD, SKEW, BS, true, 0><<<n_blocks, NT, TmaTile<FD, BS, 0>::SMEM, ctx->stream>>>(a, tm->m[0], tm->m[1], tm->m[2]);
                else stage_kernel_tma<FD, SKEW, BS, false, 0><<<n_blocks, NT, TmaTile<FD, BS, 0>::SMEM, ctx->stream>>>(a, tm->m[0], tm->m[1], tm->m[2]);
            } else {
                if (a.geom) stage_kernel_tma<FD, SKEW, BS, true, 1><<<n_blocks, NT, TmaTile<FD, BS, 1>::SMEM, ctx->stream>>>(a, tm->m[0], tm->m[1], tm->m[2]);
                else stage_kernel_tma<FD, SKEW, BS, false, 1><<<n_blocks, NT, TmaTile<FD, BS, 1>::SMEM, ctx->stream>>>(a, tm->m[0], tm->m[1], tm->m[2]);
            }
            ctx->launches++;
            WGPU_CHECK(ctx, cudaGetLastError());
            rest.skip_plain = 1;
            rest_needed = a.plain_hint <= 0;
        }
    }
    if (rest_needed) {
        if constexpr (FD == 4 && BS == 16) {
            if (rest.skip_plain) {
                static bool sconf = false;
                if (!sconf) {
                    WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel<FD, SKEW, BS, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
                    WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel<FD, SKEW, BS, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM));
                    sconf = true;
                }
                if (a.geom) stage_kernel<FD, SKEW, BS, true, true><<<n_blocks, T::NT, T::SMEM, ctx->stream>>>(rest);
                else stage_kernel<FD, SKEW, BS, false, true><<<n_blocks, T::NT, T::SMEM, ctx->stream>>>(rest);
                ctx->launches++;
                rest_needed = false;
            }
        }
    }
    if (rest_needed) {
        if (a.geom) stage_kernel<FD, SKEW, BS, (FD == 4)><<<n_blocks, T::NT, T::SMEM, ctx->stream>>>(rest);
        else stage_kernel<FD, SKEW, BS, false><<<n_blocks, T::NT, T::SMEM, ctx->stream>>>(rest);
        ctx->launches++;
    }
    if (prof) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n++ + 1], ctx->stream);
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// any other even block size: run-time Bs, xy tiles of at most 16 x 16 points
template <int FD, bool SKEW>
int32_t launch_stage_any(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    const int Bs = ctx->cfg.Bs[0];
    if (Bs != ctx->cfg.Bs[1] || Bs != ctx->cfg.Bs[2]) {
        ctx->err = "3-D stage kernel: cubic blocks only";
        return WGPU_ERR_UNSUPPORTED;
    }
    constexpr int H = St<FD>::H;
    const int ntile = (Bs + 15) / 16, T = (Bs + ntile - 1) / ntile;
    const size_t smem = sizeof(double) * (size_t)(2 * H + 1) * 4 * (T + 2 * H) * (T + 2 * H);
    static bool configured = false;
    if (!configured) {
        WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel_any<FD, SKEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        configured = true;
    }
    const bool prof = ctx->profiling && ctx->prof_n < (int)ctx->prof_ev.size() / 2;
    if (prof) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream);
    const long long grid = (long long)n_blocks * ntile * ntile;
    stage_kernel_any<FD, SKEW><<<(unsigned)grid, 256, smem, ctx->stream>>>(a, Bs, T, ntile);
    if (prof) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n++ + 1], ctx->stream);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

template <int FD, bool SKEW>
int32_t launch_stage_bs(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    if (getenv("WGPU_STAGE_GENERIC")) return launch_stage_any<FD, SKEW>(ctx, a, n_blocks);   // tests compare the two paths
    switch (ctx->cfg.Bs[0]) {
    case 16: return launch_stage_t<FD, SKEW, 16>(ctx, a, n_blocks);
    case 18: return launch_stage_t<FD, SKEW, 18>(ctx, a, n_blocks);
    case 20: return launch_stage_t<FD, SKEW, 20>(ctx, a, n_blocks);
    default: return launch_stage_any<FD, SKEW>(ctx, a, n_blocks);
    }
}

template <int FD>
int32_t launch_stage_skew(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    return ctx->cfg.skew_symmetry ? launch_stage_bs<FD, true>(ctx, a, n_blocks) : launch_stage_bs<FD, false>(ctx, a, n_blocks);
}

}  // namespace

int32_t wgpu_launch_pack(wgpu_ctx *ctx, const double *src)
{
    if (ctx->n_send == 0) return WGPU_OK;
    const int H = ctx->cfg.fd == 2 ? 1 : (ctx->cfg.fd == 4 ? 2 : 3);   // halo depth the stage kernel gathers
    pack_kernel<<<ctx->n_send, 256, 0, ctx->stream>>>(src, ctx->d_send_buf, ctx->d_send_blk, ctx->d_send_dir, ctx->nc, ctx->cfg.Bs[0], H);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_pack_put(wgpu_ctx *ctx, const double *src, int parity, unsigned seq, cudaStream_t st)
{
    if (ctx->n_send == 0) return WGPU_OK;
    const int H = ctx->cfg.fd == 2 ? 1 : (ctx->cfg.fd == 4 ? 2 : 3);
    pack_put_kernel<<<ctx->n_send, 256, 0, st>>>(src, ctx->d_send_blk, ctx->d_send_dir, ctx->d_send_peer, ctx->d_send_idx,
                                                          ctx->d_put_base + (size_t)parity * ctx->comm_world, ctx->d_put_flag, ctx->d_done,
                                                          ctx->d_n_to_peer, seq, ctx->nc, ctx->cfg.Bs[0], H);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_wait_flags(wgpu_ctx *ctx, unsigned seq, cudaStream_t st)
{
    wait_flags_kernel<<<1, ((ctx->comm_world + 31) / 32) * 32, 0, st>>>((const unsigned *)ctx->p2p_mem, ctx->d_recv_cnt, ctx->comm_world, seq, ctx->d_flags + 5);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

namespace {

template <int FD, bool SKEW>
int32_t launch_stage_2d_t(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    const int Bs = ctx->cfg.Bs[0], H = St<FD>::H;
    const size_t smem = sizeof(double) * 3 * (size_t)(Bs + 2 * H) * (Bs + 2 * H);
    static size_t configured = 0;
    if (smem > configured) {
        if (smem > 227 * 1024) {
            ctx->err = "2-D stage kernel: block too large for shared memory";
            return WGPU_ERR_UNSUPPORTED;
        }
        WGPU_CHECK(ctx, cudaFuncSetAttribute(stage_kernel_2d<FD, SKEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const bool prof = ctx->profiling && ctx->prof_n < (int)ctx->prof_ev.size() / 2;
    if (prof) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n], ctx->stream);
    stage_kernel_2d<FD, SKEW><<<n_blocks, 256, smem, ctx->stream>>>(a, Bs);
    if (prof) cudaEventRecord(ctx->prof_ev[2 * ctx->prof_n++ + 1], ctx->stream);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

template <int FD>
int32_t launch_stage_2d(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    return ctx->cfg.skew_symmetry ? launch_stage_2d_t<FD, true>(ctx, a, n_blocks) : launch_stage_2d_t<FD, false>(ctx, a, n_blocks);
}

}  // namespace

int32_t wgpu_launch_stage(wgpu_ctx *ctx, const StageArgs &a, int n_blocks)
{
    if (n_blocks == 0) return WGPU_OK;
    if (ctx->cfg.dim == 2) {
        if (ctx->cfg.Bs[0] != ctx->cfg.Bs[1]) {
            ctx->err = "2-D stage kernel: square blocks only";
            return WGPU_ERR_UNSUPPORTED;
        }
        switch (ctx->cfg.fd) {
        case 2: return launch_stage_2d<2>(ctx, a, n_blocks);
        case 4: return launch_stage_2d<4>(ctx, a, n_blocks);
        case 6: return launch_stage_2d<6>(ctx, a, n_blocks);
        case 40: return launch_stage_2d<40>(ctx, a, n_blocks);
        }
        ctx->err = "unknown order_discretization id";
        return WGPU_ERR_UNSUPPORTED;
    }
    switch (ctx->cfg.fd) {
    case 2: return launch_stage_skew<2>(ctx, a, n_blocks);
    case 4: return launch_stage_skew<4>(ctx, a, n_blocks);
    case 6: return launch_stage_skew<6>(ctx, a, n_blocks);
    case 40: return launch_stage_skew<40>(ctx, a, n_blocks);
    }
    ctx->err = "unknown order_discretization id";
    return WGPU_ERR_UNSUPPORTED;
}

int32_t wgpu_launch_dtmin(wgpu_ctx *ctx, const double *u, unsigned long long *dtmin_bits)
{
    if (ctx->n_active == 0) return WGPU_OK;
    StageArgs a;
    memset(&a, 0, sizeof(a));
    for (int l = 0; l < WGPU_MAX_LEVELS; ++l)
        for (int d = 0; d < 3; ++d) a.dx_lvl[l][d] = ldexp(1.0, -l) * ctx->cfg.domain[d] / (double)ctx->cfg.Bs[d];
    a.c0 = ctx->cfg.c0;
    a.CFL = ctx->cfg.CFL;
    a.nu = ctx->cfg.nu;
    a.CFL_nu = ctx->cfg.CFL_nu;
    dtmin_kernel<<<ctx->n_active, 256, 0, ctx->stream>>>(u, ctx->d_active, ctx->d_level, a, ctx->nc, ctx->blk_elems, ctx->cfg.dim, dtmin_bits);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_dt_finalize(wgpu_ctx *ctx, double time, const unsigned long long *dtmin_bits, unsigned long long *dtmin_next)
{
    const wgpu_config &c = ctx->cfg;
    DtArgs p;
    p.time = time;
    p.time_dev = ctx->time_on_device ? ctx->d_time : nullptr;
    p.dt_fixed = c.dt_fixed;
    p.dt_max = c.dt_max;
    p.time_max = c.time_max;
    p.write_time = c.write_time;
    p.write_time_first = c.write_time_first;
    p.tsave_stats = c.tsave_stats;
    p.CFL_eta = c.CFL_eta;
    p.gamma_p = c.gamma_p;
    p.C_eta = c.C_eta;
    p.C_sponge = c.C_sponge;
    p.penalization = c.penalization;
    p.use_sponge = c.use_sponge;
    p.write_fixed_time = c.write_method_fixed_time;
    dt_finalize_kernel<<<1, 1, 0, ctx->stream>>>(p, dtmin_bits, dtmin_next, ctx->d_dt, nullptr);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// CTAs of a layout kernel: a few per SM when the other side is page-locked host memory (PCIe-bound; WGPU_XFER_CTAS overrides), many
// when it is the device staging buffer (HBM-bound)
static unsigned xfer_ctas(long long total, int by_id)
{
    static int host_cap = 0;
    if (!host_cap) {
        const char *e = getenv("WGPU_XFER_CTAS");
        host_cap = e && atoi(e) > 0 ? atoi(e) : 148 * 4;
    }
    const long long cap = by_id ? host_cap : 148 * 32;
    return (unsigned)(total < cap ? total : cap);
}

int32_t wgpu_launch_extract(wgpu_ctx *ctx, const double *staged, double *dst, const int *d_ids, int n, int ncomp_dst, int ncomp_host, int by_id)
{
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    const int Bz = c.dim == 3 ? c.Bs[2] : 1, gz = c.dim == 3 ? c.g : 0;
    const int nc = ncomp_dst < ncomp_host ? ncomp_dst : ncomp_host;
    const long long total = ((ctx->blk_elems + 255) / 256) * nc * n;
    extract_kernel<<<xfer_ctas(total, by_id), 256, 0, ctx->stream>>>(staged, dst, d_ids, n, nc, ncomp_dst, ncomp_host, c.Bs[0], c.Bs[1], Bz, c.g, gz,
                                                                     by_id);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_export(wgpu_ctx *ctx, const double *src, double *staged, const int *d_ids, int n, int ncomp_src, int ncomp_host,
                           int g_sync, int by_id)
{
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    const int Bz = c.dim == 3 ? c.Bs[2] : 1, gz = c.dim == 3 ? c.g : 0, gsz = c.dim == 3 ? g_sync : 0;
    const int nc = ncomp_src < ncomp_host ? ncomp_src : ncomp_host;
    const long long npts = (long long)(c.Bs[0] + 2 * g_sync) * (c.Bs[1] + 2 * g_sync) * (Bz + 2 * gsz);
    const long long total = ((npts + 255) / 256) * nc * n;
    export_kernel<<<xfer_ctas(total, by_id), 256, 0, ctx->stream>>>(src, staged, d_ids, ctx->d_nbr, n, nc, ncomp_src, ncomp_host, c.Bs[0], c.Bs[1], Bz,
                                                                    c.g, gz, g_sync, gsz, by_id);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_span_unpack(wgpu_ctx *ctx, const double *stg, double *dst, const int *d_ids, int n, int nc, long long pitch, cudaStream_t st)
{
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    const long long total = (long long)((c.Bs[0] * c.Bs[1] + 255) / 256) * c.Bs[2] * nc * n;
    span_unpack_kernel<<<(unsigned)std::min<long long>(total, 148 * 16), 256, 0, st>>>(stg, dst, d_ids, n, nc, c.Bs[0], c.Bs[1], c.Bs[2], c.Bs[0] + 2 * c.g, pitch);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_span_pack(wgpu_ctx *ctx, const double *src, double *stg, const int *d_ids, int n, int nc, long long pitch, cudaStream_t st)
{
    if (n == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    const int nx = c.Bs[0] + 2 * c.g;
    const long long total = (long long)(((c.Bs[1] - 1) * nx + c.Bs[0] + 255) / 256) * c.Bs[2] * nc * n;
    span_pack_kernel<<<(unsigned)std::min<long long>(total, 148 * 16), 256, 0, st>>>(src, stg, d_ids, ctx->d_nbr, n, nc, c.Bs[0], c.Bs[1], c.Bs[2], nx, pitch);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// RungeKuttaChebychev stage update (LIB/TIME/runge_kutta_chebychev.f90:90-127): elementwise over the interiors of the active blocks,
//   mode 0 (Euler start): out = y0 + (mu~ dt) F0
//   mode 1 (stage i):     out = (1 - mu - nu) y00 + mu y1 + nu y0 + (mu~ dt) F1 + (gamma~ dt) F0      (left to right, never contracted)
// dt is read from the device.  The right-hand sides come from the stage kernel (ghost synchronisation fused); this first implementation
// keeps the update a separate streaming pass (5 reads + 1 write per point) instead of a third epilogue of the stage kernel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rkc_combine_kernel(double *__restrict__ out, const double *__restrict__ y00, const double *__restrict__ y1,
                                                          const double *__restrict__ y0, const double *__restrict__ f1, const double *__restrict__ f0,
                                                          const int *__restrict__ active, long long per_block, double cA, double cB, double cC, double cD,
                                                          double cE, const double *__restrict__ dt_ptr, int mode)
{
    const long long base = (long long)active[blockIdx.y] * per_block;
    const double dt = *dt_ptr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_block; i += (long long)gridDim.x * blockDim.x) {
        const long long g = base + i;
        double v;
        if (mode == 0) v = __dadd_rn(y0[g], __dmul_rn(__dmul_rn(cD, dt), f0[g]));
        else {
            v = __dmul_rn(cA, y00[g]);
            v = __dadd_rn(v, __dmul_rn(cB, y1[g]));
            v = __dadd_rn(v, __dmul_rn(cC, y0[g]));
            v = __dadd_rn(v, __dmul_rn(__dmul_rn(cD, dt), f1[g]));
            v = __dadd_rn(v, __dmul_rn(__dmul_rn(cE, dt), f0[g]));
        }
        out[g] = v;
    }
}

int32_t wgpu_launch_rkc_combine(wgpu_ctx *ctx, double *out, const double *y00, const double *y1, const double *y0, const double *f1, const double *f0,
                                double cA, double cB, double cC, double cD, double cE, int mode)
{
    if (ctx->n_active == 0) return WGPU_OK;
    const long long per_block = (long long)ctx->nc * ctx->blk_elems;
    for (int s0 = 0; s0 < ctx->n_active; s0 += 32768) {      // grid.y limit
        dim3 grid((unsigned)std::min<long long>((per_block + 255) / 256, 16), (unsigned)std::min(32768, ctx->n_active - s0));
        rkc_combine_kernel<<<grid, 256, 0, ctx->stream>>>(out, y00, y1, y0, f1, f0, ctx->d_active + s0, per_block, cA, cB, cC, cD, cE, ctx->d_dt, mode);
        ctx->launches++;
    }
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

void wgpu_tma_release(wgpu_ctx *ctx)
{
    if (ctx->tma_cache) delete (std::unordered_map<const void *, TmaMaps> *)ctx->tma_cache;
    ctx->tma_cache = nullptr;
}
