// Wavelet side of the block hot path on sm_100a: biorthogonal CDF decomposition / reconstruction with the ghost
// synchronisation fused in (halo gathered from the neighbours' interiors), detail norms, thresholding flags and the
// component-wise Linfty norm.
//
// Reference: waveletDecomposition_optimized_block / waveletReconstruction_optimized_block
//            (LIB/WAVELETS/wavelet_decomposition_reconstruction.f90:23-389, 426-840), setup_wavelet
//            (LIB/WAVELETS/module_wavelets.f90:1031-1417), wavelet_renorm_block (:1848-1960), threshold_block
//            (LIB/INDICATORS/threshold_block.f90:1-130), componentWiseNorm_tree (LIB/OPERATORS/componentWiseNorm_tree.f90:63-197).
//
// Arithmetic that feeds refinement flags (FWT -> renorm -> max|wc| -> compare) is written with __dmul_rn/__dadd_rn so
// that nvcc cannot contract it: the reference build has no FMA (LIB/fortran.mk:72-84) and flags must be bit-exact.
// Every filter is one product per non-zero tap, summed in increasing tap order (the reference's hard-coded and generic
// branches coincide under this rule).
//
// Kernel shape: one CTA per (block, component); the (Bs+2f)^2 input planes (f = filter half width) stream through
// shared memory (cp.async), each plane is transformed in x then y, the xy-transformed planes sit in a ring of
// 2f+2 planes from which the z transform emits two output planes every second input plane.  HBM-bound: reads the
// (Bs+2f)^3 box once (mostly L2 hits for the halo: neighbours are launched back to back in space-filling-curve order),
// writes Bs^3.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "wgpu_internal.cuh"

namespace {

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

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// sum over non-zero taps, increasing index, product then add, never contracted
__device__ __forceinline__ double filt(const double *p, int stride, const double *F, int lo, int hi)
{
    double acc = 0.0;
    bool first = true;
    for (int k = lo; k <= hi; ++k) {
        const double c = F[k + WGPU_FMAX];
        if (c == 0.0) continue;
        const double t = __dmul_rn(p[k * stride], c);
        acc = first ? t : __dadd_rn(acc, t);
        first = false;
    }
    return acc;
}

struct WaveArgs {
    const double *src;
    double *dst;
    const int *active;
    const int *nbr;      // d_nbr, or d_wnbr on grids with level jumps (codes <= -2: patch of the wavelet jump pool)
    const double *wpool;
    const long long *woff;
    int fw;              // depth of the pool patches (wjump_depth >= f)
    int nc, Bs, f;       // f = halo depth gathered = max filter half width of the transform direction
    int inverse;         // 0: decomposition (HD/GD), 1: reconstruction (HR/GR on zero-stuffed SC/WC), 2: plain filter (blockFilterXYZ_vct with
                         //    the stencil in w.HD[hd_lo..hd_hi]: every point, x then y then z, sums from 0 in increasing shift order)
    WaveFilters w;
    // mode 2 (filter_wrapper, LIB/TIME/filter_wrapper.f90): which blocks / components are filtered; the others are copied
    const signed char *level;
    int Jmax, level_mode;    // 0 every block, 1 only blocks on Jmax (filter_only_maxlevel), 2 all except those (filter_all_except_maxlevel)
    unsigned comp_mask;      // bit c: component c is filtered (filter_component)
};

// source of a ghosted coordinate c in [-f, Bs+f): direction -1/0/+1 and local coordinate
__device__ __forceinline__ void split(int c, int Bs, int &d, int &l)
{
    d = c < 0 ? -1 : (c >= Bs ? 1 : 0);
    l = c - d * Bs;
}

__global__ void __launch_bounds__(256) wavelet_kernel(const WaveArgs a)
{
    extern __shared__ __align__(16) double sm[];
    const int Bs = a.Bs, f = a.f, n = Bs + 2 * f;
    const int R = 2 * f + 2;                 // ring of xy-transformed planes
    double *in0 = sm;                        // [2][n*n] input planes (double buffered)
    double *xs = in0 + 2 * n * n;            // [n][Bs]  after the x transform
    double *ring = xs + n * Bs;              // [R][Bs*Bs]
    __shared__ int s_code[WGPU_NDIR];

    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = a.active[blockIdx.x], c = blockIdx.y;
    if (tid < WGPU_NDIR) s_code[tid] = tid == 13 ? b : a.nbr[b * WGPU_NDIR + tid];
    __syncthreads();
    const long long CS = (long long)Bs * Bs * Bs;
    if (a.inverse == 2) {
        const int lv = a.level[b];
        const bool skip = !((a.comp_mask >> c) & 1u) || (a.level_mode == 1 && lv < a.Jmax) || (a.level_mode == 2 && lv == a.Jmax);
        if (skip) {      // this block / component is not filtered: it keeps its values
            const double *sp = a.src + ((long long)b * a.nc + c) * CS;
            double *dp = a.dst + ((long long)b * a.nc + c) * CS;
            for (long long i = tid; i < CS; i += nt) dp[i] = sp[i];
            return;
        }
    }

    auto load_plane = [&](int zp, double *dstp) {
        int dz, lz;
        split(zp, Bs, dz, lz);
        for (int i = tid; i < n * n; i += nt) {
            const int y = i / n - f, x = i % n - f;
            int dy, ly, dx, lx;
            split(y, Bs, dy, ly);
            split(x, Bs, dx, lx);
            const int sb = s_code[(dz + 1) * 9 + (dy + 1) * 3 + (dx + 1)];
            if (sb >= 0) cp_async8(dstp + i, a.src + ((long long)sb * a.nc + c) * CS + ((long long)lz * Bs + ly) * Bs + lx);
            else if (sb <= -2) {
                // ghost patch across a level jump (jump_fill_kernel: decimation / prediction / coarse-extension values), fw deep, laid out
                // like the ghost region: extents (dx ? fw : Bs, ...), origin at Bs - fw on the low sides
                const int fw = a.fw;
                const int ex = dx ? fw : Bs, ey = dy ? fw : Bs, ez = dz ? fw : Bs;
                const int ox = dx < 0 ? Bs - fw : 0, oy = dy < 0 ? Bs - fw : 0, oz = dz < 0 ? Bs - fw : 0;
                cp_async8(dstp + i, a.wpool + a.woff[-2 - sb] + (long long)c * ex * ey * ez + ((long long)(lz - oz) * ey + (ly - oy)) * ex + (lx - ox));
            } else dstp[i] = 0.0;   // no neighbour (non-periodic boundary)
        }
    };

    const bool inv = a.inverse == 1;
    const double *F0 = inv ? a.w.HR : a.w.HD, *F1 = inv ? a.w.GR : a.w.GD;
    const int lo0 = inv ? a.w.hr_lo : a.w.hd_lo, hi0 = inv ? a.w.hr_hi : a.w.hd_hi;
    const int lo1 = inv ? a.w.gr_lo : a.w.gd_lo, hi1 = inv ? a.w.gr_hi : a.w.gd_hi;

    // 1-D transform of a line at interior offset o (0..Bs-1); p points at the line element of offset o.
    // decomposition: even offsets -> HD (scaling), odd -> GD (wavelet).
    // reconstruction: u(o) = sum_{k: o+k even} sc(o+k) HR(k) + sum_{k: o+k odd} wc(o+k) GR(k)   (zero stuffing)
    auto line = [&](const double *p, int stride, int o) -> double {
        if (a.inverse == 2) {
            double s = 0.0;
            for (int k = lo0; k <= hi0; ++k) s = __dadd_rn(s, __dmul_rn(p[k * stride], F0[k + WGPU_FMAX]));
            return s;
        }
        if (!a.inverse) return (o & 1) ? filt(p, stride, F1, lo1, hi1) : filt(p, stride, F0, lo0, hi0);
        double s0 = 0.0, s1 = 0.0;
        for (int k = lo0; k <= hi0; ++k)
            if (((o + k) & 1) == 0) s0 = __dadd_rn(s0, __dmul_rn(p[k * stride], F0[k + WGPU_FMAX]));
        for (int k = lo1; k <= hi1; ++k)
            if (((o + k) & 1) != 0) s1 = __dadd_rn(s1, __dmul_rn(p[k * stride], F1[k + WGPU_FMAX]));
        return __dadd_rn(s0, s1);
    };

    load_plane(-f, in0);
    cp_async_commit();
    for (int q = 0; q < n; ++q) {             // q-th input plane, zp = q - f
        const int zp = q - f;
        double *cur = in0 + (q & 1) * n * n;
        if (q + 1 < n) load_plane(zp + 1, in0 + ((q + 1) & 1) * n * n);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        // x transform: rows y = -f..Bs+f-1, outputs at interior x
        for (int i = tid; i < n * Bs; i += nt) {
            const int r = i / Bs, o = i % Bs;
            xs[i] = line(cur + r * n + f + o, 1, o);
        }
        __syncthreads();
        // y transform on interior x
        double *rp = ring + (q % R) * Bs * Bs;
        for (int i = tid; i < Bs * Bs; i += nt) {
            const int o = i / Bs, x = i % Bs;
            rp[i] = line(xs + (o + f) * Bs + x, Bs, o);
        }
        __syncthreads();
        // z transform: once the plane zp = k + f + 1 is in the ring (k even), output planes k and k+1 are complete
        const int k = zp - f - 1;
        if (k >= 0 && k < Bs && (k & 1) == 0) {
            for (int i = tid; i < 2 * Bs * Bs; i += nt) {
                const int o = k + i / (Bs * Bs), xy = i % (Bs * Bs);
                // ring index of interior plane z is (z + f) % R; walk with explicit modulo
                double acc;
                {
                    // gather the taps through the ring: emulate `line` with a modular stride
                    const double *F = nullptr;
                    int lo, hi;
                    if (a.inverse == 2) {
                        double s = 0.0;
                        for (int t = lo0; t <= hi0; ++t) s = __dadd_rn(s, __dmul_rn(ring[((o + t + f) % R) * Bs * Bs + xy], F0[t + WGPU_FMAX]));
                        acc = s;
                    } else if (!a.inverse) {
                        F = (o & 1) ? F1 : F0;
                        lo = (o & 1) ? lo1 : lo0;
                        hi = (o & 1) ? hi1 : hi0;
                        double s = 0.0;
                        bool first = true;
                        for (int t = lo; t <= hi; ++t) {
                            const double cf = F[t + WGPU_FMAX];
                            if (cf == 0.0) continue;
                            const double v = __dmul_rn(ring[((o + t + f) % R) * Bs * Bs + xy], cf);
                            s = first ? v : __dadd_rn(s, v);
                            first = false;
                        }
                        acc = s;
                    } else {
                        double s0 = 0.0, s1 = 0.0;
                        for (int t = lo0; t <= hi0; ++t)
                            if (((o + t) & 1) == 0) s0 = __dadd_rn(s0, __dmul_rn(ring[((o + t + f) % R) * Bs * Bs + xy], F0[t + WGPU_FMAX]));
                        for (int t = lo1; t <= hi1; ++t)
                            if (((o + t) & 1) != 0) s1 = __dadd_rn(s1, __dmul_rn(ring[((o + t + f) % R) * Bs * Bs + xy], F1[t + WGPU_FMAX]));
                        acc = __dadd_rn(s0, s1);
                    }
                }
                a.dst[((long long)b * a.nc + c) * CS + (long long)o * Bs * Bs + xy] = acc;
            }
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Two-dimensional blocks (dim = 2, TESTING/acm 2-D cases): the same transform without the z pass.  One CTA per (block, component):
// the ghosted tile (halo depth f) is gathered from the same-level neighbours' interiors or, across level jumps, from the patches of
// the wavelet jump pool (jump_fill_kernel: decimation / prediction / coarse-extension values in ghost-region layout), then the x pass
// and the y pass run in shared memory.  Arithmetic as above: one product per non-zero tap, summed in increasing tap order.
// ---------------------------------------------------------------------------------------------
struct Wave2dArgs {
    const double *src;
    double *dst;
    const int *active;
    const int *nbr;            // d_wnbr (codes <= -2: patch of the wavelet jump pool) or d_nbr
    const double *wpool;
    const long long *woff;
    int nc, Bs, f, inverse;
    WaveFilters w;
};

__global__ void __launch_bounds__(256) wavelet2d_kernel(const Wave2dArgs a)
{
    extern __shared__ __align__(16) double sm[];
    const int Bs = a.Bs, f = a.f, n = Bs + 2 * f;
    double *tile = sm;                       // [n][n]
    double *xs = tile + n * n;               // [n][Bs]
    __shared__ const double *s_ptr[9];
    __shared__ int s_sy[9];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = a.active[blockIdx.x], c = blockIdx.y;
    const long long CS = (long long)Bs * Bs;
    if (tid < 9) {
        const int D = 9 + tid;               // dz = 0 plane of the 27 directions
        const int sb = D == 13 ? b : a.nbr[b * WGPU_NDIR + D];
        const double *ptr = nullptr;
        int sy = Bs;
        if (sb >= 0) ptr = a.src + ((long long)sb * a.nc + c) * CS;
        else if (sb <= -2) {
            const int d[2] = {tid % 3 - 1, tid / 3 - 1};
            const int ex = d[0] ? f : Bs, ey = d[1] ? f : Bs;
            const int ox = d[0] < 0 ? Bs - f : 0, oy = d[1] < 0 ? Bs - f : 0;
            sy = ex;
            ptr = a.wpool + a.woff[-2 - sb] + (long long)c * ex * ey - ((long long)oy * ex + ox);
        }
        s_ptr[tid] = ptr;
        s_sy[tid] = sy;
    }
    __syncthreads();
    for (int i = tid; i < n * n; i += nt) {
        const int y = i / n - f, x = i % n - f;
        int dy, ly, dx, lx;
        split(y, Bs, dy, ly);
        split(x, Bs, dx, lx);
        const int D = (dy + 1) * 3 + (dx + 1);
        const double *p = s_ptr[D];
        tile[i] = p ? p[(long long)ly * s_sy[D] + lx] : 0.0;
    }
    __syncthreads();
    const double *F0 = a.inverse ? a.w.HR : a.w.HD, *F1 = a.inverse ? a.w.GR : a.w.GD;
    const int lo0 = a.inverse ? a.w.hr_lo : a.w.hd_lo, hi0 = a.inverse ? a.w.hr_hi : a.w.hd_hi;
    const int lo1 = a.inverse ? a.w.gr_lo : a.w.gd_lo, hi1 = a.inverse ? a.w.gr_hi : a.w.gd_hi;
    auto line = [&](const double *p, int stride, int o) -> double {
        if (!a.inverse) return (o & 1) ? filt(p, stride, F1, lo1, hi1) : filt(p, stride, F0, lo0, hi0);
        double s0 = 0.0, s1 = 0.0;
        for (int k = lo0; k <= hi0; ++k)
            if (((o + k) & 1) == 0) s0 = __dadd_rn(s0, __dmul_rn(p[k * stride], F0[k + WGPU_FMAX]));
        for (int k = lo1; k <= hi1; ++k)
            if (((o + k) & 1) != 0) s1 = __dadd_rn(s1, __dmul_rn(p[k * stride], F1[k + WGPU_FMAX]));
        return __dadd_rn(s0, s1);
    };
    for (int i = tid; i < n * Bs; i += nt) {
        const int r = i / Bs, o = i % Bs;
        xs[i] = line(tile + r * n + f + o, 1, o);
    }
    __syncthreads();
    double *out = a.dst + ((long long)b * a.nc + c) * CS;
    for (int i = tid; i < Bs * Bs; i += nt) {
        const int o = i / Bs, x = i % Bs;
        out[i] = line(xs + (o + f) * Bs + x, Bs, o);
    }
}

// ---------------------------------------------------------------------------------------------
// Fast path: the same transform with everything the compiler can know fixed at compile time -- wavelet (taps become
// literals, zero taps vanish), block size, halo depth.  One CTA per (block, component), 256 threads.  Every thread produces
// a (scaling, wavelet) pair of neighbouring outputs from one register window of 2F+2 inputs, so each pass costs
// ~(taps_HD + taps_GD) multiply-adds per output pair and two vector shared-memory loads per tap pair.  Arithmetic order and
// rounding are those of the generic kernel above (and of the reference): one product per non-zero tap, summed in increasing
// tap order, never contracted.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr double cdf_interp(int order, int i)   // interpolation stencil, i in -(order-1)..(order-1)
{
    const int a = i < 0 ? -i : i;
    if (a == 0) return 1.0;
    if (order == 2) return a == 1 ? 0.5 : 0.0;
    if (order == 4) return a == 1 ? 9.0 / 16.0 : (a == 3 ? -1.0 / 16.0 : 0.0);
    if (order == 6) return a == 1 ? 150.0 / 256.0 : (a == 3 ? -25.0 / 256.0 : (a == 5 ? 3.0 / 256.0 : 0.0));
    return 0.0;
}
__host__ __device__ constexpr double cdf_HR(int X, int k) { return (k < -(X - 1) || k > X - 1) ? 0.0 : cdf_interp(X, k); }
__host__ __device__ constexpr double cdf_HD(int X, int Y, int k)
{
    if (Y == 0) return k == 0 ? 1.0 : 0.0;
    double v = k == 0 ? 1.0 : 0.0;
    for (int j = -(X - 1); j <= X - 1; ++j) {
        const int d = k - j;
        if (d < -(Y - 1) || d > Y - 1 || d == 0) continue;
        v = v + ((j % 2 == 0) ? 1.0 : -1.0) * cdf_HR(X, j) * cdf_interp(Y, d) / 2.0;
    }
    return v;
}
__host__ __device__ constexpr double cdf_GD(int X, int k) { return ((k % 2 == 0) ? 1.0 : -1.0) * cdf_HR(X, k); }
__host__ __device__ constexpr double cdf_GR(int X, int Y, int k) { return ((k % 2 == 0) ? 1.0 : -1.0) * cdf_HD(X, Y, k); }
__host__ __device__ constexpr int cdf_hd_half(int X, int Y) { return Y == 0 ? 0 : X - 1 + Y - 1; }

// the two outputs at offsets o (even) and o+1 from the window w[j] = in(o - F + j), j = 0 .. 2F+1
template <int X, int Y, bool INV, int F>
__device__ __forceinline__ void pair_out(const double (&w)[2 * F + 2], double &out0, double &out1)
{
    constexpr int HDH = cdf_hd_half(X, Y), HRH = X - 1;
    if (!INV) {
        double s = 0.0;
        bool first = true;
#pragma unroll
        for (int k = -HDH; k <= HDH; ++k) {          // scaling coefficient at o: HD
            const double c = cdf_HD(X, Y, k);
            if (c != 0.0) {
                const double t = c == 1.0 ? w[F + k] : __dmul_rn(w[F + k], c);   // x*1 == x exactly
                s = first ? t : __dadd_rn(s, t);
                first = false;
            }
        }
        out0 = s;
        s = 0.0;
        first = true;
#pragma unroll
        for (int k = -HRH; k <= HRH; ++k) {          // wavelet coefficient at o+1: GD
            const double c = cdf_GD(X, k);
            if (c != 0.0) {
                const double t = c == 1.0 ? w[F + 1 + k] : __dmul_rn(w[F + 1 + k], c);
                s = first ? t : __dadd_rn(s, t);
                first = false;
            }
        }
        out1 = s;
    } else {
#pragma unroll
        for (int p = 0; p < 2; ++p) {                // u(o+p) = sum_{k == p mod 2} HR(k) sc(o+p+k) + sum_{k != p mod 2} GR(k) wc(o+p+k)
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = -HRH; k <= HRH; ++k)
                if (((k + p) & 1) == 0) s0 = __dadd_rn(s0, cdf_HR(X, k) == 1.0 ? w[F + p + k] : __dmul_rn(w[F + p + k], cdf_HR(X, k)));
#pragma unroll
            for (int k = -HDH; k <= HDH; ++k)
                if (((k + p) & 1) != 0) s1 = __dadd_rn(s1, __dmul_rn(w[F + p + k], cdf_GR(X, Y, k)));
            (p ? out1 : out0) = __dadd_rn(s0, s1);
        }
    }
}

// one of the two outputs only (P = 0: offset o, P = 1: offset o+1), same arithmetic as pair_out
template <int X, int Y, bool INV, int F, int P>
__device__ __forceinline__ void pair_out_one(const double (&w)[2 * F + 2], double &out)
{
    constexpr int HDH = cdf_hd_half(X, Y), HRH = X - 1;
    if (!INV) {
        double s = 0.0;
        bool first = true;
        if (P == 0) {
#pragma unroll
            for (int k = -HDH; k <= HDH; ++k) {
                const double c = cdf_HD(X, Y, k);
                if (c != 0.0) {
                    const double t = c == 1.0 ? w[F + k] : __dmul_rn(w[F + k], c);   // x*1 == x exactly
                    s = first ? t : __dadd_rn(s, t);
                    first = false;
                }
            }
        } else {
#pragma unroll
            for (int k = -HRH; k <= HRH; ++k) {
                const double c = cdf_GD(X, k);
                if (c != 0.0) {
                    const double t = c == 1.0 ? w[F + 1 + k] : __dmul_rn(w[F + 1 + k], c);
                    s = first ? t : __dadd_rn(s, t);
                    first = false;
                }
            }
        }
        out = s;
    } else {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = -HRH; k <= HRH; ++k)
            if (((k + P) & 1) == 0) s0 = __dadd_rn(s0, cdf_HR(X, k) == 1.0 ? w[F + P + k] : __dmul_rn(w[F + P + k], cdf_HR(X, k)));
#pragma unroll
        for (int k = -HDH; k <= HDH; ++k)
            if (((k + P) & 1) != 0) s1 = __dadd_rn(s1, __dmul_rn(w[F + P + k], cdf_GR(X, Y, k)));
        out = __dadd_rn(s0, s1);
    }
}

template <int X, int Y, int BS, bool INV>
struct FastCfg {
    static constexpr int HDH = cdf_hd_half(X, Y), HRH = X - 1;
    static constexpr int F = HDH > HRH ? HDH : HRH;   // halo depth = widest filter of the transform
    static constexpr int N = BS + 2 * F;
    static constexpr int R = 2 * F + 2;               // planes a z output pair looks at
    static constexpr int NT = ((BS * BS + 31) / 32) * 32;   // one thread per (x, y) column of the block
    static constexpr int NLD = (F % 2 == 0) ? N * (N / 2) : N * N;   // 16-byte chunks (F even) or 8-byte elements per input plane
    static constexpr int LPT = (NLD + NT - 1) / NT;   // loads per thread and plane
    // row pitches with an odd number of 16-byte slots modulo 128 bytes: two rows handled by one quarter warp of 128-bit accesses (x pass: two
    // output pairs per thread) then hit disjoint banks
    static constexpr int NP = N + 2 + (((N + 2) / 2) % 2 == 0 ? 2 : 0);      // input planes
    static constexpr int XP = BS + 2 + (((BS + 2) / 2) % 2 == 0 ? 2 : 0);    // x-pass rows
    static constexpr size_t SMEM = sizeof(double) * ((size_t)6 * N * NP + (size_t)4 * N * XP);   // three plane pairs (one read, two in flight) + the x-pass rows of two pairs
};

// One CTA per (block, component).  Input planes (xy halo included) stream through a double buffer (cp.async); the x pass
// writes rows to a second double buffer; every thread then owns one (x, y) column: it computes the y pass for its column and
// keeps the last 2F+2 y-pass results in REGISTERS (the plane loop is fully unrolled, so the sliding window is pure register
// renaming), from which the z pass emits a (scaling, wavelet) output pair every second plane -- no shared-memory traffic for z.
// Threads are mapped to columns so that a warp has one y parity (all scaling rows or all wavelet rows: no divergence).
#ifndef WFAST_MINB_BIG
#define WFAST_MINB_BIG 2   // Bs = 18, 20 (352 / 416 threads per CTA): two resident CTAs instead of the one ptxas would settle for
#endif
#ifndef WFAST_XPT
#define WFAST_XPT 2    // output pairs per thread in the x pass (one window of 2F + 2 XPT values): 4 needs more than 80 registers
#endif
#ifndef WFAST_MINB
#define WFAST_MINB 3   // minimum resident CTAs per SM requested for Bs = 16 (register cap <= 80): measured best for the two-planes-per-barrier loop
#endif
template <int X, int Y, int BS, bool INV>
__global__ void __launch_bounds__(FastCfg<X, Y, BS, INV>::NT, (BS == 16 ? WFAST_MINB : (BS <= 20 ? WFAST_MINB_BIG : 1))) wavelet_fast_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                                                                 const int *__restrict__ active, const int *__restrict__ nbr, int nc,
                                                                                 double *__restrict__ det_abs, double *__restrict__ det_sq,
                                                                                 const double *__restrict__ wpool, const long long *__restrict__ woff)
{
    using C = FastCfg<X, Y, BS, INV>;
    constexpr int F = C::F, N = C::N, R = C::R, NT = C::NT, LPT = C::LPT, HALF = BS * BS / 2, NP = C::NP, XP = C::XP;
    extern __shared__ __align__(16) double sm[];
    double *in0 = sm;                       // [3 pairs][2 planes][N][NP]
    double *xs0 = in0 + 6 * N * NP;         // [2 pairs][2 planes][N][XP]
    // source of the ghost region of each direction: pointer such that the point with neighbour-local coordinates (lx, ly, lz)
    // sits at ptr + lz*sz + ly*sy + lx -- the neighbour's interior (same level), or a patch of the wavelet jump pool (level
    // jumps: decimated / predicted values in the layout of the ghost region), or null (no neighbour)
    __shared__ const double *s_ptr[WGPU_NDIR];
    __shared__ int s_sy[WGPU_NDIR], s_sz[WGPU_NDIR];
    const int tid = threadIdx.x;
    const int b = active[blockIdx.x], c = blockIdx.y;
    constexpr long long CS = (long long)BS * BS * BS;
    if (tid < WGPU_NDIR) {
        const int sb = tid == 13 ? b : nbr[b * WGPU_NDIR + tid];
        const double *ptr = nullptr;
        int sy = BS, sz = BS * BS;
        if (sb >= 0) ptr = src + ((long long)sb * nc + c) * CS;
        else if (sb <= -2) {
            const int d[3] = {tid % 3 - 1, (tid / 3) % 3 - 1, tid / 9 - 1};
            const int ex = d[0] ? F : BS, ey = d[1] ? F : BS, ez = d[2] ? F : BS;
            const int ox = d[0] < 0 ? BS - F : 0, oy = d[1] < 0 ? BS - F : 0, oz = d[2] < 0 ? BS - F : 0;
            sy = ex;
            sz = ex * ey;
            ptr = wpool + woff[-2 - sb] + (long long)c * ex * ey * ez - ((long long)(oz * ey + oy) * ex + ox);
        }
        s_ptr[tid] = ptr;
        s_sy[tid] = sy;
        s_sz[tid] = sz;
    }
    // per-thread load descriptors (the same for every plane): destination in the plane buffer, xy part of the direction, offset in the source plane
    int ld_dst[LPT], ld_src[LPT], ld_dir[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
        const int i = tid + j * NT;
        int r, x;
        if (F % 2 == 0) { r = i / (N / 2); x = 2 * (i % (N / 2)) - F; }
        else { r = i / N; x = i % N - F; }
        const int y = r - F;
        const int dy = y < 0 ? -1 : (y >= BS ? 1 : 0), dx = x < 0 ? -1 : (x >= BS ? 1 : 0);
        ld_dst[j] = i < C::NLD ? r * NP + x + F : -1;
        ld_src[j] = ((y - dy * BS) << 8) | (x - dx * BS);   // neighbour-local (ly, lx)
        ld_dir[j] = (dy + 1) * 3 + (dx + 1);
    }
    __syncthreads();
    // the planes of the block's own z range (dz = 0: BS of the N planes) read through per-thread pointers prepared once
    const double *ld_g0[LPT];
    int ld_sz0[LPT];
#pragma unroll
    for (int j = 0; j < LPT; ++j) {
        const int D = 9 + ld_dir[j];
        const double *base = ld_dst[j] >= 0 ? s_ptr[D] : nullptr;
        ld_g0[j] = base ? base + (ld_src[j] >> 8) * s_sy[D] + (ld_src[j] & 255) : nullptr;
        ld_sz0[j] = s_sz[D];
    }

    auto load_plane = [&](int zp, double *dstp) {
        const int dz = zp < 0 ? -1 : (zp >= BS ? 1 : 0);
        const int lz = zp - dz * BS;
#pragma unroll
        for (int j = 0; j < LPT; ++j) {
            if (ld_dst[j] < 0) continue;
            double *d = dstp + ld_dst[j];
            const double *g;
            if (dz == 0) g = ld_g0[j] ? ld_g0[j] + lz * ld_sz0[j] : nullptr;
            else {
                const int D = (dz + 1) * 9 + ld_dir[j];
                const double *base = s_ptr[D];
                g = base ? base + lz * s_sz[D] + (ld_src[j] >> 8) * s_sy[D] + (ld_src[j] & 255) : nullptr;
            }
            if (g) {
                const unsigned sa = (unsigned)__cvta_generic_to_shared(d);
                if (F % 2 == 0) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(g) : "memory");
                else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(g) : "memory");
            } else {
                d[0] = 0.0;
                if (F % 2 == 0) d[1] = 0.0;
            }
        }
    };

    // column of this thread: warps hold one y parity
    const bool has_col = tid < BS * BS;
    const int par = tid / HALF, idx = tid % HALF;
    const int cx = idx % BS, cy = 2 * (idx / BS) + par;
    double win[R];                          // y-pass results of planes zp-R+1 .. zp of this column; plane z sits in win[(z + F) % R]
#pragma unroll
    for (int j = 0; j < R; ++j) win[j] = 0.0;
    double m0 = 0.0;                        // Linfty detail of this (block, component), see below
    const bool pure_col = !(cx & 1) && !(cy & 1);   // o0 of this column is a pure scaling coefficient: not a detail
    double *outc = dst + ((long long)b * nc + c) * CS + cy * BS + cx;

    // Plane loop, two planes per iteration and ONE block barrier per iteration: iteration P runs the x pass of plane pair P (input ring ->
    // x-pass rows) and, software-pipelined behind it, the y and z passes of pair P-1 (x-pass rows of the previous iteration -> register
    // window -> output).  The x-pass rows are double-buffered, the input planes triple-buffered: the barrier at the top of an iteration
    // orders (i) the cp.async data of pair P, (ii) the x-pass rows written in iteration P-1 before their readers, and (iii) the end of every
    // thread's reads of the buffers that are refilled next -- pair P+2 is issued AFTER the barrier into the buffer pair P-1 occupied.
    // Rounds of R planes are fully unrolled so that the window slots and the z ordering are compile-time.
    load_plane(-F, in0);
    load_plane(-F + 1, in0 + N * NP);
    cp_async_commit();
    load_plane(-F + 2, in0 + 2 * N * NP);
    load_plane(-F + 3, in0 + 3 * N * NP);
    cp_async_commit();
#pragma unroll 1
    for (int q0 = 0; q0 <= N; q0 += R) {
#pragma unroll
    for (int jq = 0; jq < R; jq += 2) {
        const int q = q0 + jq;
        if (q > N) break;
        const int zp = q - F;
        const int P = q >> 1;
        cp_async_wait<1>();
        __syncthreads();
        if (q + 4 < N) {
            double *nxt = in0 + ((P + 2) % 3) * 2 * N * NP;
            load_plane(zp + 4, nxt);
            load_plane(zp + 5, nxt + N * NP);
        }
        cp_async_commit();
        if (q < N) {
            // x: rows y = -F .. BS+F-1 of both planes of pair P; every thread forms XPT adjacent output pairs from one window of 2F + 2 XPT values
            // (F + XPT 128-bit loads instead of XPT (F+1): the kernel is bound by shared-memory wavefronts, profiles/r5_wavelet_fast_kernel.txt)
            const double *cur = in0 + (P % 3) * 2 * N * NP;
            double *xw = xs0 + (P & 1) * 2 * N * XP;
            constexpr int PR = BS / 2, XPT = WFAST_XPT, QR = (PR + XPT - 1) / XPT;       // output pairs per row, pairs per thread, items per row
#pragma unroll
            for (int i0 = 0; i0 < 2 * N * QR; i0 += NT) {
                const int i = i0 + tid;
                if (i < 2 * N * QR) {
                    const int pl = i / (N * QR), ii = i % (N * QR);
                    const int r = ii / QR, pi = XPT * (ii % QR), o = 2 * pi;
                    double w[2 * F + 2 * XPT];
                    const double2 *p2 = reinterpret_cast<const double2 *>(cur + pl * N * NP + r * NP + o);
#pragma unroll
                    for (int j = 0; j < F + XPT; ++j) {
                        if (o + 2 * j < NP) {                       // the last item of a row may hold fewer than XPT pairs: stay inside the row
                            const double2 v = p2[j];
                            w[2 * j] = v.x;
                            w[2 * j + 1] = v.y;
                        } else w[2 * j] = w[2 * j + 1] = 0.0;
                    }
#pragma unroll
                    for (int t = 0; t < XPT; ++t) {
                        if (PR % XPT == 0 || pi + t < PR) {
                            double wt[2 * F + 2];
#pragma unroll
                            for (int j = 0; j < 2 * F + 2; ++j) wt[j] = w[j + 2 * t];
                            double2 out;
                            pair_out<X, Y, INV, F>(wt, out.x, out.y);
                            *reinterpret_cast<double2 *>(xw + pl * N * XP + r * XP + o + 2 * t) = out;
                        }
                    }
                }
            }
        }
        if (has_col && q >= 2) {
            // y of pair P-1 (planes q-2, q-1): one output per plane of this thread's column (scaling row if cy is even, wavelet row if odd)
            const double *xr = xs0 + ((P - 1) & 1) * 2 * N * XP;
            const int jy = (jq + R - 2) % R;              // window slots of planes q-2, q-1 (compile-time after unrolling)
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
                double w[2 * F + 2], o0, o1;
                const double *colp = xr + pl * N * XP + (cy - par) * XP + cx;     // window of the pair (cy - par, cy - par + 1)
                if (par == 0) {
#pragma unroll
                    for (int j = 0; j < 2 * F + 1; ++j) w[j] = colp[j * XP];
                    w[2 * F + 1] = 0.0;
                    pair_out_one<X, Y, INV, F, 0>(w, o0);
                    win[jy + pl] = o0;
                } else {
                    w[0] = 0.0;
#pragma unroll
                    for (int j = 1; j < 2 * F + 2; ++j) w[j] = colp[j * XP];
                    pair_out_one<X, Y, INV, F, 1>(w, o1);
                    win[jy + pl] = o1;
                }
            }
            // z: once plane (q-2) + 1 = k + F + 1 is in the window (k even), output planes k and k+1 of this column are complete
            const int k = q - 2 - 2 * F;       // its plane k - F sits in slot (jy + 2) % R = jq % R
            if (k >= 0 && k < BS) {
                double w[R], o0, o1;
#pragma unroll
                for (int j = 0; j < R; ++j) w[j] = win[(jq + j) % R];
                pair_out<X, Y, INV, F>(w, o0, o1);
                outc[(long long)k * BS * BS] = o0;
                outc[(long long)(k + 1) * BS * BS] = o1;
                if (!INV) {
                    // threshold_block's Linfty detail, fused: max |wc| over everything but the pure scaling positions (wavelet_renorm_block is
                    // the identity for eps_norm = Linfty); max sqrt(wc*wc) = sqrt(fl(max|wc|^2)) (squaring and rounding are monotone): formed
                    // once from m0 at the end
                    m0 = fmax(m0, fabs(o1));
                    if (!pure_col) m0 = fmax(m0, fabs(o0));
                }
            }
        }
    }
    }
    cp_async_wait<0>();
    if (!INV && det_abs) {
        __shared__ double s0[NT / 32];
        m0 = warp_max(m0);
        if ((tid & 31) == 0) s0[tid >> 5] = m0;
        __syncthreads();
        if (tid == 0) {
            for (int i = 1; i < NT / 32; ++i) m0 = fmax(m0, s0[i]);
            det_abs[(long long)b * nc + c] = m0;
            det_sq[(long long)b * nc + c] = sqrt(__dmul_rn(m0, m0));
        }
    }
}

template <int X, int Y, int BS, bool INV>
int32_t launch_fast_t(wgpu_ctx *ctx, const double *src, double *dst)
{
    using C = FastCfg<X, Y, BS, INV>;
    static bool configured = false;
    if (!configured) {
        WGPU_CHECK(ctx, cudaFuncSetAttribute(wavelet_fast_kernel<X, Y, BS, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    dim3 grid(ctx->n_active, ctx->nc);
    if (ctx->has_jumps && C::F != ctx->wjump_depth) {
        ctx->err = "wavelet kernel: halo depth differs from the depth of the level-jump patches";
        return WGPU_ERR_UNSUPPORTED;
    }
    wavelet_fast_kernel<X, Y, BS, INV><<<grid, C::NT, C::SMEM, ctx->stream>>>(src, dst, ctx->d_active, ctx->has_jumps ? ctx->d_wnbr : ctx->d_nbr, ctx->nc,
                                                                             ctx->d_det_abs, ctx->d_det_sq, ctx->d_wpool, ctx->d_woff);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    if (!INV) ctx->det_cached_for = dst;   // Linfty details of `dst` are in d_det_abs / d_det_sq until the array is written again
    return WGPU_OK;
}

template <int X, int Y>
int32_t launch_fast_bs(wgpu_ctx *ctx, const double *src, double *dst, int inverse, bool &handled)
{
    handled = true;
    switch (ctx->cfg.Bs[0]) {
    case 16: return inverse ? launch_fast_t<X, Y, 16, true>(ctx, src, dst) : launch_fast_t<X, Y, 16, false>(ctx, src, dst);
    case 18: return inverse ? launch_fast_t<X, Y, 18, true>(ctx, src, dst) : launch_fast_t<X, Y, 18, false>(ctx, src, dst);
    case 20: return inverse ? launch_fast_t<X, Y, 20, true>(ctx, src, dst) : launch_fast_t<X, Y, 20, false>(ctx, src, dst);
    case 24: return inverse ? launch_fast_t<X, Y, 24, true>(ctx, src, dst) : launch_fast_t<X, Y, 24, false>(ctx, src, dst);
    }
    handled = false;
    return WGPU_OK;
}

int32_t launch_fast(wgpu_ctx *ctx, const double *src, double *dst, int inverse, bool &handled)
{
    const int X = ctx->wavelet.X, Y = ctx->wavelet.Y;
    handled = false;
    if (getenv("WGPU_WAVELET_GENERIC")) return WGPU_OK;   // tests compare the two paths
    if (X == 2 && Y == 0) return launch_fast_bs<2, 0>(ctx, src, dst, inverse, handled);
    if (X == 2 && Y == 2) return launch_fast_bs<2, 2>(ctx, src, dst, inverse, handled);
    if (X == 4 && Y == 0) return launch_fast_bs<4, 0>(ctx, src, dst, inverse, handled);
    if (X == 4 && Y == 2) return launch_fast_bs<4, 2>(ctx, src, dst, inverse, handled);
    if (X == 4 && Y == 4) return launch_fast_bs<4, 4>(ctx, src, dst, inverse, handled);
    if (X == 6 && Y == 0) return launch_fast_bs<6, 0>(ctx, src, dst, inverse, handled);
    if (X == 6 && Y == 2) return launch_fast_bs<6, 2>(ctx, src, dst, inverse, handled);
    return WGPU_OK;
}

// ---------------------------------------------------------------------------------------------
// wavelet_renorm_block + threshold_block's detail norms on a decomposed (spaghetti-ordered) array:
// per block and component max|wc| (pure scaling positions removed), both as max(abs) and max(sqrt(x*x)).
// ---------------------------------------------------------------------------------------------
struct DetailArgs {
    const double *wd;
    const int *active;
    const signed char *level;
    double *det_abs, *det_sq;   // [max_blocks][nc]
    int nc, Bx, By, Bz, dim;
    int eps_norm;               // 0 Linfty, 1 L1, 2 L2, 3 H1
    double fac_lvl[WGPU_MAX_LEVELS];
    double fdir;
    int fdir_div;
};

__global__ void __launch_bounds__(256) detail_kernel(const DetailArgs a)
{
    __shared__ double s0[8], s1[8];
    const int b = a.active[blockIdx.x], c = blockIdx.y;
    const long long CS = (long long)a.Bx * a.By * a.Bz;
    const double *p = a.wd + ((long long)b * a.nc + c) * CS;
    const double fac = a.fac_lvl[a.level[b]];
    const bool scale = a.eps_norm != 0 && !(a.eps_norm == 3 && a.dim != 3);
    double m0 = -INFINITY, m1 = -INFINITY;
    for (long long i = threadIdx.x; i < CS; i += blockDim.x) {
        const int x = (int)(i % a.Bx), y = (int)((i / a.Bx) % a.By), z = (int)(i / ((long long)a.Bx * a.By));
        const bool px = !(x & 1), py = !(y & 1), pz = a.dim == 3 ? !(z & 1) : true;
        double v = p[i];
        if (px && py && pz) v = 0.0;
        if (scale) {
            v = __dmul_rn(v, fac);
            if (px) v = a.fdir_div ? __ddiv_rn(v, a.fdir) : __dmul_rn(v, a.fdir);
            if (py) v = a.fdir_div ? __ddiv_rn(v, a.fdir) : __dmul_rn(v, a.fdir);
            if (a.dim == 3 && pz) v = a.fdir_div ? __ddiv_rn(v, a.fdir) : __dmul_rn(v, a.fdir);
        }
        m0 = fmax(m0, fabs(v));
        m1 = fmax(m1, sqrt(__dmul_rn(v, v)));
    }
    m0 = warp_max(m0);
    m1 = warp_max(m1);
    if ((threadIdx.x & 31) == 0) {
        s0[threadIdx.x >> 5] = m0;
        s1[threadIdx.x >> 5] = m1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
            m0 = fmax(m0, s0[i]);
            m1 = fmax(m1, s1[i]);
        }
        a.det_abs[(long long)b * a.nc + c] = m0;
        a.det_sq[(long long)b * a.nc + c] = m1;
    }
}

// threshold_block with `indices` (LIB/INDICATORS/threshold_block.f90:30-44) for the security zone of adapt_tree (addSecurityZone_CE_tree,
// LIB/MESH/securityZone_tree.f90:140-298): Linfty detail of a decomposed block inside the strip that faces one neighbour direction
// (get_indices_of_modify_patch with Nwcl / Nwcr), pure scaling positions removed.  One CTA per (pair, component).
// wavelet_renorm_block inside the strip (eps_norm L1 / L2 / H1; module_wavelets.f90:1900-1945): the per-level factor and the per-direction
// factor of the pure directions, exactly as detail_kernel applies them to the whole block
struct PatchNorm {
    int eps_norm, fdir_div;
    double fdir;
    double fac_lvl[WGPU_MAX_LEVELS];
};

__global__ void __launch_bounds__(128) patch_detail_kernel(const double *__restrict__ wd, const int *__restrict__ blk, const int *__restrict__ dir,
                                                           double *__restrict__ out, int nc, int Bs, int dim, int Nl, int Nr,
                                                           const signed char *__restrict__ level, const PatchNorm pn)
{
    __shared__ double s0[4];
    const int b = blk[blockIdx.x], dc = dir[blockIdx.x], c = blockIdx.y;
    const int d[3] = {dc % 3 - 1, (dc / 3) % 3 - 1, dc / 9 - 1};
    const long long CS = (long long)Bs * Bs * (dim == 3 ? Bs : 1);
    int lo[3], ext[3];
    for (int k = 0; k < 3; ++k) {
        const int B = k < dim ? Bs : 1;
        lo[k] = d[k] > 0 ? B - Nr : 0;
        ext[k] = d[k] < 0 ? Nl : (d[k] > 0 ? Nr : B);
        if (lo[k] < 0) { ext[k] += lo[k]; lo[k] = 0; }
        if (ext[k] > B) ext[k] = B;
    }
    const int npts = ext[0] * ext[1] * ext[2];
    const double *p = wd + ((long long)b * nc + c) * CS;
    const bool scale = pn.eps_norm != 0 && !(pn.eps_norm == 3 && dim != 3);
    const double fac = scale ? pn.fac_lvl[level[b]] : 1.0;
    double m = -INFINITY;
    for (int i = threadIdx.x; i < npts; i += blockDim.x) {
        const int x = lo[0] + i % ext[0], y = lo[1] + (i / ext[0]) % ext[1], z = lo[2] + i / (ext[0] * ext[1]);
        const bool px = !(x & 1), py = !(y & 1), pz = dim == 3 ? !(z & 1) : true;
        double v = (px && py && pz) ? 0.0 : p[((long long)z * Bs + y) * Bs + x];
        if (scale) {
            v = __dmul_rn(v, fac);
            if (px) v = pn.fdir_div ? __ddiv_rn(v, pn.fdir) : __dmul_rn(v, pn.fdir);
            if (py) v = pn.fdir_div ? __ddiv_rn(v, pn.fdir) : __dmul_rn(v, pn.fdir);
            if (dim == 3 && pz) v = pn.fdir_div ? __ddiv_rn(v, pn.fdir) : __dmul_rn(v, pn.fdir);
        }
        m = fmax(m, fabs(v));
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s0[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, s0[i]);
        out[(long long)blockIdx.x * nc + c] = m;
    }
}

// threshold_block.f90:96-121: detail per component (own / joint group / ignored), status = -1 iff all(detail <= eps*norm)
struct FlagArgs {
    const int *active;
    const double *det_abs, *det_sq;
    int *status;          // [n_active]
    double *detail_out;   // [n_active][nc] or nullptr
    int n_active, nc;
    int thresh_comp[16];
    double eps_use[16];   // eps * norm
};

__global__ void flags_kernel(const FlagArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_active) return;
    const int b = a.active[i];
    double det[16];
    int maxgrp = 0;
    for (int c = 0; c < a.nc; ++c) {
        det[c] = -1.0;
        maxgrp = a.thresh_comp[c] > maxgrp ? a.thresh_comp[c] : maxgrp;
    }
    for (int l = 2; l <= maxgrp; ++l) {
        double m = -INFINITY;
        for (int c = 0; c < a.nc; ++c)
            if (a.thresh_comp[c] == l) m = fmax(m, a.det_sq[(long long)b * a.nc + c]);
        for (int c = 0; c < a.nc; ++c)
            if (a.thresh_comp[c] == l) det[c] = m;
    }
    int status = -1;
    for (int c = 0; c < a.nc; ++c) {
        if (a.thresh_comp[c] == 1) det[c] = a.det_abs[(long long)b * a.nc + c];
        if (a.thresh_comp[c] == 0) det[c] = 0.0;
        if (!(det[c] <= a.eps_use[c])) status = 0;
        if (a.detail_out) a.detail_out[(long long)i * a.nc + c] = det[c];
    }
    a.status[i] = status;
}

// componentWiseNorm_tree, Linfty: max |u| per component over all active blocks (atomicMax on the bit pattern)
__global__ void __launch_bounds__(256) linfty_kernel(const double *__restrict__ u, const int *__restrict__ active, int nc, long long CS,
                                                     unsigned long long *out)
{
    __shared__ double s[8];
    const int b = active[blockIdx.x], c = blockIdx.y;
    const double *p = u + ((long long)b * nc + c) * CS;
    double m = 0.0;
    for (long long i = threadIdx.x; i < CS; i += blockDim.x) m = fmax(m, fabs(p[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, s[i]);
        atomicMax(out + c, (unsigned long long)__double_as_longlong(m));
    }
}

// componentWiseNorm_tree, L1 / L2: per (block, component) sum of |u| or u^2 in a fixed order (strided partial sums, shuffle tree,
// warp partials in order), so the result does not depend on scheduling; the host adds the block sums weighted by the cell volume
// in the order of hvy_active, as the reference's loop does (componentWiseNorm_tree.f90:150-197)
__global__ void __launch_bounds__(256) blocksum_kernel(const double *__restrict__ u, const int *__restrict__ active, int nc, long long CS, int squared,
                                                       double *__restrict__ out)
{
    __shared__ double s[8];
    const int b = active[blockIdx.x], c = blockIdx.y;
    const double *p = u + ((long long)b * nc + c) * CS;
    double m = 0.0;
    for (long long i = threadIdx.x; i < CS; i += blockDim.x) m += squared ? p[i] * p[i] : fabs(p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m += s[i];
        out[(long long)blockIdx.x * nc + c] = m;
    }
}

}  // namespace

static void renorm_factors(const wgpu_config &c, int eps_norm, int level_ref, double *fac_lvl, double *fdir, int *fdir_div)
{
    *fdir = 1.0;
    *fdir_div = 0;
    for (int l = 0; l < WGPU_MAX_LEVELS; ++l) {
        double fac = 1.0;   // module_wavelets.f90:1900-1945
        if (eps_norm == 1) fac = pow(2.0, (double)((level_ref - l - 1) * c.dim));
        if (eps_norm == 2) fac = pow(2.0, (double)((level_ref - l - 1) * c.dim) / 2.0);
        if (eps_norm == 3) fac = pow(2.0, (double)(level_ref - l) * (2.0 - c.dim) / 2.0);
        fac_lvl[l] = fac;
    }
    if (eps_norm == 1) { *fdir = 4.0; *fdir_div = 1; }
    if (eps_norm == 2) { *fdir = 2.0; *fdir_div = 1; }
    if (eps_norm == 3) *fdir = pow(2.0, 2.0 * (c.dim - 2.0) / 3.0);
}

int32_t wgpu_launch_patch_detail(wgpu_ctx *ctx, const double *wd, const int *d_blk, const int *d_dir, int n, int Nl, int Nr, double *d_out, int eps_norm,
                                 int level_ref)
{
    PatchNorm pn;
    pn.eps_norm = eps_norm;
    renorm_factors(ctx->cfg, eps_norm, level_ref, pn.fac_lvl, &pn.fdir, &pn.fdir_div);
    for (int s0 = 0; s0 < n; s0 += 32768) {
        dim3 grid(std::min(32768, n - s0), ctx->nc);
        patch_detail_kernel<<<grid, 128, 0, ctx->stream>>>(wd, d_blk + s0, d_dir + s0, d_out + (long long)s0 * ctx->nc, ctx->nc, ctx->cfg.Bs[0],
                                                         ctx->cfg.dim, Nl, Nr, ctx->d_level, pn);
        ctx->launches++;
        WGPU_CHECK(ctx, cudaGetLastError());
    }
    return WGPU_OK;
}


int32_t wgpu_launch_blocksum(wgpu_ctx *ctx, const double *u, int squared, double *d_out)
{
    if (ctx->n_active == 0) return WGPU_OK;
    dim3 grid(ctx->n_active, ctx->nc);
    blocksum_kernel<<<grid, 256, 0, ctx->stream>>>(u, ctx->d_active, ctx->nc, ctx->blk_elems, squared, d_out);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// ce_coarse != nullptr: reconstruction inside wavelet_reconstruct_full_tree_CEoptimized -- the ghost patches that face a coarser leaf
// hold that leaf's values (array ce_coarse) at the scaling positions and zero wavelet coefficients (sync_SCWC_from_MC +
// coarse_extension_modify, LIB/MESH/adapt_tree.f90:686-987) instead of restricted / predicted values of `src`
static size_t g_wavelet_kernel_smem = 0;   // largest dynamic shared memory wavelet_kernel has been configured for (two launchers share the kernel)

int32_t wgpu_launch_wavelet(wgpu_ctx *ctx, const double *src, double *dst, int inverse, const double *ce_coarse)
{
    if (ctx->n_active == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    WaveArgs a;
    a.src = src;
    a.dst = dst;
    a.active = ctx->d_active;
    a.nbr = ctx->d_nbr;
    a.wpool = ctx->d_wpool;
    a.woff = ctx->d_woff;
    a.fw = 0;
    a.nc = ctx->nc;
    a.Bs = c.Bs[0];
    a.inverse = inverse;
    a.w = ctx->wavelet;
    const WaveFilters &w = ctx->wavelet;
    int f = 0;
    const int b4[4] = {inverse ? -w.hr_lo : -w.hd_lo, inverse ? w.hr_hi : w.hd_hi, inverse ? -w.gr_lo : -w.gd_lo, inverse ? w.gr_hi : w.gd_hi};
    for (int i = 0; i < 4; ++i) f = b4[i] > f ? b4[i] : f;
    a.f = f;
    if (f > a.Bs) {
        ctx->err = "wavelet filter wider than the block";
        return WGPU_ERR_UNSUPPORTED;
    }
    if (c.dim == 2) {
        if (ctx->has_jumps) {
            int32_t rcj = wgpu_launch_wjump_fill(ctx, src, ce_coarse);
            if (rcj) return rcj;
        }
        Wave2dArgs q;
        q.src = src;
        q.dst = dst;
        q.active = ctx->d_active;
        q.nbr = ctx->has_jumps ? ctx->d_wnbr : ctx->d_nbr;
        q.wpool = ctx->d_wpool;
        q.woff = ctx->d_woff;
        q.nc = ctx->nc;
        q.Bs = c.Bs[0];
        q.f = ctx->has_jumps ? ctx->wjump_depth : f;     // the patches of the wavelet jump pool are wjump_depth deep
        q.inverse = inverse;
        q.w = ctx->wavelet;
        if (q.f < f || q.f > q.Bs) {
            ctx->err = "wavelet2d: halo depth does not fit";
            return WGPU_ERR_UNSUPPORTED;
        }
        const int n2 = q.Bs + 2 * q.f;
        const size_t smem2 = sizeof(double) * ((size_t)n2 * n2 + (size_t)n2 * q.Bs);
        static size_t configured2 = 0;
        if (smem2 > configured2) {
            WGPU_CHECK(ctx, cudaFuncSetAttribute(wavelet2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            configured2 = smem2;
        }
        dim3 grid2(ctx->n_active, ctx->nc);
        wavelet2d_kernel<<<grid2, 256, smem2, ctx->stream>>>(q);
        ctx->launches++;
        WGPU_CHECK(ctx, cudaGetLastError());
        return WGPU_OK;
    }
    {
        bool handled = false;
        if (ctx->has_jumps) {   // ghost values across level jumps (all 26 relations): decimation / prediction into the wavelet jump pool
            int32_t rcj = wgpu_launch_wjump_fill(ctx, src, ce_coarse);
            if (rcj) return rcj;
        }
        int32_t rc = launch_fast(ctx, src, dst, inverse, handled);
        if (rc || handled) return rc;
        // any other (wavelet, even Bs): the table-driven kernel, ghost patches across level jumps read from the same pool
        if (ctx->has_jumps) {
            a.nbr = ctx->d_wnbr;
            a.fw = ctx->wjump_depth;
            if (a.fw < f) {
                ctx->err = "wavelet transform: the ghost patches are shallower than the filter";
                return WGPU_ERR_UNSUPPORTED;
            }
        }
    }
    const int n = a.Bs + 2 * f;
    const size_t smem = sizeof(double) * ((size_t)2 * n * n + (size_t)n * a.Bs + (size_t)(2 * f + 2) * a.Bs * a.Bs);
    if (smem > 227 * 1024) {
        ctx->err = "wavelet transform: block too large for the shared-memory ring of the table-driven kernel";
        return WGPU_ERR_UNSUPPORTED;
    }
    if (smem > g_wavelet_kernel_smem) {
        WGPU_CHECK(ctx, cudaFuncSetAttribute(wavelet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_wavelet_kernel_smem = smem;
    }
    dim3 grid(ctx->n_active, ctx->nc);
    wavelet_kernel<<<grid, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

// filter_wrapper's block loop (LIB/TIME/filter_wrapper.f90:62-77): blockFilterXYZ_vct with a symmetric stencil of half width `half` on the selected
// blocks / components, src -> dst (everything else copied); ghost values as sync_ghosts_tree leaves them (same level: neighbour interiors, level
// jumps: the wavelet jump pool).  3-D, table-driven kernel (the filter runs once per filter_freq time steps).
int32_t wgpu_launch_blockfilter(wgpu_ctx *ctx, const double *src, double *dst, const double *stencil, int half, unsigned comp_mask, int level_mode)
{
    if (ctx->n_active == 0) return WGPU_OK;
    const wgpu_config &c = ctx->cfg;
    if (c.dim != 3) {
        ctx->err = "filter: three-dimensional blocks only";
        return WGPU_ERR_UNSUPPORTED;
    }
    if (half < 1 || half > WGPU_FMAX || half > c.Bs[0]) {
        ctx->err = "filter: stencil half width out of range";
        return WGPU_ERR_UNSUPPORTED;
    }
    WaveArgs a;
    memset(&a, 0, sizeof(a));
    a.src = src;
    a.dst = dst;
    a.active = ctx->d_active;
    a.nbr = ctx->d_nbr;
    a.wpool = ctx->d_wpool;
    a.woff = ctx->d_woff;
    a.nc = ctx->nc;
    a.Bs = c.Bs[0];
    a.f = half;
    a.inverse = 2;
    a.w.hd_lo = -half;
    a.w.hd_hi = half;
    for (int k = -half; k <= half; ++k) a.w.HD[k + WGPU_FMAX] = stencil[k + half];
    a.level = ctx->d_level;
    a.Jmax = c.Jmax;
    a.level_mode = level_mode;
    a.comp_mask = comp_mask;
    if (ctx->has_jumps) {
        if (!ctx->wavelet_set) {
            ctx->err = "filter on a grid with level jumps: call wgpu_set_wavelet first (the ghost synchronisation is the wavelet's)";
            return WGPU_ERR_ARG;
        }
        int32_t rcj = wgpu_launch_wjump_fill(ctx, src, nullptr);
        if (rcj) return rcj;
        a.nbr = ctx->d_wnbr;
        a.fw = ctx->wjump_depth;
        if (a.fw < half) {
            ctx->err = "filter: the ghost patches of the wavelet are shallower than the filter stencil";
            return WGPU_ERR_UNSUPPORTED;
        }
    }
    const int n = a.Bs + 2 * half;
    const size_t smem = sizeof(double) * ((size_t)2 * n * n + (size_t)n * a.Bs + (size_t)(2 * half + 2) * a.Bs * a.Bs);
    if (smem > 227 * 1024) {
        ctx->err = "filter: block too large for the shared-memory ring";
        return WGPU_ERR_UNSUPPORTED;
    }
    if (smem > g_wavelet_kernel_smem) {
        WGPU_CHECK(ctx, cudaFuncSetAttribute(wavelet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_wavelet_kernel_smem = smem;
    }
    dim3 grid(ctx->n_active, ctx->nc);
    wavelet_kernel<<<grid, 256, smem, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_detail(wgpu_ctx *ctx, const double *wd, int eps_norm, int level_ref)
{
    if (ctx->n_active == 0) return WGPU_OK;
    if (eps_norm == 0 && ctx->det_cached_for == wd) return WGPU_OK;   // computed by the decomposition kernel itself
    ctx->det_cached_for = nullptr;
    const wgpu_config &c = ctx->cfg;
    DetailArgs a;
    a.wd = wd;
    a.active = ctx->d_active;
    a.level = ctx->d_level;
    a.det_abs = ctx->d_det_abs;
    a.det_sq = ctx->d_det_sq;
    a.nc = ctx->nc;
    a.Bx = c.Bs[0];
    a.By = c.Bs[1];
    a.Bz = c.dim == 3 ? c.Bs[2] : 1;
    a.dim = c.dim;
    a.eps_norm = eps_norm;
    renorm_factors(c, eps_norm, level_ref, a.fac_lvl, &a.fdir, &a.fdir_div);
    dim3 grid(ctx->n_active, ctx->nc);
    detail_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_flags(wgpu_ctx *ctx, const int32_t *thresh_comp, const double *eps_use, int *d_status, double *d_detail_out)
{
    if (ctx->n_active == 0) return WGPU_OK;
    FlagArgs a;
    a.active = ctx->d_active;
    a.det_abs = ctx->d_det_abs;
    a.det_sq = ctx->d_det_sq;
    a.status = d_status;
    a.detail_out = d_detail_out;
    a.n_active = ctx->n_active;
    a.nc = ctx->nc;
    for (int i = 0; i < 16; ++i) {
        a.thresh_comp[i] = i < ctx->nc ? thresh_comp[i] : 0;
        a.eps_use[i] = i < ctx->nc ? eps_use[i] : 0.0;
    }
    flags_kernel<<<(ctx->n_active + 127) / 128, 128, 0, ctx->stream>>>(a);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_linfty(wgpu_ctx *ctx, const double *u, unsigned long long *d_out)
{
    if (ctx->n_active == 0) return WGPU_OK;
    dim3 grid(ctx->n_active, ctx->nc);
    linfty_kernel<<<grid, 256, 0, ctx->stream>>>(u, ctx->d_active, ctx->nc, ctx->blk_elems, d_out);
    ctx->launches++;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}
