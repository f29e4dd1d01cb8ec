// Krylov exponential integrator (time_step_method = Krylov): the vector operations of krylov_time_stepper (LIB/TIME/krylov.f90:1-190) on the
// resident block interiors, and its matrix exponential on the host.
//
//   kry_dot_kernel / kry_sum_kernel   scalarproduct / wabbit_norm (krylov.f90:500-595): sum over the interiors of the active blocks, all
//                                     components; one partial sum per block, added in list order (deterministic)
//   kry_axpy_kernel                   the five element-wise updates of the Arnoldi loop (operations evaluated as the Fortran expressions:
//                                     a division stays a division, products left to right)
//   wgpu_expm_pade                    expM_pade -> DGPADM (krylov.f90:193-396; Expokit, R. Sidje, ACM TOMS 24 (1998)), restated from the published
//                                     algorithm: scaling and squaring around the irreducible (6, 6) Pade fraction, Gaussian elimination with
//                                     partial pivoting for the one linear solve.  Host code: the matrices are (M + 2)^2 <= 64^2.
#include <math.h>

#include <vector>

#include "wgpu_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) kry_dot_kernel(const double *__restrict__ x, const double *__restrict__ y, const int *__restrict__ active,
                                                      long long per_block, double *__restrict__ part)
{
    __shared__ double red[8];
    const long long off = (long long)active[blockIdx.x] * per_block;
    double s = 0.0;
    for (long long e = threadIdx.x; e < per_block; e += blockDim.x) s += x[off + e] * y[off + e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        part[blockIdx.x] = t;
    }
}

// out[0] = sum of part[0..n) -- 256 strided partial sums, then a fixed-order tree: the same value on every run
__global__ void __launch_bounds__(256) kry_sum_kernel(const double *__restrict__ part, int n, double *__restrict__ out)
{
    __shared__ double sh[256];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += part[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// op 0: dst = x / a            (first Krylov vector, normalisation)
//    1: dst = x + a * y        (perturbed state)
//    2: dst = (x - y) / a      (finite-difference Jacobian action)
//    3: dst = x - a * y        (Gram-Schmidt)
//    4: dst = x + (a * y) * b  (the new state: beta * v * phi, left to right)
__global__ void __launch_bounds__(256) kry_axpy_kernel(double *__restrict__ dst, const double *x, const double *y, const int *__restrict__ active,
                                                       long long per_block, int op, double a, double b)
{
    const long long off = (long long)active[blockIdx.y] * per_block;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < per_block; e += (long long)gridDim.x * blockDim.x) {
        const long long i = off + e;
        double r;
        switch (op) {
        case 0: r = x[i] / a; break;
        case 1: r = __dadd_rn(x[i], __dmul_rn(a, y[i])); break;
        case 2: r = __dsub_rn(x[i], y[i]) / a; break;
        case 3: r = __dsub_rn(x[i], __dmul_rn(a, y[i])); break;
        default: r = __dadd_rn(x[i], __dmul_rn(__dmul_rn(a, y[i]), b)); break;
        }
        dst[i] = r;
    }
}

}  // namespace

int32_t wgpu_launch_kry_dot(wgpu_ctx *ctx, const double *x, const double *y, double *d_part, double *d_out)
{
    const long long per_block = (long long)ctx->nc * ctx->blk_elems;
    if (ctx->n_active) kry_dot_kernel<<<ctx->n_active, 256, 0, ctx->stream>>>(x, y, ctx->d_active, per_block, d_part);
    kry_sum_kernel<<<1, 256, 0, ctx->stream>>>(d_part, ctx->n_active, d_out);
    ctx->launches += 2;
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

int32_t wgpu_launch_kry_axpy(wgpu_ctx *ctx, double *dst, const double *x, const double *y, int op, double a, double b)
{
    if (ctx->n_active == 0) return WGPU_OK;
    const long long per_block = (long long)ctx->nc * ctx->blk_elems;
    const int gx = (int)std::min<long long>((per_block + 255) / 256, 64);
    for (int s0 = 0; s0 < ctx->n_active; s0 += 32768) {       // grid.y limit
        const int m = std::min(32768, ctx->n_active - s0);
        kry_axpy_kernel<<<dim3(gx, m), 256, 0, ctx->stream>>>(dst, x, y, ctx->d_active + s0, per_block, op, a, b);
        ctx->launches++;
    }
    WGPU_CHECK(ctx, cudaGetLastError());
    return WGPU_OK;
}

extern "C" int32_t wgpu_expm_pade(const double *H, int32_t m, double *E)
{
    if (!H || !E || m < 1) return WGPU_ERR_ARG;
    const int ideg = 6;
    const size_t mm = (size_t)m * m;
    auto at = [m](std::vector<double> &A, int i, int j) -> double & { return A[(size_t)i * m + j]; };      // row-major
    auto matmul = [m](const std::vector<double> &A, const std::vector<double> &B, std::vector<double> &C) {
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) {
                double s = 0.0;
                for (int k = 0; k < m; ++k) s += A[(size_t)i * m + k] * B[(size_t)k * m + j];
                C[(size_t)i * m + j] = s;
            }
    };
    double hnorm = 0.0;
    for (int i = 0; i < m; ++i) {
        double r = 0.0;
        for (int j = 0; j < m; ++j) r += fabs(H[(size_t)i * m + j]);
        if (!(r <= 1.0e300)) return WGPU_ERR_ARG;                       // NaN / overflow in the Hessenberg matrix
        hnorm = r > hnorm ? r : hnorm;
    }
    if (hnorm == 0.0) {
        for (size_t i = 0; i < mm; ++i) E[i] = 0.0;
        for (int i = 0; i < m; ++i) E[(size_t)i * m + i] = 1.0;
        return WGPU_OK;
    }
    int ns = (int)(log(hnorm) / log(2.0)) + 2;
    if (ns < 0) ns = 0;
    const double scale = 1.0 / ldexp(1.0, ns);
    std::vector<double> A(mm), A2(mm), ev(mm, 0.0), od(mm, 0.0), T(mm);
    for (size_t i = 0; i < mm; ++i) A[i] = H[i] * scale;
    double c[ideg + 1];
    c[0] = 1.0;
    for (int k = 1; k <= ideg; ++k) c[k] = c[k - 1] * (double)(ideg + 1 - k) / (double)(k * (2 * ideg + 1 - k));
    matmul(A, A, A2);
    // Horner in A^2: even part c0 + c2 A^2 + c4 A^4 + c6 A^6, odd part (c1 + c3 A^2 + c5 A^4) A
    for (int i = 0; i < m; ++i) {
        at(ev, i, i) = c[6];
        at(od, i, i) = c[5];
    }
    for (int k = 4; k >= 0; k -= 2) {
        matmul(ev, A2, T);
        ev.swap(T);
        for (int i = 0; i < m; ++i) at(ev, i, i) += c[k];
    }
    for (int k = 3; k >= 1; k -= 2) {
        matmul(od, A2, T);
        od.swap(T);
        for (int i = 0; i < m; ++i) at(od, i, i) += c[k];
    }
    matmul(od, A, T);
    od.swap(T);
    // X = (ev - od)^-1 od by Gaussian elimination with partial pivoting; exp(A) ~ I + 2 X
    std::vector<double> Q(mm), X(od);
    for (size_t i = 0; i < mm; ++i) Q[i] = ev[i] - od[i];
    for (int k = 0; k < m; ++k) {
        int piv = k;
        for (int i = k + 1; i < m; ++i)
            if (fabs(at(Q, i, k)) > fabs(at(Q, piv, k))) piv = i;
        if (at(Q, piv, k) == 0.0) return 240917;                       // "Problem in DGESV (within DGPADM)"
        if (piv != k)
            for (int j = 0; j < m; ++j) {
                std::swap(at(Q, k, j), at(Q, piv, j));
                std::swap(at(X, k, j), at(X, piv, j));
            }
        for (int i = k + 1; i < m; ++i) {
            const double f = at(Q, i, k) / at(Q, k, k);
            if (f == 0.0) continue;
            for (int j = k; j < m; ++j) at(Q, i, j) -= f * at(Q, k, j);
            for (int j = 0; j < m; ++j) at(X, i, j) -= f * at(X, k, j);
        }
    }
    for (int k = m - 1; k >= 0; --k)
        for (int j = 0; j < m; ++j) {
            double s = at(X, k, j);
            for (int i = k + 1; i < m; ++i) s -= at(Q, k, i) * at(X, i, j);
            at(X, k, j) = s / at(Q, k, k);
        }
    for (size_t i = 0; i < mm; ++i) X[i] = 2.0 * X[i];
    for (int i = 0; i < m; ++i) at(X, i, i) += 1.0;
    for (int s = 0; s < ns; ++s) {
        matmul(X, X, T);
        X.swap(T);
    }
    for (size_t i = 0; i < mm; ++i) E[i] = X[i];
    return WGPU_OK;
}
