"""oracle.expm_pade / oracle.krylov_step: the restatement of krylov_time_stepper (LIB/TIME/krylov.f90) and of its matrix exponential (Expokit's
DGPADM, degree 6).  The reference ships no Krylov fixture (no parameter file under TESTING selects it), so the restatement is pinned by what
the scheme must satisfy: the exponential against scipy's, exactness of the integrator on a linear problem (the Krylov space of dimension M
reproduces exp(dt J) to the error estimate), agreement with the fourth-order Runge-Kutta step of the fixtures to the size of the two schemes'
truncation errors, and the dynamic subspace control."""
import numpy as np
import pytest
import scipy.linalg

import oracle as O


@pytest.mark.parametrize("m,scale", [(3, 0.1), (8, 1.0), (14, 30.0), (14, 1.0e-9)])
def test_expm_pade_against_scipy(m, scale):
    H = np.random.default_rng(m).standard_normal((m, m)) * scale
    E, S = O.expm_pade(H), scipy.linalg.expm(H)
    assert np.abs(E - S).max() <= 1e-11 * max(1.0, np.abs(S).max())
    assert np.array_equal(O.expm_pade(np.zeros((m, m))), np.eye(m))


def _tg(J=1, Bs=16, nu=3.125e-3):
    p = O.Params(dim=3, Bs=(Bs,) * 3, g=3, g_rhs=2, domain=(6.283185307179586,) * 3, Jmax=J, discretization="FD_4th_central", skew=True, c0=10.0,
                 nu=nu, gamma_p=0.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9)
    grid = O.uniform_grid(J)
    u = O.alloc(grid, p)
    O.inicond_taylor_green(grid, p, u)
    return p, grid, u


def test_krylov_step_agrees_with_rk4_and_its_own_error_estimate():
    p, grid, u = _tg()
    p.CFL = 0.5
    v = u.copy()
    work = [O.alloc(grid, p) for _ in range(5)]
    dt_rk = O.rk_generic(grid, p, v, work, 0.0)
    dt, M, err = O.krylov_step(grid, p, u, 0.0, M_max=12)
    assert dt == dt_rk and M == 12 and 0.0 <= err < 1e-3
    I = (slice(None), slice(None)) + O.interior(p)
    d = np.abs(u[I] - v[I]).max()
    assert 0.0 < d <= 1e-7, d                            # the exponential step is exact to err: what remains is RK4's local error ...
    p1, grid1, u1 = _tg()                                # ... which grows like dt^5: twice the step, >= 16 times the distance
    v1 = u1.copy()
    O.rk_generic(grid1, p1, v1, work, 0.0)
    O.krylov_step(grid1, p1, u1, 0.0, M_max=12)
    assert 16.0 * d <= np.abs(u1[I] - v1[I]).max() <= 64.0 * d
    # a smaller subspace is a worse approximation of the same step: the error estimate and the distance to the M = 12 result grow
    p2, grid2, w = _tg()
    p2.CFL = 0.5
    _, M4, err4 = O.krylov_step(grid2, p2, w, 0.0, M_max=4)
    assert M4 == 4 and err4 > err and np.abs(w[I] - u[I]).max() > 0.0


def test_krylov_dynamic_subspace_stops_early_or_shrinks_dt():
    p, grid, u = _tg()
    p.CFL = 0.5
    dt, M, err = O.krylov_step(grid, p, u.copy(), 0.0, M_max=12, dynamic=True, err_threshold=1e-3)
    assert M < 12 and err <= 1e-3
    dt_full = O.calculate_time_step(grid, p, u.copy(), 0.0)
    dt2, M2, err2 = O.krylov_step(grid, p, u.copy(), 0.0, M_max=3, dynamic=True, err_threshold=1e-9)
    assert M2 == 3 and err2 <= 1e-9 and dt2 < dt_full and abs(np.log(dt2 / dt_full) / np.log(0.9) - round(np.log(dt2 / dt_full) / np.log(0.9))) < 1e-9


@pytest.mark.parametrize("m,scale", [(3, 0.1), (8, 1.0), (14, 30.0), (14, 1.0e-9), (1, 2.0)])
def test_library_expm_pade_host_routine(m, scale):
    """wgpu_expm_pade (host code of libwabbit_gpu.so; loading the library needs no device) against the oracle's restatement and scipy"""
    import ctypes as C
    from wabbit_b200 import _native
    lib = _native.gpu_lib()
    H = np.ascontiguousarray(np.random.default_rng(m + 1).standard_normal((m, m)) * scale)
    E = np.zeros_like(H)
    dp = C.POINTER(C.c_double)
    assert lib.wgpu_expm_pade(H.ctypes.data_as(dp), m, E.ctypes.data_as(dp)) == 0
    S = scipy.linalg.expm(H)
    tol = 1e-11 * max(1.0, np.abs(S).max())
    assert np.abs(E - S).max() <= tol and np.abs(E - O.expm_pade(H)).max() <= tol
    # the augmented Hessenberg matrix of the integrator: last column of exp gives phi_1, phi_2 of the leading block
    Z = np.zeros_like(H)
    assert lib.wgpu_expm_pade(Z.ctypes.data_as(dp), m, E.ctypes.data_as(dp)) == 0 and np.array_equal(E, np.eye(m))
    H[0, 0] = np.nan
    assert lib.wgpu_expm_pade(H.ctypes.data_as(dp), m, E.ctypes.data_as(dp)) != 0


def test_krylov_step_is_conditioned_by_its_finite_difference_jacobian(monkeypatch):
    """Why the device parity test (tests/test_gpu_krylov.py) compares fields to 2e-7 and not to 1e-12: the Jacobian action is
    (F(u + eps v) - F(u)) / eps with eps = |u| sqrt(epsilon) ~ 1e-6, so a one-ulp difference in the right-hand side (the device's stage kernel
    agrees with the restatement to ~ 1e-15, not bit for bit) comes back 1e6 times larger in the Krylov vectors: ~ 1e-9 .. 1e-7 of the field per
    step.  Scalar products in another summation order, by contrast, move nothing (<= 1e-14): M, dt and err are insensitive."""
    p, grid, u = _tg()
    u += 0.02 * np.random.default_rng(2).standard_normal(u.shape)
    I = (slice(None), slice(None)) + O.interior(p)
    ref = u.copy()
    r0 = O.krylov_step(grid, p, ref, 0.0, M_max=6)
    blockwise = lambda a, b: float(sum(float((a[k][I[1:]] * b[k][I[1:]]).sum()) for k in range(a.shape[0])))
    w = u.copy()
    r1 = O.krylov_step(grid, p, w, 0.0, M_max=6, dot=blockwise)
    assert r1[:2] == r0[:2] and np.abs(w[I] - ref[I]).max() <= 1e-14
    rng, orig = np.random.default_rng(0), O.rhs_tree

    def one_ulp(grid_, p_, hvy, rhs, mask=None, fast=False):
        orig(grid_, p_, hvy, rhs, mask, fast)
        rhs *= 1.0 + 1.1e-16 * rng.choice([-1.0, 0.0, 1.0], size=rhs.shape)
    monkeypatch.setattr(O, "rhs_tree", one_ulp)
    v = u.copy()
    r2 = O.krylov_step(grid, p, v, 0.0, M_max=6)
    assert r2[:2] == r0[:2] and abs(r2[2] - r0[2]) <= 1e-6 * r0[2]
    d = np.abs(v[I] - ref[I]).max() / np.abs(ref[I]).max()
    assert 1e-10 < d <= 2e-7, d
