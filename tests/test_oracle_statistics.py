"""oracle.statistics_acm (the numpy restatement of STATISTICS_ACM's integral stage, LIB/EQUATION/ACMnew/statistics_ACM.f90:138-368) against
the analytic integrals of the Taylor-Green vortex: e_kin = (2 pi)^3 / 8, zero mean flow, divergence-free, max |u|^2 = 1."""
import numpy as np

import oracle as O


def test_statistics_of_taylor_green():
    p = O.Params(dim=3, Bs=(16, 16, 16), g=3, g_rhs=2, domain=(6.283185307179586,) * 3, Jmax=2, discretization="FD_4th_central", skew=True,
                 c0=10.0, nu=1e-2, gamma_p=0.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9)
    grid = O.uniform_grid(2)
    u = O.alloc(grid, p)
    O.inicond_taylor_green(grid, p, u)
    O.sync_ghosts_same_level(grid, p, u, 3, 3)
    s = O.statistics_acm(grid, p, u)
    assert abs(s["e_kin"] - (2.0 * np.pi) ** 3 / 8.0) <= 1e-12 * s["e_kin"]
    assert max(abs(s["meanflow_x"]), abs(s["meanflow_y"]), abs(s["meanflow_z"])) <= 1e-12
    assert s["umag"] == 1.0 and s["div_max"] <= 1e-13 and s["div_min"] >= -1e-13
    # ACM energy = e_kin + 0.5 int p^2 / c0^2
    assert s["ACM_energy"] > s["e_kin"] and s["mask_volume"] == 0.0


def test_vorticity_statistics_of_taylor_green():
    """omega = (-cos x sin y sin z, -sin x cos y sin z, 2 sin x sin y cos z): enstrophy = 3 pi^3, zero helicity, max |omega| = 2,
    dissipation = -nu int u . lap u = 3 nu int |u|^2 = 2 nu * enstrophy (fourth-order stencils on 64 points per direction)"""
    nu = 1e-2
    p = O.Params(dim=3, Bs=(16, 16, 16), g=3, g_rhs=2, domain=(6.283185307179586,) * 3, Jmax=2, discretization="FD_4th_central", skew=True,
                 c0=10.0, nu=nu, gamma_p=0.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9)
    grid = O.uniform_grid(2)
    u = O.alloc(grid, p)
    O.inicond_taylor_green(grid, p, u)
    O.sync_ghosts_same_level(grid, p, u, 3, 3)
    s = O.vorticity_statistics_acm(grid, p, u)
    Z = 3.0 * np.pi ** 3
    assert abs(s["enstrophy"] - Z) <= 2e-4 * Z                 # O(h^4), h = 2 pi / 64
    assert abs(s["helicity"]) <= 1e-12 * Z
    assert abs(s["max_vort"] - 2.0) <= 1e-4
    assert abs(s["dissipation"] - 2.0 * nu * Z) <= 2e-4 * 2.0 * nu * Z
    # second order on the same grid is visibly worse, sixth order better: the stencil tables are the module's
    err = {}
    for disc, g in (("FD_2nd_central", 3), ("FD_6th_central", 3)):
        q = O.Params(dim=3, Bs=(16, 16, 16), g=g, g_rhs=g, domain=p.domain, Jmax=2, discretization=disc, nu=nu, u_mean_set=(0.0, 0.0, 0.0))
        v = O.alloc(grid, q)
        O.inicond_taylor_green(grid, q, v)
        O.sync_ghosts_same_level(grid, q, v, g, g)
        err[disc] = abs(O.vorticity_statistics_acm(grid, q, v)["enstrophy"] - Z) / Z
    assert err["FD_6th_central"] < abs(s["enstrophy"] - Z) / Z < err["FD_2nd_central"]


def test_vorticity_statistics_2d():
    """u = (sin x cos y, -cos x sin y): omega = v_x - u_y = 2 sin x sin y, enstrophy = 0.5 * 4 * pi^2 = 2 pi^2, no helicity in 2-D"""
    p = O.Params(dim=2, Bs=(32, 32, 1), g=3, g_rhs=2, n_eqn=3, domain=(6.283185307179586,) * 3, Jmax=2, discretization="FD_4th_central", nu=1e-3,
                 u_mean_set=(0.0, 0.0, 0.0))
    grid = O.uniform_grid(2, 2)
    u = O.alloc(grid, p)
    for b in range(grid.n):
        x0, dx = grid.spacing_origin(p, b)
        x = (np.arange(32 + 6) - 3) * dx[0] + x0[0]
        y = (np.arange(32 + 6) - 3) * dx[1] + x0[1]
        Y, X = np.meshgrid(y, x, indexing="ij")
        u[b, 0, 0], u[b, 1, 0] = np.sin(X) * np.cos(Y), -np.cos(X) * np.sin(Y)
    s = O.vorticity_statistics_acm(grid, p, u)
    Z = 2.0 * np.pi ** 2
    assert abs(s["enstrophy"] - Z) <= 1e-5 * Z and s["helicity"] == 0.0 and abs(s["max_vort"] - 2.0) <= 1e-5
    assert abs(s["dissipation"] - 2.0 * 1e-3 * Z) <= 1e-5 * 2e-3 * Z
