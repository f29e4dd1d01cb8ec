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
