"""oracle.rkc_step (restatement of RungeKuttaChebychev, LIB/TIME/runge_kutta_chebychev.f90:6-146) with the reference's own coefficient tables
(tests/golden/rkc_coefficients.npz, extracted from setup_RKC_coefficients by tests/golden/make_golden.py).  The reference ships no RKC
regression fixture, so the restatement is pinned by what the scheme must satisfy: consistency of the tables (c_s = 1, c_1 = mu~_1, the
second-order conditions) and second-order convergence of one step towards the fourth-order RungeKuttaGeneric step of the same oracle."""
import os

import numpy as np
import pytest

import oracle as O

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rkc_coefficients.npz"))


def coeffs(s):
    return tuple(GOLD[f"s{s}_{n}"] for n in ("mu", "mu_tilde", "nu", "gamma_tilde", "c"))


@pytest.mark.parametrize("s", [4, 6, 10, 20])
def test_tables_are_a_consistent_second_order_rkc_scheme(s):
    mu, mut, nu, gt, c = coeffs(s)
    assert len(mu) == s and abs(c[-1] - 1.0) <= 1e-14 and abs(c[0] - mut[0]) <= 1e-15
    # the abscissae follow the same three-term recursion as the stages (Sommeijer et al. 1997, eq. 2.3): c_j = mu_j c_{j-1} + nu_j c_{j-2} + mu~_j + gamma~_j
    cc = [0.0, c[0]]
    for j in range(1, s):
        cc.append(mu[j] * cc[-1] + nu[j] * cc[-2] + mut[j] + gt[j])
    assert np.abs(np.array(cc[1:]) - c).max() <= 1e-13


@pytest.mark.parametrize("s", [4, 10])
def test_rkc_step_converges_with_second_order_towards_rk4(s):
    def run(dt, rkc):
        p = O.Params(dim=3, Bs=(16, 16, 16), g=3, g_rhs=2, domain=(6.283185307179586,) * 3, Jmax=1, discretization="FD_4th_central", skew=True,
                     c0=5.0, nu=1.0e-2, gamma_p=1.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9, dt_fixed=dt)
        grid = O.uniform_grid(1)
        u = O.alloc(grid, p)
        O.inicond_taylor_green(grid, p, u)
        if rkc:
            assert O.rkc_step(grid, p, u, 0.0, *coeffs(s)) == dt
        else:
            work = [O.alloc(grid, p) for _ in range(5)]
            assert O.rk_generic(grid, p, u, work, 0.0) == dt
        return u[(slice(None), slice(None)) + O.interior(p)]
    h = 4.0e-3
    e1 = np.abs(run(h, True) - run(h, False)).max()
    e2 = np.abs(run(h / 2, True) - run(h / 2, False)).max()
    assert 1e-12 < e2 < e1 < 1e-4 and 6.0 <= e1 / e2 <= 10.0, (e1, e2, e1 / e2)      # local error O(dt^3)


def test_superviscosity_stencils_known_answers():
    """generate_superviscosity_stencil + the sign convention and the added identity of filter_wrapper.f90:28-62"""
    assert list(O.superviscosity_stencil("explicit_3pt").values()) == [0.25, 0.5, 0.25]
    assert list(O.superviscosity_stencil("explicit_5pt").values()) == [-1 / 16, 4 / 16, 10 / 16, 4 / 16, -1 / 16]
    assert O.superviscosity_stencil("superviscosity_6th") == O.superviscosity_stencil("explicit_7pt")
    for t in ("explicit_9pt", "explicit_13pt", "explicit_21pt"):
        c = O.superviscosity_stencil(t)
        assert abs(sum(c.values()) - 1.0) <= 1e-15 and c[0] > 0.5 and all(abs(c[k] - c[-k]) == 0.0 for k in c)
    with pytest.raises(ValueError):
        O.superviscosity_stencil("explicit_4pt")


def test_closed_form_of_the_tabulated_scheme():
    """wabbit_b200.params.rkc2_coefficients (what the Python mirror uses without RKC_custom_scheme): the damped second-order scheme of
    Sommeijer, Shampine & Verwer with eps = 10 IS the reference's table (setup_RKC_coefficients) -- to 3e-15 on the sampled rows"""
    import os
    from wabbit_b200.params import rkc2_coefficients
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rkc_coefficients.npz"))
    for s in (4, 6, 10, 20):
        for name, mine in zip(("mu", "mu_tilde", "nu", "gamma_tilde", "c"), rkc2_coefficients(s)):
            assert np.abs(mine - gold[f"s{s}_{name}"]).max() <= 3e-15, (s, name)
    # consistency (order conditions of the scheme): c_s = 1, c_j increasing
    for s in (5, 12, 33):
        c = rkc2_coefficients(s)[4]
        assert abs(c[-1] - 1.0) <= 1e-13 and np.all(np.diff(c[1:]) > 0.0)
