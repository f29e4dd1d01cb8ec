"""RungeKuttaChebychev on the device (wgpu_rkc_step) against the oracle's restatement (oracle.rkc_step, pinned on the CPU in
tests/test_oracle_rkc.py) with the reference's own coefficient tables: identical dt, fields <= 1e-12 relative -- equidistant and graded grids,
with penalization (mask from hvy_mask)."""
import os

import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, WabbitGPU

from util import graded_blocks, orc_grid, orc_params, relerr, tg_params

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rkc_coefficients.npz"))


def coeffs(s):
    return tuple(GOLD[f"s{s}_{n}"] for n in ("mu", "mu_tilde", "nu", "gamma_tilde", "c"))


@pytest.mark.parametrize("s,graded,Bs", [(4, False, 16), (10, False, 16), (6, True, 16), (4, False, 22)])
def test_rkc_steps_parity(s, graded, Bs):
    if graded:
        lv, ix = graded_blocks(3, 1, 3, seed=6)
        forest = Forest.from_blocks(3, 3, lv, ix)
    else:
        forest = Forest.uniform(3, 2, Jmax=3)
    p = tg_params(Bs=Bs, J=3)
    po, grid = orc_params(p), orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet("CDF40")
    sol.set_forest(forest)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    u += 0.05 * np.random.default_rng(2).standard_normal(u.shape)
    sol.upload(u)
    nbr = forest.neighbors(0)[:, :grid.n]
    sync = (lambda h: O.sync_ghosts_leaf(grid, po, h, nbr, p.g_rhs, p.g_rhs, 4, True)) if graded else None
    t = 0.0
    for it in range(2):
        dt = sol.RungeKuttaChebychev(t, it, *coeffs(s))
        assert dt == O.rkc_step(grid, po, u, t, *coeffs(s), sync=sync)
        t += dt
    out = np.zeros_like(u)
    sol.download(out, g_sync=0)
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(out[I], u[I]) <= 1e-12
    # a Runge-Kutta-Generic step continues from the state the Chebychev steps left (the CFL candidate is recomputed)
    work = [O.alloc(grid, po) for _ in range(5)]
    kw = {"sync": sync} if graded else {}
    assert sol.RungeKuttaGeneric(t, 2) == O.rk_generic(grid, po, u, work, t, **kw)
    sol.close()


def test_rkc_rejects_fewer_than_four_stages():
    from wabbit_b200 import WabbitAbort
    forest = Forest.uniform(3, 1, Jmax=1)
    p = tg_params(Bs=16, J=1)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    sol.upload(np.zeros(sol.host_shape()))
    with pytest.raises(WabbitAbort) as e:
        sol.RungeKuttaChebychev(0.0, 0, *[np.ones(3)] * 5)
    assert e.value.code == 1715929
    sol.close()


def test_time_step_tree_dispatches_on_the_parameter_file_keys():
    """timeStep_tree with time_step_method = RungeKuttaChebychev (custom-scheme rows) and a filter every 2nd iteration == the direct calls"""
    from wabbit_b200 import WabbitAbort
    forest = Forest.uniform(3, 1, Jmax=2)
    outs = []
    for mode in ("params", "direct"):
        p = tg_params(Bs=16, J=2)
        if mode == "params":
            p.time_step_method, p.rkc_s, p.RKC_custom_scheme = "RungeKuttaChebychev", 4, True
            p.RKC_mu, p.RKC_mu_tilde, p.RKC_nu, p.RKC_gamma_tilde, p.RKC_c = (tuple(v) for v in coeffs(4))
            p.filter_type, p.filter_freq, p.filter_component = "explicit_5pt", 2, (1, 1, 1, 0)
        sol = WabbitGPU(p, max_blocks=forest.n_blocks)
        sol.setup_wavelet("CDF40")
        sol.set_forest(forest)
        u = np.zeros(sol.host_shape())
        u[:] = np.random.default_rng(3).standard_normal(u.shape) * 0.1
        sol.upload(u)
        t, it = 0.0, 0
        for _ in range(2):
            if mode == "params":
                t, it, dt = sol.timeStep_tree(t, it)
            else:
                dt = sol.RungeKuttaChebychev(t, it, *coeffs(4))
                t, it = t + dt, it + 1
                if it % 2 == 0:
                    sol.filter_wrapper("explicit_5pt", [1, 1, 1, 0])
        out = np.zeros_like(u)
        sol.download(out, g_sync=0)
        outs.append((t, it, out))
        if mode == "direct":
            sol.params.time_step_method = "Leapfrog"
            with pytest.raises(WabbitAbort):
                sol.timeStep_tree(t, it)
        sol.close()
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] == 2
    assert np.array_equal(outs[0][2], outs[1][2])
