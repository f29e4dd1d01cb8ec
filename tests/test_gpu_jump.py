"""GPU parity on graded grids with level jumps: the stage kernel fed by the restriction / prediction face patches
(wabbit_b200/csrc/jump.cu) against the oracle's sync_ghosts_generic("full_leaf", ignore_Filter) + RHS_3D_acm + RungeKuttaGeneric
(oracle/orc_sync.c follows the reference's patch index tables; the GPU resolves ghost points geometrically)."""
import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, WabbitGPU
from wabbit_b200.solver import HVY_WORK

from util import graded_blocks, orc_grid, orc_params, relerr, tg_params

pytestmark = pytest.mark.gpu


def _setup(wavelet, Bs, J0, Jmax, seed, discretization="FD_4th_central", skew=True, frac=0.3):
    w = O.setup_wavelet(wavelet)
    p = tg_params(Bs=Bs, J=Jmax, wavelet_g=w.g_default, discretization=discretization, skew=skew)
    p.g_rhs = {"FD_2nd_central": 1, "FD_4th_central": 2, "FD_6th_central": 3, "FD_4th_central_optimized": 3}[discretization]
    p.wavelet = wavelet
    lv, ix = graded_blocks(3, J0, Jmax, seed, frac)
    forest = Forest.from_blocks(3, Jmax, lv, ix)
    assert not forest.is_uniform
    grid = orc_grid(forest)
    po = orc_params(p)
    nbr = forest.neighbors(0)[:, :grid.n]
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    rng = np.random.default_rng(seed + 17)
    u = O.alloc(grid, po)
    # smooth + noise: the prediction weights then matter at leading order and at round-off
    O.inicond_taylor_green(grid, po, u)
    u += 0.1 * rng.standard_normal(u.shape)
    return w, p, po, forest, grid, nbr, sol, u


@pytest.mark.parametrize("wavelet,Bs,disc,skew", [("CDF40", 16, "FD_4th_central", True), ("CDF44", 16, "FD_4th_central", False),
                                                   ("CDF20", 16, "FD_2nd_central", True), ("CDF62", 20, "FD_6th_central", True),
                                                   ("CDF44", 18, "FD_4th_central_optimized", True), ("CDF44", 22, "FD_4th_central", True),
                                                   ("CDF40", 26, "FD_6th_central", False)])
def test_rhs_with_level_jumps(wavelet, Bs, disc, skew):
    w, p, po, forest, grid, nbr, sol, u = _setup(wavelet, Bs, 1, 3, seed=5, discretization=disc, skew=skew)
    sol.upload(u)
    sol.RHS_wrapper(0.0, dst_slot=2)
    got = np.zeros_like(u)
    sol.download(got, HVY_WORK, 2, g_sync=0)
    ref_u = u.copy()
    n = O.sync_ghosts_leaf(grid, po, ref_u, nbr, p.g_rhs, p.g_rhs, w.X, bool(w.lifted))
    assert n > 0
    rhs = np.zeros_like(u)
    O.rhs_tree(grid, po, ref_u, rhs)
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(got[I], rhs[I]) <= 1e-12
    # block by block: every block (whatever its neighbour configuration) agrees
    for b in range(grid.n):
        assert relerr(got[b][I[1:]], rhs[b][I[1:]]) <= 1e-11, b
    sol.close()


def test_rk4_step_with_level_jumps():
    w, p, po, forest, grid, nbr, sol, u = _setup("CDF44", 16, 1, 3, seed=11)
    sol.upload(u)
    dts = []
    t = 0.0
    for it in range(2):
        dt = sol.RungeKuttaGeneric(t, it)
        dts.append(dt)
        t += dt
    got = np.zeros_like(u)
    sol.download(got, g_sync=0)
    work = [O.alloc(grid, po) for _ in range(5)]
    sync = lambda h: O.sync_ghosts_leaf(grid, po, h, nbr, p.g_rhs, p.g_rhs, w.X, bool(w.lifted))
    t = 0.0
    for it in range(2):
        dt = O.rk_generic(grid, po, u, work, t, sync=sync)
        assert dt == dts[it]
        t += dt
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(got[I], u[I]) <= 1e-12
    sol.close()


def test_three_levels_deep_and_uniform_fallback():
    # levels 1..4 present; also a uniform grid after a jump grid on the same context (tables are rebuilt)
    w, p, po, forest, grid, nbr, sol, u = _setup("CDF40", 16, 1, 4, seed=3, frac=0.2)
    assert grid.level.max() - grid.level.min() >= 2
    sol.upload(u)
    sol.RHS_wrapper(0.0, dst_slot=2)
    got = np.zeros_like(u)
    sol.download(got, HVY_WORK, 2, g_sync=0)
    ref_u = u.copy()
    O.sync_ghosts_leaf(grid, po, ref_u, nbr, p.g_rhs, p.g_rhs, w.X, bool(w.lifted))
    rhs = np.zeros_like(u)
    O.rhs_tree(grid, po, ref_u, rhs)
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(got[I], rhs[I]) <= 1e-12
    sol.close()
