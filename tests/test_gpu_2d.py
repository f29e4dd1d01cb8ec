"""GPU parity of the 2-D path (RHS_2D_acm, LIB/EQUATION/ACMnew/rhs_ACM.f90:292-922): RHS for every discretisation x skew x
penalization/sponge against the oracle, a Runge-Kutta step with identical dt, and the reference's own 2-D regression case
TESTING/acm/3vortices/3vorticesEquiFD4_CDF40 (restart at t = 10, 3073 steps to t = 20) on the GPU."""
import os

import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, Params, WabbitGPU
from wabbit_b200.solver import HVY_MASK, HVY_WORK

from util import orc_grid, orc_params, relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def params_2d(Bs, J, g, discretization="FD_4th_central", skew=True, **kw):
    kw.setdefault("c0", 7.0)
    kw.setdefault("nu", 5.0e-3)
    kw.setdefault("gamma_p", 1.0)
    return Params(dim=2, domain=(6.283185307179586,) * 3, Bs=(Bs, Bs, 1), g=g, g_rhs=2, n_eqn=3, Jmax=J, discretization=discretization,
                  skew_symmetry=skew, CFL=1.0, u_mean_set=(0.3, -0.1, 0.0), time_max=1.0e9, **kw).finalize()


@pytest.mark.parametrize("disc", ["FD_2nd_central", "FD_4th_central", "FD_6th_central", "FD_4th_central_optimized"])
@pytest.mark.parametrize("skew", [True, False])
@pytest.mark.parametrize("Bs", [26, 32])
def test_rhs_2d(disc, skew, Bs):
    p = params_2d(Bs, 2, 4, disc, skew)
    forest = Forest.uniform(2, 2)
    grid, po = orc_grid(forest), orc_params(p)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    rng = np.random.default_rng(3)
    u = rng.standard_normal(sol.host_shape())
    sol.upload(u)
    sol.RHS_wrapper(0.0, dst_slot=2)
    got = np.zeros_like(u)
    sol.download(got, HVY_WORK, 2, g_sync=0)
    ref = u.copy()
    O.sync_ghosts_same_level(grid, po, ref, p.g_rhs, p.g_rhs)
    rhs = np.zeros_like(u)
    O.rhs_tree(grid, po, ref, rhs)
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(got[I], rhs[I]) <= 1e-12
    sol.close()


def test_rhs_2d_penalization_and_sponge():
    p = params_2d(26, 2, 6, penalization=True, use_sponge=True, C_eta=1.34e-3, C_sponge=2.0e-2, n_mask=6)
    forest = Forest.uniform(2, 2)
    grid, po = orc_grid(forest), orc_params(p)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    rng = np.random.default_rng(4)
    u = rng.standard_normal(sol.host_shape())
    mask = rng.random(sol.host_shape(6))
    mask[:, 4] = rng.integers(0, 3, size=mask[:, 4].shape).astype(np.float64)   # colour
    sol.upload(u)
    sol.upload(mask, HVY_MASK)
    sol.RHS_wrapper(0.0, dst_slot=2)
    got = np.zeros_like(u)
    sol.download(got, HVY_WORK, 2, g_sync=0)
    ref = u.copy()
    O.sync_ghosts_same_level(grid, po, ref, p.g_rhs, p.g_rhs)
    rhs = np.zeros_like(u)
    O.rhs_tree(grid, po, ref, rhs, mask)
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(got[I], rhs[I]) <= 1e-12
    sol.close()


def test_rk4_2d_step_matches_oracle():
    p = params_2d(32, 3, 3)
    forest = Forest.uniform(2, 3)
    grid, po = orc_grid(forest), orc_params(p)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    rng = np.random.default_rng(5)
    u = 0.3 * rng.standard_normal(sol.host_shape())
    sol.upload(u)
    work = [O.alloc(grid, po) for _ in range(5)]
    t = 0.0
    for it in range(3):
        dt = sol.RungeKuttaGeneric(t, it)
        dt_ref = O.rk_generic(grid, po, u, work, t)
        assert dt == dt_ref
        t += dt
    got = np.zeros_like(u)
    sol.download(got, g_sync=0)
    I = (slice(None), slice(None)) + O.interior(po)
    assert relerr(got[I], u[I]) <= 1e-12
    sol.close()


def test_three_vortices_fixture_on_gpu():
    inp = np.load(os.path.join(GOLD, "three_vortices_t10.npz"))
    gold = np.load(os.path.join(GOLD, "three_vortices_FD4_CDF40.npz"))
    p = Params(dim=2, domain=(6.283185307179586,) * 3, Bs=(32, 32, 1), g=3, g_rhs=2, n_eqn=3, Jmax=3, discretization="FD_4th_central",
               skew_symmetry=True, c0=7.0, nu=5.0e-5, gamma_p=1.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=20.0,
               write_method="fixed_time", write_time=10.0).finalize()
    forest = Forest.uniform(2, 3)
    hvy, lvl, ixyz, _ = forest.active(0)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    src = {tuple(v): k for k, v in enumerate(inp["ixy"])}
    u = np.zeros(sol.host_shape())
    for h, ix in zip(hvy, ixyz):
        u[h - 1, :, 0, 3:35, 3:35] = inp["u"][src[(int(ix[0]), int(ix[1]))]]
    sol.upload(u)
    t, it = float(inp["time"][0]), int(inp["iteration"][0])
    while t < p.time_max:
        t, it, _ = sol.timeStep_tree(t, it)
    assert it == int(gold["iteration"][0]) and t == float(gold["time"][0])
    sol.download(u, g_sync=0)
    s = int(gold["stride"][0])
    where = {(int(ix[0]), int(ix[1])): h - 1 for h, ix in zip(hvy, ixyz)}
    got = np.stack([u[where[tuple(int(q) for q in v)], :, 0, 3:35:s, 3:35:s] for v in gold["ixy"]])
    err = np.abs(got - gold["u"]).max()
    # 3073 steps of a chaotic flow; FMA contraction on the GPU vs none in the reference: record, require round-off accumulation level
    print("three vortices GPU vs reference fields: max abs err", err)
    assert err <= 1e-9, err
