"""Pin the CPU oracle against the reference's own regression fields (TESTING/acm/taylorGreen/*): the oracle
advances the analytic Taylor-Green initial condition to t = 10 with the reference's parameters and must land on
the fields the reference Fortran code wrote, with the same number of time steps.

This pins RHS_3D_acm (FD2/FD4/FD6, skew-symmetric), RungeKuttaGeneric, calculate_time_step (incl. the tsave_stats
clipping), GET_DT_BLOCK_ACM and the same-level ghost synchronisation.
"""
import os

import numpy as np
import pytest

import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")

CASES = {
    "FD4_CDF40": dict(discretization="FD_4th_central", g=3, g_rhs=2),
    "FD2_CDF20": dict(discretization="FD_2nd_central", g=1, g_rhs=1),
    "FD6_CDF60": dict(discretization="FD_6th_central", g=5, g_rhs=3),
}


def tg_params(case):
    c = CASES[case]
    return O.Params(dim=3, Bs=(20, 20, 20), g=c["g"], g_rhs=c["g_rhs"], domain=(6.283185307179586,) * 3, Jmax=1,
                    discretization=c["discretization"], skew=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0,
                    time_max=10.0, write_method="fixed_time", write_time=10.0, tsave_stats=0.20, u_mean_set=(0, 0, 0))


def sample(u, p, grid, gold_ixyz, stride):
    g = p.g
    out = []
    for ix in gold_ixyz:
        b = [k for k in range(grid.n) if (grid.ixyz[k] == ix).all()][0]
        out.append(u[b, :, g:g + p.Bs[2]:stride, g:g + p.Bs[1]:stride, g:g + p.Bs[0]:stride])
    return np.stack(out)


def _tg_run(case):
    """the run of one Taylor-Green case (a worker process started by conftest.py at collection time, or inline): the sampled initial
    condition, the number of steps, the final time and the sampled final fields"""
    gold = np.load(os.path.join(GOLD, f"taylor_green_{case}.npz"))
    p = tg_params(case)
    grid = O.uniform_grid(1)
    u = O.alloc(grid, p)
    O.inicond_taylor_green(grid, p, u)
    stride = int(gold["stride"][0])
    s0 = sample(u, p, grid, gold["t0_ixyz"], stride)
    work = [O.alloc(grid, p) for _ in range(5)]
    t, it = 0.0, 0
    while t < p.time_max:
        t += O.rk_generic(grid, p, u, work, t)
        it += 1
    return s0, it, t, sample(u, p, grid, gold["t1_ixyz"], stride)


@pytest.mark.parametrize("case", list(CASES))
def test_taylor_green_fixture(case):
    from conftest import background
    gold = np.load(os.path.join(GOLD, f"taylor_green_{case}.npz"))
    s0, it, t, s1 = background("taylor_green", case, _tg_run)
    # initial condition (inicond_ACM.f90:371-389)
    assert np.abs(s0 - gold["t0"]).max() <= 1e-15
    assert it == int(gold["t1_iteration"][0])
    assert abs(t - float(gold["t1_time"][0])) < 1e-12
    err = np.abs(s1 - gold["t1"]).max()
    # the restatement reproduces the Fortran output to round-off; 1e-12 is the north-star field tolerance
    assert err <= 1e-12, err


# ---------------------------------------------------------------------------------------------------------------------
# 2-D: TESTING/acm/3vortices/3vorticesEqui*: restart from the stored t = 10 fields (64 blocks of 32^2 on level 3), run to
# t = 20 (fixed_time output clips the last step).  Pins RHS_2D_acm (FD2/FD4/FD6, skew-symmetric, gamma_p), the 2-D
# same-level synchronisation (8 relations), RungeKuttaGeneric and calculate_time_step in two dimensions.
TV_CASES = {"FD4_CDF40": ("FD_4th_central", 3, 2), "FD2_CDF20": ("FD_2nd_central", 1, 1), "FD6_CDF60": ("FD_6th_central", 5, 3)}


def three_vortices_setup(case):
    inp = np.load(os.path.join(GOLD, "three_vortices_t10.npz"))
    disc, g, g_rhs = TV_CASES[case]
    p = O.Params(dim=2, Bs=(32, 32, 1), g=g, g_rhs=g_rhs, n_eqn=3, domain=(6.283185307179586,) * 3, Jmax=3, discretization=disc,
                 skew=True, c0=7.0, nu=5.0e-5, gamma_p=1.0, CFL=1.0, time_max=20.0, write_method="fixed_time", write_time=10.0,
                 u_mean_set=(0.0, 0.0, 0.0))
    ixyz = np.concatenate([inp["ixy"], np.zeros((len(inp["ixy"]), 1), np.int32)], axis=1).astype(np.int64)
    grid = O.Grid(level=inp["level"].astype(np.int64), ixyz=ixyz, dim=2)
    u = O.alloc(grid, p)
    u[:, :, 0, g:g + 32, g:g + 32] = inp["u"]
    return p, grid, u, float(inp["time"][0]), int(inp["iteration"][0])


@pytest.mark.parametrize("case", list(TV_CASES))
def test_three_vortices_2d_fixture(case):
    gold = np.load(os.path.join(GOLD, f"three_vortices_{case}.npz"))
    p, grid, u, t, it = three_vortices_setup(case)
    nbr, dxb = O.nbr_table(grid), O.dx_table(grid, p)
    work = np.zeros((5,) + u.shape)
    times = []
    while t < p.time_max:
        t += O.rk_step_c(grid, p, u, work, t, nbr, dxb)
        it += 1
        times.append(t)
    assert it == int(gold["iteration"][0])
    assert t == float(gold["time"][0])
    # the reference's own log of this run (log.original.txt: "RUN: it= .. time= 10.003239697 ..") -- the time after EVERY step, printed to
    # nine decimals: calculate_time_step is pinned along the whole run, not only by the final iteration counter
    log = np.load(os.path.join(GOLD, "three_vortices_log_times.npz"))
    assert len(times) == len(log[f"{case}_time"]) and int(log[f"{case}_iteration"][-1]) == it
    assert np.abs(np.array(times) - log[f"{case}_time"]).max() <= 5.0e-10 + 1e-13
    s, g = int(gold["stride"][0]), p.g
    order = {tuple(v): k for k, v in enumerate(grid.ixyz[:, :2])}
    got = np.stack([u[order[tuple(v)], :, 0, g:g + 32:s, g:g + 32:s] for v in gold["ixy"]])
    err = np.abs(got - gold["u"]).max()
    assert err <= 1e-12, err
