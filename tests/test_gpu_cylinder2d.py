"""The reference's adaptive cylinder regression TESTING/acm/acm_CDF44 (acm_cyl.ini: BASELINE's "2D ACM test case from TESTING/acm", 1 rank)
on the GPU: adaptive initial condition, then refine everywhere -> RK4 with penalization + sponge -> adapt_tree (CDF44, coarse extension,
security zone, threshold_mask, force_maxlevel_dealiasing) to t = 0.1, heavy data on the device, the mask function generated on the host
(wabbit_b200.mask) and uploaded for every RHS grid.

Compared through the C ABI with the fields the reference Fortran code wrote (tests/golden/cylinder_adapt_CDF44.npz): grids at t = 0, 0.05 and
0.1 identical (block lists, refinement statuses, iteration counters 40 / 82, times), the mask function bit for bit, ux / uy / p <= 1e-11;
and in lockstep with the oracle (oracle/adaptive.py, pinned by the same fixtures) over the first steps.
"""
import numpy as np
import pytest

import cylinder_case as CC
from wabbit_b200 import Forest, Params, WabbitGPU
from wabbit_b200.mask import CylinderMask2D
from wabbit_b200.solver import HVY_BLOCK
from wabbit_b200.timeloop import AdaptiveLoop

pytestmark = pytest.mark.gpu

MAXB = 1600


def set_inicond(loop):
    """inicond = meanflow: (ux, uy, p) = (0, -1, 0)"""
    hvy, _, _, _ = loop.forest.active(0)
    host = np.zeros(loop.sol.host_shape())
    host[hvy - 1, 1] = -1.0
    loop.sol.upload(host, HVY_BLOCK, 0, hvy_ids=hvy)


def make_loop(case="CDF44"):
    c = CC.CASES[case]
    p = Params(wavelet=c["wavelet"], eps=CC.EPS, eps_normalized=True, eps_norm="Linfty", Jmin=CC.JMIN, force_maxlevel_dealiasing=True, adapt_tree=True,
               refinement_indicator=c["indicator"], **CC.ini(case)).finalize()
    forest = Forest.uniform(2, CC.JMIN, Jmax=p.Jmax, max_blocks=MAXB)
    sol = WabbitGPU(p, max_blocks=MAXB)
    sol.setup_wavelet(c["wavelet"])
    sol.set_forest(forest)
    loop = AdaptiveLoop(sol, forest, 0.0, 0, mask=CylinderMask2D(p), threshold_mask=True, thresh_comp=c["thresh_comp"])
    set_inicond(loop)
    loop.adaptive_inicond(set_inicond)
    return loop


def state(loop):
    hvy, lvl, pos, _ = loop.forest.active(0)
    g = loop.sol.params.g
    got = np.zeros(loop.sol.host_shape())
    loop.sol.download(got, g_sync=0)
    u = got[hvy - 1][:, :, 0, g:g + CC.BS, g:g + CC.BS]
    return lvl, pos, loop.status, u, loop.iteration, loop.time, loop.mask.chi(lvl, pos)


@pytest.mark.parametrize("case", list(CC.CASES))
def test_cylinder_fixture_2d(case):
    c, gd = CC.CASES[case], CC.gold(case)
    chk = case != "CDF40"      # see test_oracle_cylinder.py
    loop = make_loop(case)
    errs = {"t0": CC.compare(gd, "t0", *state(loop), check_status=chk)}
    assert errs["t0"] == 0.0
    stops = {k: t for k, t in c["files"].items() if t > 0.0}
    while loop.time < loop.sol.params.time_max:
        loop.step()
        for k, t in stops.items():
            if abs(loop.time - t) <= 1e-15:
                errs[k] = CC.compare(gd, k, *state(loop), check_status=chk)
    print(f"\nacm_{case} cylinder on the GPU: {loop.iteration} adaptive steps, {loop.forest.n_blocks} blocks at t = {loop.time}, "
          f"max |u - reference| = {errs}, blocks on the RHS grid max {max(r[2] for r in loop.log)}")
    assert set(errs) == set(c["files"]) and max(errs.values()) <= 1e-11, errs
    loop.sol.close()


@pytest.mark.parametrize("case", ["CDF44", "CDF40", "significant_CDF44"])
def test_cylinder_lockstep_2d(case):
    from test_oracle_cylinder import make_run
    loop, run = make_loop(case), make_run(case)
    g = run.p.g
    for _ in range(8):
        dt_g, dt_o = loop.step(), run.step()
        assert loop.log[-1][2:4] == run.log[-1][2:4]
        assert abs(dt_g - dt_o) <= 1e-13 * dt_o
        lvl, pos, st, u, _, _, _ = state(loop)
        okey = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(run.grid.level, run.grid.ixyz))}
        o = np.array([okey[(int(l), int(x[0]), int(x[1]))] for l, x in zip(lvl, pos)])
        assert np.array_equal(np.asarray(st), run.status[o])
        ou = run.u[:, :, 0, g:g + CC.BS, g:g + CC.BS][o]
        assert np.abs(u - ou).max() <= 1e-12
    loop.sol.close()
