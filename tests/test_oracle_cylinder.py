"""Pin the oracle against the reference's own adaptive cylinder regressions TESTING/acm/acm_CDF44 (and its variants acm_CDF40,
acm_norm_CDF44, acm_significant_CDF44, see cylinder_case.py; fields written by the reference Fortran code; tests/golden/cylinder_adapt_*.npz): the grid after the adaptive initial condition (52 blocks on levels 2-5), at t = 0.05
(iteration 40, 124 blocks) and at t = 0.1 (iteration 82, 160 blocks) -- block lists, refinement statuses, iteration counters and times
identical, the mask function bit for bit, ux / uy / p <= 1e-12.

On top of what the 3vortices fixtures pin (test_oracle_adaptive.py) this pins, in the adaptive loop: volume penalization and the sponge
in RHS_2D_acm, create_mask_2D_ACM (circle with cosine smoothing, p-norm sponge), threshold_mask (coarseningIndicatorMask_tree),
force_maxlevel_dealiasing, the CFL_eta time-step limit, the adaptive initial condition (setInitialCondition_tree) and the lifted CDF44
wavelet with Bs = 26.
"""
import numpy as np

import adaptive as A
import cylinder_case as CC
import oracle as O


def make_run(case="CDF44"):
    c = CC.CASES[case]
    p = O.Params(skew=False, **CC.ini(case))
    mask = A.CylinderMask2D(p)
    grid = O.uniform_grid(CC.JMIN, 2)
    run = A.AdaptiveRun(p, c["wavelet"], grid, O.alloc(grid, p), 0.0, 0, CC.EPS, Jmin=CC.JMIN, refinement_indicator=c["indicator"],
                        force_maxlevel_dealiasing=True, mask=mask, threshold_mask=True, fd_half_width=2, thresh_comp=c["thresh_comp"])

    def inicond(r):                      # inicond = meanflow (inicond_ACM.f90:285-288)
        r.u[:] = 0.0
        r.u[:, 1] = -1.0
    inicond(run)
    run.adaptive_inicond(inicond)
    return run


def state(run):
    g = run.p.g
    u = run.u[:, :, 0, g:g + CC.BS, g:g + CC.BS]
    chi = np.stack([run.mask.block(int(l), x)[0, 0, g:g + CC.BS, g:g + CC.BS] for l, x in zip(run.grid.level, run.grid.ixyz)])
    return run.grid.level, run.grid.ixyz, run.status, u, run.iteration, run.time, chi


def _run_case(case):
    c, gd = CC.CASES[case], CC.gold(case)
    run = make_run(case)
    chk = case != "CDF40"      # the stored statuses of acm_CDF40 are all 0 (the file does not carry REF_UNSIGNIFICANT_STAY); grids and fields are compared
    errs = {"t0": CC.compare(gd, "t0", *state(run), check_status=chk)}
    stops = {k: t for k, t in c["files"].items() if t > 0.0}
    while run.time < run.p.time_max:
        run.step()
        for k, t in stops.items():
            if abs(run.time - t) <= 1e-15:
                errs[k] = CC.compare(gd, k, *state(run), check_status=chk)
    return case, errs, max(r[2] for r in run.log)


def test_cylinder_fixtures():
    """the four variants run in worker processes started by conftest.py (1 - 2 minutes of CPU each)"""
    from conftest import background
    results = [background("cylinder", c, _run_case) for c in CC.CASES]               # started at collection time (conftest.py)
    for case, errs, nb_rhs_max in results:
        assert set(errs) == set(CC.CASES[case]["files"]), (case, errs)
        assert errs["t0"] == 0.0 and max(errs.values()) <= 1e-12, (case, errs)
        if CC.CASES[case]["nb_rhs_max"]:
            assert nb_rhs_max == CC.CASES[case]["nb_rhs_max"]
