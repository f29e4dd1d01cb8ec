"""Pin the oracle against the reference's own adaptive cylinder regression TESTING/acm/acm_CDF44 (fields written by the reference Fortran
code; tests/golden/cylinder_adapt_CDF44.npz): the grid after the adaptive initial condition (52 blocks on levels 2-5), at t = 0.05
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


def make_run():
    p = O.Params(skew=False, **CC.INI)
    mask = A.CylinderMask2D(p)
    grid = O.uniform_grid(CC.JMIN, 2)
    run = A.AdaptiveRun(p, "CDF44", grid, O.alloc(grid, p), 0.0, 0, CC.EPS, Jmin=CC.JMIN, refinement_indicator="everywhere",
                        force_maxlevel_dealiasing=True, mask=mask, threshold_mask=True, fd_half_width=2)

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


def test_cylinder_fixture():
    gd = CC.gold()
    run = make_run()
    assert CC.compare(gd, "t0", *state(run)) == 0.0
    seen = 0
    while run.time < run.p.time_max:
        run.step()
        if abs(run.time - 0.05) <= 1e-15:
            assert CC.compare(gd, "t1", *state(run)) <= 1e-12
            seen += 1
    assert seen == 1
    assert CC.compare(gd, "t2", *state(run)) <= 1e-12
    assert max(r[2] for r in run.log) == 640
