"""Pin the oracle's ADAPTIVE path against the reference's own regression fixtures TESTING/acm/3vortices/3vorticesAdaptFD{2,4,6}_CDF{20,22,40,42,
60,62} -- all six (fields written by the reference Fortran code; sampled into tests/golden/three_vortices_adapt_*.npz by
tests/golden/make_golden.py).  The six parameter files differ in the wavelet and the order of the discretization only.

  t = 10 file: the grid after adapt_inicond -- ONE adapt_tree on the 64-block restart.  Pins the full-tree decomposition, the
      coarsening indicator (Linfty, normalised), the security zone, completeness / gradedness, the coarse extension on the lasting
      interfaces and the CE-optimised reconstruction, for the lifted CDF22 / 42 / 62 and the unlifted CDF20 / 40 / 60 (which take the
      same path with Nsc = 0, useCoarseExtension = 1 in the parameter file): block list and the stored refinement statuses
      (0 / REF_UNSIGNIFICANT_STAY) identical, fields to round-off.  All six cases, every run.
  t = 15 file: 2280 / 2281 passes of main.f90's loop (sync_ghosts_tree -> refine_tree("significant") -> RungeKuttaGeneric -> adapt_tree).
      Pins, in addition, the refinement indicator, gradedness of the refinement, refineBlock, the level-jump ghost synchronisation
      inside the time stepper (FD2 / FD4 / FD6) and the time-step control on a graded grid: iteration counter (5334 / 5335) and final
      time identical, block list (22 ... 109 blocks on levels 2-4) and statuses identical, fields <= 1e-12 (measured: <= 1e-14).
      All six cases, every run (WABBIT_FEWER_FIXTURES=1 keeps FD4_CDF40 / FD4_CDF42 only, for a machine with few cores).

The runs happen in worker processes that conftest.py starts at collection time (3 - 4 minutes of CPU each), in parallel with the rest of the
CPU suite (about five minutes in all on 8 cores).
"""
import os

import numpy as np
import pytest

import adaptive_case as AC


def make_run(case):
    import oracle as O
    import adaptive as A
    case = case if case in AC.CASES else "FD4_" + case
    wavelet, disc = AC.CASES[case]
    lev, ixyz, u0, t, it = AC.restart_fields()
    g = AC.CASE_G[case]
    assert g == max(O.setup_wavelet(wavelet).g_default, AC.FD_HALF_WIDTH[disc])
    p = O.Params(g=g, skew=True, **AC.case_ini(case))
    grid = O.Grid(level=lev, ixyz=ixyz, dim=2)
    u = O.alloc(grid, p)
    u[:, :, 0, g:g + AC.BS, g:g + AC.BS] = u0
    run = A.AdaptiveRun(p, wavelet, grid, u, t, it, AC.EPS, Jmin=AC.JMIN, refinement_indicator="significant", use_coarse_extension=True,
                        use_security_zone=True, fd_half_width=AC.FD_HALF_WIDTH[disc])
    run.sync_ghosts_tree()
    run.adapt_tree()                       # setInitialCondition_tree: read_from_files + adapt_inicond
    return run


def interiors(run):
    g = run.p.g
    return run.u[:, :, 0, g:g + AC.BS, g:g + AC.BS]


@pytest.mark.parametrize("case", list(AC.CASES))
def test_adapt_inicond_fixture(case):
    run = make_run(case)
    err = AC.compare(AC.gold(case), "t10", run.grid.level, run.grid.ixyz, run.status, interiors(run), run.iteration, run.time)
    assert err <= 1e-15, err


def _full_run(case):
    run = make_run(case)
    while run.time < run.p.time_max:
        run.step()
    return case, run.grid.level, run.grid.ixyz, run.status, np.ascontiguousarray(interiors(run)), run.iteration, run.time, \
        max(r[2] for r in run.log)


def full_run_cases():
    """the cases whose t = 15 file is reproduced in this run, longest first"""
    cases = ["FD2_CDF20", "FD2_CDF22", "FD6_CDF62", "FD4_CDF42", "FD4_CDF40", "FD6_CDF60"]
    if os.environ.get("WABBIT_FEWER_FIXTURES"):          # a machine with few cores: one lifted and one unlifted wavelet
        cases = ["FD4_CDF42", "FD4_CDF40"]
    return cases


def test_adaptive_run_fixture():
    from conftest import background
    results = [background("adaptive", c, _full_run) for c in full_run_cases()]     # started at collection time (conftest.py)
    for case, level, ixyz, status, u, iteration, time, nb_rhs_max in results:
        err = AC.compare(AC.gold(case), "t15", level, ixyz, status, u, iteration, time)
        assert err <= 1e-12, (case, err)
        assert nb_rhs_max > len(level)     # the grid was refined before every step
