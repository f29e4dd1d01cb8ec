"""Pin the oracle's ADAPTIVE path against the reference's own regression fixtures TESTING/acm/3vortices/3vorticesAdaptFD4_CDF4{0,2}
(fields written by the reference Fortran code; sampled into tests/golden/three_vortices_adapt_*.npz by tests/golden/make_golden.py).

  t = 10 file: the grid after adapt_inicond -- ONE adapt_tree on the 64-block restart.  Pins the full-tree decomposition, the
      coarsening indicator (Linfty, normalised), the security zone, completeness / gradedness, the coarse extension on the lasting
      interfaces and the CE-optimised reconstruction: block list and the stored refinement statuses (0 / REF_UNSIGNIFICANT_STAY)
      identical, fields to round-off.
  t = 15 file: 2281 passes of main.f90's loop (sync_ghosts_tree -> refine_tree("significant") -> RungeKuttaGeneric -> adapt_tree).
      Pins, in addition, the refinement indicator, gradedness of the refinement, refineBlock, the level-jump ghost synchronisation
      inside the time stepper and the time-step control on a graded grid: iteration counter (5335) and final time identical, block
      list (61 / 58 blocks on levels 2-4) and statuses identical, fields <= 1e-12.

Both wavelets run in worker processes that conftest.py starts at collection time (about 4 minutes of CPU each), in parallel with the
rest of the CPU suite.
"""
import numpy as np
import pytest

import adaptive_case as AC


def make_run(wavelet):
    import oracle as O
    import adaptive as A
    lev, ixyz, u0, t, it = AC.restart_fields()
    g = AC.WAVELET_G[wavelet]
    p = O.Params(g=g, skew=True, **AC.INI)
    grid = O.Grid(level=lev, ixyz=ixyz, dim=2)
    u = O.alloc(grid, p)
    u[:, :, 0, g:g + AC.BS, g:g + AC.BS] = u0
    run = A.AdaptiveRun(p, wavelet, grid, u, t, it, AC.EPS, Jmin=AC.JMIN, refinement_indicator="significant", use_coarse_extension=True,
                        use_security_zone=True, fd_half_width=2)
    run.sync_ghosts_tree()
    run.adapt_tree()                       # setInitialCondition_tree: read_from_files + adapt_inicond
    return run


def interiors(run):
    g = run.p.g
    return run.u[:, :, 0, g:g + AC.BS, g:g + AC.BS]


@pytest.mark.parametrize("wavelet", ["CDF40", "CDF42"])
def test_adapt_inicond_fixture(wavelet):
    run = make_run(wavelet)
    err = AC.compare(AC.gold(wavelet), "t10", run.grid.level, run.grid.ixyz, run.status, interiors(run), run.iteration, run.time)
    assert err <= 1e-15, err


def _full_run(wavelet):
    run = make_run(wavelet)
    while run.time < run.p.time_max:
        run.step()
    return wavelet, run.grid.level, run.grid.ixyz, run.status, np.ascontiguousarray(interiors(run)), run.iteration, run.time, \
        max(r[2] for r in run.log)


def test_adaptive_run_fixture():
    from conftest import background
    results = [background("adaptive", w, _full_run) for w in ("CDF40", "CDF42")]     # started at collection time (conftest.py)
    for wavelet, level, ixyz, status, u, iteration, time, nb_rhs_max in results:
        err = AC.compare(AC.gold(wavelet), "t15", level, ixyz, status, u, iteration, time)
        assert err <= 1e-12, (wavelet, err)
        assert nb_rhs_max > len(level)     # the grid was refined before every step
