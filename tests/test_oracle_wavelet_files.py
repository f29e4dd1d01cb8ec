"""The reference's regression group "adaptive" (TESTING/runtests.py: wabbit-post --refine-everywhere, then --coarsen-everywhere, on
TESTING/wavelets/vor_000020000000.h5 -- 2-D, Bs = 32, 112 blocks on levels 2 - 5 -- for CDF20 / 22 / 40 / 42 / 44 / 60 / 62), WHOLE files
(tests/golden/wavelet_files.npz, made by tests/golden/make_golden.py):

  refine everywhere  (sync_ghosts_tree + refine_tree("everywhere"), sparse_to_dense.f90:192-206): 448 blocks; the oracle's refineBlock with
      the ghost nodes its level-jump synchronisation leaves reproduces the reference's file BIT FOR BIT (SHA-256 of the interiors) for the
      predictor orders 2, 4, 6;
  coarsen everywhere (adapt_tree("everywhere") on that file, :209, useCoarseExtension = useSecurityZone = isLiftedWavelet): the 112-block
      grid comes back; the full-tree decomposition, the coarse extension and the reconstruction on a four-level graded grid agree with the
      reference's output bit for bit for the unlifted wavelets and to 4e-15 for CDF22 / 42 / 44 / 62.

This is the one reference fixture that pins CDF44 -- the headline wavelet of the compression leg -- and CDF62 through adapt_tree on a graded
grid with more than one level jump."""
import hashlib
import os

import numpy as np
import pytest

import adaptive as A
import fulltree as FT
import oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wavelet_files.npz"))
BS = int(G["Bs"][0])
WAVELETS = ("CDF20", "CDF22", "CDF40", "CDF42", "CDF44", "CDF60", "CDF62")


def _keys(level, ixy):
    return [(int(l), int(x[0]), int(x[1])) for l, x in zip(level, ixy)]


def _refined(wavelet):
    """the reference's first command: read, sync_ghosts_tree, refine_tree("everywhere")"""
    w = O.setup_wavelet(wavelet)
    g = w.g_default
    lv, ixy = G["in_level"].astype(np.int64), G["in_ixy"].astype(np.int64)
    p = O.Params(dim=2, Bs=(BS, BS, 1), g=g, g_rhs=1, n_eqn=1, domain=(6.283185307179586,) * 3, Jmax=int(lv.max()) + 1)
    grid = O.Grid(level=lv, ixyz=np.concatenate([ixy, np.zeros((len(ixy), 1), np.int64)], axis=1), dim=2)
    u = O.alloc(grid, p)
    u[:, 0, 0, g:g + BS, g:g + BS] = G["in_blocks"]
    run = A.AdaptiveRun(p, wavelet, grid, u, 0.0, 0, 0.0, Jmin=1, refinement_indicator="everywhere", use_coarse_extension=bool(w.lifted),
                        use_security_zone=bool(w.lifted))
    run.sync_ghosts_tree()
    run.refine_tree("everywhere")
    return run, p, w, g


def _interiors_in_file_order(grid, u, g, level, ixy):
    mine = {k: b for b, k in enumerate(_keys(grid.level, grid.ixyz))}
    ref = _keys(level, ixy)
    assert set(mine) == set(ref)
    return np.stack([u[mine[k]][0, 0, g:g + BS, g:g + BS] for k in ref])


@pytest.mark.parametrize("wavelet", WAVELETS)
def test_refine_then_coarsen_everywhere_reproduces_the_reference_files(wavelet):
    run, p, w, g = _refined(wavelet)
    X = w.X
    fine = _interiors_in_file_order(run.grid, run.u, g, G[f"refined_X{X}_level"], G[f"refined_X{X}_ixy"])
    assert fine.shape[0] == 448
    assert np.array_equal(fine[:, ::4, ::4], G[f"refined_X{X}_sample"])
    assert hashlib.sha256(np.ascontiguousarray(fine).astype("<f8").tobytes()).digest() == G[f"refined_X{X}_sha256"].tobytes()      # bit for bit
    # second command: the refined file is read (interiors only), synchronised, adapt_tree("everywhere")
    I = (slice(None), slice(None)) + O.interior(p)
    u0 = np.zeros_like(run.u)
    u0[I] = run.u[I]
    gc, uc, _ = FT.adapt_tree(run.p, w, run.grid, u0, eps=0.0, Jmin=1, indicator="everywhere", use_security_zone=bool(w.lifted),
                              use_coarse_extension=bool(w.lifted))
    coarse = _interiors_in_file_order(gc, uc, g, G[f"coarsened_{wavelet}_level"], G[f"coarsened_{wavelet}_ixy"])
    assert coarse.shape[0] == 112 and sorted(_keys(gc.level, gc.ixyz)) == sorted(_keys(G["in_level"], G["in_ixy"]))     # the input grid is back
    err = float(np.abs(coarse[:, ::2, ::2] - G[f"coarsened_{wavelet}_sample"]).max())
    assert err <= (4.0e-15 if w.lifted else 0.0), err


# what the reference printed when it set up the wavelets of its equidistant 3vortices runs (TESTING/acm/3vortices/3vorticesEqui*/
# log.original.txt:131-141 -- "Increased Nwc to consider FD-stencil size", "Coarse extension will copy SC / delete WC (L,R)", the filters)
SETUP_LOG = {
    ("CDF20", 1): dict(Nsc=(0, 0), Nwc=(2, 2), GD=(-1, [-5.0e-1, 1.0, -5.0e-1]), HR=(-1, [5.0e-1, 1.0, 5.0e-1])),
    ("CDF40", 2): dict(Nsc=(0, 0), Nwc=(4, 4), GD=(-3, [6.25e-2, 0.0, -5.625e-1, 1.0, -5.625e-1, 0.0, 6.25e-2]),
                       HR=(-3, [-6.25e-2, 0.0, 5.625e-1, 1.0, 5.625e-1, 0.0, -6.25e-2])),
    ("CDF60", 3): dict(Nsc=(0, 0), Nwc=(6, 6),
                       GD=(-5, [-1.1719e-2, 0.0, 9.7656e-2, 0.0, -5.8594e-1, 1.0, -5.8594e-1, 0.0, 9.7656e-2, 0.0, -1.1719e-2]),
                       HR=(-5, [1.1719e-2, 0.0, -9.7656e-2, 0.0, 5.8594e-1, 1.0, 5.8594e-1, 0.0, -9.7656e-2, 0.0, 1.1719e-2])),
}


@pytest.mark.parametrize("wavelet,fd_half", list(SETUP_LOG))
def test_wavelet_setup_matches_what_the_reference_logged(wavelet, fd_half):
    w, ref = O.setup_wavelet(wavelet), SETUP_LOG[(wavelet, fd_half)]
    assert (w.Nscl, w.Nscr) == ref["Nsc"]
    assert (max(w.Nwcl, 2 * fd_half), max(w.Nwcr, 2 * fd_half)) == ref["Nwc"]          # the widening of module_wavelets.f90:1404-1417
    F = (len(w.GD) - 1) // 2
    for name, lo_hi in (("GD", (w.gd_lo, w.gd_hi)), ("HR", (w.hr_lo, w.hr_hi))):
        lo, vals = ref[name]
        assert lo_hi == (lo, -lo)
        mine = [getattr(w, name)[k + F] for k in range(lo, -lo + 1)]
        assert np.allclose(mine, vals, rtol=0.0, atol=5.1e-6)                           # printed with five significant digits
    assert (w.hd_lo, w.hd_hi, w.HD[F]) == (0, 0, 1.0) and (w.gr_lo, w.gr_hi, w.GR[F]) == (0, 0, 1.0)
