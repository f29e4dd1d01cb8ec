"""Pin the wavelet part of the oracle: (1) known-answer filter coefficients derived from the reference source
(SURVEY 8c), (2) the reference's own unit-test property IWT(FWT(u)) = u <= 1e-14 (unit_test_waveletDecomposition.f90),
(3) the reference's regression fields TESTING/wavelets/*: `prediction`/refineBlock and the low-pass decomposition filter +
decimation alignment, on blocks whose neighbourhood is uniform (tests/golden/wavelet_blocks.npz)."""
import os
from fractions import Fraction as F

import numpy as np
import pytest

import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_known_answer_filters():
    w = O.setup_wavelet("CDF44")
    hd = w.taps("HD")
    exp = [F(-1, 512), 0, F(9, 256), F(-1, 32), F(-63, 512), F(9, 32), F(87, 128), F(9, 32), F(-63, 512), F(-1, 32), F(9, 256), 0, F(-1, 512)]
    assert [hd[k] for k in range(-6, 7)] == [float(x) for x in exp]
    assert [w.taps("GD")[k] for k in range(-3, 4)] == [1 / 16, 0, -9 / 16, 1, -9 / 16, 0, 1 / 16]
    assert [w.taps("HR")[k] for k in range(-3, 4)] == [-1 / 16, 0, 9 / 16, 1, 9 / 16, 0, -1 / 16]
    assert all(w.taps("GR")[k] == (-1) ** k * hd[k] for k in range(-6, 7))
    assert (w.g_default, w.Nscl, w.Nscr, w.Nwcl, w.Nwcr, w.Nreconl, w.Nreconr) == (6, 5, 6, 8, 9, 14, 15)
    w = O.setup_wavelet("CDF42")
    assert [w.taps("HD")[k] for k in range(-4, 5)] == [1 / 64, 0, -1 / 8, 1 / 4, 23 / 32, 1 / 4, -1 / 8, 0, 1 / 64] and w.g_default == 4
    w = O.setup_wavelet("CDF22")
    assert [w.taps("HD")[k] for k in range(-2, 3)] == [-1 / 8, 1 / 4, 3 / 4, 1 / 4, -1 / 8]
    assert [w.taps("GD")[k] for k in range(-1, 2)] == [-1 / 2, 1, -1 / 2] and w.g_default == 2
    w = O.setup_wavelet("CDF40")
    assert w.taps("HD") == {0: 1.0} and w.g_default == 3 and not w.lifted
    for name in ("CDF20", "CDF22", "CDF40", "CDF42", "CDF44", "CDF60", "CDF62", "CDF64", "CDF66"):
        w = O.setup_wavelet(name)   # sums printed by the reference, module_wavelets.f90:1551-1554
        assert abs(sum(w.taps("HD").values()) - 1) < 1e-15 and abs(sum(w.taps("HR").values()) - 2) < 1e-15
        assert abs(sum(w.taps("GD").values())) < 1e-15
        assert abs(sum(w.taps("GR").values()) - (0 if w.lifted else 1)) < 1e-15   # unlifted: GR = delta


@pytest.mark.parametrize("name", ["CDF20", "CDF22", "CDF40", "CDF42", "CDF44", "CDF60", "CDF62"])
@pytest.mark.parametrize("dim", [2, 3])
def test_fwt_iwt_invertible(name, dim):
    """unit_test_waveletDecomposition: random data on an equidistant periodic grid, rel. L2 error <= 1e-14"""
    w = O.setup_wavelet(name)
    Bs = 20 if dim == 3 else 24
    p = O.Params(dim=dim, Bs=(Bs, Bs, Bs if dim == 3 else 1), g=w.g_default, g_rhs=w.g_default, n_eqn=2, Jmax=1)
    grid = O.uniform_grid(1, dim)
    rng = np.random.default_rng(11)
    u = O.alloc(grid, p)
    u[...] = rng.random(u.shape)
    O.sync_ghosts_same_level(grid, p, u, p.g, p.g)
    wd = np.zeros_like(u)
    O.fwt_tree(w, p, u, wd)
    O.sync_ghosts_same_level(grid, p, wd, p.g, p.g)
    r = np.zeros_like(u)
    O.iwt_tree(w, p, wd, r)
    it = (slice(None), slice(None)) + O.interior(p)
    err = np.sqrt(((r[it] - u[it]) ** 2).sum() / (u[it] ** 2).sum())
    assert err <= 1e-14, err


def test_threshold_block_semantics():
    """details = max |WC| with pure SC positions removed; status -1 iff all(detail <= eps*norm) (threshold_block.f90:96-121)"""
    w = O.setup_wavelet("CDF40")
    p = O.Params(dim=3, Bs=(8, 8, 8), g=3, n_eqn=2, Jmax=1)
    wd = np.zeros((1, 2, 14, 14, 14))
    wd[0, :, 3:11:2, 3:11:2, 3:11:2] = 50.0      # pure scaling coefficients must not count
    wd[0, 0, 4, 3, 3] = -0.25                     # a detail of component 0
    wd[0, 1, 3, 4, 4] = 0.5
    st, det = O.threshold_tree(p, wd, [1], eps=0.3)
    assert det.tolist() == [[0.25, 0.5]] and st[0] == 0
    st, det = O.threshold_tree(p, wd, [1], eps=0.5)
    assert st[0] == -1
    st, det = O.threshold_tree(p, wd, [1], eps=0.3, thresh_comp=[1, 0])
    assert det.tolist() == [[0.25, 0.0]] and st[0] == -1
    st, det = O.threshold_tree(p, wd, [1], eps=1.0, norm=[0.2, 1.0])
    assert st[0] == 0                             # 0.25 > 1.0*0.2
    st, det = O.threshold_tree(p, wd, [1], eps=0.3, thresh_comp=[2, 2])
    assert det.tolist() == [[0.5, 0.5]]
    # L2 renormalisation: factor 2^((Jref-J-1)d/2), then /2 for every direction in which the position is an SC position
    st, det = O.threshold_tree(p, wd, [1], eps=1.0, eps_norm="L2")
    f = 2.0 ** (-2 * 3 / 2.0)
    assert det.tolist() == [[0.25 * f / 2 / 2, 0.5 * f / 2]]


def test_prediction_and_decomposition_against_reference_fields():
    gold = np.load(os.path.join(GOLD, "wavelet_blocks.npz"))
    Bs = int(gold["Bs"][0])
    assert int(gold["n"][0]) >= 2
    for k in range(int(gold["n"][0])):
        # --- refineBlock with the CDF4x predictor (TESTING/wavelets/adaptive_CDF40/vor_00100.h5)
        p = O.Params(dim=2, Bs=(Bs, Bs, 1), g=3, n_eqn=1, Jmax=7)
        d = O.refine_block(4, p, gold[f"mother{k}"][None, None])
        for kd in range(4):
            assert np.array_equal(d[kd, 0, 0, 3:-3, 3:-3], gold[f"daughters{k}"][kd]), (k, kd)
        # --- decomposition: the mother's values are the scaling coefficients of the four daughters
        for name in ("CDF22", "CDF42", "CDF44", "CDF62"):
            fine = gold[f"fine_X{name[3]}_{k}"]   # (2Bs+12)^2 : the four daughters with a 6-wide ring of neighbour data
            w = O.setup_wavelet(name)
            g = w.g_default
            pw = O.Params(dim=2, Bs=(Bs, Bs, 1), g=g, n_eqn=1, Jmax=7)
            ref = gold[f"coarse_{name}_{k}"]
            for kd in range(4):
                ox, oy = (kd // 2) % 2, kd % 2
                y0, x0 = 6 + oy * Bs - g, 6 + ox * Bs - g
                blk = np.ascontiguousarray(fine[y0:y0 + Bs + 2 * g, x0:x0 + Bs + 2 * g])[None, None, None]
                wd = np.zeros_like(blk)
                O.fwt_tree(w, pw, blk, wd)
                sc = wd[0, 0, 0, g:g + Bs:2, g:g + Bs:2]
                exp = ref[oy * Bs // 2:(oy + 1) * Bs // 2, ox * Bs // 2:(ox + 1) * Bs // 2]
                assert np.abs(sc - exp).max() <= 1e-13 * np.abs(exp).max(), (k, name, kd, np.abs(sc - exp).max())
