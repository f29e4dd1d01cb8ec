"""BASELINE config 5 as the reference specifies it (LIB/POSTPROCESSING/post_compression_unit_test.f90:107-215): one component, domain 2,
Gauss blob of set_block_testing_data, adapt_tree with the full wavelet transformation (coarse extension + security zone exactly for the
lifted wavelets), Nb, refineToEquidistant_tree, relative L2 / Linfty errors -- for CDF40, CDF42 and CDF44 and a spread of the protocol's
51 thresholds, on the GPU through wabbit_b200.compression against the same sequence assembled from the oracle (oracle/fulltree.py
adapt_tree, sync_ghosts_leaf + refine_block): identical Nb, grids and fields bit for bit, identical error curves."""
import numpy as np
import pytest

import fulltree as OFT
import oracle as O
from wabbit_b200 import Forest, WabbitGPU
from wabbit_b200 import compression as CP

from util import orc_grid, orc_params

pytestmark = pytest.mark.gpu


def _key_list(level, ixyz):
    return [(int(l), int(a), int(b), int(c)) for l, (a, b, c) in zip(level, ixyz)]


def _oracle_protocol(w, po, Jmax, Bs, eps, H):
    """steps 1-4 of the protocol on the host; returns Nb, the kept grid's keys, final field {key: interior}, (err_L2, err_Linfty)"""
    uni = Forest.uniform(3, Jmax, Jmax=Jmax)
    grid = orc_grid(uni)
    u = O.alloc(grid, po)
    I = O.interior(po)
    u[(slice(None), 0) + I] = CP.set_block_testing_data(Bs, grid.level, grid.ixyz)
    norm = O.norm_linfty_tree(po, u)
    og, od, _ = OFT.adapt_tree(po, w, grid, u, eps, Jmin=1, norm=norm, level_ref=Jmax, fd_half_width=H,
                               use_security_zone=bool(w.lifted), use_coarse_extension=bool(w.lifted))
    kept = _key_list(og.level, og.ixyz)
    nb = og.n
    data = {k: od[b][(slice(None),) + I].copy() for b, k in enumerate(kept)}
    while min(k[0] for k in data) < Jmax:                               # refineToEquidistant_tree
        ks = sorted(data)
        f = Forest.from_blocks(3, Jmax, np.array([k[0] for k in ks], np.int32), np.array([k[1:] for k in ks], np.int32))
        _, lv, ix, _ = f.active(0)
        fk = _key_list(lv, ix)
        g = O.Grid(level=lv.astype(np.int64), ixyz=ix.astype(np.int64), dim=3)
        uu = O.alloc(g, po)
        for b, k in enumerate(fk):
            uu[b][(slice(None),) + I] = data[k]
        O.sync_ghosts_leaf(g, po, uu, f.neighbors(0)[:, :g.n], po.g, po.g, w.X, bool(w.lifted), ignore_filter=False, w=w)
        new = {}
        for b, (L, x, y, z) in enumerate(fk):
            if L >= Jmax:
                new[(L, x, y, z)] = data[(L, x, y, z)]
                continue
            d = O.refine_block(w.X, po, uu[b])
            for q in range(8):
                qq = ((q >> 1) & 1, q & 1, (q >> 2) & 1)
                new[(L + 1, 2 * x + qq[0], 2 * y + qq[1], 2 * z + qq[2])] = d[q][(slice(None),) + I].copy()
        data = new
    ks = sorted(data)
    exact = CP.set_block_testing_data(Bs, np.array([k[0] for k in ks]), np.array([k[1:] for k in ks]))
    got = np.stack([data[k][0] for k in ks])
    e2 = np.sqrt(((got - exact) ** 2).sum()) / np.sqrt((exact ** 2).sum())
    einf = np.abs(got - exact).max() / np.abs(exact).max()
    return nb, sorted(kept), data, (float(e2), float(einf))


@pytest.mark.parametrize("wavelet,Bs", [("CDF40", 16), ("CDF42", 16), ("CDF44", 16), ("CDF44", 18)])
def test_compression_protocol_matches_oracle(wavelet, Bs):
    Jmax = 3
    w = O.setup_wavelet(wavelet)
    p = CP.compression_params(wavelet, Bs, Jmax)
    po = orc_params(p)
    assert po.n_eqn == 1 and po.g == w.g_default
    H = 2                                                               # FD_4th_central (the default order_discretization)
    uniform = Forest.uniform(3, Jmax, Jmax=Jmax, max_blocks=2 * 8 ** Jmax)
    sol = WabbitGPU(p, max_blocks=uniform.max_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(uniform)
    test = CP.CompressionTest(sol, uniform)
    eps_list = CP.EPS_SWEEP[[0, 20, 30, 35, 40, 45, 50]]
    recs = test.run(eps_list)
    # the device evaluates exp() with its own library: compare the curve with the oracle's on host data below, here only its shape
    nbs = [r["Nb"] for r in recs]
    assert nbs == sorted(nbs, reverse=True) and nbs[-1] < nbs[0] <= 8 ** Jmax
    # bit-for-bit leg: the same protocol with the HOST's data uploaded (numpy exp on both sides)
    hvy, lvl, ixyz, _ = uniform.active(0)
    host0 = np.zeros(sol.host_shape())
    I = O.interior(po)
    host0[(slice(0, len(hvy)), 0) + I] = CP.set_block_testing_data(Bs, lvl, ixyz)
    curve = []
    for eps in eps_list[1:6]:
        nb_o, kept_o, data_o, err_o = _oracle_protocol(w, po, Jmax, Bs, float(eps), H)
        sol.set_forest(uniform)
        sol.upload(host0)
        forest, n0, nb = sol.adapt_tree(uniform, eps=float(eps), Jmin=1, full_tree=True)
        assert nb == nb_o and sorted(_key_list(*forest.active(0)[1:3])) == kept_o, eps
        forest = CP.refineToEquidistant_tree(sol, forest, Jmax)
        assert forest.n_blocks == 8 ** Jmax
        got = np.zeros(sol.host_shape())
        sol.download(got, g_sync=0)
        h2, l2, x2, _ = forest.active(0)
        for h, k in zip(h2, _key_list(l2, x2)):
            assert np.array_equal(got[h - 1][(slice(None),) + I], data_o[k]), (eps, k)
        exact = CP.set_block_testing_data(Bs, l2, x2)
        mine = got[h2 - 1][(slice(None), 0) + I]
        e2 = float(np.sqrt(((mine - exact) ** 2).sum()) / np.sqrt((exact ** 2).sum()))
        einf = float(np.abs(mine - exact).max() / np.abs(exact).max())
        assert abs(e2 - err_o[0]) <= 1e-13 * max(err_o[0], 1e-300) + 1e-18 and einf == err_o[1]
        curve.append((float(eps), nb, e2, einf))
        # the device-data run of the same eps: same Nb, errors equal to rounding of exp()
        r = [r for r in recs if r["eps"] == float(eps)][0]
        assert r["Nb"] == nb and abs(r["err_Linfty"] - einf) <= 1e-9 * max(einf, 1e-12) + 1e-14, (r, einf)
    assert all(a[1] >= b[1] for a, b in zip(curve, curve[1:]))          # fewer blocks for larger thresholds
    sol.close()
