"""filter_wrapper on the device (wgpu_filter) against the oracle's restatement (oracle.filter_wrapper: generate_superviscosity_stencil +
blockFilterXYZ_vct, LIB/TIME/filter_wrapper.f90, LIB/WAVELETS/module_wavelets.f90:307-401) on ghost-synchronised data: bit for bit, on
equidistant and graded grids, with component and level selection."""
import numpy as np
import pytest

import oracle as O
from wabbit_b200 import Forest, WabbitAbort, WabbitGPU

from util import graded_blocks, orc_grid, orc_params, tg_params

pytestmark = pytest.mark.gpu


def test_stencils_preserve_constants():
    for t in ("explicit_3pt", "explicit_5pt", "explicit_7pt", "explicit_9pt", "superviscosity_10th"):
        c = O.superviscosity_stencil(t)
        assert abs(sum(c.values()) - 1.0) <= 1e-15 and c[0] > 0.0 and len(c) % 2 == 1


@pytest.mark.parametrize("filter_type,graded,Bs,sel", [("explicit_5pt", False, 16, "all"), ("explicit_7pt", False, 18, "comp"), ("explicit_3pt", True, 16, "all"),
                                                      ("explicit_5pt", True, 16, "maxlevel"), ("explicit_7pt", True, 22, "notmax")])
def test_filter_wrapper_bit_exact(filter_type, graded, Bs, sel):
    wavelet = "CDF44"
    w = O.setup_wavelet(wavelet)
    if graded:
        lv, ix = graded_blocks(3, 1, 3, seed=8)
        J = int(lv.max())                      # Jmax = the finest level present: "only_maxlevel" then selects a strict subset
        forest = Forest.from_blocks(3, J, lv, ix)
    else:
        J = 2
        forest = Forest.uniform(3, 2, Jmax=2)
    p = tg_params(Bs=Bs, J=J, wavelet_g=w.g_default)
    p.wavelet = wavelet
    po, grid = orc_params(p), orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    u = O.alloc(grid, po)
    u[:] = np.random.default_rng(4).standard_normal(u.shape)
    sol.upload(u)
    comp = [1, 0, 1, 0] if sel == "comp" else None
    kw = dict(only_maxlevel=sel == "maxlevel", all_except_maxlevel=sel == "notmax")
    sol.filter_wrapper(filter_type, comp, **kw)
    got = np.zeros_like(u)
    sol.download(got, g_sync=0)
    ref = u.copy()
    O.sync_ghosts_leaf(grid, po, ref, forest.neighbors(0)[:, :grid.n], po.g, po.g, w.X, bool(w.lifted), ignore_filter=False, w=w)    # sync_ghosts_tree
    O.filter_wrapper(grid, po, ref, filter_type, comp, **kw)
    I = (slice(None), slice(None)) + O.interior(po)
    assert np.array_equal(got[I], ref[I])
    assert not np.array_equal(ref[I], u[I])
    if sel == "comp":
        assert np.array_equal(got[I][:, 1], u[I][:, 1])
    sol.close()


def test_filter_errors():
    forest = Forest.uniform(3, 1, Jmax=1)
    p = tg_params(Bs=16, J=1)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    for args, code in ((("explicit_4pt",), 251107), (("explicit_9pt",), 251108), (("explicit_5pt", None, True, True), 251106)):
        with pytest.raises(WabbitAbort) as e:
            sol.filter_wrapper(*args)
        assert e.value.code == code, (args, e.value.code)
    sol.close()
