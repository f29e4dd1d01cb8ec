"""Topology on the device (wgpu_set_grid / wgpu_set_active) against the table-driven route: the gather tables and
patch lists derived on the GPU from the block positions must equal what wgpu_set_topology derives from the host forest's 168-slot
hvy_neighbor table (find_neighbors semantics), entry for entry."""
import numpy as np
import pytest

from util import graded_blocks
from wabbit_b200 import Forest, Params, WabbitGPU

pytestmark = pytest.mark.gpu


def _params(dim, Bs, Jmax, wavelet="CDF44", periodic=(1, 1, 1)):
    X, Y = int(wavelet[3]), int(wavelet[4])
    return Params(dim=dim, domain=(1.0,) * 3, Bs=(Bs, Bs, Bs if dim == 3 else 1), wavelet=wavelet, g=X - 1 + max(Y - 1, 0), g_rhs=2,
                  n_eqn=dim + 1, Jmax=Jmax, discretization="FD_4th_central", periodic=tuple(periodic)).finalize()


def _tables(sol):
    nbr, wnbr, counts, lists = sol.topology_tables()
    act = sol.hvy_active - 1
    return nbr[act], wnbr[act], counts, lists


@pytest.mark.parametrize("dim,J0,Jmax,seed,periodic", [(3, 1, 4, 3, (1, 1, 1)), (3, 2, 4, 7, (1, 1, 1)), (2, 2, 5, 5, (1, 1, 1)),
                                                      (3, 1, 3, 11, (0, 1, 0)), (2, 1, 4, 2, (0, 0, 1)), (3, 3, 3, 0, (1, 1, 1))])
def test_device_topology_equals_the_table_route(dim, J0, Jmax, seed, periodic):
    lv, ix = graded_blocks(dim, J0, Jmax, seed, 0.3)
    forest = Forest.from_blocks(dim, Jmax, lv, ix, periodic=periodic)
    p = _params(dim, 16, Jmax, periodic=periodic)
    out = []
    for host_tables in (True, False):
        sol = WabbitGPU(p, max_blocks=forest.n_blocks + 8, device=0)
        sol.setup_wavelet("CDF44")
        sol.set_forest(forest, host_tables=host_tables)
        out.append(_tables(sol))
        sol.close()
    (n0, w0, c0, l0), (n1, w1, c1, l1) = out
    assert c0 == c1, (c0, c1)
    assert np.array_equal(n0, n1)
    if c0["has_jumps"]:
        assert np.array_equal(w0, w1)
    for name in ("jump", "wjump", "ce", "rst", "int", "bnd"):
        assert np.array_equal(l0[name][0], l1[name][0]), name
        assert np.array_equal(l0[name][1], l1[name][1]), name
    if Jmax > J0:
        assert c0["has_jumps"] and c0["n_jump"] > 0 and c0["n_ce"] > 0 and c0["n_rst"] > 0


def test_set_active_names_a_pass_of_the_registered_blocks():
    """registering once and naming sub-lists afterwards (the passes of adapt_tree's full-tree transformation) gives, for the named blocks,
    the rows the same call with only that list active gives"""
    lv, ix = graded_blocks(3, 1, 3, 4, 0.35)
    forest = Forest.from_blocks(3, 3, lv, ix)
    hvy, lvl, _, tc = forest.active(0)
    p = _params(3, 16, 3)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks, device=0)
    sol.setup_wavelet("CDF44")
    sol.set_grid(hvy, lvl, tc)
    full = _tables(sol)
    sub = hvy[lvl == lvl.max()]
    sol.set_active(sub)
    nbr, wnbr, counts, lists = _tables(sol)
    assert counts["n_active"] == len(sub)
    sel = np.flatnonzero(lvl == lvl.max())
    same = full[0][sel] >= 0
    assert np.array_equal(np.where(same, nbr, 0), np.where(same, full[0][sel], 0))
    assert set(lists["ce"][0].tolist()) <= set((sub - 1).tolist())
    sol.set_active(hvy)
    again = _tables(sol)
    assert np.array_equal(again[0], full[0]) and np.array_equal(again[1], full[1]) and again[2] == full[2]
    sol.close()


def test_duplicate_position_is_rejected():
    from wabbit_b200.solver import WabbitAbort
    forest = Forest.uniform(3, 1)
    hvy, lvl, _, tc = forest.active(0)
    p = _params(3, 16, 1)
    sol = WabbitGPU(p, max_blocks=16, device=0)
    tc2 = tc.copy()
    tc2[1] = tc2[0]
    with pytest.raises(WabbitAbort):
        sol.set_grid(hvy, lvl, tc2)
    sol.set_grid(hvy, lvl, tc)      # the context stays usable
    sol.close()
