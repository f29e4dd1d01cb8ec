"""The C/OpenMP tree loops (oracle/orc_tree.c, used as CPU baseline) must equal the NumPy restatement bit for bit."""
import numpy as np

import oracle as O


def test_c_tree_loops_equal_numpy_loops():
    p = O.Params(dim=3, Bs=(16, 16, 16), g=3, g_rhs=2, domain=(6.283185307179586,) * 3, Jmax=2, discretization="FD_4th_central",
                 skew=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0, time_max=1e9, tsave_stats=0.03, u_mean_set=(0, 0, 0))
    grid = O.uniform_grid(2)
    u1 = O.alloc(grid, p)
    O.inicond_taylor_green(grid, p, u1)
    rng = np.random.default_rng(1)
    u1 += 0.01 * rng.random(u1.shape)
    u2 = u1.copy()
    w1 = [O.alloc(grid, p) for _ in range(5)]
    w2 = np.zeros((5,) + u2.shape)
    nbr, dxb = O.nbr_table(grid), O.dx_table(grid, p)
    t = 0.0
    for it in range(3):
        d1 = O.rk_generic(grid, p, u1, w1, t)
        d2 = O.rk_step_c(grid, p, u2, w2, t, nbr, dxb)
        assert d1 == d2
        t += d1
    assert np.array_equal(u1, u2)
    # the timing build (-O3 -march=native, FMA allowed) stays within round-off of the parity build
    u3 = u2.copy()
    O.rk_step_c(grid, p, u2, w2, t, nbr, dxb, fast=False)
    O.rk_step_c(grid, p, u3, w2, t, nbr, dxb, fast=True)
    assert np.abs(u2 - u3).max() < 1e-12


def test_neighbour_table_for_all_blocks_at_once_equals_the_block_loop():
    """oracle.neighbor_table168 (sorted keys + binary search) against the block-by-block restatement of find_neighbor, on graded 2-D / 3-D
    grids, with Jmax at and above the finest level present, periodic and not"""
    import numpy as np
    import oracle as O
    from util import graded_blocks
    for dim, J0, Jm, seed in [(2, 1, 4, 1), (3, 1, 3, 2), (2, 2, 5, 7), (3, 1, 2, 9)]:
        lv, ix = graded_blocks(dim, J0, Jm, seed)
        g = O.Grid(level=lv.astype(np.int64), ixyz=ix.astype(np.int64), dim=dim)
        for Jmax in (int(lv.max()), int(lv.max()) + 3):
            for per in ((1, 1, 1), (0, 1, 0)):
                assert np.array_equal(O.neighbor_table168(g, Jmax, per), O.neighbor_table168_loop(g, Jmax, per)), (dim, seed, Jmax, per)
    for g, J in ((O.uniform_grid(0, 3), 3), (O.uniform_grid(2, 2), 2)):
        assert np.array_equal(O.neighbor_table168(g, J), O.neighbor_table168_loop(g, J))
