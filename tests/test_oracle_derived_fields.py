"""The derived fields the reference saved next to its regression states are golden vectors for the operators behind the device statistics
(wgpu_statistics: divergence extrema, enstrophy / max vorticity / helicity) -- and, on the graded grids, one more pin of sync_ghosts_tree
across level jumps, since every stencil at a block boundary reads the synchronised ghost nodes:

  TESTING/acm/3vortices/3vorticesAdapt*/{vor,div}_*.h5   2-D vorticity and divergence (FD2 / FD4 / FD6 stencils) on the adaptive grids at
                                                           t = 10 (all six cases) and t = 15 (FD4_CDF42, FD6_CDF62): lifted wavelets
                                                           restrict through the HD filter, unlifted ones by decimation
  TESTING/acm/bumblebeeFlowEquiFD4_CDF40/vorabs_*.h5      |vorticity| in 3-D (compute_vorticity_abs, FD4)

oracle.vorticity_block / oracle.divergence_block reproduce all of them BIT FOR BIT."""
import os

import numpy as np
import pytest

import adaptive_case as AC
import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _derived(case, level, ixy, vel):
    wavelet, disc = AC.CASES[case]
    w, g = O.setup_wavelet(wavelet), AC.CASE_G[case]
    p = O.Params(g=g, skew=True, **AC.case_ini(case))
    grid = O.Grid(level=level.astype(np.int64), ixyz=np.concatenate([ixy, np.zeros((len(ixy), 1), ixy.dtype)], axis=1).astype(np.int64), dim=2)
    u = O.alloc(grid, p)
    u[:, :2, 0, g:g + AC.BS, g:g + AC.BS] = vel
    # what save_data does before PREPARE_SAVE_DATA: sync_ghosts_tree (all g ghost nodes, the wavelet's restriction filter)
    O.sync_ghosts_leaf(grid, p, u, O.neighbor_table168(grid, p.Jmax), g, g, w.X, bool(w.lifted), ignore_filter=not w.lifted, w=w)
    vor, div = [], []
    for b in range(grid.n):
        dx = [2.0 ** (-float(level[b])) * p.domain[d] / float(AC.BS) for d in range(2)]
        vor.append(O.vorticity_block(p, u[b], dx)[0][0])
        div.append(O.divergence_block(p, u[b], dx)[0])
    return np.stack(vor), np.stack(div), grid


@pytest.mark.parametrize("case", list(AC.CASES))
def test_vorticity_and_divergence_files_of_the_adaptive_cases(case):
    gd = AC.gold(case)
    times = [("t10", gd["t10_u"][:, :2])]
    if "t15_velocity" in gd:
        times.append(("t15", gd["t15_velocity"]))
    for key, vel in times:
        vor, div, grid = _derived(case, gd[f"{key}_level"], gd[f"{key}_ixy"], vel)
        assert len(np.unique(grid.level)) >= 2                                  # a graded grid: the stencils cross level jumps
        assert np.array_equal(vor[:, ::2, ::2], gd[f"{key}_vor"]), (case, key, float(np.abs(vor[:, ::2, ::2] - gd[f"{key}_vor"]).max()))
        assert np.array_equal(div[:, ::2, ::2], gd[f"{key}_div"]), (case, key, float(np.abs(div[:, ::2, ::2] - gd[f"{key}_div"]).max()))
        assert np.abs(gd[f"{key}_vor"]).max() > 2.5


def test_vorticity_magnitude_file_3d():
    gd = np.load(os.path.join(GOLD, "vorabs_3d.npz"))
    Bs, H, dx = int(gd["Bs"][0]), int(gd["H"][0]), float(gd["dx"][0])
    p = O.Params(dim=3, Bs=(Bs,) * 3, g=H, g_rhs=H, n_eqn=4, domain=(3.0, 3.0, 3.0), Jmax=1, discretization="FD_4th_central")
    assert abs(dx - 3.0 / (2 * Bs)) < 1e-15
    for j in range(2):
        u = np.zeros((4,) + gd[f"u{j}"].shape[1:])
        u[:3] = gd[f"u{j}"]                                                    # the block with its two-point halo = a ghosted block with g = 2
        vor = O.vorticity_block(p, u, [dx] * 3)
        mag = np.sqrt(vor[0] ** 2 + vor[1] ** 2 + vor[2] ** 2)
        assert np.array_equal(mag, gd[f"vorabs{j}"]) and mag.max() > 10.0
