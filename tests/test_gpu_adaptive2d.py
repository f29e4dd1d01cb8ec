"""The reference's 2-D adaptive regression case TESTING/acm/3vortices/3vorticesAdaptFD4_CDF4{0,2} on the GPU (BASELINE config 1's kind of
run: 2-D ACM, adaptive every step, one rank): restart from the stored t = 10 fields, adapt_inicond, then main.f90's loop
refine_tree("significant") -> RungeKuttaGeneric -> adapt_tree to t = 15, all heavy data on the device (wabbit_b200.timeloop.AdaptiveLoop).

Compared (through the C ABI) with
  * the fields the reference Fortran code wrote (tests/golden/three_vortices_adapt_*.npz): grid after adapt_inicond and grid at t = 15
    identical (block lists and refinement statuses), iteration counter identical, fields within 1e-10 (FMA contraction on the GPU, none
    in the reference build; the run is 2281 steps long);
  * the oracle (oracle/adaptive.py, itself pinned by the same fixtures) in lockstep over the first steps: block lists, statuses and dt
    identical after every step, fields <= 1e-12.
Also: the 2-D wavelet kernels (decomposition / reconstruction, uniform and graded grids) and the 2-D level-jump ghost patches of the
stage kernel against the oracle.
"""
import numpy as np
import pytest

import adaptive_case as AC
import oracle as O
from wabbit_b200 import Forest, Params, WabbitGPU
from wabbit_b200.solver import HVY_BLOCK, HVY_WORK
from wabbit_b200.timeloop import AdaptiveLoop

from test_oracle_adaptive import interiors, make_run
from util import orc_grid, orc_params

pytestmark = pytest.mark.gpu

MAXB = 400


def gpu_params(wavelet):
    p = Params(wavelet=wavelet, g=AC.WAVELET_G[wavelet], skew_symmetry=True, eps=AC.EPS, eps_normalized=True, eps_norm="Linfty", Jmin=AC.JMIN,
               useCoarseExtension=1, useSecurityZone=1, adapt_tree=True, refinement_indicator="significant", **AC.INI)
    return p.finalize()


def upload_blocks(sol, forest, level, ixyz, u_int):
    """interiors [nb, nc, Bs, Bs] given per (level, ixyz) -> device, at the forest's slots"""
    hvy, lvl, pos, _ = forest.active(0)
    at = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(level, ixyz))}
    g = sol.params.g
    host = np.zeros(sol.host_shape())
    for h, l, x in zip(hvy, lvl, pos):
        host[h - 1, :, 0, g:g + AC.BS, g:g + AC.BS] = u_int[at[(int(l), int(x[0]), int(x[1]))]]
    sol.upload(host, hvy_ids=hvy)


def download_blocks(sol, forest):
    hvy, lvl, pos, _ = forest.active(0)
    g = sol.params.g
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    return lvl, pos, got[hvy - 1][:, :, 0, g:g + AC.BS, g:g + AC.BS]


def make_loop(wavelet):
    lev, ixyz, u0, t, it = AC.restart_fields()
    p = gpu_params(wavelet)
    forest = Forest.from_blocks(2, p.Jmax, lev.astype(np.int32), ixyz.astype(np.int32), max_blocks=MAXB)
    sol = WabbitGPU(p, max_blocks=MAXB)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    upload_blocks(sol, forest, lev, ixyz, u0)
    loop = AdaptiveLoop(sol, forest, t, it)
    loop.adapt_tree()                       # setInitialCondition_tree: read_from_files + adapt_inicond
    return loop


def same_grid(loop, run):
    lvl, pos, u = download_blocks(loop.sol, loop.forest)
    okey = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(run.grid.level, run.grid.ixyz))}
    keys = [(int(l), int(x[0]), int(x[1])) for l, x in zip(lvl, pos)]
    assert sorted(keys) == sorted(okey)
    ou = interiors(run)
    o = np.array([okey[k] for k in keys])
    assert np.array_equal(np.asarray(loop.status), run.status[o])
    return float(np.abs(u - ou[o]).max() / np.abs(ou).max())


@pytest.mark.parametrize("wavelet", ["CDF40", "CDF42"])
def test_adapt_inicond_2d(wavelet):
    loop = make_loop(wavelet)
    lvl, pos, u = download_blocks(loop.sol, loop.forest)
    err = AC.compare(AC.gold(wavelet), "t10", lvl, pos, loop.status, u, loop.iteration, loop.time)
    assert err <= 1e-15, err
    run = make_run(wavelet)
    assert same_grid(loop, run) == 0.0      # wavelet arithmetic is not contracted on the device: bit for bit
    loop.sol.close()


@pytest.mark.parametrize("wavelet", ["CDF40", "CDF42"])
def test_adaptive_lockstep_2d(wavelet):
    loop, run = make_loop(wavelet), make_run(wavelet)
    for _ in range(12):
        dt_g = loop.step()
        dt_o = run.step()
        assert loop.log[-1][2:4] == run.log[-1][2:4], (loop.log[-1], run.log[-1])      # blocks on the RHS grid, blocks after adapt_tree
        assert abs(dt_g - dt_o) <= 1e-13 * dt_o
        assert same_grid(loop, run) <= 1e-12
    assert loop.log[-1][2] > loop.log[-1][3]
    loop.sol.close()


@pytest.mark.parametrize("wavelet", ["CDF40", "CDF42"])
def test_adaptive_run_fixture_2d(wavelet):
    """the whole regression run of the reference on the device: 2281 adaptive steps to t = 15"""
    loop = make_loop(wavelet)
    p = loop.sol.params
    while loop.time < p.time_max:
        loop.step()
    lvl, pos, u = download_blocks(loop.sol, loop.forest)
    err = AC.compare(AC.gold(wavelet), "t15", lvl, pos, loop.status, u, loop.iteration, loop.time)
    print(f"\n3vorticesAdaptFD4_{wavelet} on the GPU: {loop.iteration - 3054} adaptive steps, {loop.forest.n_blocks} blocks at t = {loop.time}, "
          f"max |u - reference| = {err:.3e}, blocks on the RHS grid max {max(r[2] for r in loop.log)}")
    assert err <= 1e-10, err
    loop.sol.close()


# ---------------------------------------------------------------------------------------------------------------------- kernels
def _graded_2d(seed, Jmax=4):
    """a graded 2-D grid: the adapted 3vortices grid, refined where the restart field is significant"""
    run = make_run("CDF40")
    rng = np.random.default_rng(seed)
    run.status = np.where(rng.random(run.grid.n) < 0.5, 0, 9)
    run.refine_tree("significant")
    return run.grid


@pytest.mark.parametrize("wavelet", ["CDF40", "CDF42", "CDF44", "CDF22", "CDF62"])
def test_wavelet_kernels_2d_graded(wavelet):
    """FWT (with the ghost synchronisation across level jumps, filtered restriction for lifted wavelets) and IWT(FWT(u)) on a graded 2-D
    grid: coefficients bit for bit those of the oracle's sync_ghosts_tree + waveletDecomposition_optimized_block"""
    grid = _graded_2d(1)
    w = O.setup_wavelet(wavelet)
    g = w.g_default
    p = Params(wavelet=wavelet, g=g, skew_symmetry=True, **AC.INI).finalize()
    po = orc_params(p)
    forest = Forest.from_blocks(2, p.Jmax, grid.level.astype(np.int32), grid.ixyz.astype(np.int32), max_blocks=MAXB)
    og = orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=MAXB)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    u = O.alloc(og, po)
    I = (slice(None), slice(None)) + O.interior(po)
    u[I] = np.random.default_rng(7).standard_normal(u[I].shape)
    host = np.zeros(sol.host_shape())
    host[:og.n] = u
    sol.upload(host, hvy_ids=np.arange(1, og.n + 1, dtype=np.int32))
    nbr = O.neighbor_table168(og, p.Jmax)
    O.sync_ghosts_leaf(og, po, u, nbr, g, g, w.X, bool(w.lifted), ignore_filter=not w.lifted, w=w)
    # ghosted download = what sync_ghosts_tree leaves in the ghost nodes
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=g)
    assert np.array_equal(got[:og.n], u)
    wd = np.zeros_like(u)
    O.fwt_tree(w, po, u, wd)
    sol.waveletDecomposition_tree((HVY_BLOCK, 0), (HVY_WORK, 2))
    gw = np.zeros(sol.host_shape())
    sol.download(gw, HVY_WORK, 2, g_sync=0)
    assert np.array_equal(gw[:og.n][I], wd[I])
    sol.close()


def test_rk4_2d_graded():
    """RungeKuttaGeneric on a graded 2-D grid (level-jump face patches of the 2-D stage kernel) against the oracle"""
    grid = _graded_2d(2)
    wavelet = "CDF40"
    p = gpu_params(wavelet)
    po = orc_params(p)
    forest = Forest.from_blocks(2, p.Jmax, grid.level.astype(np.int32), grid.ixyz.astype(np.int32), max_blocks=MAXB)
    og = orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=MAXB)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    u = O.alloc(og, po)
    I = (slice(None), slice(None)) + O.interior(po)
    rng = np.random.default_rng(11)
    x = np.linspace(0, 1, AC.BS)
    u[I] = 0.3 * rng.standard_normal((og.n, 3, 1, 1, 1)) + 0.1 * np.sin(2 * np.pi * x)[None, None, None, None, :] + \
        0.01 * rng.standard_normal(u[I].shape)
    host = np.zeros(sol.host_shape())
    host[:og.n] = u
    sol.upload(host, hvy_ids=np.arange(1, og.n + 1, dtype=np.int32))
    nbr = O.neighbor_table168(og, p.Jmax)
    w = O.setup_wavelet(wavelet)
    work = [O.alloc(og, po) for _ in range(5)]
    sync = lambda h: O.sync_ghosts_leaf(og, po, h, nbr, p.g_rhs, p.g_rhs, w.X, bool(w.lifted))
    t = 10.0
    for it in range(2):
        dt_o = O.rk_generic(og, po, u, work, t, sync=sync)
        dt_g = sol.RungeKuttaGeneric(t, it)
        assert abs(dt_g - dt_o) <= 1e-14 * dt_o
        t += dt_o
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    err = np.abs(got[:og.n][I] - u[I]).max() / np.abs(u[I]).max()
    assert err <= 1e-12, err
    sol.close()
