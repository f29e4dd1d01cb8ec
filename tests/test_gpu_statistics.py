"""Device-side mask function and statistics: wgpu_create_mask against the host / oracle generators of create_mask_2D_ACM
(cylinder + p-norm sponge) and create_mask_3D_ACM (sphere), wgpu_statistics against the oracle's numpy restatement of STATISTICS_ACM's
integral stage (oracle.statistics_acm; pinned on the CPU by the analytic Taylor-Green integrals, tests/test_oracle_statistics.py)."""
import numpy as np
import pytest

import adaptive as OA
import oracle as O
from wabbit_b200 import Forest, Params, WabbitGPU
from wabbit_b200.mask import CylinderMask2D, SphereMask3D
from wabbit_b200.solver import HVY_MASK

from util import graded_blocks, orc_grid, orc_params, tg_params

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-12):
    return abs(a - b) <= tol * max(abs(a), abs(b), 1e-30) + 1e-13


def test_sphere_mask_and_statistics_3d():
    lv, ix = graded_blocks(3, 1, 3, seed=4)
    forest = Forest.from_blocks(3, 3, lv, ix)
    p = tg_params(Bs=16, J=3)
    p.penalization, p.C_eta = True, 1.0e-2
    p = p.finalize()
    po = orc_params(p)
    grid = orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet("CDF40")                 # the predictor order of the level-jump ghost patches
    sol.set_forest(forest)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    u += 0.05 * np.random.default_rng(1).standard_normal(u.shape)
    sol.upload(u)
    sph = SphereMask3D(p, center=(3.0, 3.1, 3.2), radius=0.9, velocity=(0.5, 0.3, -0.2))
    osph = OA.SphereMask3D(po, center=(3.0, 3.1, 3.2), radius=0.9, velocity=(0.5, 0.3, -0.2))
    t = 0.37
    sph.fill_device(sol, t)
    got = np.zeros((grid.n, 6) + u.shape[2:])
    sol.download(got, HVY_MASK, g_sync=0)
    ref = np.stack([osph.block(int(l), x, t) for l, x in zip(grid.level, grid.ixyz)])
    I = (slice(None), slice(None)) + O.interior(po)
    assert np.abs(got[I] - ref[I]).max() <= 1e-14 and ref[I][:, 0].max() == 1.0 and 0.0 < ref[I][:, 0].mean() < 0.2
    # statistics: the oracle reads the ghost-synchronised state and the same mask
    synced = u.copy()
    O.sync_ghosts_leaf(grid, po, synced, forest.neighbors(0)[:, :grid.n], po.g, po.g, 4, True)
    want = O.statistics_acm(grid, po, synced, ref)
    have = sol.statistics_ACM(t, with_divergence=True)
    for k in want:
        tol = 1e-9 if k.startswith("div") else 1e-12          # the divergence is recovered from the pressure row of the RHS (c0^2 = 100)
        assert _close(have[k], want[k], tol), (k, have[k], want[k])
    assert want["mask_volume"] > 1.0 and abs(want["force_x"]) > 0.0 and want["div_max"] > 0.0
    _check_vorticity_entries(sol, grid, po, synced, t)
    sol.close()


def _check_vorticity_entries(sol, grid, po, synced, t=0.0):
    """enstrophy / max_vort / helicity / dissipation (WGPU_STAT_VORTICITY) against oracle.vorticity_statistics_acm on the synchronised state;
    the other 19 entries do not change with the flag"""
    want = O.vorticity_statistics_acm(grid, po, synced)
    have = sol.statistics_ACM(t, with_divergence=False, with_vorticity=True)
    base = sol.statistics_ACM(t, with_divergence=False)
    assert set(have) == set(base) | set(want) and all(have[k] == base[k] for k in base)
    scale = want["enstrophy"]
    assert scale > 0.0 and want["max_vort"] > 0.0 and (po.dim == 2 or want["helicity"] != 0.0) and want["dissipation"] != 0.0
    for k in want:
        assert abs(have[k] - want[k]) <= 1e-12 * max(abs(want[k]), scale if k == "helicity" else 0.0), (k, have[k], want[k])


@pytest.mark.parametrize("disc,g,Bs", [("FD_6th_central", 3, 16), ("FD_2nd_central", 3, 18), ("FD_4th_central", 3, 22)])
def test_vorticity_statistics_on_equidistant_grids(disc, g, Bs):
    forest = Forest.uniform(3, 2, Jmax=2)
    p = tg_params(Bs=Bs, J=2)
    p.discretization, p.g, p.g_rhs = disc, g, g
    p = p.finalize()
    po, grid = orc_params(p), orc_grid(forest)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    u += 0.05 * np.random.default_rng(5).standard_normal(u.shape)
    sol.upload(u)
    synced = u.copy()
    O.sync_ghosts_same_level(grid, po, synced, po.g, po.g)
    _check_vorticity_entries(sol, grid, po, synced)
    sol.close()


def test_cylinder_mask_sponge_and_statistics_2d():
    p = Params(dim=2, domain=(20.0, 20.0, 0.0), Bs=(26, 26, 1), wavelet="CDF44", g=6, g_rhs=2, n_eqn=3, Jmax=4, discretization="FD_4th_central",
               skew_symmetry=True, c0=20.0, nu=1.0e-2, gamma_p=1.0, CFL=1.0, u_mean_set=(1.0, 0.0, 0.0), time_max=1.0e9)
    p.penalization, p.C_eta, p.use_sponge, p.C_sponge = True, 1.0e-3, True, 1.0e-2
    p = p.finalize()
    forest = Forest.uniform(2, 3, Jmax=4)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.set_forest(forest)
    hvy, lvl, ixyz, _ = forest.active(0)
    cyl = CylinderMask2D(p, x_cntr=(9.0, 10.5), R_cyl=1.0, C_smooth=1.5, L_sponge=2.0, p_sponge=20.0)
    cyl.fill_device(sol, 0.0)
    got = np.zeros((len(hvy), 6, 1, 26 + 12, 26 + 12))
    sol.download(got, HVY_MASK, g_sync=0)
    ref = cyl.fill(lvl, ixyz)
    g = p.g
    assert np.abs(got[:, :, :, g:-g, g:-g] - ref[:, :, :, g:-g, g:-g]).max() <= 1e-13
    assert ref[:, 5].max() == 1.0 and ref[:, 0].max() == 1.0
    po = orc_params(p)
    grid = orc_grid(forest)
    u = O.alloc(grid, po)
    rng = np.random.default_rng(3)
    u[:] = rng.standard_normal(u.shape) * 0.1
    u[:, 0] += 1.0
    sol.upload(u)
    synced = u.copy()
    O.sync_ghosts_same_level(grid, po, synced, g, g)
    want = O.statistics_acm(grid, po, synced, ref)
    have = sol.statistics_ACM(0.0, with_divergence=True)
    for k in want:
        tol = 1e-8 if k.startswith("div") else 1e-12
        assert _close(have[k], want[k], tol), (k, have[k], want[k])
    assert want["sponge_volume"] > 10.0 and want["penal_power_sponge"] != 0.0 and want["force_x"] > 0.0
    _check_vorticity_entries(sol, grid, po, synced)
    sol.close()


def test_time_loop_writes_the_t_files(tmp_path):
    """AdaptiveLoop.statistics (main.f90:388-397): every nsave_stats iterations the device statistics go to the *.t files -- rows in the
    reference's format whose numbers are the ones statistics_ACM returns for the state after the time step"""
    from wabbit_b200.timeloop import AdaptiveLoop
    p = tg_params(Bs=16, J=3, wavelet_g=6)
    p.wavelet, p.penalization, p.C_eta, p.nsave_stats, p.eps = "CDF44", True, 1.0e-2, 1, 1.0e-3
    p = p.finalize()
    forest = Forest.uniform(3, 2, Jmax=3, max_blocks=600)
    sol = WabbitGPU(p, max_blocks=600)
    sol.setup_wavelet("CDF44")
    sol.set_forest(forest)
    po, grid = orc_params(p), orc_grid(forest)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    host = np.zeros(sol.host_shape())
    host[:grid.n] = u
    sol.upload(host)
    loop = AdaptiveLoop(sol, forest, 0.0, 0, refinement_indicator="everywhere", mask=SphereMask3D(p, center=(3.0, 3.1, 3.2), radius=0.9))
    loop.stats_dir = str(tmp_path)
    seen = []
    orig = loop.statistics
    loop.statistics = lambda dt: seen.append(orig(dt)) or seen[-1]
    for _ in range(2):
        loop.step()
    assert len(seen) == 2 and all(s is not None for s in seen)
    rows = {f.name: [[float(x) for x in line.split(";")] for line in f.read_text().splitlines()] for f in tmp_path.iterdir()}
    assert set(rows) == {"umag.t", "CFL.t", "meanflow.t", "div.t", "forces.t", "mask_volume.t", "penal_power.t", "u_residual.t", "e_kin.t",
                         "enstrophy.t", "helicity.t", "dissipation.t"}
    assert all(len(r) == 2 for r in rows.values())
    for k, s in enumerate(seen):
        assert abs(rows["e_kin.t"][k][1] - s["e_kin"]) <= 1e-8 * s["e_kin"] and abs(rows["enstrophy.t"][k][1] - s["enstrophy"]) <= 1e-8 * s["enstrophy"]
        assert abs(rows["mask_volume.t"][k][1] - s["mask_volume"]) <= 1e-8 * s["mask_volume"]
    vol = 4.0 / 3.0 * np.pi * 0.9 ** 3
    assert abs(seen[0]["mask_volume"] - vol) <= 0.15 * vol                     # smoothed sphere on the level-3 lattice
    assert abs(seen[0]["e_kin"] - (2.0 * np.pi) ** 3 / 8.0) <= 0.05 * (2.0 * np.pi) ** 3 / 8.0 and seen[0]["enstrophy"] > 0.0
    assert abs(rows["e_kin.t"][0][0] - loop.log[0][1]) <= 1e-8 * loop.log[0][1] and rows["e_kin.t"][1][0] > rows["e_kin.t"][0][0]
    sol.close()
