"""SURVEY 8d config 4's stand-in on the GPU: 3-D ACM with the penalization term of a translating sphere evaluated INSIDE the stage kernel
(wgpu_set_mask_sphere: no mask array is generated, uploaded or read), against the oracle, which builds the six mask components of
create_mask_3D_ACM / draw_sphere at every stage time on the host and runs RHS_3D_acm on them:
  * RungeKuttaGeneric on an equidistant grid with the sphere moving through it: dt identical, fields <= 1e-12;
  * analytic mask vs. the same mask uploaded to hvy_mask (sphere at rest): the two device paths agree to round-off;
  * the adaptive loop (adaptive initial condition, refine everywhere -> RK4 -> adapt_tree with CDF44, threshold_mask on the moving sphere,
    force_maxlevel_dealiasing) in lockstep with the oracle: block lists, refinement statuses and dt identical, fields <= 1e-12.
"""
import numpy as np
import pytest

import adaptive as A
import oracle as O
import sphere_case as SC
from wabbit_b200 import Forest, Params, WabbitGPU
from wabbit_b200.mask import SphereMask3D
from wabbit_b200.solver import HVY_BLOCK, HVY_MASK
from wabbit_b200.timeloop import AdaptiveLoop

from util import orc_grid, orc_params

pytestmark = pytest.mark.gpu


def _params(**kw):
    d = dict(SC.INI, wavelet="CDF44", skew_symmetry=True, eps=SC.EPS, eps_normalized=True, eps_norm="Linfty", Jmin=SC.JMIN,
             force_maxlevel_dealiasing=True, adapt_tree=True, refinement_indicator="everywhere")
    d.update(kw)
    return Params(**d).finalize()


def _field(og, po, seed):
    u = O.alloc(og, po)
    I = (slice(None), slice(None)) + O.interior(po)
    rng = np.random.default_rng(seed)
    u[I] = 0.05 * rng.standard_normal(u[I].shape)
    u[:, 0] += 1.0
    return u, I


@pytest.mark.parametrize("velocity", [(0.6, 0.2, -0.1), (0.0, 0.0, 0.0)])
def test_rk4_analytic_sphere(velocity):
    p = _params(Jmax=2)
    po = orc_params(p)
    forest = Forest.uniform(3, 2, Jmax=2)
    og = orc_grid(forest)
    sph = dict(SC.SPHERE, radius=0.2, velocity=velocity)
    om = A.SphereMask3D(po, **sph)
    sol = WabbitGPU(p, max_blocks=forest.n_blocks)
    sol.setup_wavelet("CDF44")
    sol.set_forest(forest)
    SphereMask3D(p, **sph).attach(sol)
    u, I = _field(og, po, 3)
    host = np.zeros(sol.host_shape())
    host[:og.n] = u
    sol.upload(host)
    work = [O.alloc(og, po) for _ in range(5)]
    mask_at = lambda t: np.stack([om.block(int(l), x, t) for l, x in zip(og.level, og.ixyz)])
    t = 0.3
    for it in range(2):
        dt_o = O.rk_generic(og, po, u, work, t, mask_at=mask_at)
        dt_g = sol.RungeKuttaGeneric(t, it)
        assert dt_g == dt_o
        t += dt_o
    got = np.zeros(sol.host_shape())
    sol.download(got, g_sync=0)
    err = np.abs(got[:og.n][I] - u[I]).max() / np.abs(u[I]).max()
    assert err <= 1e-12, err
    chi = mask_at(t)[:, 0]
    assert 0.0 < (chi > 0).mean() < 0.2 and chi.max() == 1.0          # the sphere is inside the grid and resolved
    if not any(velocity):
        # the same sphere through hvy_mask: switch the analytic mask off, upload the oracle's mask arrays, repeat the two steps
        sol.set_mask_sphere(None)
        u2, _ = _field(og, po, 3)
        host[:og.n] = u2
        sol.upload(host)
        mh = np.zeros(sol.host_shape(6))
        mh[:og.n] = mask_at(0.0)
        sol.upload(mh, HVY_MASK)
        t = 0.3
        for it in range(2):
            t += sol.RungeKuttaGeneric(t, it)
        got2 = np.zeros(sol.host_shape())
        sol.download(got2, g_sync=0)
        assert np.abs(got2[:og.n][I] - got[:og.n][I]).max() <= 1e-13
    sol.close()


def test_adaptive_sphere_lockstep():
    p = _params()
    po = O.Params(skew=True, **SC.INI)
    MAXB = 2400
    # oracle
    grid = O.uniform_grid(SC.JMIN, 3)
    run = A.AdaptiveRun(po, "CDF44", grid, O.alloc(grid, po), 0.0, 0, SC.EPS, Jmin=SC.JMIN, refinement_indicator="everywhere",
                        force_maxlevel_dealiasing=True, mask=A.SphereMask3D(po, **SC.SPHERE), threshold_mask=True, mask_time_dependent=True,
                        fd_half_width=2)

    def inicond(r):
        r.u[:] = 0.0
        r.u[:, 0] = 1.0
    inicond(run)
    run.adaptive_inicond(inicond)
    # device
    forest = Forest.uniform(3, SC.JMIN, Jmax=p.Jmax, max_blocks=MAXB)
    sol = WabbitGPU(p, max_blocks=MAXB)
    sol.setup_wavelet("CDF44")
    sol.set_forest(forest)
    loop = AdaptiveLoop(sol, forest, 0.0, 0, mask=SphereMask3D(p, **SC.SPHERE), threshold_mask=True)
    assert loop.mask_time_dependent

    def set_inicond(lp):
        hvy, _, _, _ = lp.forest.active(0)
        host = np.zeros((int(hvy.max()),) + sol.host_shape()[1:])
        host[hvy - 1, 0] = 1.0
        sol.upload(host, HVY_BLOCK, 0, hvy_ids=hvy)
    set_inicond(loop)
    loop.adaptive_inicond(set_inicond)

    def same():
        hvy, lvl, pos, _ = loop.forest.active(0)
        okey = {(int(l),) + tuple(int(v) for v in x): b for b, (l, x) in enumerate(zip(run.grid.level, run.grid.ixyz))}
        keys = [(int(l),) + tuple(int(v) for v in x) for l, x in zip(lvl, pos)]
        assert sorted(keys) == sorted(okey)
        o = np.array([okey[k] for k in keys])
        assert np.array_equal(np.asarray(loop.status), run.status[o])
        got = np.zeros((int(hvy.max()),) + sol.host_shape()[1:])
        sol.download(got, g_sync=0)
        I = (slice(None), slice(None)) + O.interior(po)
        return float(np.abs(got[hvy - 1][I] - run.u[o][I]).max())
    assert same() == 0.0 and len(np.unique(run.grid.level)) > 1
    for _ in range(2):
        dt_g, dt_o = loop.step(), run.step()
        assert loop.log[-1][2:4] == run.log[-1][2:4], (loop.log[-1], run.log[-1])
        assert abs(dt_g - dt_o) <= 1e-13 * dt_o
        assert same() <= 1e-12
    sol.close()


@pytest.mark.parametrize("world", [2, 3])
def test_adaptive_sphere_across_ranks_equals_single_rank(world):
    """the adaptive loop with the moving-sphere mask (threshold_mask) and the "significant" refinement indicator with the blocks
    partitioned over `world` ranks (threads driving one device context each, collectives through ThreadTransport): same grids, refinement
    statuses and dt, bit-identical data as the single-rank loop, which the oracle pins above"""
    import threading
    import torch
    from wabbit_b200.multi import DistributedWabbit, ThreadTransport
    from wabbit_b200.timeloop import DistributedAdaptiveLoop
    p = _params(refinement_indicator="significant")
    MAXB = 2400
    forest = Forest.uniform(3, SC.JMIN, Jmax=p.Jmax, max_blocks=MAXB)
    sol = WabbitGPU(p, max_blocks=MAXB)
    sol.setup_wavelet("CDF44")
    sol.set_forest(forest)
    loop = AdaptiveLoop(sol, forest, 0.0, 0, mask=SphereMask3D(p, **SC.SPHERE), threshold_mask=True)

    def set_inicond(lp):
        hvy, _, _, _ = lp.forest.active(0)
        host = np.zeros((int(hvy.max()),) + sol.host_shape()[1:])
        host[hvy - 1, 0] = 1.0
        sol.upload(host, HVY_BLOCK, 0, hvy_ids=hvy)
    set_inicond(loop)
    loop.adaptive_inicond(set_inicond)
    loop.step()                                      # a flow with structure around the sphere
    hvy, l0, x0, _ = loop.forest.active(0)
    u0 = np.zeros((len(hvy),) + sol.host_shape()[1:])
    sol.download(u0, g_sync=0)
    st0, t0, it0 = loop.status.copy(), loop.time, loop.iteration
    nsteps = 2
    for _ in range(nsteps):
        loop.step()
    _, lf, xf, _ = loop.forest.active(0)
    ref = np.zeros((len(lf),) + sol.host_shape()[1:])
    sol.download(ref, g_sync=0)
    ref_log, ref_status = list(loop.log[-nsteps:]), loop.status.copy()
    assert len(np.unique(lf)) > 1 and (ref_status == 0).any() and (ref_status == 9).any()
    sol.close()

    fw = Forest.from_blocks(3, p.Jmax, l0, x0, n_ranks=world, max_blocks=MAXB)
    offs = np.concatenate([[0], np.cumsum([fw.n_active(r) for r in range(world)])])
    sols = []
    for _ in range(world):
        s = WabbitGPU(p, max_blocks=MAXB)
        s.setup_wavelet("CDF44")
        sols.append(s)
    shared = ThreadTransport.Shared(world)
    res, errs = [None] * world, []

    def worker(r):
        try:
            torch.cuda.set_device(0)
            d = DistributedWabbit(sols[r], fw, r, world, transport=ThreadTransport(shared, r), overlap=False)
            n = fw.n_active(r)
            h = np.zeros((max(n, 1),) + sols[r].host_shape()[1:])
            h[:n] = u0[offs[r]:offs[r] + n]
            sols[r].upload(h, hvy_ids=np.arange(1, n + 1, dtype=np.int32))
            lp = DistributedAdaptiveLoop(d, t0, it0, mask=SphereMask3D(p, **SC.SPHERE), threshold_mask=True)
            lp.status = st0.copy()
            for _ in range(nsteps):
                lp.step()
            _, l, x, _ = d.forest.active(r)
            out = np.zeros((max(len(l), 1),) + sols[r].host_shape()[1:])
            sols[r].download(out, g_sync=0, hvy_ids=np.arange(1, len(l) + 1, dtype=np.int32))
            res[r] = (list(lp.log), lp.status.copy(), l, x, out[:len(l)].copy())
        except BaseException as e:      # noqa: BLE001
            errs.append((r, repr(e)))
            shared.barrier.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs
    for r in range(world):
        assert res[r][0] == ref_log, (r, res[r][0], ref_log)
        assert np.array_equal(res[r][1], ref_status)
    assert np.array_equal(np.concatenate([res[r][2] for r in range(world)]), lf)
    assert np.array_equal(np.concatenate([res[r][3] for r in range(world)]), xf)
    got = np.concatenate([res[r][4] for r in range(world)])
    g = p.g
    assert np.array_equal(got[:, :, g:-g, g:-g, g:-g], ref[:, :, g:-g, g:-g, g:-g])
    for s in sols:
        s.close()
