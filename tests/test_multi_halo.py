"""Halo-block mode of the multi-GPU path (wabbit_b200/multi.py: HaloPlan, HaloStepper; include/wabbit_gpu.h: wgpu_set_halo).
CPU: the plans derived independently on every rank agree (who mirrors what, in which order), also through a real
all-to-all (gloo, world_size 2).  GPU (ranks as contexts on one device, in lockstep): RK4 on a graded grid partitioned over
2 / 3 ranks and the wavelet decomposition across partition boundaries reproduce the oracle exactly like the single-rank path."""
import os
import sys

import numpy as np
import pytest

import oracle as O
from util import graded_blocks, orc_params, relerr, tg_params
from wabbit_b200 import Forest
from wabbit_b200.multi import HaloPlan, halo_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def graded_forest(world, seed=11, J0=1, Jmax=3, frac=0.3, slack=3.0):
    lv, ix = graded_blocks(3, J0, Jmax, seed, frac)
    mb = int(slack * -(-len(lv) // world)) + 8
    return Forest.from_blocks(3, Jmax, lv, ix, n_ranks=world, max_blocks=mb)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_halo_plans_are_consistent(world):
    forest = graded_forest(world, slack=6.0 if world == 8 else 3.0)
    N = forest.max_blocks
    plans = [HaloPlan(forest, r, world) for r in range(world)]
    act = [forest.active(r) for r in range(world)]
    for r, p in enumerate(plans):
        assert p.n_halo > 0 and p.n_own + p.n_halo <= N
        assert sum(p.recv_counts) == p.n_halo and sum(p.send_counts) == p.n_send and p.send_counts[r] == 0
        for q in range(world):
            assert p.send_counts[q] == plans[q].recv_counts[r]
        # the mirrored blocks are exactly the remote ids of the neighbour table; levels and treecodes are the owner's
        nb = forest.neighbors(r)
        remote = {int(v) for v in nb[nb >= 1] if (int(v) - 1) // N != r}
        assert remote == {int(v) for v in p.halo_lgt}
        for k, lgt in enumerate(p.halo_lgt):
            o, h = (int(lgt) - 1) // N, (int(lgt) - 1) % N
            assert act[o][1][h] == p.halo_level[k] and act[o][3][h] == p.halo_tc[k]
        assert (np.diff(p.halo_hvy) == 1).all() and p.halo_hvy[0] == p.n_own + 1
    # what r sends to q, in order, is what q expects from r, in order
    for r in range(world):
        so = np.concatenate([[0], np.cumsum(plans[r].send_counts)])
        for q in range(world):
            mine = plans[r].send_hvy[so[q]:so[q + 1]]
            ro = int(np.sum(plans[q].recv_counts[:r]))
            theirs = plans[q].halo_lgt[ro:ro + plans[q].recv_counts[r]]
            assert np.array_equal(r * N + mine, theirs)


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        forest = graded_forest(world)
        N = forest.max_blocks
        plan = HaloPlan(forest, rank, world)
        blk = 6                                                    # doubles per "block": stamped with the owner's lgt id
        own = np.repeat((rank * N + np.arange(1, plan.n_own + 1, dtype=np.float64))[:, None], blk, 1) + 0.125 * np.arange(blk)
        send = own[plan.send_hvy - 1].ravel()
        recv = torch.zeros(plan.n_halo * blk, dtype=torch.float64)
        dist.all_to_all_single(recv, torch.from_numpy(send.copy()), [c * blk for c in plan.recv_counts], [c * blk for c in plan.send_counts])
        got = recv.numpy().reshape(plan.n_halo, blk)
        exp = plan.halo_lgt[:, None].astype(np.float64) + 0.125 * np.arange(blk)
        ret[rank] = (bool(np.array_equal(got, exp)), plan.n_halo, plan.n_send)
    finally:
        dist.destroy_process_group()


def test_halo_exchange_gloo_world2():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for r in range(2):
        ok, nh, ns = ret[r]
        assert ok and nh > 0 and ns > 0


# ------------------------------------------------------------------------------------------------------------------ GPU
def _global_oracle(forest, world, p):
    po = orc_params(p)
    lv, ix = [], []
    for r in range(world):
        _, lvl, ixyz, _ = forest.active(r)
        lv.append(lvl)
        ix.append(ixyz)
    grid = O.Grid(level=np.concatenate(lv).astype(np.int64), ixyz=np.concatenate(ix).astype(np.int64), dim=3)
    return po, grid, O.neighbor_table168(grid, forest.Jmax)


def _make_ranks(forest, world, p, wavelet):
    from wabbit_b200 import WabbitGPU
    from wabbit_b200.multi import HaloLockstepGroup
    sols = []
    for _ in range(world):
        s = WabbitGPU(p, max_blocks=forest.max_blocks)
        s.setup_wavelet(wavelet)
        sols.append(s)
    return sols, HaloLockstepGroup(sols, forest)


def _scatter(sols, forest, u):
    off = 0
    for r, s in enumerate(sols):
        n = forest.n_active(r)
        host = np.zeros(s.host_shape())
        host[:n] = u[off:off + n]
        s.upload(host, hvy_ids=np.arange(1, n + 1, dtype=np.int32))
        off += n


@pytest.mark.gpu
@pytest.mark.parametrize("world,split,wavelet", [(2, False, "CDF44"), (3, True, "CDF40"), (2, True, "CDF44")])
def test_rk4_on_a_graded_grid_across_ranks(world, split, wavelet):
    w = O.setup_wavelet(wavelet)
    forest = graded_forest(world)
    p = tg_params(Bs=16, J=forest.Jmax, wavelet_g=w.g_default)
    p.wavelet = wavelet
    sols, grp = _make_ranks(forest, world, p, wavelet)
    assert all(s.n_bnd > 0 for s in grp.st)
    po, grid, nbr = _global_oracle(forest, world, p)
    rng = np.random.default_rng(3)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    u += 0.1 * rng.standard_normal(u.shape)
    _scatter(sols, forest, u)
    work = [O.alloc(grid, po) for _ in range(5)]
    sync = lambda h: O.sync_ghosts_leaf(grid, po, h, nbr, p.g_rhs, p.g_rhs, w.X, bool(w.lifted))
    t = 0.0
    for it in range(2):
        dt = grp.step(t, split=split)
        assert dt == O.rk_generic(grid, po, u, work, t, sync=sync)
        t += dt
    off, g = 0, p.g
    for r, s in enumerate(sols):
        n = forest.n_active(r)
        out = np.zeros(s.host_shape())
        s.download(out, g_sync=0, hvy_ids=np.arange(1, n + 1, dtype=np.int32))
        assert relerr(out[:n, :, g:-g, g:-g, g:-g], u[off:off + n, :, g:-g, g:-g, g:-g]) <= 1e-12
        off += n
    for s in sols:
        s.close()


@pytest.mark.gpu
def test_diffusion_limited_dt_with_the_finest_level_on_one_rank_only():
    """GET_DT_BLOCK applies every limit per block before the MPI_MIN (module_ACM.f90:617-691): with a large viscosity the limit CFL_nu dx^2 / nu
    of the finest blocks binds, and those blocks live on ONE rank here -- every rank must still take the same, oracle-identical dt."""
    from wabbit_b200 import Forest
    world, wavelet = 2, "CDF40"
    w = O.setup_wavelet(wavelet)
    # level 1 everywhere, one corner block refined to level 2 and one of its children to level 3: the finest blocks sit at the start of the curve
    leaves = {(1, x, y, z) for x in range(2) for y in range(2) for z in range(2)}

    def refine(k):
        leaves.remove(k)
        L, x, y, z = k
        for c in range(8):
            leaves.add((L + 1, 2 * x + (c & 1), 2 * y + ((c >> 1) & 1), 2 * z + ((c >> 2) & 1)))
    refine((1, 0, 0, 0))
    refine((2, 0, 0, 0))
    ks = sorted(leaves)
    forest = Forest.from_blocks(3, 3, np.array([k[0] for k in ks], np.int32), np.array([k[1:] for k in ks], np.int32), n_ranks=world,
                                max_blocks=4 * len(ks) + 64)
    finest = [int((forest.active(r)[1] == 3).sum()) for r in range(world)]
    assert min(finest) == 0 < max(finest), finest
    p = tg_params(Bs=16, J=3, wavelet_g=w.g_default)
    p.wavelet, p.nu = wavelet, 0.5
    p = p.finalize()
    sols, grp = _make_ranks(forest, world, p, wavelet)
    po, grid, nbr = _global_oracle(forest, world, p)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    _scatter(sols, forest, u)
    work = [O.alloc(grid, po) for _ in range(5)]
    sync = lambda h: O.sync_ghosts_leaf(grid, po, h, nbr, p.g_rhs, p.g_rhs, w.X, bool(w.lifted))
    t = 0.0
    for it in range(2):
        dt = grp.step(t)
        dt_ref = O.rk_generic(grid, po, u, work, t, sync=sync)
        dx_min = p.domain[0] / (2 ** 3 * 16)
        assert dt == dt_ref and abs(dt - p.CFL_nu * dx_min ** 2 / p.nu) <= 1e-15, (dt, dt_ref)     # the diffusion limit of level 3 binds
        t += dt
    for s in sols:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,wavelet,ignore_filter", [(2, "CDF40", True), (3, "CDF44", True), (2, "CDF44", False), (3, "CDF42", False)])
def test_wavelet_side_across_ranks(world, wavelet, ignore_filter):
    """halo copies refreshed for hvy_block, then on every rank: download with a synchronised ghost shell (all 26 relations, level
    jumps included) and the wavelet decomposition + flags == the oracle on the global grid, bit for bit.  ignore_filter=False with a
    lifted wavelet: the HD-filtered copies of finer neighbours owned by other ranks are exchanged as well (wgpu_set_halo_restrict)."""
    from wabbit_b200.solver import HVY_TMP
    w = O.setup_wavelet(wavelet)
    forest = graded_forest(world, seed=21)
    p = tg_params(Bs=16, J=forest.Jmax, wavelet_g=w.g_default)
    p.wavelet = wavelet
    sols, grp = _make_ranks(forest, world, p, wavelet)
    po, grid, nbr = _global_oracle(forest, world, p)
    rng = np.random.default_rng(8)
    u = O.alloc(grid, po)
    u[:] = rng.standard_normal(u.shape)
    _scatter(sols, forest, u)
    for s in sols:
        s.set_ghost_filter(ignore_filter)
    grp.exchange_array(0, 0)
    ref = u.copy()
    O.sync_ghosts_leaf(grid, po, ref, nbr, po.g, po.g, w.X, bool(w.lifted), ignore_filter=ignore_filter, w=w)
    wd_ref = np.zeros_like(ref)
    O.fwt_tree(w, po, ref, wd_ref)
    st_ref, det_ref = O.threshold_tree(po, wd_ref, grid.level, 0.5)
    I = (slice(None), slice(None)) + O.interior(po)
    off = 0
    for r, s in enumerate(sols):
        n = forest.n_active(r)
        ids = np.arange(1, n + 1, dtype=np.int32)
        got = np.zeros(s.host_shape())
        got[:n] = u[off:off + n]
        s.download(got, g_sync=p.g, hvy_ids=ids)
        assert np.array_equal(got[:n], ref[off:off + n]), r
        s.waveletDecomposition_tree((0, 0), (HVY_TMP, 0))
        wd = np.zeros(s.host_shape())
        s.download(wd, HVY_TMP, g_sync=0, hvy_ids=ids)
        assert np.array_equal(wd[:n][I], wd_ref[off:off + n][I]), r
        st, det = s.threshold_tree((HVY_TMP, 0), eps=0.5, want_detail=True)
        assert np.array_equal(st, st_ref[off:off + n]) and np.array_equal(det, det_ref[off:off + n])
        off += n
    for s in sols:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,overlap", [(2, True), (3, False)])
def test_adaptive_cycle_across_ranks_equals_single_rank(world, overlap):
    """refine_tree("everywhere") -> RK4 step -> adapt_tree (CDF40) with the blocks partitioned over `world` ranks (threads driving one
    device context each, collectives through ThreadTransport) gives the same grids and bit-identical data as the single-rank driver,
    which the oracle pins (tests/test_gpu_cycle.py)."""
    import threading
    import torch
    from wabbit_b200 import WabbitGPU
    from wabbit_b200.multi import DistributedWabbit, ThreadTransport
    wavelet, Jmax = "CDF40", 4
    w = O.setup_wavelet(wavelet)
    lv, ix = graded_blocks(3, 1, 3, seed=5, frac=0.25)
    MB = 8 * len(lv) + 64
    p = tg_params(Bs=16, J=Jmax, wavelet_g=w.g_default)
    p.wavelet = wavelet
    p.eps = 1.0e-2
    f1 = Forest.from_blocks(3, Jmax, lv, ix, n_ranks=1, max_blocks=MB)
    fw = Forest.from_blocks(3, Jmax, lv, ix, n_ranks=world, max_blocks=MB)
    po = orc_params(p)
    _, l1, x1, _ = f1.active(0)
    grid = O.Grid(level=l1.astype(np.int64), ixyz=x1.astype(np.int64), dim=3)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    # strong small-scale content in the low-x half of the domain only: the sweep coarsens part of the grid
    amp = np.where(x1[:, 0] * 2 < 2 ** l1, 0.05, 1.0e-6)
    u += amp[:, None, None, None, None] * np.random.default_rng(2).standard_normal(u.shape)

    def run_sequence(drv_refine, drv_step, drv_adapt):
        out = [drv_refine()]
        t, it, dt = drv_step(0.0, 0)
        out.append(dt)
        out.append(drv_adapt())
        return out

    # ---- single rank
    s1 = WabbitGPU(p, max_blocks=MB)
    s1.setup_wavelet(wavelet)
    s1.set_forest(f1)
    host = np.zeros(s1.host_shape())
    host[:grid.n] = u
    s1.upload(host)
    state = {"f": f1}

    def r1():
        state["f"] = s1.refine_tree(state["f"])
        return state["f"].n_blocks

    def a1():
        state["f"], n0, n1 = s1.adapt_tree(state["f"], eps=p.eps, Jmin=1)
        return (n0, n1)

    ref_seq = run_sequence(r1, s1.timeStep_tree, a1)
    assert 8 < ref_seq[2][1] < ref_seq[2][0]                   # the sweep coarsens part of the grid
    _, lf, xf, _ = state["f"].active(0)
    ref_data = np.zeros(s1.host_shape())
    s1.download(ref_data, g_sync=0)
    ref_data = ref_data[:len(lf)].copy()
    s1.close()

    # ---- `world` ranks
    sols = []
    for _ in range(world):
        s = WabbitGPU(p, max_blocks=MB)
        s.setup_wavelet(wavelet)
        sols.append(s)
    shared = ThreadTransport.Shared(world)
    offs = np.concatenate([[0], np.cumsum([fw.n_active(r) for r in range(world)])])
    res, errs = [None] * world, []

    def worker(r):
        try:
            torch.cuda.set_device(0)
            d = DistributedWabbit(sols[r], fw, r, world, transport=ThreadTransport(shared, r), overlap=overlap)
            n = fw.n_active(r)
            h = np.zeros(sols[r].host_shape())
            h[:n] = u[offs[r]:offs[r] + n]
            sols[r].upload(h)
            seq = run_sequence(lambda: d.refine_tree().n_blocks, d.timeStep_tree, lambda: d.adapt_tree(eps=p.eps, Jmin=1)[1:])
            _, l, x, _ = d.forest.active(r)
            out = np.zeros(sols[r].host_shape())
            sols[r].download(out, g_sync=0)
            res[r] = (seq, l, x, out[:len(l)].copy())
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
        assert res[r][0][0] == ref_seq[0] and res[r][0][1] == ref_seq[1] and tuple(res[r][0][2]) == tuple(ref_seq[2]), (r, res[r][0], ref_seq)
    assert np.array_equal(np.concatenate([res[r][1] for r in range(world)]), lf)
    assert np.array_equal(np.concatenate([res[r][2] for r in range(world)]), xf)
    got = np.concatenate([res[r][3] for r in range(world)])
    g = p.g
    assert np.array_equal(got[:, :, g:-g, g:-g, g:-g], ref_data[:, :, g:-g, g:-g, g:-g])
    for s in sols:
        s.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,Bs,sz", [(2, 16, True), (3, 18, True), (2, 18, False)])
def test_lifted_adapt_tree_across_ranks_equals_single_rank(world, Bs, sz):
    """adapt_tree with the full wavelet transformation (CDF44: coarse extension, security zone, CE-optimised reconstruction; Bs=16 level-wise,
    Bs=18 leaf-first variant) with the tree's blocks owned by `world` ranks: same grid and bit-identical data as the single-rank driver,
    which the oracle pins (tests/test_gpu_fulltree.py)."""
    import threading
    import torch
    from wabbit_b200 import WabbitGPU
    from wabbit_b200.multi import DistributedWabbit, ThreadTransport
    wavelet, Jmax = "CDF44", 3
    w = O.setup_wavelet(wavelet)
    lv, ix = graded_blocks(3, 1, 3, seed=5, frac=0.3)
    MB = 3 * len(lv) + 64
    p = tg_params(Bs=Bs, J=Jmax, wavelet_g=w.g_default)
    p.wavelet = wavelet
    f1 = Forest.from_blocks(3, Jmax, lv, ix, n_ranks=1, max_blocks=MB)
    fw = Forest.from_blocks(3, Jmax, lv, ix, n_ranks=world, max_blocks=MB)
    po = orc_params(p)
    _, l1, x1, _ = f1.active(0)
    grid = O.Grid(level=l1.astype(np.int64), ixyz=x1.astype(np.int64), dim=3)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    amp = np.where(x1[:, 0] * 2 < 2 ** l1, 0.05, 1.0e-7)
    u += amp[:, None, None, None, None] * np.random.default_rng(2).standard_normal(u.shape)
    eps = 0.01
    s1 = WabbitGPU(p, max_blocks=MB)
    s1.setup_wavelet(wavelet)
    s1.set_forest(f1)
    host = np.zeros(s1.host_shape())
    host[:grid.n] = u
    s1.upload(host)
    new1, n0, n1 = s1.adapt_tree(f1, eps=eps, Jmin=1, useSecurityZone=sz)
    assert 8 <= n1 < n0
    _, lf, xf, _ = new1.active(0)
    ref = np.zeros(s1.host_shape())
    s1.download(ref, g_sync=0)
    ref = ref[:n1].copy()
    s1.close()

    sols = []
    for _ in range(world):
        s = WabbitGPU(p, max_blocks=MB)
        s.setup_wavelet(wavelet)
        sols.append(s)
    shared = ThreadTransport.Shared(world)
    offs = np.concatenate([[0], np.cumsum([fw.n_active(r) for r in range(world)])])
    res, errs = [None] * world, []

    def worker(r):
        try:
            torch.cuda.set_device(0)
            d = DistributedWabbit(sols[r], fw, r, world, transport=ThreadTransport(shared, r), overlap=False)
            n = fw.n_active(r)
            h = np.zeros(sols[r].host_shape())
            h[:n] = u[offs[r]:offs[r] + n]
            sols[r].upload(h)
            new, m0, m1 = d.adapt_tree(eps=eps, Jmin=1, useSecurityZone=sz)
            _, l, x, _ = d.forest.active(r)
            out = np.zeros(sols[r].host_shape())
            sols[r].download(out, g_sync=0)
            res[r] = (m0, m1, l, x, out[:len(l)].copy())
        except BaseException as e:      # noqa: BLE001
            import traceback
            errs.append((r, traceback.format_exc()))
            shared.barrier.abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs[0][1]
    assert all(res[r][0] == n0 and res[r][1] == n1 for r in range(world))
    assert np.array_equal(np.concatenate([res[r][2] for r in range(world)]), lf)
    assert np.array_equal(np.concatenate([res[r][3] for r in range(world)]), xf)
    got = np.concatenate([res[r][4] for r in range(world)])
    g = p.g
    assert np.array_equal(got[:, :, g:-g, g:-g, g:-g], ref[:, :, g:-g, g:-g, g:-g])
    for s in sols:
        s.close()
