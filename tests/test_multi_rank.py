"""Multi-rank path.
CPU (gloo, world_size 2): the exchange plan derived independently on each rank moves exactly the ghost values the
oracle's synchronisation produces (pack emulated in NumPy -- the CUDA pack kernel is covered by the GPU test below).
GPU (one device, several contexts in lockstep): pack kernel + patch pool + interior/boundary split against the oracle.
"""
import os
import sys

import numpy as np
import pytest

import oracle as O
from util import orc_grid, orc_params, relerr, tg_params
from wabbit_b200 import Forest
from wabbit_b200.multi import FACES, ExchangePlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def np_pack(u_compact, hvy, d, H):
    """strip of the sender block facing direction d (sender -> receiver), laid out as the receiver's ghost strip"""
    dx, dy, dz = d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1
    B = u_compact.shape[-1]
    sl = lambda s: slice(B - H, B) if s > 0 else (slice(0, H) if s < 0 else slice(0, B))
    return np.ascontiguousarray(u_compact[hvy - 1][:, sl(dz), sl(dy), sl(dx)]).ravel()


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        Bs, H, J = 8, 2, 2
        forest = Forest.uniform(3, J, n_ranks=world)
        plan = ExchangePlan(forest, rank, world)
        # global field: every rank builds the same one, keeps only its blocks
        rng = np.random.default_rng(5)
        nb = forest.n_blocks
        full = {}
        for r in range(world):
            hvy, lvl, ixyz, _ = forest.active(r)
            for k in range(len(hvy)):
                full[(r, int(hvy[k]))] = tuple(ixyz[k])
        field = rng.random((2 ** J, 2 ** J, 2 ** J, 4, Bs, Bs, Bs))     # [bz,by,bx,c,z,y,x]
        hvy, lvl, ixyz, _ = forest.active(rank)
        mine = np.stack([field[i[2], i[1], i[0]] for i in ixyz])          # compact [blk,c,z,y,x]
        pd = 4 * H * Bs * Bs
        send = np.concatenate([np_pack(mine, int(h), int(d), H) for h, d in zip(plan.send_hvy, plan.send_dir)]) if plan.n_send else np.zeros(0)
        recv = torch.zeros(plan.n_recv * pd, dtype=torch.float64)
        dist.all_to_all_single(recv, torch.from_numpy(send), [c * pd for c in plan.recv_counts], [c * pd for c in plan.send_counts])
        recv = recv.numpy().reshape(plan.n_recv, -1)
        # expected: the neighbour block's strip, straight from the global field
        n = 2 ** J
        ok = True
        for k in range(plan.n_recv):
            h, d = int(plan.recv_hvy[k]), int(plan.recv_dir[k])
            dx, dy, dz = d % 3 - 1, (d // 3) % 3 - 1, d // 9 - 1
            me = ixyz[h - 1]
            src = field[(me[2] + dz) % n, (me[1] + dy) % n, (me[0] + dx) % n]
            sl = lambda s: slice(Bs - H, Bs) if s < 0 else (slice(0, H) if s > 0 else slice(0, Bs))
            exp = src[:, sl(dz), sl(dy), sl(dx)].ravel()
            ok = ok and np.array_equal(recv[k], exp)
        # dt all-reduce (MIN) as in calculate_time_step.f90:48
        t = torch.tensor([0.5 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ret[rank] = (ok, plan.n_recv, plan.n_send, float(t.item()))
    finally:
        dist.destroy_process_group()


def test_exchange_plan_gloo_world2():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert len(ret) == 2
    for r in range(2):
        ok, nr, ns, dtmin = ret[r]
        assert ok and nr > 0 and nr == ns and dtmin == 0.5


def test_plan_symmetry_many_ranks():
    forest = Forest.uniform(3, 3, n_ranks=8)
    plans = [ExchangePlan(forest, r, 8) for r in range(8)]
    for r in range(8):
        for q in range(8):
            assert plans[r].send_counts[q] == plans[q].recv_counts[r]
        assert plans[r].send_counts[r] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("world,disc", [(2, "FD_4th_central"), (4, "FD_4th_central"), (3, "FD_6th_central"), (2, "FD_2nd_central")])
def test_multi_rank_lockstep_on_one_gpu(world, disc):
    """`world` ranks as contexts on one device: RK4 steps with the pack kernel / patch pool must reproduce the
    oracle exactly like the single-rank path does (same dt, fields within 1e-12)."""
    from wabbit_b200 import WabbitGPU
    from wabbit_b200.multi import LockstepGroup
    J = 2
    p = tg_params(Bs=16, J=J, wavelet_g=3, discretization=disc, skew=True)
    forest = Forest.uniform(3, J, n_ranks=world)
    sols = [WabbitGPU(p, max_blocks=forest.max_blocks) for _ in range(world)]
    grp = LockstepGroup(sols, forest)
    assert sum(s.n_bnd for s in grp.st) > 0
    # oracle on the global grid (rank-major block order)
    po = orc_params(p)
    lv, ix = [], []
    for r in range(world):
        hvy, lvl, ixyz, _ = forest.active(r)
        lv.append(lvl); ix.append(ixyz)
    grid = O.Grid(level=np.concatenate(lv).astype(np.int64), ixyz=np.concatenate(ix).astype(np.int64), dim=3)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    off = 0
    for r, s in enumerate(sols):
        n = forest.n_active(r)
        host = np.zeros(s.host_shape())
        host[:n] = u[off:off + n]
        s.upload(host)
        off += n
    work = [O.alloc(grid, po) for _ in range(5)]
    t = 0.0
    for it in range(3):
        dt = grp.step(t)
        dt_ref = O.rk_generic(grid, po, u, work, t)
        assert dt == dt_ref
        t += dt
    off = 0
    g = p.g
    for r, s in enumerate(sols):
        n = forest.n_active(r)
        out = np.zeros(s.host_shape())
        s.download(out, g_sync=0)
        a = out[:n, :, g:-g, g:-g, g:-g]
        b = u[off:off + n, :, g:-g, g:-g, g:-g]
        assert relerr(a, b) <= 1e-12
        off += n
    for s in sols:
        s.close()
