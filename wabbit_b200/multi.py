"""Multi-GPU time stepping: one process per GPU, blocks partitioned by the space-filling curve (the reference's
balanceLoad_tree decomposition), ghost patches of partition-boundary blocks exchanged once per Runge-Kutta stage.

The reference packs patches per destination rank, posts MPI_Isend/Irecv and unpacks (LIB/MPI/xfer_block_data.f90:10-99),
overlapping the transfer with the copies between blocks of the same rank.  Here:
  pack kernel -> per-peer contiguous send regions -> ONE all-to-all over NCCL (torch.distributed) -> the receive
  buffer IS the patch pool the stage kernel gathers from (no unpack pass);
  blocks whose neighbours are all local ("interior") are advanced while the transfer is in flight, the
  partition-boundary blocks right after it.
The time step needs one more exchange: MPI_Allreduce(MIN) of dt (LIB/TIME/calculate_time_step.f90:48) -> all_reduce
of the device scalar.

Because the topology (light data) is replicated on every rank, both sides derive the same patch order from it and no
size/metadata handshake is needed (the reference encodes metadata into the message, xfer_block_data.f90:52-66).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .forest import Forest

# direction index (dz+1)*9 + (dy+1)*3 + (dx+1) of the six faces and their same-level slot in hvy_neighbor (0-based)
FACES: List[Tuple[int, Tuple[int, int, int], int]] = [
    (4, (0, 0, -1), 16), (10, (0, -1, 0), 8), (12, (-1, 0, 0), 0), (14, (1, 0, 0), 4), (16, (0, 1, 0), 12), (22, (0, 0, 1), 20)]


def opposite(d: int) -> int:
    return 26 - d


def recv_list(forest: Forest, rank: int) -> np.ndarray:
    """Face patches `rank` receives: rows (peer, hvy_receiver, dir_receiver, hvy_sender), sorted by
    (peer, hvy_receiver, dir_receiver) -- the order patches sit in the receive buffer."""
    N = forest.max_blocks
    hvy, _, _, _ = forest.active(rank)
    nb = forest.neighbors(rank)
    rows = []
    for d, _vec, slot in FACES:
        lgt = nb[slot, hvy - 1]
        ok = lgt >= 1
        r = np.where(ok, (lgt - 1) // N, -1)
        h = np.where(ok, (lgt - 1) % N + 1, -1)
        sel = ok & (r != rank)
        if sel.any():
            rows.append(np.stack([r[sel], hvy[sel], np.full(sel.sum(), d), h[sel]], axis=1))
    if not rows:
        return np.zeros((0, 4), np.int64)
    a = np.concatenate(rows).astype(np.int64)
    order = np.lexsort((a[:, 2], a[:, 1], a[:, 0]))
    return a[order]


class ExchangePlan:
    """Who sends which face patch to whom, from the replicated topology."""

    def __init__(self, forest: Forest, rank: int, world: int):
        self.rank, self.world = rank, world
        mine = recv_list(forest, rank)
        self.recv_hvy = mine[:, 1].astype(np.int32)
        self.recv_dir = mine[:, 2].astype(np.int32)
        self.recv_counts = [int((mine[:, 0] == p).sum()) for p in range(world)]
        send_h, send_d, self.send_counts = [], [], []
        for p in range(world):
            if p == rank:
                self.send_counts.append(0)
                continue
            theirs = recv_list(forest, p)
            t = theirs[theirs[:, 0] == rank]            # patches peer p expects from me, in p's receive order
            send_h.append(t[:, 3])
            send_d.append(26 - t[:, 2])                  # direction from the sender (me) towards the receiver
            self.send_counts.append(len(t))
        self.send_hvy = (np.concatenate(send_h) if send_h else np.zeros(0)).astype(np.int32)
        self.send_dir = (np.concatenate(send_d) if send_d else np.zeros(0)).astype(np.int32)

    @property
    def n_recv(self) -> int:
        return len(self.recv_hvy)

    @property
    def n_send(self) -> int:
        return len(self.send_hvy)


class _DevPtr:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class MultiGPUStepper:
    """RungeKuttaGeneric across ranks.  `exchange(send, recv, send_counts, recv_counts) -> handle-with-wait()` moves
    the patches; the default uses torch.distributed all_to_all_single (NCCL).  `allreduce_min(tensor)` reduces dt."""

    def __init__(self, sol, forest: Forest, rank: int, world: int, exchange: Optional[Callable] = None,
                 allreduce_min: Optional[Callable] = None, overlap: bool = True):
        import torch
        self.torch = torch
        self.sol, self.rank, self.world, self.overlap = sol, rank, world, overlap
        self.plan = ExchangePlan(forest, rank, world)
        lib, ctx = sol._lib, sol._ctx
        hvy, lvl, _, _ = forest.active(rank)
        sol.set_topology(hvy, lvl, forest.neighbors(rank), rank)
        self.pd = int(lib.wgpu_patch_doubles(ctx))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.pool = torch.zeros(max(self.plan.n_recv, 1) * self.pd, dtype=torch.float64, device=dev)
        self.send = torch.zeros(max(self.plan.n_send, 1) * self.pd, dtype=torch.float64, device=dev)
        sol._check(lib.wgpu_set_exchange(ctx, self.plan.n_recv, _i32(self.plan.recv_hvy), _i32(self.plan.recv_dir),
                                         C.c_void_p(self.pool.data_ptr()), self.plan.n_send, _i32(self.plan.send_hvy),
                                         _i32(self.plan.send_dir), C.c_void_p(self.send.data_ptr())))
        self.in_splits = [c * self.pd for c in self.plan.send_counts]
        self.out_splits = [c * self.pd for c in self.plan.recv_counts]
        self._exchange = exchange or self._nccl_exchange
        self._allreduce_min = allreduce_min or self._nccl_min
        self.n_int = lib.wgpu_block_count(ctx, 1)
        self.n_bnd = lib.wgpu_block_count(ctx, 2)

    # -- default transports (NCCL through torch.distributed)
    def _nccl_exchange(self, send, recv, in_splits, out_splits):
        import torch.distributed as dist
        return dist.all_to_all_single(recv[:sum(out_splits)], send[:sum(in_splits)], out_splits, in_splits, async_op=True)

    def _nccl_min(self, t):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MIN)

    def _dtmin_tensor(self):
        p = C.c_void_p()
        self.sol._check(self.sol._lib.wgpu_dtmin_pointer(self.sol._ctx, C.byref(p)))
        return self.torch.as_tensor(_DevPtr(p.value, 1), device=self.pool.device)

    def step(self, time: float, iteration: int = 0) -> float:
        lib, ctx, chk = self.sol._lib, self.sol._ctx, self.sol._check
        chk(lib.wgpu_rk_begin(ctx, float(time)))
        if self.world > 1 and not self.sol.params.dt_fixed > 0.0:
            self._allreduce_min(self._dtmin_tensor())       # positive doubles: MIN of values == MIN of bit patterns
        chk(lib.wgpu_rk_dt(ctx, float(time)))
        for j in range(1, self.sol.params.n_stages + 1):
            chk(lib.wgpu_pack_halo(ctx, j))
            work = self._exchange(self.send, self.pool, self.in_splits, self.out_splits)
            if self.overlap and self.n_bnd:
                chk(lib.wgpu_rk_stage(ctx, j, 1))            # interior blocks while the patches are in flight
                if work is not None:
                    work.wait()
                chk(lib.wgpu_rk_stage(ctx, j, 2))            # partition-boundary blocks
            else:
                if work is not None:
                    work.wait()
                chk(lib.wgpu_rk_stage(ctx, j, 0))
        dt = C.c_double()
        chk(lib.wgpu_rk_end(ctx, C.byref(dt)))
        return dt.value

    def timeStep_tree(self, time: float, iteration: int):
        dt = self.step(time, iteration)
        return time + dt, iteration + 1, dt


def attach_exchange(sol, forest: Forest, rank: int, world: int, **kw) -> MultiGPUStepper:
    """Set the topology of `rank`, allocate the exchange buffers and route sol.timeStep_tree through the
    multi-GPU stepper."""
    st = MultiGPUStepper(sol, forest, rank, world, **kw)
    sol.timeStep_tree = st.timeStep_tree
    sol.RungeKuttaGeneric = st.step
    sol.stepper = st
    return st


class LockstepGroup:
    """Several ranks driven by ONE process (each rank a WabbitGPU context, on the same or on different devices),
    advanced in lockstep; patches move by device-to-device copies.  Used by the single-GPU parity test of the
    exchange path and usable as the `1 process x N devices` mode."""

    def __init__(self, sols, forest: Forest):
        import torch
        self.torch = torch
        self.world = len(sols)
        self.st = [MultiGPUStepper(s, forest, r, self.world, exchange=lambda *a: None, allreduce_min=lambda t: None, overlap=False)
                   for r, s in enumerate(sols)]

    def _move(self):
        W = self.world
        for r in range(W):
            so = np.concatenate([[0], np.cumsum(self.st[r].in_splits)])
            for p in range(W):
                n = self.st[r].in_splits[p]
                if n == 0:
                    continue
                ro = int(np.sum(self.st[p].out_splits[:r]))
                self.st[p].pool[ro:ro + n].copy_(self.st[r].send[int(so[p]):int(so[p]) + n])

    def step(self, time: float) -> float:
        torch = self.torch
        for s in self.st:
            s.sol._check(s.sol._lib.wgpu_rk_begin(s.sol._ctx, float(time)))
        if not self.st[0].sol.params.dt_fixed > 0.0:
            ts = [s._dtmin_tensor() for s in self.st]
            m = torch.stack([t.to(ts[0].device) for t in ts]).min()
            for t in ts:
                t.fill_(m.item())
        for s in self.st:
            s.sol._check(s.sol._lib.wgpu_rk_dt(s.sol._ctx, float(time)))
        for j in range(1, self.st[0].sol.params.n_stages + 1):
            for s in self.st:
                s.sol._check(s.sol._lib.wgpu_pack_halo(s.sol._ctx, j))
            self._move()
            for s in self.st:
                s.sol._check(s.sol._lib.wgpu_rk_stage(s.sol._ctx, j, 1))
                s.sol._check(s.sol._lib.wgpu_rk_stage(s.sol._ctx, j, 2))
        dts = []
        for s in self.st:
            dt = C.c_double()
            s.sol._check(s.sol._lib.wgpu_rk_end(s.sol._ctx, C.byref(dt)))
            dts.append(dt.value)
        assert all(d == dts[0] for d in dts)
        return dts[0]


# ----------------------------------------------------------------------------------------------------------------------
# Halo blocks: the multi-GPU mode for grids with level jumps and for the wavelet side (include/wabbit_gpu.h, wgpu_set_halo)
# ----------------------------------------------------------------------------------------------------------------------
def halo_list(forest: Forest, rank: int) -> np.ndarray:
    """lgt ids (1-based, owner*max_blocks + hvy) of the blocks of other ranks that appear in the 168 neighbour relations of
    `rank`'s blocks, ascending = sorted by (owner, hvy): the order they arrive in and the order of the halo slots."""
    N = forest.max_blocks
    nb = forest.neighbors(rank)
    ids = np.unique(nb[nb >= 1]).astype(np.int64)
    return ids[(ids - 1) // N != rank]


class HaloPlan:
    """Which blocks every rank mirrors and which of its own it sends, derived from the replicated light data (no handshake)."""

    def __init__(self, forest: Forest, rank: int, world: int):
        N = forest.max_blocks
        self.rank, self.world = rank, world
        self.n_own = forest.n_active(rank)
        lists = [halo_list(forest, r) for r in range(world)]
        mine = lists[rank]
        self.halo_lgt = mine.astype(np.int32)
        self.halo_hvy = (self.n_own + 1 + np.arange(len(mine))).astype(np.int32)
        if self.n_own + len(mine) > N:
            raise MemoryError(f"rank {rank}: {self.n_own} own blocks + {len(mine)} halo copies exceed max_blocks = {N}")
        owner = (mine - 1) // N
        self.recv_counts = [int((owner == p).sum()) for p in range(world)]
        lvl = np.zeros(len(mine), np.int32)
        tc = np.zeros(len(mine), np.int64)
        for p in range(world):
            sel = owner == p
            if sel.any():
                hvy_p, lvl_p, _, tc_p = forest.active(p)
                assert (hvy_p == np.arange(1, len(hvy_p) + 1)).all()
                k = (mine[sel] - 1) % N
                lvl[sel], tc[sel] = lvl_p[k], tc_p[k]
        self.halo_level, self.halo_tc = lvl, tc
        send, self.send_counts = [], []
        for p in range(world):
            t = lists[p][(lists[p] - 1) // N == rank] if p != rank else np.zeros(0, np.int64)
            send.append((t - 1) % N + 1)
            self.send_counts.append(len(t))
        self.send_hvy = np.concatenate(send).astype(np.int32)

    @property
    def n_halo(self) -> int:
        return len(self.halo_lgt)

    @property
    def n_send(self) -> int:
        return len(self.send_hvy)


class HaloStepper:
    """One rank of a multi-GPU run on a grid with level jumps: halo copies of the neighbouring blocks of other ranks are refreshed by
    ONE all-to-all of whole blocks per Runge-Kutta stage (received straight into the halo slots of the stage input), overlapped with
    the stage kernel on the blocks that have no halo neighbour.  `exchange_array` does the same for a named array before a
    wavelet-side call (waveletDecomposition_tree, refine_tree, download with ghosts)."""

    def __init__(self, sol, forest: Forest, rank: int, world: int, exchange: Optional[Callable] = None,
                 allreduce_min: Optional[Callable] = None, overlap: bool = True):
        import torch
        self.torch = torch
        self.sol, self.rank, self.world, self.overlap = sol, rank, world, overlap
        self.plan = plan = HaloPlan(forest, rank, world)
        lib, ctx = sol._lib, sol._ctx
        p = sol.params
        self.blk = p.n_eqn * int(np.prod([p.Bs[d] for d in range(p.dim)]))
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.send = torch.zeros(max(plan.n_send, 1) * self.blk, dtype=torch.float64, device=self.dev)
        sol._check(lib.wgpu_set_halo(ctx, plan.n_halo, _i32(plan.halo_lgt), _i32(plan.halo_hvy), _i32(plan.halo_level), plan.n_send,
                                     _i32(plan.send_hvy), C.c_void_p(self.send.data_ptr())))
        hvy, lvl, _, tc = forest.active(rank)
        sol.set_treecodes(np.concatenate([hvy, plan.halo_hvy]), np.concatenate([lvl, plan.halo_level]), np.concatenate([tc, plan.halo_tc]))
        sol.set_topology(hvy, lvl, forest.neighbors(rank), rank)
        self.in_splits = [c * self.blk for c in plan.send_counts]
        self.out_splits = [c * self.blk for c in plan.recv_counts]
        self._exchange = exchange or MultiGPUStepper._nccl_exchange.__get__(self)
        self._allreduce_min = allreduce_min or MultiGPUStepper._nccl_min.__get__(self)
        self.n_int = lib.wgpu_block_count(ctx, 1)
        self.n_bnd = lib.wgpu_block_count(ctx, 2)
        self._views: Dict[int, object] = {}

    def _view(self, ptr: int, n: int):
        if n == 0:
            return self.torch.zeros(0, dtype=self.torch.float64, device=self.dev)
        v = self._views.get(ptr)
        if v is None or v.numel() != n:
            v = self._views[ptr] = self.torch.as_tensor(_DevPtr(ptr, n), device=self.dev)
        return v

    def _dtmin_tensor(self):
        p = C.c_void_p()
        self.sol._check(self.sol._lib.wgpu_dtmin_pointer(self.sol._ctx, C.byref(p)))
        return self.torch.as_tensor(_DevPtr(p.value, 1), device=self.dev)

    def stage_halo(self, j: int):
        p, n = C.c_void_p(), C.c_int64()
        self.sol._check(self.sol._lib.wgpu_rk_stage_halo_pointer(self.sol._ctx, j, C.byref(p), C.byref(n)))
        return self._view(p.value or 0, n.value)

    def array_halo(self, array_id: int, slot: int = 0):
        p, n = C.c_void_p(), C.c_int64()
        self.sol._check(self.sol._lib.wgpu_halo_pointer(self.sol._ctx, array_id, slot, C.byref(p), C.byref(n)))
        return self._view(p.value or 0, n.value)

    def exchange_array(self, array_id: int = 0, slot: int = 0):
        """refresh the halo copies of a named array (blocking)"""
        self.sol._check(self.sol._lib.wgpu_pack_blocks(self.sol._ctx, array_id, slot))
        work = self._exchange(self.send, self.array_halo(array_id, slot), self.in_splits, self.out_splits)
        if work is not None:
            work.wait()

    def step(self, time: float, iteration: int = 0) -> float:
        lib, ctx, chk = self.sol._lib, self.sol._ctx, self.sol._check
        chk(lib.wgpu_rk_begin(ctx, float(time)))
        if self.world > 1 and not self.sol.params.dt_fixed > 0.0:
            self._allreduce_min(self._dtmin_tensor())
        chk(lib.wgpu_rk_dt(ctx, float(time)))
        for j in range(1, self.sol.params.n_stages + 1):
            chk(lib.wgpu_pack_halo(ctx, j))
            work = self._exchange(self.send, self.stage_halo(j), self.in_splits, self.out_splits)
            if self.overlap and self.n_bnd and self.n_int:
                chk(lib.wgpu_rk_stage(ctx, j, 1))            # blocks without a halo neighbour while the blocks are in flight
                if work is not None:
                    work.wait()
                chk(lib.wgpu_rk_stage(ctx, j, 2))            # partition-boundary blocks
            else:
                if work is not None:
                    work.wait()
                chk(lib.wgpu_rk_stage(ctx, j, 0))
        dt = C.c_double()
        chk(lib.wgpu_rk_end(ctx, C.byref(dt)))
        return dt.value

    def timeStep_tree(self, time: float, iteration: int):
        dt = self.step(time, iteration)
        return time + dt, iteration + 1, dt


def attach_halo(sol, forest: Forest, rank: int, world: int, **kw) -> HaloStepper:
    """Halo-block mode for `rank`: declare the halo slots, upload the topology and route the time step through HaloStepper."""
    st = HaloStepper(sol, forest, rank, world, **kw)
    sol.timeStep_tree = st.timeStep_tree
    sol.RungeKuttaGeneric = st.step
    sol.stepper = st
    return st


class HaloLockstepGroup:
    """Several ranks in halo mode driven by ONE process (contexts on one device), advanced in lockstep; blocks move by device copies.
    The single-GPU parity test of the halo path."""

    def __init__(self, sols, forest: Forest):
        import torch
        self.torch = torch
        self.world = len(sols)
        self.st = [HaloStepper(s, forest, r, self.world, exchange=lambda *a: None, allreduce_min=lambda t: None, overlap=False)
                   for r, s in enumerate(sols)]

    def _move(self, views):
        W = self.world
        for r in range(W):
            so = np.concatenate([[0], np.cumsum(self.st[r].in_splits)])
            for p in range(W):
                n = self.st[r].in_splits[p]
                if n == 0:
                    continue
                ro = int(np.sum(self.st[p].out_splits[:r]))
                views[p][ro:ro + n].copy_(self.st[r].send[int(so[p]):int(so[p]) + n])

    def exchange_array(self, array_id: int = 0, slot: int = 0):
        for s in self.st:
            s.sol._check(s.sol._lib.wgpu_pack_blocks(s.sol._ctx, array_id, slot))
        self._move([s.array_halo(array_id, slot) for s in self.st])
        self.torch.cuda.synchronize()

    def step(self, time: float, split: bool = False) -> float:
        torch = self.torch
        for s in self.st:
            s.sol._check(s.sol._lib.wgpu_rk_begin(s.sol._ctx, float(time)))
        if not self.st[0].sol.params.dt_fixed > 0.0:
            ts = [s._dtmin_tensor() for s in self.st]
            m = torch.stack([t.to(ts[0].device) for t in ts]).min()
            for t in ts:
                t.fill_(m.item())
        for s in self.st:
            s.sol._check(s.sol._lib.wgpu_rk_dt(s.sol._ctx, float(time)))
        for j in range(1, self.st[0].sol.params.n_stages + 1):
            for s in self.st:
                s.sol._check(s.sol._lib.wgpu_pack_halo(s.sol._ctx, j))
            if split:                                        # the overlap order: interior blocks before the halos arrive
                for s in self.st:
                    s.sol._check(s.sol._lib.wgpu_rk_stage(s.sol._ctx, j, 1))
            self._move([s.stage_halo(j) for s in self.st])
            for s in self.st:
                s.sol._check(s.sol._lib.wgpu_rk_stage(s.sol._ctx, j, 2 if split else 0))
        dts = []
        for s in self.st:
            dt = C.c_double()
            s.sol._check(s.sol._lib.wgpu_rk_end(s.sol._ctx, C.byref(dt)))
            dts.append(dt.value)
        assert all(d == dts[0] for d in dts)
        return dts[0]
