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
from typing import Callable, Dict, List, Optional, Sequence, Tuple

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


# phase timers of the multi-rank drivers (WABBIT_MG_TIMING=1): wall time per named phase, device synchronised at both ends
import os as _os
import time as _time
TIMING = {} if _os.environ.get("WABBIT_MG_TIMING") else None


class tick:
    def __init__(self, name, sol=None):
        self.name, self.sol = name, sol

    def __enter__(self):
        if TIMING is not None:
            if self.sol is not None:
                self.sol.synchronize()
            self.t0 = _time.perf_counter()
        return self

    def __exit__(self, *a):
        if TIMING is not None:
            if self.sol is not None:
                self.sol.synchronize()
            TIMING[self.name] = TIMING.get(self.name, 0.0) + (_time.perf_counter() - self.t0)
        return False


def _require_torch_stream(sol, torch):
    """torch.distributed orders a collective against torch's CURRENT stream, the library issues its pack / stage kernels on the context's
    stream: the two must be the same stream (or both the legacy default stream), otherwise pack -> all-to-all -> stage would race."""
    cur = int(torch.cuda.current_stream().cuda_stream)
    mine = int(getattr(sol, "stream", 0))
    if mine != cur:
        raise RuntimeError(f"the WabbitGPU context works on stream {mine:#x} but torch's current stream is {cur:#x}: create the context with "
                           "stream=torch.cuda.current_stream().cuda_stream (collectives are ordered against torch's current stream)")


class MultiGPUStepper:
    """RungeKuttaGeneric across ranks.  `exchange(send, recv, send_counts, recv_counts) -> handle-with-wait()` moves
    the patches; the default uses torch.distributed all_to_all_single (NCCL).  `allreduce_min(tensor)` reduces dt."""

    def __init__(self, sol, forest: Forest, rank: int, world: int, exchange: Optional[Callable] = None,
                 allreduce_min: Optional[Callable] = None, overlap: bool = True):
        import torch
        self.torch = torch
        self.sol, self.rank, self.world, self.overlap = sol, rank, world, overlap
        if exchange is None and world > 1 and getattr(sol, "comm_world", 1) != world:
            _require_torch_stream(sol, torch)
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
        # the context owns an NCCL communicator (WabbitGPU.comm_init): the whole step runs inside the library (wgpu_rk_steps)
        self.in_library = exchange is None and getattr(sol, "comm_world", 1) == world and world > 1
        if self.in_library:
            sol.comm_set_counts(self.plan.send_counts, self.plan.recv_counts)
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

    def steps(self, time: float, n_steps: int):
        """n_steps back to back inside the library (device-resident time and dt, one host read-back): (time after, last dt)"""
        if self.in_library or self.world == 1:
            return self.sol.RungeKuttaSteps(time, n_steps)
        dt = 0.0
        for _ in range(n_steps):
            dt = self.step(time)
            time += dt
        return time, dt

    def step(self, time: float, iteration: int = 0) -> float:
        if self.in_library:
            return self.sol.RungeKuttaSteps(time, 1)[1]
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


def fine_halo_list(forest: Forest, rank: int) -> np.ndarray:
    """lgt ids of the FINER neighbours (relations 113..168) of `rank`'s blocks that other ranks own, ascending"""
    N = forest.max_blocks
    nb = forest.neighbors(rank)[112:168]
    ids = np.unique(nb[nb >= 1]).astype(np.int64)
    return ids[(ids - 1) // N != rank]


class HaloPlan:
    """Which blocks every rank mirrors and which of its own it sends, derived from the replicated light data (no handshake): computed in
    libwabbit_host.so from the block positions (whost_halo_plan; relations are symmetric, so a rank's own blocks give both lists)."""

    def __init__(self, forest: Forest, rank: int, world: int):
        from ._native import host_lib
        N = forest.max_blocks
        self.rank, self.world = rank, world
        self.n_own = forest.n_active(rank)
        ntot, W = forest.n_blocks, forest.n_ranks
        assert W == world
        cap_h, cap_s = max(ntot - self.n_own, 1), max(self.n_own * max(W - 1, 1), 1)
        lgt, lvl, fine = np.zeros(cap_h, np.int32), np.zeros(cap_h, np.int32), np.zeros(cap_h, np.int32)
        tc = np.zeros(cap_h, np.int64)
        send, fsend = np.zeros(cap_s, np.int32), np.zeros(cap_s, np.int32)
        rc_, sc_, fc_ = np.zeros(W, np.int32), np.zeros(W, np.int32), np.zeros(W, np.int32)
        nh, ns, nf = C.c_int32(), C.c_int32(), C.c_int32()
        rc = host_lib().whost_halo_plan(forest._h, rank, C.byref(nh), _i32(lgt), _i32(lvl), tc.ctypes.data_as(C.POINTER(C.c_int64)), _i32(fine),
                                        _i32(rc_), C.byref(ns), _i32(send), _i32(sc_), C.byref(nf), _i32(fsend), _i32(fc_))
        if rc:
            raise RuntimeError(f"whost_halo_plan: {rc}")
        nh, ns, nf = nh.value, ns.value, nf.value
        if self.n_own + nh > N:
            raise MemoryError(f"rank {rank}: {self.n_own} own blocks + {nh} halo copies exceed max_blocks = {N}")
        self.halo_lgt, self.halo_level, self.halo_tc = lgt[:nh].copy(), lvl[:nh].copy(), tc[:nh].copy()
        self.halo_hvy = (self.n_own + 1 + np.arange(nh)).astype(np.int32)
        self.recv_counts = [int(v) for v in rc_]
        self.send_hvy, self.send_counts = send[:ns].copy(), [int(v) for v in sc_]
        # filtered copies of finer neighbours (lifted wavelets): who needs whose
        isf = fine[:nh] != 0
        self.fine_lgt = self.halo_lgt[isf].astype(np.int64)
        self.fine_recv_hvy = self.halo_hvy[isf].astype(np.int32)
        owner = (self.halo_lgt.astype(np.int64) - 1) // N
        self.fine_recv_counts = [int((isf & (owner == p)).sum()) for p in range(world)]
        self.fine_send_hvy, self.fine_send_counts = fsend[:nf].copy(), [int(v) for v in fc_]

    @property
    def n_halo(self) -> int:
        return len(self.halo_lgt)

    @property
    def n_send(self) -> int:
        return len(self.send_hvy)


class HaloPlanFromTables:
    """The same plan read off the 168-slot hvy_neighbor tables of every rank (the formulation the C++ plan is checked against in
    tests/test_host.py)."""

    def __init__(self, forest: Forest, rank: int, world: int):
        N = forest.max_blocks
        self.rank, self.world = rank, world
        self.n_own = forest.n_active(rank)
        lists = [halo_list(forest, r) for r in range(world)]
        fl = [fine_halo_list(forest, r) for r in range(world)]
        self.fine_lgt = fl[rank]
        self.fine_recv_counts = [int(((fl[rank] - 1) // N == p).sum()) for p in range(world)]
        fs = [fl[p][(fl[p] - 1) // N == rank] if p != rank else np.zeros(0, np.int64) for p in range(world)]
        self.fine_send_counts = [len(t) for t in fs]
        self.fine_send_hvy = ((np.concatenate(fs) - 1) % N + 1).astype(np.int32)
        mine = lists[rank]
        self.halo_lgt = mine.astype(np.int32)
        self.halo_hvy = (self.n_own + 1 + np.arange(len(mine))).astype(np.int32)
        owner = (mine - 1) // N
        self.recv_counts = [int((owner == p).sum()) for p in range(world)]
        self.fine_recv_hvy = self.halo_hvy[np.searchsorted(mine, self.fine_lgt)].astype(np.int32)
        send, self.send_counts = [], []
        for p in range(world):
            t = lists[p][(lists[p] - 1) // N == rank] if p != rank else np.zeros(0, np.int64)
            send.append((t - 1) % N + 1)
            self.send_counts.append(len(t))
        self.send_hvy = np.concatenate(send).astype(np.int32)


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
        self.in_library = exchange is None and getattr(sol, "comm_world", 1) == world and world > 1
        if exchange is None and world > 1 and not self.in_library:
            _require_torch_stream(sol, torch)
        self.plan = plan = HaloPlan(forest, rank, world)
        lib, ctx = sol._lib, sol._ctx
        p = sol.params
        self.blk = p.n_eqn * int(np.prod([p.Bs[d] for d in range(p.dim)]))
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.send = torch.zeros(max(plan.n_send, 1) * self.blk, dtype=torch.float64, device=self.dev)
        sol._check(lib.wgpu_set_halo(ctx, plan.n_halo, _i32(plan.halo_lgt), _i32(plan.halo_hvy), _i32(plan.halo_level), plan.n_send,
                                     _i32(plan.send_hvy), C.c_void_p(self.send.data_ptr())))
        hvy, lvl, _, tc = forest.active(rank)
        # own blocks + halo copies are the resident blocks; the neighbour relations are derived on the device (wgpu_set_grid);
        # WABBIT_HOST_TOPOLOGY=1: the Fortran-facing route through the 168-slot hvy_neighbor table instead
        import os
        if os.environ.get("WABBIT_HOST_TOPOLOGY"):
            sol.set_treecodes(np.concatenate([hvy, plan.halo_hvy]), np.concatenate([lvl, plan.halo_level]), np.concatenate([tc, plan.halo_tc]))
            sol.set_topology(hvy, lvl, forest.neighbors(rank), rank)
        else:
            sol.set_grid(np.concatenate([hvy, plan.halo_hvy]), np.concatenate([lvl, plan.halo_level]), np.concatenate([tc, plan.halo_tc]), hvy)
        self.in_splits = [c * self.blk for c in plan.send_counts]
        self.out_splits = [c * self.blk for c in plan.recv_counts]
        self._exchange = exchange or MultiGPUStepper._nccl_exchange.__get__(self)
        self._allreduce_min = allreduce_min or MultiGPUStepper._nccl_min.__get__(self)
        self.n_int = lib.wgpu_block_count(ctx, 1)
        self.n_bnd = lib.wgpu_block_count(ctx, 2)
        self._views: Dict[int, object] = {}
        # lifted wavelet: the filtered copies of finer neighbours on other ranks travel too (restrict_copy_at_CE needs the owner)
        w = getattr(p, "wavelet", "CDF40")
        self.lifted = len(w) == 5 and w[4] != "0"
        self.rblk = p.n_eqn * int(np.prod([p.Bs[d] // 2 for d in range(p.dim)]))
        if self.lifted and world > 1:
            self.rsend = torch.zeros(max(len(plan.fine_send_hvy), 1) * self.rblk, dtype=torch.float64, device=self.dev)
            sol._check(lib.wgpu_set_halo_restrict(ctx, len(plan.fine_recv_hvy), _i32(plan.fine_recv_hvy), len(plan.fine_send_hvy),
                                                  _i32(plan.fine_send_hvy), C.c_void_p(self.rsend.data_ptr())))
            self.r_in = [c * self.rblk for c in plan.fine_send_counts]
            self.r_out = [c * self.rblk for c in plan.fine_recv_counts]
        if self.in_library:
            if self.lifted:
                sol.comm_set_counts(plan.send_counts, plan.recv_counts, plan.fine_send_counts, plan.fine_recv_counts)
            else:
                sol.comm_set_counts(plan.send_counts, plan.recv_counts)

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

    def exchange_array(self, array_id: int = 0, slot: int = 0, filtered: bool = True):
        """refresh the halo copies of a named array (blocking); with a lifted wavelet also the filtered copies of the finer neighbours
        other ranks own (what the next wavelet-side synchronisation of this array restricts from)"""
        if self.in_library:
            self.sol._check(self.sol._lib.wgpu_exchange_array(self.sol._ctx, array_id, slot, int(bool(filtered))))
            return
        self.sol._check(self.sol._lib.wgpu_pack_blocks(self.sol._ctx, array_id, slot))
        work = self._exchange(self.send, self.array_halo(array_id, slot), self.in_splits, self.out_splits)
        if work is not None:
            work.wait()
        if self.lifted and filtered and self.world > 1:
            self.sol._check(self.sol._lib.wgpu_restrict_pack(self.sol._ctx, array_id, slot))
            p, n = C.c_void_p(), C.c_int64()
            self.sol._check(self.sol._lib.wgpu_restrict_halo_pointer(self.sol._ctx, C.byref(p), C.byref(n)))
            work = self._exchange(self.rsend, self._view(p.value or 0, n.value), self.r_in, self.r_out)
            if work is not None:
                work.wait()

    def steps(self, time: float, n_steps: int):
        if self.in_library or self.world == 1:
            return self.sol.RungeKuttaSteps(time, n_steps)
        dt = 0.0
        for _ in range(n_steps):
            dt = self.step(time)
            time += dt
        return time, dt

    def step(self, time: float, iteration: int = 0) -> float:
        if self.in_library:
            return self.sol.RungeKuttaSteps(time, 1)[1]
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
        if self.st[0].lifted and self.world > 1:          # filtered copies of finer neighbours on other ranks
            views = []
            for s in self.st:
                s.sol._check(s.sol._lib.wgpu_restrict_pack(s.sol._ctx, array_id, slot))
                p, n = C.c_void_p(), C.c_int64()
                s.sol._check(s.sol._lib.wgpu_restrict_halo_pointer(s.sol._ctx, C.byref(p), C.byref(n)))
                views.append(s._view(p.value or 0, n.value))
            W = self.world
            for r in range(W):
                so = np.concatenate([[0], np.cumsum(self.st[r].r_in)])
                for q in range(W):
                    n = self.st[r].r_in[q]
                    if n:
                        ro = int(np.sum(self.st[q].r_out[:r]))
                        views[q][ro:ro + n].copy_(self.st[r].rsend[int(so[q]):int(so[q]) + n])
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


# ----------------------------------------------------------------------------------------------------------------------
# Grid adaptation across ranks: the heavy data of refine_tree / adapt_tree with blocks partitioned over several GPUs
# ----------------------------------------------------------------------------------------------------------------------
class NcclTransport:
    """Collectives of one rank (one process per GPU, torch.distributed / NCCL)."""

    def __init__(self, rank: int, world: int):
        self.rank, self.world = rank, world

    def alltoall(self, send, recv, in_splits, out_splits, async_op: bool = False):
        import torch.distributed as dist
        return dist.all_to_all_single(recv[:sum(out_splits)], send[:sum(in_splits)], out_splits, in_splits, async_op=async_op)

    def allreduce_min_(self, t):
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MIN)

    def allreduce_max_np(self, a: np.ndarray) -> np.ndarray:
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    def allreduce_sum_np(self, a: np.ndarray) -> np.ndarray:
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allgather_np(self, a: np.ndarray, counts: Sequence[int]) -> np.ndarray:
        """concatenation over ranks of int32 arrays whose lengths (counts) every rank knows"""
        import torch
        import torch.distributed as dist
        m = max(max(counts), 1)
        t = torch.zeros(m, dtype=torch.int32).cuda()
        t[:len(a)] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).cuda()
        out = [torch.zeros(m, dtype=torch.int32).cuda() for _ in range(self.world)]
        dist.all_gather(out, t)
        return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)])


class LibTransport:
    """The collectives of one rank on the NCCL communicator the library owns (WabbitGPU.comm_init): no torch.distributed on the data path."""

    def __init__(self, sol):
        self.sol, self.rank, self.world = sol, sol.comm_rank, sol.comm_world

    def _reduce(self, a, op):
        a = np.ascontiguousarray(a, dtype=np.float64).copy()
        out = []
        for s0 in range(0, a.size, 4096):                     # the library reduces at most 4096 doubles per call
            part = np.ascontiguousarray(a.ravel()[s0:s0 + 4096])
            self.sol._check(self.sol._lib.wgpu_comm_allreduce(self.sol._ctx, part.ctypes.data_as(C.POINTER(C.c_double)), part.size, op))
            out.append(part)
        return np.concatenate(out).reshape(a.shape) if out else a

    def allreduce_max_np(self, a):
        return self._reduce(a, 0)

    def allreduce_sum_np(self, a):
        return self._reduce(a, 2)

    def allreduce_min_(self, t):
        raise RuntimeError("the dt reduction runs inside wgpu_rk_steps")

    def allgather_np(self, a, counts):
        a = np.ascontiguousarray(a, dtype=np.int32)
        cnt = np.ascontiguousarray(counts, dtype=np.int32)
        out = np.zeros(int(cnt.sum()), np.int32)
        self.sol._check(self.sol._lib.wgpu_comm_allgatherv_i32(self.sol._ctx, _i32(a), _i32(cnt), _i32(out)))
        return out

    def alltoall(self, *a, **k):
        raise RuntimeError("block transport runs inside wgpu_ship_blocks / wgpu_exchange_array")


class ThreadTransport:
    """The same collectives between `world` host threads of ONE process (every rank a device context on the same GPU): the
    single-GPU test harness of the multi-rank drivers."""

    class Shared:
        def __init__(self, world: int):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    def __init__(self, shared: "ThreadTransport.Shared", rank: int):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def _sync(self):
        import torch
        torch.cuda.synchronize()
        self.sh.barrier.wait()

    def alltoall(self, send, recv, in_splits, out_splits, async_op: bool = False):
        self.sh.slots[self.rank] = (send, in_splits)
        self._sync()
        ro = 0
        for q in range(self.world):
            s, sp = self.sh.slots[q]
            so = int(sum(sp[:self.rank]))
            n = sp[self.rank]
            assert n == out_splits[q]
            if n:
                recv[ro:ro + n].copy_(s[so:so + n])
            ro += n
        self._sync()
        return None

    def allreduce_min_(self, t):
        self.sh.slots[self.rank] = t
        self._sync()
        m = min(float(x.item()) for x in self.sh.slots)
        self._sync()
        t.fill_(m)

    def allreduce_max_np(self, a):
        self.sh.slots[self.rank] = np.asarray(a, dtype=np.float64)
        self._sync()
        out = np.max(np.stack(self.sh.slots), axis=0)
        self._sync()
        return out

    def allreduce_sum_np(self, a):
        self.sh.slots[self.rank] = np.asarray(a, dtype=np.float64)
        self._sync()
        out = np.sum(np.stack(self.sh.slots), axis=0)      # rank order: the same on every rank
        self._sync()
        return out

    def allgather_np(self, a, counts):
        self.sh.slots[self.rank] = np.asarray(a, dtype=np.int32)
        self._sync()
        out = np.concatenate(self.sh.slots)
        self._sync()
        return out


class DistributedWabbit:
    """One rank of a multi-GPU adaptive run: time step (HaloStepper) plus refine_tree and one coarsening sweep of adapt_tree with the
    blocks partitioned by the space-filling curve (balanceLoad_tree) over `world` GPUs.  Light data are replicated: every rank derives
    the same new grid, partition and transfer lists from the all-gathered refinement flags, as WABBIT does from synchronize_lgt_data.
    Heavy data move as whole blocks: block_xfer (LIB/MPI/block_xfer_nonblocking.f90:16) -> gather kernel + all-to-all + scatter kernel."""

    HVY_BLOCK, HVY_WORK = 0, 1

    def __init__(self, sol, forest: Forest, rank: int, world: int, transport=None, overlap: bool = True):
        import torch
        self.torch = torch
        self.sol, self.rank, self.world = sol, rank, world
        self.in_library = transport is None and getattr(sol, "comm_world", 1) == world and world > 1
        if transport is None and world > 1 and not self.in_library:
            _require_torch_stream(sol, torch)
        self.tr = transport or (LibTransport(sol) if self.in_library else NcclTransport(rank, world))
        self.overlap = overlap
        self.dev = torch.device("cuda", torch.cuda.current_device())
        p = sol.params
        self.blk = p.n_eqn * int(np.prod([p.Bs[d] for d in range(p.dim)]))
        self.attach(forest)

    # ------------------------------------------------------------------ topology
    def attach(self, forest: Forest):
        self.forest = forest
        if self.in_library:
            self.stepper = HaloStepper(self.sol, forest, self.rank, self.world, overlap=self.overlap)
        else:
            self.stepper = HaloStepper(self.sol, forest, self.rank, self.world,
                                       exchange=lambda s, r, i, o: self.tr.alltoall(s, r, i, o, async_op=True),
                                       allreduce_min=self.tr.allreduce_min_, overlap=self.overlap)
        self.counts = [forest.n_active(r) for r in range(self.world)]
        self.off = np.concatenate([[0], np.cumsum(self.counts)]).astype(np.int64)

    def timeStep_tree(self, time: float, iteration: int):
        return self.stepper.timeStep_tree(time, iteration)

    def _global_blocks(self, forest: Forest):
        lv, ix = [], []
        for r in range(self.world):
            _, l, x, _ = forest.active(r)
            lv.append(l)
            ix.append(x)
        return np.concatenate(lv), np.concatenate(ix)

    def _shadow(self, forest: Forest) -> Forest:
        """the same leaves on ONE rank: its block order is the global space-filling-curve order = rank-major order of `forest`"""
        lv, ix = self._global_blocks(forest)
        sh = Forest.from_blocks(forest.dim, forest.Jmax, lv, ix, block_dist=forest.block_dist, n_ranks=1, max_blocks=len(lv),
                                periodic=forest.periodic)
        _, l2, x2, _ = sh.active(0)
        assert np.array_equal(l2, lv) and np.array_equal(x2, ix), "global block order is not the space-filling-curve order"
        return sh

    def _partition(self, shadow: Forest) -> Forest:
        _, lv, ix, _ = shadow.active(0)
        return Forest.from_blocks(shadow.dim, shadow.Jmax, lv, ix, block_dist=shadow.block_dist, n_ranks=self.world,
                                  max_blocks=self.forest.max_blocks, periodic=shadow.periodic)

    # ------------------------------------------------------------------ block transport
    def _ship(self, array, src_rank, src_slot, dst_rank, first_free: int):
        """Item k: a block that lives on src_rank[k] in slot src_slot[k] (1-based) of `array` = (array_id, slot) and is needed on
        dst_rank[k]; the item order is the same on every rank.  Remote blocks are received into free slots from `first_free` on.
        Returns (local slot of every item with dst_rank == me, in item order; next free slot)."""
        me, W, torch = self.rank, self.world, self.torch
        src_rank, src_slot, dst_rank = (np.asarray(a) for a in (src_rank, src_slot, dst_rank))
        if self.in_library:                     # block_xfer inside the library: gather -> ncclSend / ncclRecv straight into the free slots
            sr, ss, dr = (np.ascontiguousarray(a, dtype=np.int32) for a in (src_rank, src_slot, dst_rank))
            loc = np.zeros(max(int((dr == me).sum()), 1), np.int32)
            nf = C.c_int32()
            self.sol._check(self.sol._lib.wgpu_ship_blocks(self.sol._ctx, array[0], array[1], len(sr), _i32(sr), _i32(ss), _i32(dr), int(first_free),
                                                           _i32(loc), C.byref(nf)))
            return loc[:int((dr == me).sum())].astype(np.int64), nf.value
        send_ids, in_splits, out_splits = [], [], []
        mine = dst_rank == me
        local = np.where(src_rank[mine] == me, src_slot[mine], 0).astype(np.int64)
        nxt = first_free
        pos_mine = np.flatnonzero(mine)
        for q in range(W):
            s = np.flatnonzero((src_rank == me) & (dst_rank == q)) if q != me else np.zeros(0, np.int64)
            send_ids.append(src_slot[s])
            in_splits.append(len(s) * self.blk)
            r = np.flatnonzero((src_rank[pos_mine] == q)) if q != me else np.zeros(0, np.int64)
            local[r] = nxt + np.arange(len(r))
            nxt += len(r)
            out_splits.append(len(r) * self.blk)
        if nxt - 1 > self.sol.max_blocks:
            raise MemoryError(f"rank {me}: block transfer needs {nxt - 1} slots, max_blocks = {self.sol.max_blocks}")
        send_ids = np.concatenate(send_ids).astype(np.int32) if send_ids else np.zeros(0, np.int32)
        n_send, n_recv = len(send_ids), nxt - first_free
        send = torch.empty(max(n_send, 1) * self.blk, dtype=torch.float64, device=self.dev)
        recv = torch.empty(max(n_recv, 1) * self.blk, dtype=torch.float64, device=self.dev)
        lib, ctx = self.sol._lib, self.sol._ctx
        if n_send:
            self.sol._check(lib.wgpu_gather_blocks(ctx, array[0], array[1], n_send, _i32(send_ids), C.c_void_p(send.data_ptr())))
        self.tr.alltoall(send, recv, in_splits, out_splits)
        if n_recv:
            ids = (first_free + np.arange(n_recv)).astype(np.int32)
            self.torch.cuda.synchronize()
            self.sol._check(lib.wgpu_scatter_blocks(ctx, array[0], array[1], n_recv, _i32(ids), C.c_void_p(recv.data_ptr())))
        return local, nxt

    def _owner(self, off, idx):
        return np.searchsorted(off[1:], idx, side="right")

    # ------------------------------------------------------------------ refine_tree
    def refine_tree(self, refine_flags: Optional[np.ndarray] = None) -> Forest:
        """refine_tree (LIB/MESH/refine_tree.f90:15) across ranks: every rank interpolates its own mothers (ghost nodes from the halo
        copies), then the new blocks move to their owners in the new partition (balanceLoad_tree).  refine_flags: one +1/0 flag per
        block of the GLOBAL grid in space-filling-curve order (None = everywhere)."""
        me, sol = self.rank, self.sol
        old, ooff = self.forest, self.off
        with tick("refine: exchange_array", sol):
            self.stepper.exchange_array(0, 0)
        try:
            with tick("refine: refine_global (host)"):
                new, mo, da, ks, kd = old.refine_global(refine_flags)
        except MemoryError as e:
            raise RuntimeError(f"refine_tree: {e}")
        t_np = tick("refine: id lists (numpy)").__enter__()
        noff = np.concatenate([[0], np.cumsum([new.n_active(r) for r in range(self.world)])]).astype(np.int64)
        nd = 2 ** old.dim
        mo, da, ks, kd = (a.astype(np.int64) - 1 for a in (mo, da, ks, kd))          # 0-based global indices
        # blocks derived from my old blocks, in new global order -> intermediate local slots 1..n_tmp
        my_m = self._owner(ooff, mo) == me
        my_k = self._owner(ooff, ks) == me
        d_new = da.reshape(-1, nd)[my_m].ravel()
        derived = np.sort(np.concatenate([kd[my_k], d_new]))
        tmp_slot = lambda j: np.searchsorted(derived, j) + 1
        if len(derived) > sol.max_blocks:
            raise MemoryError("refine_tree: the refined blocks of this rank do not fit max_blocks")
        m_loc = (mo[my_m] - ooff[me] + 1).astype(np.int32)
        t_np.__exit__()
        with tick("refine: wgpu_refine", sol):
            sol._check(sol._lib.wgpu_refine(sol._ctx, len(m_loc), _i32(m_loc), _i32(tmp_slot(d_new).astype(np.int32)), int(my_k.sum()),
                                            _i32((ks[my_k] - ooff[me] + 1).astype(np.int32)), _i32(tmp_slot(kd[my_k]).astype(np.int32))))
        t_np = tick("refine: id lists (numpy)").__enter__()
        # every new block: who holds it now (the owner of the old block it derives from) and in which intermediate slot
        holder = np.empty(new.n_blocks, np.int64)
        holder[kd] = self._owner(ooff, ks)
        holder[da] = np.repeat(self._owner(ooff, mo), nd)
        slot = np.empty(new.n_blocks, np.int64)
        for r in range(self.world):                         # intermediate slots are ranks' private orderings of contiguous index sets
            sel = np.flatnonzero(holder == r)
            slot[sel] = np.arange(1, len(sel) + 1)
        dst = self._owner(noff, np.arange(new.n_blocks))
        n_tmp = len(derived)
        t_np.__exit__()
        with tick("refine: ship", sol):
            local, _ = self._ship((0, 0), holder, slot, dst, n_tmp + 1)
        mine_new = np.flatnonzero(dst == me)
        with tick("refine: move_blocks", sol):
            sol._check(sol._lib.wgpu_move_blocks(sol._ctx, len(mine_new), _i32(local.astype(np.int32)),
                                                 _i32((mine_new - noff[me] + 1).astype(np.int32))))
        with tick("refine: attach (halo plan + set_grid)", sol):
            self.attach(new)
        return new

    # ------------------------------------------------------------------ adapt_tree (one coarsening sweep, unlifted wavelets)
    def global_norm(self, eps_norm: str = "Linfty", thresh_comp=None) -> np.ndarray:
        """componentWiseNorm_tree over all ranks: MPI_MAX for Linfty, MPI_SUM of the ranks' partial sums for L1 / L2 / H1
        (componentWiseNorm_tree.f90:283-326), then the threshold_state_vector_component treatment shared with the single-rank driver."""
        from .solver import threshold_norm
        loc = self.sol.componentWiseNorm_tree((0, 0), eps_norm)
        if eps_norm == "Linfty":
            nrm = self.tr.allreduce_max_np(loc)
        elif eps_norm == "L1":
            nrm = self.tr.allreduce_sum_np(loc)
        else:                                                # L2 / H1: sqrt of the summed squares
            nrm = np.sqrt(self.tr.allreduce_sum_np(loc * loc))
        return threshold_norm(nrm, thresh_comp)

    def adapt_tree(self, eps: Optional[float] = None, eps_normalized: bool = True, Jmin: int = 1, force_maxlevel_dealiasing: bool = False,
                   thresh_comp=None, useSecurityZone: Optional[bool] = None, mask_keeps=None, full_tree: Optional[bool] = None,
                   eps_norm: str = "Linfty"):
        """One coarsening sweep of adapt_tree (LIB/MESH/adapt_tree.f90:11) across ranks, indicator "threshold-state-vector", Linfty norm,
        unlifted wavelets (as WabbitGPU.adapt_tree): norm -> all-reduce MAX; halo refresh; decomposition + flags per rank; flags
        all-gathered (synchronize_lgt_data); completeness / gradedness on the replicated light data; sister blocks gathered on the
        mother's new owner and blocks that stay moved to theirs (block_xfer); mothers assembled from the scaling coefficients.
        Returns (new forest, number of blocks before, after)."""
        me, sol = self.rank, self.sol
        old, ooff = self.forest, self.off
        w = sol.params.wavelet
        lifted = not (len(w) == 5 and w[4] == "0")
        use_ce = lifted if sol.params.useCoarseExtension < 0 else bool(sol.params.useCoarseExtension)
        self.refinement_status = None
        if use_ce if full_tree is None else full_tree:
            # the reference's full-tree algorithm (coarse extension and security zone for lifted wavelets; fulltree.py)
            from .fulltree import DistributedFullTree
            with tick("adapt: norm", sol):
                norm = self.global_norm(eps_norm, thresh_comp) if eps_normalized else None
            n0 = old.n_blocks
            sz = (lifted if sol.params.useSecurityZone < 0 else bool(sol.params.useSecurityZone)) if useSecurityZone is None else bool(useSecurityZone)
            ft = DistributedFullTree(self, Jmin=Jmin)
            new, _ = ft.adapt(eps=sol.params.eps if eps is None else eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp,
                              force_maxlevel_dealiasing=force_maxlevel_dealiasing, use_security_zone=sz, mask_keeps=mask_keeps)
            self.refinement_status = ft.leaf_status              # global space-filling-curve order, the same on every rank
            return new, n0, new.n_blocks
        if mask_keeps is not None:
            raise ValueError("adapt_tree: threshold_mask needs the full-tree algorithm")
        WD = (self.HVY_WORK, 2)
        norm = self.global_norm(eps_norm, thresh_comp) if eps_normalized else None
        self.stepper.exchange_array(0, 0)
        sol.waveletDecomposition_tree((0, 0), WD)
        st = sol.threshold_tree(WD, eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, level_ref=old.Jmax)
        st = self.tr.allgather_np(st, self.counts)
        lv, _ = self._global_blocks(old)
        if force_maxlevel_dealiasing:
            st = np.where(lv == old.Jmax, -1, st)
        n0 = old.n_blocks
        if not (st == -1).any():
            return old, n0, n0
        new, st_final, mo, da, ks, kd = old.coarsen_global(st, Jmin)
        if not (st_final == -1).any():
            return old, n0, n0
        noff = np.concatenate([[0], np.cumsum([new.n_active(r) for r in range(self.world)])]).astype(np.int64)
        nd = 2 ** old.dim
        mo, da, ks, kd = (a.astype(np.int64) - 1 for a in (mo, da, ks, kd))
        n_own = self.counts[me]
        # blocks that stay: hvy_block of the old owner -> the new owner
        loc_k, nxt = self._ship((0, 0), self._owner(ooff, ks), ks - ooff[self._owner(ooff, ks)] + 1, self._owner(noff, kd), n_own + 1)
        # daughters: decomposed data of the old owner -> the mother's new owner
        m_rank = np.repeat(self._owner(noff, mo), nd)
        loc_d, _ = self._ship(WD, self._owner(ooff, da), da - ooff[self._owner(ooff, da)] + 1, m_rank, nxt)
        kd_mine = kd[self._owner(noff, kd) == me]
        sol._check(sol._lib.wgpu_move_blocks(sol._ctx, len(kd_mine), _i32(loc_k.astype(np.int32)), _i32((kd_mine - noff[me] + 1).astype(np.int32))))
        mo_mine = mo[self._owner(noff, mo) == me]
        sol._check(sol._lib.wgpu_coarsen(sol._ctx, len(mo_mine), _i32((mo_mine - noff[me] + 1).astype(np.int32)), _i32(loc_d.astype(np.int32)),
                                         WD[0], WD[1]))
        self.attach(new)
        return new, n0, new.n_blocks
