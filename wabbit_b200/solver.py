"""Host-side mirror of WABBIT's tree-level routines for the block hot path, on top of the C ABI.

Method names and argument meaning follow the reference routines they replace
(LIB/TIME/runge_kutta_generic.f90, LIB/TIME/RHS_wrapper.f90, LIB/TIME/calculate_time_step.f90,
LIB/TIME/timeStep_tree.f90, LIB/MPI/synchronize_ghosts_generic.f90); errors surface as `WabbitAbort`
carrying the reference's integer abort code (LIB/MODULE/module_globals.f90:131-161).

Heavy arrays on the host are NumPy arrays `hvy[number_blocks, ncomp, nz, ny, nx]` (C order), i.e. exactly
the memory of the Fortran array hvy(nx,ny,nz,ncomp,number_blocks).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from ._native import WgpuConfig, gpu_lib
from .forest import Forest
from .params import Params

HVY_BLOCK, HVY_WORK, HVY_MASK, HVY_TMP = 0, 1, 2, 3


def threshold_norm(norm, thresh_comp=None):
    """The norm adapt_tree divides the details by (eps_normalized): componentWiseNorm_tree's treatment of
    threshold_state_vector_component (componentWiseNorm_tree.f90:119-163: components with 0 are not computed, components of a group >= 2
    share the group's maximum) and coarseningIndicator_tree.f90:165-167 (norm <= 1e-9 -> 1).  `norm` is the plain per-component norm,
    already reduced over all ranks.  One helper for the single-rank and the multi-rank driver."""
    norm = np.array(norm, dtype=np.float64, copy=True)
    if thresh_comp is not None:
        tc = np.asarray(thresh_comp)
        norm = np.where(tc == 0, -1.0, norm)
        for l in range(2, int(tc.max()) + 1):
            if (tc == l).any():
                norm[tc == l] = norm[tc == l].max()
    norm[norm <= 1.0e-9] = 1.0
    return norm


class WabbitAbort(RuntimeError):
    """The library's equivalent of `call abort(code, msg)`."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class WabbitGPU:
    """One rank's device-resident forest (one process per GPU)."""

    def __init__(self, params: Params, max_blocks: int, device: int = 0, stream: Optional[int] = None):
        self.params = params.finalize()
        self.max_blocks = int(max_blocks)
        self._lib = gpu_lib()
        self._cfg: WgpuConfig = params.to_config(self.max_blocks, device)
        self._ctx = C.c_void_p()
        rc = self._lib.wgpu_create(C.byref(self._cfg), C.byref(self._ctx))
        if rc:
            buf = C.create_string_buffer(512)
            self._lib.wgpu_last_error(None, buf, 512)
            self._ctx = None
            raise WabbitAbort(rc, buf.value.decode())
        self.stream = 0 if stream is None else int(stream)          # cudaStream_t all work of this context is issued on (0 = legacy default)
        if stream is not None:
            self._check(self._lib.wgpu_set_stream(self._ctx, C.c_void_p(stream)))
        self.hvy_active = np.zeros(0, np.int32)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int):
        if rc:
            buf = C.create_string_buffer(512)
            self._lib.wgpu_last_error(self._ctx, buf, 512)
            raise WabbitAbort(rc, buf.value.decode())

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.wgpu_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._check(self._lib.wgpu_synchronize(self._ctx))

    def profile(self, enable: bool = True):
        self._check(self._lib.wgpu_profile(self._ctx, int(enable)))

    def profile_read(self):
        """(number of stage-kernel launches recorded, their summed duration in ms)"""
        n, ms = C.c_int32(), C.c_double()
        self._check(self._lib.wgpu_profile_read(self._ctx, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    @property
    def launch_count(self) -> int:
        return int(self._lib.wgpu_launch_count(self._ctx))

    @property
    def device_bytes(self) -> int:
        return int(self._lib.wgpu_device_bytes(self._ctx))

    def host_shape(self, ncomp: Optional[int] = None):
        p = self.params
        nz = p.Bs[2] + 2 * p.g if p.dim == 3 else 1
        return (self.max_blocks, p.n_eqn if ncomp is None else ncomp, nz, p.Bs[1] + 2 * p.g, p.Bs[0] + 2 * p.g)

    # ------------------------------------------------------------------ topology
    def set_topology(self, hvy_active: np.ndarray, level: np.ndarray, hvy_neighbor: np.ndarray, rank: int = 0):
        """Upload what updateMetadata_tree produced: hvy_active (1-based), level per active block and
        hvy_neighbor[168, max_blocks] (lgt ids)."""
        hvy_active = np.ascontiguousarray(hvy_active, dtype=np.int32)
        level = np.ascontiguousarray(level, dtype=np.int32)
        hvy_neighbor = np.ascontiguousarray(hvy_neighbor, dtype=np.int32)
        assert hvy_neighbor.shape[0] == 168
        self._check(self._lib.wgpu_set_topology(self._ctx, len(hvy_active), _i32(hvy_active), _i32(level), _i32(hvy_neighbor),
                                                hvy_neighbor.shape[1], rank))
        self.hvy_active = hvy_active.copy()

    def set_treecodes(self, hvy_active: np.ndarray, level: np.ndarray, treecode: np.ndarray):
        """Block positions (numerical binary treecodes, module_treelib.f90:837) -- needed for grids with level jumps."""
        hvy_active = np.ascontiguousarray(hvy_active, dtype=np.int32)
        level = np.ascontiguousarray(level, dtype=np.int32)
        treecode = np.ascontiguousarray(treecode, dtype=np.int64)
        self._check(self._lib.wgpu_set_treecodes(self._ctx, len(hvy_active), _i32(hvy_active), _i32(level),
                                                 treecode.ctypes.data_as(C.POINTER(C.c_int64))))

    def set_grid(self, hvy_ids: np.ndarray, level: np.ndarray, treecode: np.ndarray, hvy_active: Optional[np.ndarray] = None):
        """Topology derived on the device (wgpu_set_grid): the resident blocks (hvy id, level, treecode) and the active list; no hvy_neighbor
        table.  hvy_active None: every resident block that is not a halo copy declared by wgpu_set_halo is active."""
        hvy_ids = np.ascontiguousarray(hvy_ids, dtype=np.int32)
        level = np.ascontiguousarray(level, dtype=np.int32)
        treecode = np.ascontiguousarray(treecode, dtype=np.int64)
        act = hvy_ids if hvy_active is None else np.ascontiguousarray(hvy_active, dtype=np.int32)
        self._check(self._lib.wgpu_set_grid(self._ctx, len(hvy_ids), _i32(hvy_ids), _i32(level), treecode.ctypes.data_as(C.POINTER(C.c_int64)),
                                            len(act), _i32(act)))
        self.hvy_active = act.copy()

    def set_active(self, hvy_active: np.ndarray):
        """the active list of a pass on the blocks registered by set_grid (wgpu_set_active)"""
        act = np.ascontiguousarray(hvy_active, dtype=np.int32)
        self._check(self._lib.wgpu_set_active(self._ctx, len(act), _i32(act)))
        self.hvy_active = act.copy()

    def topology_tables(self):
        """(nbr27, wnbr27, counts dict, lists dict) of the current topology, read back from the device (tests)"""
        N = self.max_blocks
        nbr = np.zeros((N, 27), np.int32)
        wnbr = np.full((N, 27), -1, np.int32)
        cnt = np.zeros(8, np.int32)
        self._check(self._lib.wgpu_topology_tables(self._ctx, _i32(nbr), _i32(wnbr), _i32(cnt)))
        names = ("n_active", "n_jump", "n_wjump", "n_ce", "n_rst", "n_int", "n_bnd", "has_jumps")
        counts = dict(zip(names, (int(v) for v in cnt)))
        lists = {}
        for which, (name, key) in enumerate((("jump", "n_jump"), ("wjump", "n_wjump"), ("ce", "n_ce"), ("rst", "n_rst"), ("int", "n_int"), ("bnd", "n_bnd"))):
            n = counts[key]
            a, b = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
            self._check(self._lib.wgpu_topology_list(self._ctx, which, n, _i32(a), _i32(b)))
            lists[name] = (a[:n], b[:n])
        return nbr, wnbr, counts, lists

    def set_forest(self, forest: Forest, rank: int = 0, host_tables: Optional[bool] = None):
        """Upload a grid.  Default: the topology is derived on the device from the block positions (wgpu_set_grid); host_tables = True (or
        WABBIT_HOST_TOPOLOGY=1) takes the Fortran-facing route instead: the 168-slot hvy_neighbor table of updateMetadata_tree through
        wgpu_set_treecodes + wgpu_set_topology."""
        import os
        hvy, lvl, _, tc = forest.active(rank)
        if host_tables is None:
            host_tables = bool(os.environ.get("WABBIT_HOST_TOPOLOGY"))
        if host_tables or forest.n_ranks > 1:
            self.set_treecodes(hvy, lvl, tc)
            self.set_topology(hvy, lvl, forest.neighbors(rank), rank)
        else:
            self.set_grid(hvy, lvl, tc)

    # ------------------------------------------------------------------ data movement
    def _ids(self, hvy_ids):
        return np.ascontiguousarray(self.hvy_active if hvy_ids is None else hvy_ids, dtype=np.int32)

    def upload(self, host: np.ndarray, array_id: int = HVY_BLOCK, slot: int = 0, hvy_ids: Optional[Sequence[int]] = None):
        assert host.dtype == np.float64 and host.flags.c_contiguous and host.ndim == 5
        ids = self._ids(hvy_ids)
        self._check(self._lib.wgpu_upload(self._ctx, array_id, slot, _i32(ids), len(ids), C.c_void_p(host.ctypes.data), host.shape[1]))

    def download(self, host: np.ndarray, array_id: int = HVY_BLOCK, slot: int = 0, hvy_ids: Optional[Sequence[int]] = None,
                 g_sync: Optional[int] = None):
        assert host.dtype == np.float64 and host.flags.c_contiguous and host.ndim == 5
        ids = self._ids(hvy_ids)
        gs = self.params.g if g_sync is None else g_sync
        self._check(self._lib.wgpu_download(self._ctx, array_id, slot, _i32(ids), len(ids), C.c_void_p(host.ctypes.data), host.shape[1], gs))

    def upload_ptr(self, host_ptr: int, ncomp: int, array_id: int = HVY_BLOCK, slot: int = 0, hvy_ids=None):
        ids = self._ids(hvy_ids)
        self._check(self._lib.wgpu_upload(self._ctx, array_id, slot, _i32(ids), len(ids), C.c_void_p(host_ptr), ncomp))

    def download_ptr(self, host_ptr: int, ncomp: int, array_id: int = HVY_BLOCK, slot: int = 0, hvy_ids=None, g_sync=None):
        ids = self._ids(hvy_ids)
        gs = self.params.g if g_sync is None else g_sync
        self._check(self._lib.wgpu_download(self._ctx, array_id, slot, _i32(ids), len(ids), C.c_void_p(host_ptr), ncomp, gs))

    def set_transfer_mode(self, upload_dma: bool = True, download_dma: bool = True):
        """page-locked 3-D host arrays: copy engines (plane spans by DMA + a layout kernel, default) or zero-copy layout kernels, per direction
        (wgpu_set_transfer_mode)"""
        self._check(self._lib.wgpu_set_transfer_mode(self._ctx, int(bool(upload_dma)), int(bool(download_dma))))

    def set_ghost_filter(self, ignore_filter: bool):
        """ignore_Filter of sync_ghosts_tree (synchronize_ghosts_generic.f90:125-153): False (default) = restriction through the HD
        filter of a lifted wavelet in download(g_sync>0) / waveletDecomposition_tree / refine_tree, True = plain decimation."""
        self._check(self._lib.wgpu_set_ghost_filter(self._ctx, int(bool(ignore_filter))))

    def set_mask_sphere(self, center0=None, velocity=(0.0, 0.0, 0.0), radius: float = 0.0, smoothing_width: float = 0.0):
        """analytic penalization mask of a (translating) sphere evaluated inside the stage kernel (wgpu_set_mask_sphere); center0 = None
        switches back to the hvy_mask array"""
        dp = C.POINTER(C.c_double)
        if center0 is None:
            self._check(self._lib.wgpu_set_mask_sphere(self._ctx, 0, None, None, 0.0, 0.0))
            return
        c = np.ascontiguousarray(center0, dtype=np.float64)
        v = np.ascontiguousarray(velocity, dtype=np.float64)
        self._check(self._lib.wgpu_set_mask_sphere(self._ctx, 1, c.ctypes.data_as(dp), v.ctypes.data_as(dp), float(radius), float(smoothing_width)))

    # ------------------------------------------------------------------ reference routines
    STAT_NAMES = ("meanflow_x", "meanflow_y", "meanflow_z", "e_kin", "ACM_energy", "mask_volume", "sponge_volume", "penal_power_solid_input",
                  "penal_power_solid_dissipation", "penal_power_sponge", "force_x", "force_y", "force_z", "umag", "div_max", "div_min",
                  "u_residual_x", "u_residual_y", "u_residual_z")

    def create_mask_device(self, time: float, geometry: str, center, velocity=(0.0, 0.0, 0.0), radius: float = 0.5, smoothing_width: float = 0.0,
                           L_sponge: float = 0.0, p_sponge: float = 20.0):
        """createMask_tree for a closed-form geometry evaluated on the device into the resident hvy_mask (wgpu_create_mask): "cylinder" (2-D,
        with the p-norm sponge) or "sphere" (3-D, optionally translating)"""
        gid = {"cylinder": 1, "circle": 1, "sphere": 2, "sphere-fixed": 2}[geometry]
        c = (C.c_double * 3)(*(list(center) + [0.0, 0.0, 0.0])[:3])
        v = (C.c_double * 3)(*(list(velocity) + [0.0, 0.0, 0.0])[:3])
        self._check(self._lib.wgpu_create_mask(self._ctx, float(time), gid, c, v, float(radius), float(smoothing_width), float(L_sponge), float(p_sponge)))

    VORT_STAT_NAMES = ("enstrophy", "max_vort", "helicity", "dissipation")

    def statistics_ACM(self, time: float = 0.0, with_divergence: bool = True, with_vorticity: bool = False) -> dict:
        """STATISTICS_ACM's integral quantities reduced on the device (and over the ranks of the communicator): wgpu_statistics.
        with_vorticity adds enstrophy / max_vort / helicity / dissipation (statistics_ACM.f90:371-387; one rank)"""
        out = (C.c_double * 23)()
        self._check(self._lib.wgpu_statistics(self._ctx, float(time), int(bool(with_divergence)) | (2 if with_vorticity else 0), out))
        names = self.STAT_NAMES + (self.VORT_STAT_NAMES if with_vorticity else ())
        return dict(zip(names, [float(x) for x in out]))

    def sync_ghosts_RHS_tree(self, g_minus: Optional[int] = None, g_plus: Optional[int] = None):
        """synchronize_ghosts_generic.f90:155-174"""
        g = self.params.g_rhs
        self._check(self._lib.wgpu_sync_ghosts(self._ctx, HVY_BLOCK, 0, g if g_minus is None else g_minus, g if g_plus is None else g_plus))

    def RHS_wrapper(self, time: float, dst_slot: int, src_slot: int = 0):
        """RHS_wrapper.f90:16 -- hvy_work(:,:,:,:,:,dst_slot) = RHS(hvy_block | hvy_work(...,src_slot))."""
        self._check(self._lib.wgpu_rhs(self._ctx, float(time), src_slot, dst_slot))

    def calculate_time_step(self, time: float) -> float:
        """calculate_time_step.f90:2"""
        dt = C.c_double()
        self._check(self._lib.wgpu_calculate_time_step(self._ctx, float(time), C.byref(dt)))
        return dt.value

    # ------------------------------------------------------------------ multi-GPU inside the library (NCCL communicator owned by the context)
    comm_rank, comm_world = 0, 1

    def comm_init(self, rank: int, world: int, broadcast=None):
        """Join the library's NCCL communicator (wgpu_comm_unique_id on rank 0 + wgpu_comm_init).  `broadcast(buf: bytearray)` must leave rank
        0's 128 bytes in buf on every rank (a Fortran host: MPI_Bcast); default: torch.distributed."""
        if world < 2:
            return
        buf = C.create_string_buffer(128)
        if rank == 0:
            self._check(self._lib.wgpu_comm_unique_id(buf))
        if broadcast is None:
            import torch
            import torch.distributed as dist
            t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
            dist.broadcast(t, 0)
            raw = bytes(t.cpu().numpy().tobytes())
        else:
            ba = bytearray(buf.raw)
            broadcast(ba)
            raw = bytes(ba)
        self._check(self._lib.wgpu_comm_init(self._ctx, raw, int(rank), int(world)))
        self.comm_rank, self.comm_world = int(rank), int(world)

    def comm_set_counts(self, send_counts, recv_counts, rsend_counts=None, rrecv_counts=None):
        a = [None if v is None else np.ascontiguousarray(v, dtype=np.int32) for v in (send_counts, recv_counts, rsend_counts, rrecv_counts)]
        self._check(self._lib.wgpu_comm_set_counts(self._ctx, *[None if v is None else _i32(v) for v in a]))

    def comm_set_transport(self, peer_stores: bool):
        """face-patch exchange inside wgpu_rk_steps: peer stores over NVLink (CUDA IPC, default) or grouped ncclSend / ncclRecv; call before the
        exchange is attached (wgpu_comm_set_transport)"""
        self._check(self._lib.wgpu_comm_set_transport(self._ctx, int(bool(peer_stores))))

    def comm_transport(self) -> str:
        return "peer stores" if self._lib.wgpu_comm_transport(self._ctx) else "nccl"

    def RungeKuttaSteps(self, time: float, n_steps: int = 1):
        """n_steps of RungeKuttaGeneric back to back with time, dt and the divergence flag resident on the device (wgpu_rk_steps: the
        N_dt_per_grid loop of performance_test.f90); across ranks if the context has a communicator.  Returns (time after, last dt)."""
        t, dt = C.c_double(), C.c_double()
        self._check(self._lib.wgpu_rk_steps(self._ctx, float(time), int(n_steps), C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def filter_wrapper(self, filter_type: str, filter_component=None, only_maxlevel: bool = False, all_except_maxlevel: bool = False):
        """filter_wrapper (LIB/TIME/filter_wrapper.f90): explicit binomial filter of the resident hvy_block (wgpu_filter)"""
        fc = None if filter_component is None else _i32(np.ascontiguousarray(filter_component, dtype=np.int32))
        self._check(self._lib.wgpu_filter(self._ctx, filter_type.encode(), fc, int(bool(only_maxlevel)), int(bool(all_except_maxlevel))))

    def krylov_time_stepper(self, time: float, iteration: int, M_krylov: int = 12, dynamic: bool = False, err_threshold: float = 1.0e-3):
        """krylov_time_stepper (LIB/TIME/krylov.f90:1) on the device (wgpu_krylov_step); returns (dt, M_iter, err) -- the last two are what
        the reference appends to krylov_err.t"""
        dt, err, M = C.c_double(), C.c_double(), C.c_int32()
        self._check(self._lib.wgpu_krylov_step(self._ctx, float(time), int(iteration), int(M_krylov), int(bool(dynamic)), float(err_threshold),
                                               C.byref(dt), C.byref(M), C.byref(err)))
        self.krylov_log = (float(dt.value), int(M.value), float(err.value))
        return self.krylov_log

    def RungeKuttaChebychev(self, time: float, iteration: int, mu, mu_tilde, nu, gamma_tilde, c) -> float:
        """RungeKuttaChebychev (runge_kutta_chebychev.f90:6) with the host's coefficient rows of length s (wgpu_rkc_step); returns dt"""
        arr = [np.ascontiguousarray(v, dtype=np.float64) for v in (mu, mu_tilde, nu, gamma_tilde, c)]
        s = len(arr[0])
        assert all(len(a) == s for a in arr)
        dt = C.c_double()
        dp = C.POINTER(C.c_double)
        self._check(self._lib.wgpu_rkc_step(self._ctx, float(time), int(iteration), s, *[a.ctypes.data_as(dp) for a in arr], C.byref(dt)))
        return dt.value

    def RungeKuttaGeneric(self, time: float, iteration: int = 0) -> float:
        """runge_kutta_generic.f90:1 -- advances hvy_block by one step on the device, returns dt."""
        dt = C.c_double()
        self._check(self._lib.wgpu_rk_step(self._ctx, float(time), int(iteration), C.byref(dt)))
        return dt.value

    # ------------------------------------------------------------------ wavelet side (adapt_tree's heavy-data loops)
    EPS_NORMS = {"Linfty": 0, "L1": 1, "L2": 2, "H1": 3}

    def setup_wavelet(self, name: Optional[str] = None):
        """setup_wavelet (module_wavelets.f90:1031): returns the default (g, g_rhs) of the wavelet."""
        g, grhs = C.c_int32(), C.c_int32()
        self._check(self._lib.wgpu_set_wavelet(self._ctx, (name or self.params.wavelet).encode(), C.byref(g), C.byref(grhs)))
        return g.value, grhs.value

    def waveletDecomposition_tree(self, src=(HVY_BLOCK, 0), dst=(HVY_TMP, 0)):
        """sync_ghosts_tree + waveletDecomposition_optimized_block on every block (adapt_tree.f90:403-446)."""
        self._check(self._lib.wgpu_fwt(self._ctx, src[0], src[1], dst[0], dst[1]))

    def waveletReconstruction_tree(self, src=(HVY_TMP, 0), dst=(HVY_BLOCK, 0)):
        """sync of SC/WC + waveletReconstruction_optimized_block on every block (adapt_tree.f90:813-843)."""
        self._check(self._lib.wgpu_iwt(self._ctx, src[0], src[1], dst[0], dst[1]))

    def coarse_extension_modify(self, wd=(HVY_TMP, 0), orig=(HVY_BLOCK, 0), clear_wc: bool = True, copy_sc: bool = True):
        """coarse_extension_modify(CE_case="tree") (LIB/MPI/reconstruction_step.f90:3) on the interiors of a decomposed array."""
        self._check(self._lib.wgpu_coarse_extension(self._ctx, wd[0], wd[1], orig[0], orig[1], int(clear_wc), int(copy_sc)))

    def waveletReconstruction_CE(self, wd=(HVY_WORK, 2), coarse=(HVY_BLOCK, 0), dst=(HVY_BLOCK, 0)):
        """sync_SCWC_from_MC + coarse_extension_modify + waveletReconstruction_optimized_block on the active blocks
        (wavelet_reconstruct_full_tree_CEoptimized, adapt_tree.f90:686-987)"""
        self._check(self._lib.wgpu_iwt_ce(self._ctx, wd[0], wd[1], coarse[0], coarse[1], dst[0], dst[1]))

    def patch_details(self, hvy_ids, dirs, array=(HVY_WORK, 2), eps_norm: str = "Linfty", level_ref: int = 0) -> np.ndarray:
        """details of decomposed blocks inside the strips facing given neighbour directions (addSecurityZone_CE_tree), renormalised for
        eps_norm: [n, n_eqn]"""
        ids = np.ascontiguousarray(hvy_ids, dtype=np.int32)
        dd = np.ascontiguousarray(dirs, dtype=np.int32)
        out = np.zeros((len(ids), self.params.n_eqn))
        self._check(self._lib.wgpu_patch_details_norm(self._ctx, array[0], array[1], self.EPS_NORMS[eps_norm], int(level_ref), len(ids), _i32(ids), _i32(dd),
                                                      out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def wavelet_filter_width(self) -> int:
        """max |tap index| of the decomposition low-pass filter HD of params.wavelet (setup_wavelet: 0 for unlifted CDFX0)"""
        w = self.params.wavelet
        X, Y = int(w[3]), int(w[4])
        return (X - 1) + (Y - 1) if Y > 0 else 0

    def move_blocks(self, src_hvy, dst_hvy):
        """block_xfer on one rank (wgpu_move_blocks): hvy_block(dst[i]) = hvy_block(src[i]) for all i at once"""
        s = np.ascontiguousarray(src_hvy, dtype=np.int32)
        d = np.ascontiguousarray(dst_hvy, dtype=np.int32)
        self._check(self._lib.wgpu_move_blocks(self._ctx, len(s), _i32(s), _i32(d)))

    def coarsen_blocks(self, mothers, daughters, decomposed=(HVY_WORK, 2)):
        """sync_D2M for an explicit list (executeCoarsening_tree.f90:125): hvy_block(mother)[octant] = scaling coefficients of the decomposed
        daughters (2^dim per mother, treecode digit order)"""
        mo = np.ascontiguousarray(mothers, dtype=np.int32)
        da = np.ascontiguousarray(daughters, dtype=np.int32)
        self._check(self._lib.wgpu_coarsen(self._ctx, len(mo), _i32(mo), _i32(da), decomposed[0], decomposed[1]))

    def componentWiseNorm_tree(self, array=(HVY_BLOCK, 0), norm: str = "Linfty") -> np.ndarray:
        out = np.zeros(self.params.n_eqn)
        self._check(self._lib.wgpu_norm(self._ctx, array[0], array[1], self.EPS_NORMS[norm], out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def threshold_tree(self, array=(HVY_TMP, 0), eps: Optional[float] = None, norm=None, eps_norm: str = "Linfty", thresh_comp=None,
                       level_ref: int = 0, want_detail: bool = False):
        """threshold_block on every active block of a decomposed array: refinement_status[n_active] (-1 coarsen / 0 keep)."""
        nc = self.params.n_eqn
        tc = np.ascontiguousarray(np.ones(nc) if thresh_comp is None else thresh_comp, dtype=np.int32)
        e = np.full(nc, self.params.eps if eps is None else eps, dtype=np.float64)
        nrm = None if norm is None else np.ascontiguousarray(norm, dtype=np.float64)
        st = np.zeros(len(self.hvy_active), dtype=np.int32)
        det = np.zeros((len(self.hvy_active), nc)) if want_detail else None
        dp = C.POINTER(C.c_double)
        self._check(self._lib.wgpu_threshold(self._ctx, array[0], array[1], self.EPS_NORMS[eps_norm], level_ref, _i32(tc),
                                             e.ctypes.data_as(dp), None if nrm is None else nrm.ctypes.data_as(dp), _i32(st),
                                             None if det is None else det.ctypes.data_as(dp)))
        return (st, det) if want_detail else st

    # ------------------------------------------------------------------ refinement / coarsening (single rank)
    def refine_tree(self, forest: Forest, refine_flags: Optional[np.ndarray] = None) -> Forest:
        """refine_tree (LIB/MESH/refine_tree.f90:15) with indicator "everywhere" (refine_flags None) or an explicit +1/0 flag per
        active block (the result of refinementIndicator_tree + ensureGradedness_tree, host logic): respectJmaxJmin_tree drops the
        flag of blocks on Jmax, refinement_execute_tree -> refineBlock runs on the device, the new grid is ordered along the
        space-filling curve (balanceLoad_tree) and uploaded.  Returns the new forest."""
        try:
            new, mo, da, ks, kd = forest.refine(refine_flags, max_blocks=self.max_blocks)
        except MemoryError as e:
            raise WabbitAbort(1909181740, f"refine_tree: {e}")                  # error_OOM
        self._check(self._lib.wgpu_refine(self._ctx, len(mo), _i32(mo), _i32(da), len(ks), _i32(ks), _i32(kd)))
        self.set_forest(new)
        return new

    def executeCoarsening_tree(self, forest: Forest, coarsen_flags: np.ndarray, decomposed=(HVY_WORK, 2), Jmin: int = 0) -> Forest:
        """executeCoarsening_tree (LIB/MESH/executeCoarsening_tree.f90): sister groups whose 2^dim members all carry -1 (and pass the
        completeness / gradedness rules of the host logic) are merged into their mother, whose octants are the scaling coefficients
        of the decomposed daughters (array `decomposed`, the output of waveletDecomposition_tree; not hvy_tmp, which serves as the
        second buffer of the block move).  Blocks that stay move to their position along the space-filling curve of the new grid
        first; the mothers are then assembled from the daughters' OLD slots of the decomposed array, which the move does not touch."""
        new, st, mo, da, ks, kd = forest.coarsen(coarsen_flags, Jmin, max_blocks=self.max_blocks)
        self._check(self._lib.wgpu_move_blocks(self._ctx, len(ks), _i32(ks), _i32(kd)))
        self._check(self._lib.wgpu_coarsen(self._ctx, len(mo), _i32(mo), _i32(da), decomposed[0], decomposed[1]))
        self.set_forest(new)
        return new

    def adapt_tree(self, forest: Forest, eps: Optional[float] = None, eps_normalized: bool = True, eps_norm: str = "Linfty",
                   Jmin: int = 1, force_maxlevel_dealiasing: bool = False, thresh_comp=None, useSecurityZone: Optional[bool] = None,
                   full_tree: Optional[bool] = None, mask_keeps=None):
        """adapt_tree (LIB/MESH/adapt_tree.f90:11) with indicator "threshold-state-vector".  full_tree (default for lifted wavelets): the
        reference's full-tree algorithm, with the coarse extension and the security zone for lifted wavelets (wabbit_b200/fulltree.py).
        Otherwise (default for unlifted wavelets CDFX0, which have no coarse extension and no security zone) one coarsening sweep:
        componentWiseNorm_tree -> ghost synchronisation + wavelet
        decomposition of every leaf -> threshold_block flags (device), then completeness / gradedness (host light data) and
        executeCoarsening (device).  The reference's current adapt_tree decomposes the full tree and can remove several levels
        in one call; this driver removes one level per call (call it again to go further) -- the per-block arithmetic is the same.
        Returns (new forest, number of blocks before, after)."""
        w = self.params.wavelet
        lifted = not (len(w) == 5 and w[4] == "0")
        use_ce = lifted if self.params.useCoarseExtension < 0 else bool(self.params.useCoarseExtension)
        self.refinement_status = None
        if use_ce if full_tree is None else full_tree:
            # the reference's full-tree algorithm (wabbit_b200/fulltree.py; for lifted wavelets with the coarse extension), which can remove
            # several levels in one call; the security zone is on unless params.useSecurityZone = 0 (the reference's default)
            from .fulltree import FullTree
            import os
            import time as _time
            _tm = {} if os.environ.get("WABBIT_FT_TIMING") else None
            _t0 = _time.perf_counter()

            def _lap(name):
                nonlocal _t0
                if _tm is not None:
                    self.synchronize()
                    _tm[name] = round((_time.perf_counter() - _t0) * 1e3, 2)
                    _t0 = _time.perf_counter()
            if eps_norm != "Linfty" and eps_normalized:
                norm_l = self.componentWiseNorm_tree((HVY_BLOCK, 0), eps_norm)
            else:
                norm_l = self.componentWiseNorm_tree((HVY_BLOCK, 0), "Linfty") if eps_normalized else None
            if norm_l is not None:
                norm_l = threshold_norm(norm_l, thresh_comp)
            n0 = forest.n_blocks
            _lap("norm")
            ft = FullTree(self, forest, Jmin=Jmin)
            _lap("FullTree init")
            sz = (lifted if self.params.useSecurityZone < 0 else bool(self.params.useSecurityZone)) if useSecurityZone is None else bool(useSecurityZone)
            new, _info = ft.adapt(eps=self.params.eps if eps is None else eps, norm=norm_l, eps_norm=eps_norm, thresh_comp=thresh_comp,
                                  force_maxlevel_dealiasing=force_maxlevel_dealiasing, want_info=False, use_security_zone=sz,
                                  mask_keeps=mask_keeps)
            self.refinement_status = ft.leaf_status          # lgt_block(:, IDX_REFINE_STS) after adapt_tree, in the order of new.active(0)
            _lap("ft.adapt")
            if _tm is not None:
                print("adapt_tree phases [ms]:", _tm, flush=True)
            return new, n0, new.n_blocks
        if mask_keeps is not None:
            raise ValueError("adapt_tree: threshold_mask needs the full-tree algorithm")
        hvy, lvl, _, _ = forest.active(0)
        norm = None
        if eps_normalized:
            norm = self.componentWiseNorm_tree((HVY_BLOCK, 0), eps_norm)
            norm[norm <= 1.0e-9] = 1.0                                        # coarseningIndicator_tree.f90:165-167
        self.waveletDecomposition_tree((HVY_BLOCK, 0), (HVY_WORK, 2))
        st = self.threshold_tree((HVY_WORK, 2), eps=eps, norm=norm, eps_norm=eps_norm, thresh_comp=thresh_comp, level_ref=forest.Jmax)
        if force_maxlevel_dealiasing:
            st = np.where(lvl == forest.Jmax, -1, st)                         # coarseningIndicator_tree.f90:307-309
        n0 = forest.n_blocks
        if not (st == -1).any():
            return forest, n0, n0
        new = self.executeCoarsening_tree(forest, st, decomposed=(HVY_WORK, 2), Jmin=Jmin)
        return new, n0, new.n_blocks

    def timeStep_tree(self, time: float, iteration: int):
        """timeStep_tree.f90:1-60 -- dispatch on time_step_method, then filter_wrapper every filter_freq iterations and before data are saved (main.f90:368-374); returns
        (time+dt, iteration+1, dt)."""
        p = self.params
        method = p.time_step_method.strip().lower()
        if method == "rungekuttageneric":
            dt = self.RungeKuttaGeneric(time, iteration)
        elif method == "rungekuttachebychev":
            dt = self.RungeKuttaChebychev(time, iteration, *p.rkc_coefficients())
        elif method == "krylov":
            dt = self.krylov_time_stepper(time, iteration, p.M_krylov, p.krylov_subspace_dimension.strip().lower() == "dynamic", p.krylov_err_threshold)[0]
        else:
            raise WabbitAbort(19101816, "time_step_method is unkown: " + p.time_step_method)
        iteration += 1
        if p.filter_type != "no_filter" and ((p.filter_freq > 0 and iteration % p.filter_freq == 0) or p.is_it_time_to_save_data(time + dt, iteration)):
            self.filter_wrapper(p.filter_type, p.filter_component or None, p.filter_only_maxlevel, p.filter_all_except_maxlevel)
        return time + dt, iteration, dt
