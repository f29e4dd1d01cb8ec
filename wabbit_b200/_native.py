"""ctypes bindings of the two native libraries.  No CPU fallback: if libwabbit_gpu.so is missing or no
CUDA device is present, creating a solver raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _build

WGPU_MAX_STAGES = 8

_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_dp = C.POINTER(C.c_double)


class WgpuConfig(C.Structure):
    """Mirror of `wgpu_config` (include/wabbit_gpu.h)."""
    _fields_ = [
        ("dim", C.c_int32), ("Bs", C.c_int32 * 3), ("g", C.c_int32), ("g_rhs", C.c_int32), ("n_eqn", C.c_int32),
        ("n_mask", C.c_int32), ("max_blocks", C.c_int32), ("Jmax", C.c_int32), ("periodic", C.c_int32 * 3),
        ("fd", C.c_int32), ("skew_symmetry", C.c_int32), ("penalization", C.c_int32), ("use_sponge", C.c_int32),
        ("n_stages", C.c_int32), ("write_method_fixed_time", C.c_int32), ("device", C.c_int32),
        ("domain", C.c_double * 3), ("c0", C.c_double), ("nu", C.c_double), ("gamma_p", C.c_double),
        ("C_eta", C.c_double), ("C_sponge", C.c_double), ("u_mean_set", C.c_double * 3),
        ("CFL", C.c_double), ("CFL_eta", C.c_double), ("CFL_nu", C.c_double),
        ("dt_fixed", C.c_double), ("dt_max", C.c_double), ("time_max", C.c_double),
        ("write_time", C.c_double), ("write_time_first", C.c_double), ("tsave_stats", C.c_double),
        ("butcher", C.c_double * ((WGPU_MAX_STAGES + 1) * (WGPU_MAX_STAGES + 1))),
    ]


# every symbol include/wabbit_gpu.h declares: name -> (restype, argtypes)
GPU_SYMBOLS = {
    "wgpu_create": (C.c_int32, [C.POINTER(WgpuConfig), C.POINTER(C.c_void_p)]),
    "wgpu_destroy": (C.c_int32, [C.c_void_p]),
    "wgpu_last_error": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int32]),
    "wgpu_set_stream": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "wgpu_synchronize": (C.c_int32, [C.c_void_p]),
    "wgpu_set_topology": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, _i32p, C.c_int32, C.c_int32]),
    "wgpu_set_treecodes": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, _i64p]),
    "wgpu_upload": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, _i32p, C.c_int32, C.c_void_p, C.c_int32]),
    "wgpu_download": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, _i32p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32]),
    "wgpu_set_transfer_mode": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "wgpu_filter": (C.c_int32, [C.c_void_p, C.c_char_p, _i32p, C.c_int32, C.c_int32]),
    "wgpu_rkc_step": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, _dp, _dp, _dp, _dp, _dp, _dp]),
    "wgpu_krylov_step": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_double, _dp, C.POINTER(C.c_int32), _dp]),
    "wgpu_expm_pade": (C.c_int32, [_dp, C.c_int32, _dp]),
    "wgpu_create_mask": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "wgpu_statistics": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, _dp]),
    "wgpu_sync_ghosts": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "wgpu_set_ghost_filter": (C.c_int32, [C.c_void_p, C.c_int32]),
    "wgpu_set_halo": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, _i32p, C.c_int32, _i32p, C.c_void_p]),
    "wgpu_pack_blocks": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "wgpu_set_halo_restrict": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, C.c_int32, _i32p, C.c_void_p]),
    "wgpu_restrict_pack": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "wgpu_restrict_halo_pointer": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), _i64p]),
    "wgpu_gather_blocks": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p, C.c_void_p]),
    "wgpu_scatter_blocks": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p, C.c_void_p]),
    "wgpu_halo_pointer": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), _i64p]),
    "wgpu_rk_stage_halo_pointer": (C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), _i64p]),
    "wgpu_set_mask_sphere": (C.c_int32, [C.c_void_p, C.c_int32, _dp, _dp, C.c_double, C.c_double]),
    "wgpu_rhs": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, C.c_int32]),
    "wgpu_calculate_time_step": (C.c_int32, [C.c_void_p, C.c_double, _dp]),
    "wgpu_rk_step": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, _dp]),
    "wgpu_set_wavelet": (C.c_int32, [C.c_void_p, C.c_char_p, _i32p, _i32p]),
    "wgpu_fwt": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "wgpu_iwt": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "wgpu_iwt_ce": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "wgpu_set_grid": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, _i64p, C.c_int32, _i32p]),
    "wgpu_set_active": (C.c_int32, [C.c_void_p, C.c_int32, _i32p]),
    "wgpu_topology_tables": (C.c_int32, [C.c_void_p, _i32p, _i32p, _i32p]),
    "wgpu_topology_list": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, _i32p, _i32p]),
    "wgpu_comm_unique_id": (C.c_int32, [C.c_char_p]),
    "wgpu_comm_init": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_int32, C.c_int32]),
    "wgpu_comm_destroy": (C.c_int32, [C.c_void_p]),
    "wgpu_comm_info": (C.c_int32, [C.c_void_p, _i32p, _i32p]),
    "wgpu_comm_set_counts": (C.c_int32, [C.c_void_p, _i32p, _i32p, _i32p, _i32p]),
    "wgpu_comm_set_transport": (C.c_int32, [C.c_void_p, C.c_int32]),
    "wgpu_comm_transport": (C.c_int32, [C.c_void_p]),
    "wgpu_rk_steps": (C.c_int32, [C.c_void_p, C.c_double, C.c_int32, _dp, _dp]),
    "wgpu_exchange_array": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "wgpu_ship_blocks": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _i32p, C.c_int32, _i32p, _i32p]),
    "wgpu_comm_allreduce": (C.c_int32, [C.c_void_p, _dp, C.c_int32, C.c_int32]),
    "wgpu_comm_allgatherv_i32": (C.c_int32, [C.c_void_p, _i32p, _i32p, _i32p]),
    "wgpu_norm": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _dp]),
    "wgpu_threshold": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p, _dp, _dp, _i32p, _dp]),
    "wgpu_patch_details": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _dp]),
    "wgpu_patch_details_norm": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _dp]),
    "wgpu_coarse_extension": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "wgpu_refine": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, C.c_int32, _i32p, _i32p]),
    "wgpu_coarsen": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, C.c_int32, C.c_int32]),
    "wgpu_move_blocks": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p]),
    "wgpu_patch_doubles": (C.c_int64, [C.c_void_p]),
    "wgpu_set_exchange": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, C.c_void_p, C.c_int32, _i32p, _i32p, C.c_void_p]),
    "wgpu_pack_halo": (C.c_int32, [C.c_void_p, C.c_int32]),
    "wgpu_rk_begin": (C.c_int32, [C.c_void_p, C.c_double]),
    "wgpu_dtmin_pointer": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "wgpu_rk_dt": (C.c_int32, [C.c_void_p, C.c_double]),
    "wgpu_rk_stage": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32]),
    "wgpu_rk_end": (C.c_int32, [C.c_void_p, _dp]),
    "wgpu_block_count": (C.c_int32, [C.c_void_p, C.c_int32]),
    "wgpu_profile": (C.c_int32, [C.c_void_p, C.c_int32]),
    "wgpu_profile_read": (C.c_int32, [C.c_void_p, _i32p, _dp]),
    "wgpu_launch_count": (C.c_int64, [C.c_void_p]),
    "wgpu_device_bytes": (C.c_int64, [C.c_void_p]),
    "wgpu_device_pointer": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), _i64p]),
}

HOST_SYMBOLS = {
    "whost_create_uniform": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p, C.POINTER(C.c_void_p)]),
    "whost_create_from_blocks": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p, C.c_int32, _i32p, _i32p,
                                             C.POINTER(C.c_void_p)]),
    "whost_destroy": (C.c_int32, [C.c_void_p]),
    "whost_n_blocks": (C.c_int32, [C.c_void_p]),
    "whost_n_active": (C.c_int32, [C.c_void_p, C.c_int32]),
    "whost_get_active": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, _i32p, _i64p]),
    "whost_get_neighbors": (C.c_int32, [C.c_void_p, C.c_int32, _i32p]),
    "whost_neighbors_ptr": (_i32p, [C.c_void_p, C.c_int32]),
    "whost_is_uniform": (C.c_int32, [C.c_void_p]),
    "whost_halo_plan": (C.c_int32, [C.c_void_p, C.c_int32, _i32p, _i32p, _i32p, _i64p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "whost_refine": (C.c_int32, [C.c_void_p, _i32p, C.c_int32, C.POINTER(C.c_void_p), _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "whost_coarsen": (C.c_int32, [C.c_void_p, _i32p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "whost_refine_global": (C.c_int32, [C.c_void_p, _i32p, C.c_int32, C.POINTER(C.c_void_p), _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "whost_coarsen_global": (C.c_int32, [C.c_void_p, _i32p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), _i32p, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "whost_ft_build": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, C.c_int32, _i32p, _i32p, _i32p, _i32p]),
    "whost_encode_many": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _i64p]),
    "whost_ft_tables": (C.c_int32, [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "whost_ft_rows": (C.c_int32, [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int64, _i32p]),
    "whost_ft_decide": (C.c_int32, [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _i32p, C.c_int32, _i32p]),
    "whost_ft_security_pairs": (C.c_int32, [C.c_int32, C.c_int32, _i32p, _i32p, C.POINTER(C.c_uint8), C.c_int32, _i32p, _i32p]),
    "whost_encode": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, _i32p]),
    "whost_decode": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _i32p]),
    "whost_sfc_key": (C.c_uint64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _i32p]),
}

_gpu = None
_host = None


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)   # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def gpu_lib() -> C.CDLL:
    """Load libwabbit_gpu.so (built in-tree).  Loading needs libcudart only, not a GPU."""
    global _gpu
    if _gpu is None:
        path = _build.GPU_LIB
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -m wabbit_b200._build` (no CPU fallback exists)")
        _gpu = _bind(C.CDLL(path), GPU_SYMBOLS)
    return _gpu


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        path = _build.build_host()        # rebuilds when missing OR older than its sources / headers (the .so is not tracked by git)
        _host = _bind(C.CDLL(path), HOST_SYMBOLS)
    return _host
