#!/usr/bin/env python
"""Benchmark of the WABBIT block hot path on B200: 3-D ACM RK4 block-updates/s.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun, one rank per GPU)
  python bench.py --impl reference ...                    the reference's CPU algorithm on the host cores

A step = one RungeKuttaGeneric time step (4 stages of RHS_3D_acm + ghost synchronisation + stage updates + dt)
over every block of a periodic, equidistant 3-D Taylor-Green grid (BASELINE.json configs[1]).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "3D ACM RK4 block-updates/sec"
UNIT = "block-updates/s"
TWO_PI = 6.283185307179586


def workload_name(a):
    return f"taylor_green_3d_periodic_uniform_J{a.level}_Bs{a.bs}_FD4_skew_RK4_CDF40"


def algorithmic_bytes_per_block_update(Bs: int, g_rhs: int = 2, nc: int = 4, n_mask: int = 0) -> int:
    """SURVEY.md 8(d) / BASELINE.md 3:  B_rk4 = 8*[23*nc*N + 8*nc*(box_r - N) + 4*n_mask*N]."""
    N = Bs ** 3
    box = (Bs + 2 * g_rhs) ** 3
    return 8 * (23 * nc * N + 8 * nc * (box - N) + 4 * n_mask * N)


def make_params(a):
    from wabbit_b200 import Params
    return Params(dim=3, domain=(TWO_PI,) * 3, Bs=(a.bs,) * 3, wavelet="CDF40", g=3, g_rhs=2, n_eqn=4, Jmax=a.level,
                  discretization="FD_4th_central", skew_symmetry=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0,
                  u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9).finalize()


def taylor_green_host(p, ixyz, level, out):
    """inicond 'taylor-green-vanRees2011' (inicond_ACM.f90:371-389) for the blocks listed, into out[k] (ghosted)."""
    import torch
    g, Bs = p.g, p.Bs[0]
    n = Bs + 2 * g
    t_out = torch.from_numpy(out)
    lv = torch.from_numpy(level.astype(np.float64))
    dx = (2.0 ** (-lv)) * p.domain[0] / float(Bs)                      # [nb]
    idx = torch.arange(n, dtype=torch.float64) - g
    chunk = 2048
    for s in range(0, len(level), chunk):
        e = min(s + chunk, len(level))
        d = dx[s:e, None]
        x0 = torch.from_numpy((ixyz[s:e] * Bs).astype(np.float64)) * d  # [m,3]
        X = (idx[None, :] * d + x0[:, 0:1])[:, None, None, :]
        Y = (idx[None, :] * d + x0[:, 1:2])[:, None, :, None]
        Z = (idx[None, :] * d + x0[:, 2:3])[:, :, None, None]
        t_out[s:e, 0] = torch.sin(X) * torch.cos(Y) * torch.cos(Z)
        t_out[s:e, 1] = -torch.cos(X) * torch.sin(Y) * torch.cos(Z)
        t_out[s:e, 2] = 0.0
        t_out[s:e, 3] = (torch.cos(2.0 * X) + torch.cos(2.0 * Y)) * (torch.cos(2.0 * Z) + 2.0) / 16.0


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        self.marks = {}

    def mark(self, name):
        self.f.flush()
        self.marks[name] = os.path.getsize(self.f.name)

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.f.flush()
        with open(self.f.name) as f:
            data = f.read()
        os.unlink(self.f.name)
        lo, hi = self.marks.get("t0", 0), self.marks.get("t1", len(data))
        rows = [r for r in data[lo:hi].splitlines() if r.count(",") >= 8]
        where = "timed_region"
        if not rows:
            rows = [r for r in data.splitlines() if r.count(",") >= 8]
            where = "whole_run"
        sm, smax, reasons = [], 0.0, set()
        for r in rows:
            c = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(c[1]))
                smax = max(smax, float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm), "sampled": where}


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_run(a, level: int, steps: int, warmup: int):
    """The reference's algorithm on the host cores (oracle C restatement, -O3 -march=native, OpenMP: one worker
    per core over contiguous SFC chunks of blocks).  Returns (block_updates_per_s, ms_per_step, cores, nblocks)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core
    po = O.Params(dim=3, Bs=(a.bs,) * 3, g=3, g_rhs=2, domain=(TWO_PI,) * 3, Jmax=level, discretization="FD_4th_central",
                  skew=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0, u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9)
    grid = O.uniform_grid(level)
    u = O.alloc(grid, po)
    O.inicond_taylor_green(grid, po, u)
    work = np.zeros((5,) + u.shape)
    nbr, dxb = O.nbr_table(grid), O.dx_table(grid, po)
    t = 0.0
    for _ in range(warmup):
        t += O.rk_step_c(grid, po, u, work, t, nbr, dxb, fast=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        t += O.rk_step_c(grid, po, u, work, t, nbr, dxb, fast=True)
    el = time.perf_counter() - t0
    return grid.n * steps / el, 1e3 * el / steps, cores, grid.n


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lvl = min(a.level, a.cpu_level)
    steps = max(1, min(a.steps, 6))
    warm = 1
    v, ms, cores, nb = cpu_run(a, lvl, steps, warm)
    sample = f"{steps} RK4 steps on {nb} blocks (level {lvl}, Bs={a.bs}) of the {workload_name(a)} workload"
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "blocks": nb, "Bs": a.bs, "note": "bounded sample of the workload on the host cores"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; oracle C restatement of the reference Fortran (no Fortran compiler in the image), "
                                            "-O3 -march=native, OpenMP static chunks of the SFC-ordered block list"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from wabbit_b200 import Forest, WabbitGPU

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    p = make_params(a)
    forest = Forest.uniform(3, a.level, block_dist="sfc_hilbert", n_ranks=world)
    nb_global = forest.n_blocks
    hvy, lvl, ixyz, _ = forest.active(rank)
    nb_local = len(hvy)
    stream = torch.cuda.current_stream()
    sol = WabbitGPU(p, max_blocks=forest.max_blocks, device=local, stream=stream.cuda_stream)
    if world > 1:
        from wabbit_b200.multi import attach_exchange
        sol.comm_init(rank, world)          # the library's own NCCL communicator: pack -> ncclSend/Recv -> stage kernels inside wgpu_rk_steps
        attach_exchange(sol, forest, rank, world)
    else:
        sol.set_forest(forest, rank)
    steps_fn = sol.stepper.steps if world > 1 else sol.RungeKuttaSteps

    # host state in the reference layout hvy_block(nx,ny,nz,4,number_blocks), pinned
    shape = sol.host_shape()
    host = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    h_np = host.numpy()
    taylor_green_host(p, ixyz, lvl, h_np)
    ids = hvy
    sol.upload_ptr(host.data_ptr(), shape[1], hvy_ids=ids)
    sol.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    t = 0.0
    it = 0
    t, _dt = steps_fn(t, a.warmup)
    # ---------------- timed region: device-resident state; the K steps are issued back to back (wgpu_rk_steps: time, dt and the
    # divergence flag stay on the device, one host read-back after the last step -- the N_dt_per_grid loop of performance_test.f90)
    sol.profile(True)
    barrier()
    if sampler:
        sampler.mark("t0")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = sol.launch_count
    ev0.record(stream)
    t, _dt = steps_fn(t, a.steps)
    ev1.record(stream)
    barrier()
    it += a.warmup + a.steps
    if sampler:
        sampler.mark("t1")
    launches = sol.launch_count - n0
    ms = ev0.elapsed_time(ev1)
    n_stage, stage_ms = sol.profile_read()
    sol.profile(False)
    tm = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    value = nb_global * a.steps / (ms * 1e-3)
    # checksum of the state after warmup + steps: Linfty norm per component (MAX over blocks and ranks: independent of the partition, so the
    # line of every GPU count must show the same bits) and the time reached
    sol.setup_wavelet("CDF40")
    cks = sol.componentWiseNorm_tree((0, 0), "Linfty")
    tck = torch.tensor(list(cks), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tck, op=dist.ReduceOp.MAX)
    checksum = {"time": t, "linfty": [float(v).hex() for v in tck.cpu().tolist()]}

    # ---------------- end to end: host buffers in, host buffers out, every step
    e2e_steps = max(1, min(a.steps, a.e2e_steps))
    per_block = int(np.prod(shape[1:])) * 8
    barrier()
    w0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        sol.upload_ptr(host.data_ptr(), shape[1], hvy_ids=ids)
        t, it, _dt = sol.timeStep_tree(t, it)
        sol.download_ptr(host.data_ptr(), shape[1], hvy_ids=ids, g_sync=0)
    e1.record(stream)
    barrier()
    e2e_wall = time.perf_counter() - w0
    te = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, e2e_wall)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = nb_global * e2e_steps / float(te.item())
    finite = bool(np.isfinite(h_np[: min(nb_local, 8)]).all())
    e2e_seq = e2e_value
    e2e_trees = 1
    if world == 1 and a.e2e_trees > 1:
        try:
            e2e_value = e2e_pipelined(a, p, forest, local, sol, host, shape, ids, e2e_steps, a.e2e_trees)
            e2e_trees = a.e2e_trees
        except Exception as e:   # keep the sequential figure
            print(f"bench: pipelined e2e failed ({e}); reporting the sequential figure", file=sys.stderr)

    clocks = sampler.stop() if sampler else None

    # ---------------- secondary figure (BASELINE config 5's kernel): CDF44 decomposition + thresholding of every block
    wavelet = None
    if world == 1 and not a.no_wavelet:
        try:
            wavelet = wavelet_leg(a, sol, nb_local, stream, barrier)
        except Exception as e:   # a secondary figure must not take the headline line down
            wavelet = {"error": str(e)}

    # ---------------- secondary figure (BASELINE config 4's kind of run): the same RK4 steps with volume penalization of a translating
    # sphere, the mask evaluated inside the stage kernel at every stage time (wgpu_set_mask_sphere)
    penalized = None
    if world == 1 and not a.no_wavelet:
        try:
            penalized = penalized_leg(a, forest, local, stream, barrier, host, shape, ids)
        except Exception as e:
            penalized = {"error": repr(e)}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "B200_PROFILING.md fallback 6650 GB/s (of fallback)"
        B = algorithmic_bytes_per_block_update(a.bs)
        # dominant kernel = stage_kernel, 4 launches per step, each processing every local block once:
        # algorithmic bytes per launch = (B_rk4 / 4) * blocks_per_gpu
        # (multi-GPU runs split a stage into an interior and a boundary launch; their durations are summed)
        n_stages = p.n_stages
        avg_launch_s = (stage_ms / (a.steps * n_stages)) * 1e-3
        if world > 1:   # interior and boundary launches overlap in time on two streams: take the stage's share of the step instead (an
            avg_launch_s = (ms / a.steps / n_stages) * 1e-3        # upper bound: it includes whatever of the exchange is exposed)
        achieved = (B / float(n_stages)) * nb_local / avg_launch_s / 1e9 if n_stage else None
        traffic = None
        tr_path = os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")
        if os.path.exists(tr_path):
            try:
                traffic = json.load(open(tr_path)).get("dram_bytes_per_block_per_launch") * nb_local
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "blocks": nb_global, "blocks_per_gpu": nb_local, "Bs": a.bs, "nc": 4, "g_rhs": 2,
                       "host_layout_g": p.g, "block_dist": "sfc_hilbert", "parallelism": f"sfc-partition x{world}",
                       "l2": "inputs larger than L2 (state array %.0f MB per GPU)" % (nb_local * 4 * a.bs ** 3 * 8 / 1e6),
                       "stepping": "K steps issued back to back inside the library (wgpu_rk_steps), time / dt device-resident, one host read-back",
                       "transport": (("pack kernel stores the face patches straight into the receivers' pools over NVLink (CUDA IPC) + a flag per peer"
                                      if sol.comm_transport() == "peer stores" else "NCCL send/recv inside libwabbit_gpu.so on its own communicator")
                                     + ", overlapped with the interior blocks" if world > 1 else "none"),
                       "finite": finite, "checksum": checksum},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * a.bs ** 3 * 8 * nb_local, "d2h_bytes_per_step": 4 * a.bs ** 3 * 8 * nb_local,
                    "steps": e2e_steps * e2e_trees, "trees_in_flight": e2e_trees, "sequential_value": e2e_seq,
                    "transfer": "copy engines: interior plane spans by cudaMemcpy3DAsync (1.35x the interior bytes at Bs=16, g=3) + layout kernels, "
                                "double-buffered in 16 MB chunks (wgpu_set_transfer_mode default)",
                    "note": "wgpu_upload(host hvy_block) + wgpu_rk_step + wgpu_download(host hvy_block, g_sync=0) per step: "
                            "RungeKuttaGeneric's contract -- interiors of the page-locked Fortran-layout host array in, interiors out "
                            "(runge_kutta_generic.f90:136-154).  value: `trees_in_flight` independent trees of the forest (each "
                            "with its own host array, device context and stream, driven by its own host thread), every step of every tree "
                            "doing its own upload and download; transfers of one direction take turns, so one tree's download runs against "
                            "another's upload on the full-duplex link; sequential_value: one tree, upload -> step -> download back to back"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic, "kernel": "stage_kernel<FD4,skew,Bs16>" + (" (interior + boundary launch per stage, step time / 4)" if world > 1 else ""), "launches_timed": n_stage,
                         "avg_launch_ms": avg_launch_s * 1e3, "algorithmic_bytes_per_launch": (B / float(n_stages)) * nb_local, "algorithmic_bytes_per_block_update": B, "peak_source": peak_src},
        }
        if wavelet is not None:
            if "value" in wavelet:
                wavelet["roofline"]["peak"] = peak
                wavelet["roofline"]["frac"] = wavelet["roofline"]["achieved"] / peak
            line["wavelet"] = wavelet
        if penalized is not None:
            if "value" in penalized:
                penalized["roofline"]["peak"] = peak
                penalized["roofline"]["frac"] = penalized["roofline"]["achieved"] / peak
            line["penalized"] = penalized
        if world == 1 and not a.no_cpu:
            v, cms, cores, nbc = cpu_run(a, min(a.level, a.cpu_level), a.cpu_steps, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{a.cpu_steps} RK4 steps on {nbc} blocks (level {min(a.level, a.cpu_level)}) of the same workload; "
                                              "oracle C restatement, -O3 -march=native, OpenMP"}
    sol.close()
    del host, h_np
    adaptive = None
    adaptive_lifted = None
    if world == 1 and not a.no_adaptive:
        try:
            adaptive = adaptive_leg(a, local, stream)
        except Exception as e:      # a secondary figure must not take the headline line down
            adaptive = {"error": repr(e)}
        try:                        # BASELINE config 3 names CDF4,4: the lifted adapt_tree (full-tree algorithm)
            adaptive_lifted = adaptive_leg(a, local, stream, wavelet="CDF44")
        except Exception as e:
            adaptive_lifted = {"error": repr(e)}
    elif world > 1 and not a.no_adaptive:
        # BASELINE configs 3 / 4 (the north-star target): 3-D adaptive ACM, CDF44, coarsening + refinement every step, all GPUs
        adaptive_lifted = {}
        J0 = a.adaptive_level if a.adaptive_level > 0 else (6 if world >= 8 else 5)     # 8^6 = 262 144 initial blocks (18 GB per resident array and rank at 4 GPUs): from 8 GPUs on
        legs = [("Bs16", 16, J0, False), ("Bs16_sphere", 16, J0, True), ("Bs18", 18, J0, False)]
        if a.adaptive_legs:
            legs = [l for l in legs if l[0] in a.adaptive_legs.split(",")]
        for name, bs, j0, sph in legs:
            adaptive_lifted[name] = adaptive_leg_multi(a, rank, world, local, stream, J0=j0, Jmax=j0 + 1, wavelet="CDF44", bs=bs, sphere_on=sph)
    weak = None
    if world == 8 and not a.no_weak:
        weak = weak_scaling_leg(a, rank, world, local, stream)
    compression = None
    if not a.no_compression:
        idx = list(range(51)) if a.compression_full else (list(range(0, 51, 5)) if world == 1 else list(range(0, 51, 10)))
        compression = {f"J{a.compression_level}": compression_leg(a, rank, world, local, stream, a.compression_level, eps_idx=idx)}
        if a.compression_full or world >= 4:      # ~10^5.4 blocks: the protocol's J = 6 needs the memory of >= 4 GPUs for the full tree
            compression[f"J{a.compression_level + 1}"] = compression_leg(a, rank, world, local, stream, a.compression_level + 1,
                                                                          wavelets=("CDF40", "CDF42", "CDF44") if a.compression_full else ("CDF44",),
                                                                          eps_idx=idx if a.compression_full else [20, 30, 40])
    if rank == 0:
        if weak is not None:
            line["weak_scaling"] = weak
        if compression is not None:
            line["compression"] = compression
        if adaptive is not None:
            line["adaptive"] = adaptive
        if adaptive_lifted is not None:
            line["adaptive_cdf44"] = adaptive_lifted
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_pipelined(a, p, forest, local, sol0, host0, shape, ids, steps, trees):
    """End-to-end throughput with `trees` independent trees in flight (WABBIT's forest holds several trees, module_forestMetaData.f90):
    tree k has its own page-locked host array, device context and stream and is driven by its own host thread; every step of every
    tree uploads its input from the host and downloads its result, exactly as the sequential leg.  Returns block-updates/s."""
    import threading
    import torch
    from wabbit_b200 import WabbitGPU
    sols, hosts = [sol0], [host0]
    for _ in range(1, trees):
        st = torch.cuda.Stream()
        sk = WabbitGPU(p, max_blocks=forest.max_blocks, device=local, stream=st.cuda_stream)
        sk.set_forest(forest, 0)
        hk = torch.empty(shape, dtype=torch.float64, pin_memory=True)
        hk.copy_(host0)
        sols.append(sk)
        hosts.append(hk)
    start = threading.Barrier(trees + 1)
    done = threading.Barrier(trees + 1)
    errs = []

    def worker(k):
        try:
            torch.cuda.set_device(local)
            sk, hk = sols[k], hosts[k]
            t, it = 0.0, 0
            sk.upload_ptr(hk.data_ptr(), shape[1], hvy_ids=ids)       # untimed warm-up step of this tree
            t, it, _ = sk.timeStep_tree(t, it)
            sk.download_ptr(hk.data_ptr(), shape[1], hvy_ids=ids, g_sync=0)
            start.wait()
            for _ in range(steps):
                sk.upload_ptr(hk.data_ptr(), shape[1], hvy_ids=ids)
                t, it, _ = sk.timeStep_tree(t, it)
                sk.download_ptr(hk.data_ptr(), shape[1], hvy_ids=ids, g_sync=0)
        except Exception as e:      # noqa: BLE001
            errs.append(e)
            start.abort()
            done.abort()
            return
        done.wait()

    th = [threading.Thread(target=worker, args=(k,)) for k in range(trees)]
    for x in th:
        x.start()
    try:
        start.wait()                 # every tree has finished its (blocking) warm-up step: the device is idle
        w0 = time.perf_counter()
        done.wait()
        torch.cuda.synchronize()
        el = time.perf_counter() - w0
    except threading.BrokenBarrierError:
        el = None
    for x in th:
        x.join()
    for sk in sols[1:]:
        sk.close()
    if errs or el is None:
        raise RuntimeError(str(errs[0]) if errs else "worker failed")
    return forest.n_blocks * steps * trees / el


def penalized_leg(a, forest, local, stream, barrier, host, shape, ids, steps=20, warmup=3):
    """RK4 on the same equidistant grid with penalization = 1 and the mask of a translating sphere (radius 0.8, smoothing 1.5 dx) computed
    in the stage kernel; algorithmic bytes per block-update are those of the run without a mask (n_mask = 0: nothing is read)."""
    import torch
    from wabbit_b200 import WabbitGPU
    p = make_params(a)
    p.penalization, p.C_eta = True, 1.0e-3
    p.finalize()                                     # hvy_mask is allocated (n_mask = 6) but never written or read
    sol = WabbitGPU(p, max_blocks=forest.max_blocks, device=local, stream=stream.cuda_stream)
    try:
        sol.set_forest(forest, 0)
        dx = p.domain[0] / (2 ** a.level * a.bs)
        sol.set_mask_sphere((3.0, 3.1, 3.2), (0.5, 0.3, -0.2), 0.8, 1.5 * dx)
        sol.upload_ptr(host.data_ptr(), shape[1], hvy_ids=ids)
        t, it = 0.0, 0
        for _ in range(warmup):
            t, it, _dt = sol.timeStep_tree(t, it)
        sol.profile(True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = sol.launch_count
        e0.record(stream)
        for _ in range(steps):
            t, it, _dt = sol.timeStep_tree(t, it)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        n_stage, stage_ms = sol.profile_read()
        sol.profile(False)
        nb = len(ids)
        B = algorithmic_bytes_per_block_update(a.bs)
        avg = stage_ms / max(n_stage, 1) * 1e-3
        chk = np.zeros((1,) + tuple(shape[1:]))
        sol.download(chk, hvy_ids=ids[:1], g_sync=0)
        return {"metric": "3D ACM RK4 block-updates/sec with volume penalization (translating sphere, mask evaluated in the stage kernel)",
                "value": nb * steps / (ms * 1e-3), "unit": UNIT, "steps": steps, "blocks": nb, "gpu_launches": int(sol.launch_count - n0), "dt": _dt,
                "C_eta": p.C_eta, "finite": bool(np.isfinite(chk).all()),
                "roofline": {"bound": "hbm", "kernel": "stage_kernel<FD4,skew,Bs16,sphere>", "achieved": (B / 4.0) * nb / avg / 1e9, "unit": "GB/s",
                             "avg_launch_ms": avg * 1e3, "algorithmic_bytes_per_block_update": B}}
    finally:
        sol.close()


def wavelet_leg(a, sol, nb, stream, barrier, wavelet="CDF44", reps=20):
    """FWT (ghost synchronisation fused) + detail norms + refinement flags of every block, the per-block work of
    coarseningIndicator_tree (SURVEY 8(d) 'wavelet side'): algorithmic bytes 8*nc*[(Bs+2g)^3 + Bs^3] + 8*nc per block."""
    import torch
    from wabbit_b200.solver import HVY_BLOCK, HVY_TMP
    g, _ = sol.setup_wavelet(wavelet)
    norm = sol.componentWiseNorm_tree((HVY_BLOCK, 0))
    norm[norm <= 1e-9] = 1.0
    for _ in range(3):
        sol.waveletDecomposition_tree((HVY_BLOCK, 0), (HVY_TMP, 0))
        st = sol.threshold_tree((HVY_TMP, 0), eps=1e-3, norm=norm)
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    fwt_ms = 0.0
    n0 = sol.launch_count
    w0 = time.perf_counter()
    for _ in range(reps):
        e0.record(stream)
        sol.waveletDecomposition_tree((HVY_BLOCK, 0), (HVY_TMP, 0))
        e1.record(stream)
        st = sol.threshold_tree((HVY_TMP, 0), eps=1e-3, norm=norm)     # synchronises (flags come back to the host)
        fwt_ms += e0.elapsed_time(e1)
    barrier()
    total_s = time.perf_counter() - w0
    nc, Bs = 4, a.bs
    bytes_fwt = 8 * nc * ((Bs + 2 * g) ** 3 + Bs ** 3) + 8 * nc
    achieved = bytes_fwt * nb / (fwt_ms / reps * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "wavelet_kernel_traffic.json")))["dram_bytes_per_block_per_launch"] * nb
    except Exception:
        pass
    return {"metric": "block-decompositions/s (FWT + threshold flags)", "value": nb * reps / total_s, "unit": "blocks/s", "wavelet": wavelet,
            "blocks": nb, "reps": reps, "coarsen_flags": int((st == -1).sum()), "gpu_launches": int(sol.launch_count - n0),
            "roofline": {"bound": "hbm", "kernel": "wavelet_fast_kernel<4,4,16,fwd> (FWT + Linfty detail)", "achieved": achieved, "unit": "GB/s", "avg_launch_ms": fwt_ms / reps,
                         "algorithmic_bytes_per_block": bytes_fwt, "traffic": traffic}}


def adaptive_leg(a, device_index, stream, eps=None, J0=5, Jmax=6, cycles=3, max_blocks=None, wavelet="CDF40"):
    """BASELINE config 3's cycle on one GPU (the protocol of performance_test.f90: refine_tree("everywhere") -> timeStep_tree ->
    adapt_tree), CDF40 (one coarsening sweep per adapt_tree call) or a lifted wavelet such as CDF44 (the reference's full-tree algorithm
    with the coarse extension, wabbit_b200/fulltree.py): Taylor-Green + three Gaussian vortex blobs
    on an equidistant level-J0 grid, coarsened by wavelet thresholding until the grid is stationary, then `cycles` timed cycles.
    block-updates/s counts the blocks the Runge-Kutta step advances (Nb after refinement, as performance.t's Nb_rhs)."""
    import torch
    from wabbit_b200 import Forest, Params, WabbitGPU
    eps = a.adaptive_eps if eps is None else eps
    max_blocks = max_blocks or int(2.5 * 8 ** J0)     # leaves after refinement + the mothers of the full tree, with room to spare
    lifted = wavelet[4] != "0"
    p = Params(dim=3, domain=(TWO_PI,) * 3, Bs=(a.bs,) * 3, wavelet=wavelet, g=int(wavelet[3]) - 1 + max(int(wavelet[4]) - 1, 0), g_rhs=2, n_eqn=4, Jmax=Jmax,
               discretization="FD_4th_central", skew_symmetry=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0,
               u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9).finalize()
    forest = Forest.uniform(3, J0, Jmax=Jmax, max_blocks=max_blocks)
    hvy, lvl, ixyz, _ = forest.active(0)
    sol = WabbitGPU(p, max_blocks=max_blocks, device=device_index, stream=stream.cuda_stream)
    sol.setup_wavelet(wavelet)
    sol.set_forest(forest)
    nb0 = len(hvy)
    shape = (nb0,) + sol.host_shape()[1:]
    host = torch.empty(shape, dtype=torch.float64, pin_memory=True)
    h_np = host.numpy()
    taylor_green_host(p, ixyz, lvl, h_np)
    # three Gaussian vortex blobs (sigma = 0.15), centres from rng seed 1
    rng = np.random.default_rng(1)
    centres = rng.random((3, 3)) * TWO_PI
    g, Bs = p.g, a.bs
    n = Bs + 2 * g
    dx = TWO_PI / (2 ** J0 * Bs)
    idx = torch.arange(n, dtype=torch.float64) - g
    for s0 in range(0, nb0, 2048):
        e = min(s0 + 2048, nb0)
        x0 = torch.from_numpy((ixyz[s0:e] * Bs).astype(np.float64)) * dx
        X = (idx[None, :] * dx + x0[:, 0:1])[:, None, None, :]
        Y = (idx[None, :] * dx + x0[:, 1:2])[:, None, :, None]
        Z = (idx[None, :] * dx + x0[:, 2:3])[:, :, None, None]
        for c in centres:
            r2 = (X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2
            blob = torch.exp(-r2 / (2 * 0.15 ** 2))
            host[s0:e, 0] += 2.0 * blob
            host[s0:e, 1] -= blob
            host[s0:e, 2] += 0.5 * blob
    sol.upload_ptr(host.data_ptr(), shape[1], hvy_ids=hvy)
    del host, h_np
    import gc
    gc.collect()                                # the page-locked initial-condition array is released before anything is timed
    torch.cuda.synchronize()
    sizes = [forest.n_blocks]
    for _ in range(Jmax):                       # coarsen until the grid is stationary
        forest, n0, n1 = sol.adapt_tree(forest, eps=eps, Jmin=1)
        sizes.append(n1)
        if n1 == n0:
            break
    t, it = 0.0, 0
    recs = []
    for cyc in range(cycles + 2):               # the first two cycles are the warm-up
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        forest = sol.refine_tree(forest)
        nb_rhs = forest.n_blocks
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        t, it, _dt = sol.timeStep_tree(t, it)
        torch.cuda.synchronize()
        w2 = time.perf_counter()
        forest, n0, n1 = sol.adapt_tree(forest, eps=eps, Jmin=1)
        torch.cuda.synchronize()
        w3 = time.perf_counter()
        recs.append((nb_rhs, n1, w1 - w0, w2 - w1, w3 - w2))
    sol.close()
    recs = recs[2:]
    tot = sum(r[2] + r[3] + r[4] for r in recs)
    return {"metric": f"adaptive block-updates/s (refine everywhere -> RK4 -> adapt, {wavelet}, 1 GPU)", "value": sum(r[0] for r in recs) / tot,
            "adapt_tree": ("full wavelet transformation with coarse extension and security zone (lifted wavelet): all levels in one call"
                           if lifted else "one coarsening sweep per call (unlifted wavelet)"),
            "unit": UNIT, "eps": eps, "Jmax": Jmax, "blocks_initial_coarsening": sizes, "cycles": len(recs),
            "blocks_rhs": [r[0] for r in recs], "blocks_after_adapt": [r[1] for r in recs],
            "ms_refine": [round(r[2] * 1e3, 2) for r in recs], "ms_rk4": [round(r[3] * 1e3, 2) for r in recs],
            "ms_adapt": [round(r[4] * 1e3, 2) for r in recs],
            "rk4_block_updates_per_s": sum(r[0] for r in recs) / sum(r[3] for r in recs),
            "note": "refine / adapt times include the host light-data stand-ins (new forest, neighbour search, topology upload), which "
                    "stay in host Fortran in a WABBIT build; ms_rk4 is the device-resident time step on the graded grid"}


def blob_initial_condition(p, sol, hvy, lvl, ixyz, J0, chunk=4096):
    """Taylor-Green + three Gaussian vortex blobs (sigma = 0.15, centres from rng seed 1) on the listed blocks of an
    equidistant level-J0 grid, evaluated ON THE DEVICE straight into the resident hvy_block (interiors, compact layout): a level-6 grid would
    need 23 GB of host memory per rank and minutes of host time otherwise.  Elementwise float64: the values do not depend on the partition."""
    import ctypes as C
    import torch
    from wabbit_b200.multi import _DevPtr
    rng = np.random.default_rng(1)
    centres = rng.random((3, 3)) * TWO_PI
    Bs = p.Bs[0]
    dx = TWO_PI / (2 ** J0 * Bs)
    ptr, n = C.c_void_p(), C.c_int64()
    sol._check(sol._lib.wgpu_device_pointer(sol._ctx, 0, 0, C.byref(ptr), C.byref(n)))
    dev = torch.device("cuda", torch.cuda.current_device())
    U = torch.as_tensor(_DevPtr(ptr.value, n.value), device=dev).view(sol.max_blocks, 4, Bs, Bs, Bs)
    idx = torch.arange(Bs, dtype=torch.float64, device=dev)
    assert (np.diff(hvy) == 1).all()
    with torch.cuda.stream(torch.cuda.ExternalStream(sol.stream) if sol.stream else torch.cuda.current_stream()):
        for s0 in range(0, len(hvy), chunk):
            e = min(s0 + chunk, len(hvy))
            x0 = torch.from_numpy((ixyz[s0:e] * Bs).astype(np.float64)).to(dev) * dx
            X = (idx[None, :] * dx + x0[:, 0:1])[:, None, None, :]
            Y = (idx[None, :] * dx + x0[:, 1:2])[:, None, :, None]
            Z = (idx[None, :] * dx + x0[:, 2:3])[:, :, None, None]
            blk = U[int(hvy[s0]) - 1:int(hvy[s0]) - 1 + (e - s0)]
            blk[:, 0] = torch.sin(X) * torch.cos(Y) * torch.cos(Z)
            blk[:, 1] = -torch.cos(X) * torch.sin(Y) * torch.cos(Z)
            blk[:, 2] = 0.0
            blk[:, 3] = (torch.cos(2.0 * X) + torch.cos(2.0 * Y)) * (torch.cos(2.0 * Z) + 2.0) / 16.0
            for c in centres:
                r2 = (X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2
                blob = torch.exp(-r2 / (2 * 0.15 ** 2))
                blk[:, 0] += 2.0 * blob
                blk[:, 1] -= blob
                blk[:, 2] += 0.5 * blob
    torch.cuda.synchronize()


def adaptive_leg_multi(a, rank, world, local, stream, eps=None, J0=5, Jmax=6, cycles=3, wavelet="CDF44", bs=16, sphere_on=False):
    """BASELINE config 3 / 4 on `world` GPUs, one process each (the protocol of performance_test.f90:188-214: refine_tree("everywhere") ->
    timeStep_tree -> adapt_tree): blocks partitioned by the Hilbert curve; halo copies of the neighbouring blocks of other ranks; block
    transport, dt reduction and light-data collectives on the library's own NCCL communicator (wabbit_b200/multi.py: DistributedWabbit).
    Returns a record; any failure is reported in it (the same on every rank: the light data are replicated) instead of raised."""
    import hashlib
    import torch
    import torch.distributed as dist
    from wabbit_b200 import Forest, Params, WabbitGPU
    from wabbit_b200.multi import DistributedWabbit
    eps = a.adaptive_eps if eps is None else eps
    X, Y = int(wavelet[3]), int(wavelet[4])
    p = Params(dim=3, domain=(TWO_PI,) * 3, Bs=(bs,) * 3, wavelet=wavelet, g=X - 1 + max(Y - 1, 0), g_rhs=2, n_eqn=4, Jmax=Jmax,
               discretization="FD_4th_central", skew_symmetry=True, c0=10.0, nu=3.125e-3, gamma_p=0.0, CFL=1.0,
               u_mean_set=(0.0, 0.0, 0.0), time_max=1.0e9)
    sphere = None
    if sphere_on:      # BASELINE config 4's kind of run: volume penalization of a translating sphere, mask evaluated in the stage kernel
        p.penalization, p.C_eta = True, 1.0e-3
    p = p.finalize()
    if sphere_on:
        from wabbit_b200.mask import SphereMask3D
        sphere = SphereMask3D(p, center=(3.0, 3.1, 3.2), radius=0.8, velocity=(0.5, 0.3, -0.2))
    nb_total = 8 ** J0
    # own blocks + mothers of the full tree + halo / scratch copies, with a wide margin: a capacity error on ONE rank would leave the others
    # waiting in a collective, so it must not happen (the light data are replicated, but slot counts are per rank)
    mb = int(2.0 * nb_total / world) + 8192
    rec = {"metric": f"adaptive block-updates/s (refine everywhere -> RK4{' + penalization of a translating sphere' if sphere_on else ''} -> adapt_tree, "
                     f"{wavelet}, Bs={bs}, {world} GPUs)", "unit": UNIT, "eps": eps, "Jmax": Jmax, "Bs": bs, "initial_level": J0}
    sol = None
    try:
        forest = Forest.uniform(3, J0, Jmax=Jmax, n_ranks=world, max_blocks=mb)
        hvy, lvl, ixyz, _ = forest.active(rank)
        sol = WabbitGPU(p, max_blocks=mb, device=local, stream=stream.cuda_stream)
        sol.setup_wavelet(wavelet)
        sol.comm_init(rank, world)
        if sphere is not None:
            sphere.attach(sol)
        drv = DistributedWabbit(sol, forest, rank, world)
        blob_initial_condition(p, sol, hvy, lvl, ixyz, J0)

        def sync():
            dist.barrier()
            torch.cuda.synchronize()

        t, it = 0.0, 0
        keeps = (lambda level, pos: sphere.keeps(level, pos, t)) if sphere is not None else None      # threshold_mask follows the sphere
        extra = dict(mask_keeps=keeps, full_tree=True) if sphere is not None else {}
        sizes = [forest.n_blocks]
        for _ in range(Jmax):
            _f, n0, n1 = drv.adapt_tree(eps=eps, Jmin=1, **extra)
            sizes.append(n1)
            if n1 == n0:
                break
        recs = []
        from wabbit_b200 import multi as _mg
        prof = None
        for cyc in range(cycles + 2):
            if cyc == 2 and _mg.TIMING is not None:
                _mg.TIMING.clear()
            if cyc == 2 and rank == 0 and os.environ.get("WABBIT_PROFILE"):      # development: where the host time of a cycle goes
                import cProfile
                prof = cProfile.Profile()
                prof.enable()
            sync()
            w0 = time.perf_counter()
            nb_rhs = drv.refine_tree().n_blocks
            sync()
            w1 = time.perf_counter()
            t, it, _dt = drv.timeStep_tree(t, it)
            sync()
            w2 = time.perf_counter()
            _f, n0, n1 = drv.adapt_tree(eps=eps, Jmin=1, **extra)
            sync()
            w3 = time.perf_counter()
            recs.append((nb_rhs, n1, w1 - w0, w2 - w1, w3 - w2))
        if prof is not None:
            import io
            import pstats
            prof.disable()
            buf = io.StringIO()
            pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(45)
            print(buf.getvalue(), file=sys.stderr, flush=True)
        if _mg.TIMING is not None and rank == 0:
            rec["phase_ms_per_cycle_rank0"] = {k: round(v * 1e3 / cycles, 2) for k, v in sorted(_mg.TIMING.items())}
        per_rank = [drv.forest.n_active(r) for r in range(world)]
        n_halo, n_int, n_bnd = drv.stepper.plan.n_halo, drv.stepper.n_int, drv.stepper.n_bnd
        # checksum of the final grid: the block list (level, treecode) in space-filling-curve order, identical for every GPU count, and the
        # Linfty norm of the state per component (MAX over ranks: independent of the partition)
        h = hashlib.sha256()
        for r in range(world):
            _, l_r, _, tc_r = drv.forest.active(r)
            h.update(l_r.astype(np.int32).tobytes())
            h.update(tc_r.astype(np.int64).tobytes())
        nrm = drv.global_norm("Linfty")
        recs = recs[2:]
        tm = torch.tensor([[r[2], r[3], r[4]] for r in recs], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        tm = tm.cpu().numpy()
        tot = float(tm.sum())
        rec.update({"value": sum(r[0] for r in recs) / tot, "blocks_initial_coarsening": sizes, "cycles": len(recs),
                    "blocks_rhs": [r[0] for r in recs], "blocks_after_adapt": [r[1] for r in recs],
                    "ms_refine": [round(v * 1e3, 2) for v in tm[:, 0]], "ms_rk4": [round(v * 1e3, 2) for v in tm[:, 1]],
                    "ms_adapt": [round(v * 1e3, 2) for v in tm[:, 2]],
                    "rk4_block_updates_per_s": sum(r[0] for r in recs) / float(tm[:, 1].sum()),
                    "cycle_over_rk4": float(tm[:, 1].sum()) / tot,
                    "blocks_per_rank_final": per_rank, "rank0_halo_blocks": n_halo, "rank0_interior_boundary": [n_int, n_bnd],
                    "checksum": {"grid_sha256": h.hexdigest()[:16], "time": t, "linfty": [float(v).hex() for v in nrm]},
                    "note": "max over ranks of every phase; refine / adapt include the replicated host light-data logic (new forest, partition, "
                            "halo plan), which stays in host Fortran in a WABBIT build; the neighbour relations are derived on the device"})
    except Exception as e:      # noqa: BLE001 -- a secondary figure must not take the headline line down
        import traceback
        rec["error"] = repr(e)
        rec["traceback"] = traceback.format_exc()[-1500:]
    finally:
        if sol is not None:
            try:
                sol.close()
            except Exception:
                pass
    return rec


def weak_scaling_leg(a, rank, world, local, stream):
    """The N = 1 workload on EVERY GPU: the equidistant grid one level finer (8^(J+1) blocks) over 8 GPUs = 32 768 blocks per GPU, the same
    per-GPU work as the headline's single-GPU run (an octree's block count only tiles the GPUs evenly for N = 1 and N = 8).  Taylor-Green
    evaluated on the device straight into the resident hvy_block; same stepper (wgpu_rk_steps), same checksum convention."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from wabbit_b200 import Forest, WabbitGPU
    from wabbit_b200.multi import _DevPtr, attach_exchange
    level = a.level + 1
    a2 = argparse.Namespace(**vars(a))
    a2.level = level
    p = make_params(a2)
    rec = {"metric": METRIC, "unit": UNIT, "scaling": "weak", "level": level, "blocks": 8 ** level, "blocks_per_gpu": 8 ** level // world, "n_gpus": world}
    sol = None
    try:
        forest = Forest.uniform(3, level, block_dist="sfc_hilbert", n_ranks=world)
        hvy, lvl, ixyz, _ = forest.active(rank)
        sol = WabbitGPU(p, max_blocks=forest.max_blocks, device=local, stream=stream.cuda_stream)
        sol.comm_init(rank, world)
        attach_exchange(sol, forest, rank, world)
        ptr, n = C.c_void_p(), C.c_int64()
        sol._check(sol._lib.wgpu_device_pointer(sol._ctx, 0, 0, C.byref(ptr), C.byref(n)))
        dev = torch.device("cuda", local)
        Bs = a.bs
        U = torch.as_tensor(_DevPtr(ptr.value, n.value), device=dev).view(sol.max_blocks, 4, Bs, Bs, Bs)
        dx = TWO_PI / (2 ** level * Bs)
        idx = torch.arange(Bs, dtype=torch.float64, device=dev)
        assert (np.diff(hvy) == 1).all()
        for s0 in range(0, len(hvy), 4096):
            e = min(s0 + 4096, len(hvy))
            x0 = torch.from_numpy((ixyz[s0:e] * Bs).astype(np.float64)).to(dev) * dx
            X = (idx[None, :] * dx + x0[:, 0:1])[:, None, None, :]
            Y = (idx[None, :] * dx + x0[:, 1:2])[:, None, :, None]
            Z = (idx[None, :] * dx + x0[:, 2:3])[:, :, None, None]
            blk = U[int(hvy[s0]) - 1:int(hvy[s0]) - 1 + (e - s0)]
            blk[:, 0] = torch.sin(X) * torch.cos(Y) * torch.cos(Z)
            blk[:, 1] = -torch.cos(X) * torch.sin(Y) * torch.cos(Z)
            blk[:, 2] = 0.0
            blk[:, 3] = (torch.cos(2.0 * X) + torch.cos(2.0 * Y)) * (torch.cos(2.0 * Z) + 2.0) / 16.0
        torch.cuda.synchronize()
        steps = max(1, min(a.steps, 20))
        t, _ = sol.stepper.steps(0.0, a.warmup)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t, _dt = sol.stepper.steps(t, steps)
        e1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
        tm = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
        sol.setup_wavelet("CDF40")
        tck = torch.tensor(list(sol.componentWiseNorm_tree((0, 0), "Linfty")), dtype=torch.float64, device=dev)
        dist.all_reduce(tck, op=dist.ReduceOp.MAX)
        rec.update({"value": 8 ** level * steps / (ms * 1e-3), "per_gpu_value": 8 ** level * steps / (ms * 1e-3) / world, "steps": steps,
                    "ms_per_step": ms / steps, "transport": sol.comm_transport(),
                    "checksum": {"time": t, "linfty": [float(v).hex() for v in tck.cpu().tolist()]},
                    "note": "weak-scaling efficiency = per_gpu_value / the N = 1 line's value (same 32 768 blocks per GPU)"})
    except Exception as e:      # noqa: BLE001 -- a secondary figure must not take the headline line down
        import traceback
        rec["error"] = repr(e)
        rec["traceback"] = traceback.format_exc()[-1200:]
    finally:
        if sol is not None:
            try:
                sol.close()
            except Exception:
                pass
    return rec


def compression_leg(a, rank, world, local, stream, J, wavelets=("CDF40", "CDF42", "CDF44"), eps_idx=None, bs=16):
    """BASELINE config 5: the protocol of post_compression_unit_test.f90:107-215 (wabbit_b200/compression.py) -- one component, Gauss blob on the
    equidistant level-J grid, adapt_tree (full wavelet transformation; coarse extension + security zone for the lifted wavelets), Nb,
    refineToEquidistant_tree, relative errors -- for a spread of the protocol's 51 thresholds (all of them with --compression-full), on
    `world` GPUs.  value: blocks of the equidistant grid pushed through adapt_tree + refineToEquidistant_tree per second."""
    import torch
    import torch.distributed as dist
    from wabbit_b200 import Forest, WabbitGPU
    from wabbit_b200 import compression as CP
    eps_idx = list(range(0, 51, 5)) if eps_idx is None else list(eps_idx)
    eps_list = [float(CP.EPS_SWEEP[i]) for i in eps_idx]
    nb0 = 8 ** J
    out = {"metric": "blocks/s through the compression protocol (adapt_tree + refineToEquidistant_tree, 1 component)", "unit": "blocks/s",
           "level": J, "blocks": nb0, "Bs": bs, "n_gpus": world, "eps": eps_list}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_all = 0.0
    for wv in wavelets:
        sol = None
        try:
            p = CP.compression_params(wv, bs, J)
            mb = int((1.4 if world == 1 else 2.0) * nb0 / world) + (0 if world == 1 else 8192)
            forest = Forest.uniform(3, J, Jmax=J, n_ranks=world, max_blocks=mb)
            sol = WabbitGPU(p, max_blocks=mb, device=local, stream=stream.cuda_stream)
            sol.setup_wavelet(wv)
            drv = None
            if world > 1:
                from wabbit_b200.multi import DistributedWabbit
                sol.comm_init(rank, world)
                drv = DistributedWabbit(sol, forest, rank, world)
            else:
                sol.set_forest(forest)
            test = CP.CompressionTest(sol, forest, drv)
            test.run(eps_list[len(eps_list) // 2:len(eps_list) // 2 + 1], sync=sync)       # warm-up
            recs = test.run(eps_list, sync=sync)
            tm = torch.tensor([[r["ms_adapt"], r["ms_refine"]] for r in recs], dtype=torch.float64, device=torch.device("cuda", local))
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            tm = tm.cpu().numpy()
            t = float(tm.sum()) * 1e-3
            t_all += t
            out[wv] = {"value": nb0 * len(recs) / t, "Nb": [r["Nb"] for r in recs], "err_L2": [float("%.6e" % r["err_L2"]) for r in recs],
                       "err_Linfty": [float("%.6e" % r["err_Linfty"]) for r in recs], "ms_adapt": [round(float(v), 1) for v in tm[:, 0]],
                       "ms_refine": [round(float(v), 1) for v in tm[:, 1]]}
        except Exception as e:      # noqa: BLE001 -- a secondary figure must not take the headline line down
            import traceback
            out[wv] = {"error": repr(e), "traceback": traceback.format_exc()[-1200:]}
        finally:
            if sol is not None:
                try:
                    sol.close()
                except Exception:
                    pass
    n_ok = sum(len(eps_list) for wv in wavelets if "value" in out.get(wv, {}))
    out["value"] = nb0 * n_ok / t_all if t_all > 0 else None
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--level", type=int, default=5, help="equidistant level J: (2^J)^3 blocks")
    ap.add_argument("--bs", type=int, default=16)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--e2e-trees", type=int, default=3, help="independent trees in flight in the end-to-end leg (1 = sequential only)")
    ap.add_argument("--cpu-level", type=int, default=4, help="level of the CPU arm's bounded sample (4: 4096 blocks, ~1 s per RK4 step on 16 cores)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-wavelet", action="store_true", help="skip the secondary FWT + threshold figure")
    ap.add_argument("--no-adaptive", action="store_true", help="skip the secondary adaptive-cycle figure")
    ap.add_argument("--no-weak", action="store_true", help="N = 8: skip the weak-scaling record (one level finer, 32 768 blocks per GPU)")
    ap.add_argument("--no-compression", action="store_true", help="skip the compression-protocol figure (BASELINE config 5)")
    ap.add_argument("--compression-level", type=int, default=5)
    ap.add_argument("--compression-full", action="store_true", help="all 51 thresholds, three wavelets, levels J and J+1")
    ap.add_argument("--adaptive-eps", type=float, default=1.0e-6)
    ap.add_argument("--adaptive-wavelet", default="CDF40", help="wavelet of --adaptive-only / --adaptive-multi")
    ap.add_argument("--adaptive-legs", default="", help="N > 1: comma list out of Bs16,Bs16_sphere,Bs18 (default: all three)")
    ap.add_argument("--adaptive-level", type=int, default=0, help="initial equidistant level of the adaptive legs (0: 5 up to 4 GPUs, 6 from 8 GPUs on)")
    ap.add_argument("--adaptive-only", action="store_true", help="run only the adaptive-cycle leg and print its record (development)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:   # torchrun exports OMP_NUM_THREADS=1: give the host light-data logic (neighbour search) this rank's share of the cores
        os.environ["OMP_NUM_THREADS"] = os.environ.get("WABBIT_OMP_THREADS") or str(max(1, (os.cpu_count() or 1) // world))
    if a.adaptive_only:
        import torch
        torch.cuda.set_device(0)
        print(json.dumps(adaptive_leg(a, 0, torch.cuda.current_stream(), wavelet=a.adaptive_wavelet, J0=a.adaptive_level or 5, Jmax=(a.adaptive_level or 5) + 1)), flush=True)
        return
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
