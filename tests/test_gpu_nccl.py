"""Parity of the real NCCL path (one process per GPU, the library's own communicator: wgpu_comm_init, wgpu_rk_steps, wgpu_exchange_array,
wgpu_ship_blocks) against the single-rank driver, bit for bit.  Needs >= 2 CUDA devices; skipped otherwise (the same drivers run over
2 / 3 ranks as threads on ONE device in tests/test_multi_halo.py and tests/test_multi_rank.py)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, *args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py"), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(_n_devices() < 2, reason="needs >= 2 CUDA devices")
@pytest.mark.parametrize("case", [("uniform",), ("uniform_nccl",), ("graded",), ("cycle", "CDF44", "16"), ("cycle", "CDF44", "18"), ("cycle", "CDF40", "16"),
                                  ("compression", "CDF42", "4", "1e-6"), ("compression", "CDF44", "4", "1.0"), ("compression", "CDF40", "4", "1e-2")],
                         ids=lambda c: "-".join(c))
def test_nccl_path_equals_single_rank(case):
    world = min(_n_devices(), 4)
    out = _run(world, *case)
    assert out.count("ok=True") == world, out
