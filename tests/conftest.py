import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu() -> bool:
    """A CUDA device, asked of the CUDA runtime itself (libcudart's cudaGetDeviceCount), so that a box with a GPU but a broken torch
    does not silently skip the parity suite; torch is the fallback probe."""
    import ctypes
    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
            n = ctypes.c_int(0)
            if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0:
                return n.value > 0
            return False
        except OSError:
            continue
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        try:
            import torch  # noqa: F401  (the GPU tests use torch for pinned memory / streams)
        except Exception as e:      # a GPU box without a working torch must fail loudly, not look green
            raise pytest.UsageError(f"a CUDA device is present but torch cannot be imported ({e}): the -m gpu parity suite cannot run")
        return
    if os.environ.get("WABBIT_REQUIRE_GPU"):
        raise pytest.UsageError("WABBIT_REQUIRE_GPU is set but no CUDA device was found")
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---------------------------------------------------------------------------------------------------------------------
# The long oracle runs against the reference's fixtures (2281-step 3vortices runs, the four cylinder runs, the Taylor-Green runs: 0.5 - 5
# minutes of CPU each) start in worker processes as soon as the collection is known and run while the other CPU tests execute; the tests that own them
# (test_oracle_adaptive.py, test_oracle_cylinder.py) collect the results.  Selecting one of them alone works the same way.
_BG = {"pool": None, "futures": {}}


def background(kind, key, fn):
    """result of the background job (kind, key); computed inline if the job was not started"""
    fut = _BG["futures"].get((kind, key))
    return fut.result() if fut is not None else fn(key)


def pytest_collection_finish(session):
    names = {it.name.split("[")[0] for it in session.items}
    want_adaptive = "test_adaptive_run_fixture" in names
    want_cylinder = "test_cylinder_fixtures" in names
    want_tg = "test_taylor_green_fixture" in names and (want_adaptive or want_cylinder)      # alone they run inline, one after the other
    if not (want_adaptive or want_cylinder) or _BG["pool"] is not None:
        return
    import concurrent.futures as cf
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    _BG["pool"] = cf.ProcessPoolExecutor(max_workers=7, mp_context=mp.get_context("spawn"))
    if want_adaptive:
        import test_oracle_adaptive as TA
        for c in TA.full_run_cases():
            _BG["futures"][("adaptive", c)] = _BG["pool"].submit(TA._full_run, c)
    if want_cylinder:
        import cylinder_case as CC
        import test_oracle_cylinder as TC
        for c in CC.CASES:
            _BG["futures"][("cylinder", c)] = _BG["pool"].submit(TC._run_case, c)
    if want_tg:                      # queued behind the long jobs: done well before the adaptive runs are
        import test_oracle_golden as TG
        for c in sorted(TG.CASES, key=lambda c: -TG.CASES[c]["g"]):
            _BG["futures"][("taylor_green", c)] = _BG["pool"].submit(TG._tg_run, c)


def pytest_sessionfinish(session, exitstatus):
    if _BG["pool"] is not None:
        _BG["pool"].shutdown(wait=False, cancel_futures=True)
        _BG["pool"] = None
