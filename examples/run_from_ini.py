#!/usr/bin/env python
"""Run an adaptive 2-D ACM case from a WABBIT parameter file on one GPU and write WABBIT field files:

    python examples/run_from_ini.py /path/to/TESTING/acm/acm_CDF44/acm_cyl.ini --out out/

What main.f90 does for `adapt_tree = 1` (LIB/MAIN/main.f90:85-443): READ_PARAMETERS, the initial condition (setInitialCondition_tree:
`inicond = meanflow` on an adaptively generated grid, or `read_from_files = 1` with `input_files` + one adapt_tree), then per time step refine_tree -> createMask_tree -> RungeKuttaGeneric -> adapt_tree,
saving `ux uy p mask` at t = 0 and at every multiple of `write_time` (write_method = fixed_time).  The same sequence as
tests/test_gpu_cylinder2d.py, which compares its results with the files the reference wrote for this .ini; `--plan` stops before the first
device call (parameter / mask / grid summary only; runs without a GPU).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wabbit_b200 import Forest, Params  # noqa: E402
from wabbit_b200 import h5io  # noqa: E402
from wabbit_b200.mask import mask_from_ini  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("ini")
    ap.add_argument("--out", default="out")
    ap.add_argument("--max-blocks", type=int, default=4000)
    ap.add_argument("--input-dir", default="", help="directory of the input_files of a restart (default: next to the .ini)")
    ap.add_argument("--plan", action="store_true", help="parse, build the mask generator and the initial grid, print the plan, stop")
    a = ap.parse_args()

    p = Params.from_ini(a.ini)
    if p.dim != 2 or not (p.read_from_files or p.inicond == "meanflow"):
        raise SystemExit("this example covers dim = 2 with inicond = meanflow or read_from_files = 1 (the TESTING/acm 2-D cases)")
    mask = mask_from_ini(a.ini, p)
    state = None
    if p.read_from_files:                        # setInitialCondition_tree, read_from_files = 1: readHDF5vct_tree(input_files)
        base = a.input_dir if a.input_dir else os.path.dirname(os.path.abspath(a.ini))
        state = h5io.read_state([os.path.join(base, f) for f in p.input_files], p.g)
        forest = Forest.from_blocks(2, p.Jmax, state["level"], state["ixyz"].astype(np.int32), max_blocks=a.max_blocks)
        print(f"restart from {p.input_files}: t = {state['time']}, iteration = {state['iteration']}, {forest.n_blocks} blocks on levels "
              f"{int(state['level'].min())}..{int(state['level'].max())}")
    else:
        forest = Forest.uniform(2, p.Jmin, Jmax=p.Jmax, max_blocks=a.max_blocks)
    print(f"{a.ini}: {p.wavelet}, Bs = {p.Bs[0]}, g = {p.g}, Jmin..Jmax = {p.Jmin}..{p.Jmax}, eps = {p.eps}, refinement {p.refinement_indicator}, "
          f"penalization = {p.penalization} ({type(mask).__name__ if mask is not None else 'no mask'}), sponge = {p.use_sponge}, "
          f"time_max = {p.time_max}, write_time = {p.write_time}; initial grid {forest.n_blocks} blocks")
    if a.plan:
        return

    from wabbit_b200 import WabbitGPU
    from wabbit_b200.solver import HVY_BLOCK
    from wabbit_b200.timeloop import AdaptiveLoop
    os.makedirs(a.out, exist_ok=True)
    sol = WabbitGPU(p, max_blocks=a.max_blocks)
    sol.setup_wavelet(p.wavelet)
    sol.set_forest(forest)
    tc = p.threshold_state_vector_component or None
    t0, it0 = (state["time"], state["iteration"]) if state is not None else (0.0, 0)
    loop = AdaptiveLoop(sol, forest, t0, it0, mask=mask, threshold_mask=mask is not None and p.threshold_mask, thresh_comp=tc)
    if p.nsave_stats != 99999999 or abs(p.tsave_stats - 9999999.9) > 1e-3:        # [Statistics] nsave_stats / tsave_stats: the *.t files
        loop.stats_dir = a.out

    def set_inicond(lp):                      # inicond = meanflow (inicond_ACM.f90:285-288): u = u_mean_set, p = 0
        hvy, _, _, _ = lp.forest.active(0)
        host = np.zeros((int(hvy.max()),) + sol.host_shape()[1:])
        for d in range(p.dim):
            host[hvy - 1, d] = p.u_mean_set[d]
        sol.upload(host, HVY_BLOCK, 0, hvy_ids=hvy)

    def save():
        hvy, lvl, pos, tc = loop.forest.active(0)
        host = np.zeros((int(hvy.max()),) + sol.host_shape()[1:])
        sol.download(host, g_sync=p.g)                         # ghost nodes as sync_ghosts_tree leaves them: the files hold one of them
        status = loop.status if loop.status is not None else np.zeros(len(hvy), np.int32)
        paths = h5io.save_data(a.out, ("ux", "uy", "p"), host[hvy - 1], lvl, pos, tc, p, loop.time, loop.iteration, refinement_status=status)
        if mask is not None:
            chi = mask.fill(lvl, pos)[:, :1]
            paths += h5io.save_data(a.out, ("mask",), chi, lvl, pos, tc, p, loop.time, loop.iteration, refinement_status=status)
        print(f"t = {loop.time:.6f} it = {loop.iteration} Nb = {loop.forest.n_blocks}: wrote {[os.path.basename(q) for q in paths]}")

    if state is not None:
        hvy, lvl, pos, _ = forest.active(0)                   # the forest orders the blocks along the space-filling curve
        at = {(int(l), int(x[0]), int(x[1])): b for b, (l, x) in enumerate(zip(state["level"], state["ixyz"]))}
        order = np.array([at[(int(l), int(x[0]), int(x[1]))] for l, x in zip(lvl, pos)])
        host = np.zeros((int(hvy.max()),) + sol.host_shape()[1:])
        host[hvy - 1] = state["hvy"][order]
        sol.upload(host, HVY_BLOCK, 0, hvy_ids=hvy)
        if p.adapt_inicond:
            loop.adapt_tree()
    else:
        set_inicond(loop)
        loop.adaptive_inicond(set_inicond)
    save()
    saved_at = loop.time
    while loop.time < p.time_max:
        loop.step()
        if p.write_method == "fixed_time" and abs(loop.time / p.write_time - round(loop.time / p.write_time)) <= 1e-12:
            save()
            saved_at = loop.time
    if saved_at != loop.time:                                  # the final state (main.f90 saves when the run ends)
        save()
    sol.close()


if __name__ == "__main__":
    main()
