"""Worker of tests/test_gpu_nccl.py: one process per GPU (torchrun), the library's own NCCL communicator.  Every rank also computes the
single-rank result on its own GPU and compares its share bit for bit:
  uniform   RK4 steps on an equidistant grid, face patches (MultiGPUStepper, wgpu_rk_steps): peer stores over NVLink (CUDA IPC) when available
  uniform_nccl   the same with grouped ncclSend / ncclRecv (wgpu_comm_set_transport(0))
  graded    RK4 steps on a graded grid, halo blocks (HaloStepper, wgpu_rk_steps), then download with a synchronised ghost shell
  cycle     refine_tree -> RK4 -> adapt_tree with the lifted full-tree algorithm (DistributedWabbit: wgpu_ship_blocks, wgpu_exchange_array)
  compression   the protocol of post_compression_unit_test.f90 (adapt_tree from the equidistant grid + refineToEquidistant_tree, one component)
Exit code 0 = all ranks agree with the single-rank driver."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import oracle as O
    from util import graded_blocks, orc_params, tg_params
    from wabbit_b200 import Forest, WabbitGPU
    from wabbit_b200.multi import DistributedWabbit, attach_exchange, attach_halo

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    what = sys.argv[1]
    ok = True

    def state(p, forest_1, seed=3, noise=0.02):
        po = orc_params(p)
        _, l1, x1, _ = forest_1.active(0)
        grid = O.Grid(level=l1.astype(np.int64), ixyz=x1.astype(np.int64), dim=3)
        u = O.alloc(grid, po)
        O.inicond_taylor_green(grid, po, u)
        u += noise * np.random.default_rng(seed).standard_normal(u.shape)
        return u

    def make(p, mb, wavelet=None):
        s = WabbitGPU(p, max_blocks=mb, device=local, stream=torch.cuda.current_stream().cuda_stream)
        if wavelet:
            s.setup_wavelet(wavelet)
        return s

    def interior(p, a):
        g = p.g
        return a[:, :, g:-g, g:-g, g:-g]

    if what in ("uniform", "uniform_nccl"):
        p = tg_params(Bs=16, J=3)
        f1, fw = Forest.uniform(3, 3), Forest.uniform(3, 3, n_ranks=world)
        u = state(p, f1)
        s1 = make(p, f1.n_blocks)
        s1.set_forest(f1)
        s1.upload(u)
        t1, dt1 = s1.RungeKuttaSteps(0.0, 3)
        ref = np.zeros_like(u)
        s1.download(ref, g_sync=0)
        t1b, dt1b = s1.RungeKuttaSteps(t1, 1)
        ref_b = np.zeros_like(u)
        s1.download(ref_b, g_sync=0)
        s1.close()
        s = make(p, fw.max_blocks)
        s.comm_init(rank, world)
        if what == "uniform_nccl":
            s.comm_set_transport(False)
        st = attach_exchange(s, fw, rank, world)
        assert st.in_library
        if what == "uniform_nccl":
            assert s.comm_transport() == "nccl"
        off = sum(fw.n_active(r) for r in range(rank))
        n = fw.n_active(rank)
        h = np.zeros(s.host_shape())
        h[:n] = u[off:off + n]
        s.upload(h)
        t2, dt2 = st.steps(0.0, 2)
        t2, dt2 = st.steps(t2, 1)          # a second call: the stage sequence numbers of the peer-store exchange carry on
        out = np.zeros(s.host_shape())
        s.download(out, g_sync=0)
        ok = (t1 == t2) and (dt1 == dt2) and np.array_equal(interior(p, out[:n]), interior(p, ref[off:off + n]))
        # the exchange declared a second time on the same communicator (pools re-exported, peers re-mapped), then one more step
        st = attach_exchange(s, fw, rank, world)
        t3, dt3 = st.steps(t2, 1)
        s.download(out, g_sync=0)
        ok = ok and (t3 == t1b) and (dt3 == dt1b) and np.array_equal(interior(p, out[:n]), interior(p, ref_b[off:off + n]))
        print(f"rank {rank}: uniform t={t3!r} dt={dt3!r} int/bnd={st.n_int}/{st.n_bnd} transport={s.comm_transport()} ok={ok}", flush=True)
        s.close()
    elif what == "graded":
        wavelet = "CDF44"
        w = O.setup_wavelet(wavelet)
        lv, ix = graded_blocks(3, 1, 3, 11, 0.3)
        mb = 4 * len(lv) + 64
        p = tg_params(Bs=16, J=3, wavelet_g=w.g_default)
        p.wavelet = wavelet
        f1 = Forest.from_blocks(3, 3, lv, ix, max_blocks=mb)
        fw = Forest.from_blocks(3, 3, lv, ix, n_ranks=world, max_blocks=mb)
        u = state(p, f1)
        s1 = make(p, mb, wavelet)
        s1.set_forest(f1)
        h1 = np.zeros(s1.host_shape())
        h1[:len(lv)] = u
        s1.upload(h1)
        t1, dt1 = s1.RungeKuttaSteps(0.0, 2)
        ref = np.zeros(s1.host_shape())
        s1.download(ref, g_sync=p.g)                        # fully synchronised ghost shell, filtered restriction
        s1.close()
        s = make(p, mb, wavelet)
        s.comm_init(rank, world)
        st = attach_halo(s, fw, rank, world)
        assert st.in_library
        off = sum(fw.n_active(r) for r in range(rank))
        n = fw.n_active(rank)
        h = np.zeros(s.host_shape())
        h[:n] = u[off:off + n]
        s.upload(h)
        t2, dt2 = st.steps(0.0, 2)
        st.exchange_array(0, 0)
        out = np.zeros(s.host_shape())
        s.download(out, g_sync=p.g, hvy_ids=np.arange(1, n + 1, dtype=np.int32))
        ok = (t1 == t2) and (dt1 == dt2) and np.array_equal(out[:n], ref[off:off + n])
        print(f"rank {rank}: graded t={t2!r} dt={dt2!r} halo={st.plan.n_halo} int/bnd={st.n_int}/{st.n_bnd} ok={ok}", flush=True)
        s.close()
    elif what == "cycle":
        wavelet, Jmax = sys.argv[2] if len(sys.argv) > 2 else "CDF44", 4
        Bs = int(sys.argv[3]) if len(sys.argv) > 3 else 16
        w = O.setup_wavelet(wavelet)
        lv, ix = graded_blocks(3, 1, 3, seed=5, frac=0.25)
        mb = 10 * len(lv) + 64
        p = tg_params(Bs=Bs, J=Jmax, wavelet_g=w.g_default)
        p.wavelet = wavelet
        p.eps = 1.0e-2
        f1 = Forest.from_blocks(3, Jmax, lv, ix, n_ranks=1, max_blocks=mb)
        fw = Forest.from_blocks(3, Jmax, lv, ix, n_ranks=world, max_blocks=mb)
        po = orc_params(p)
        _, l1, x1, _ = f1.active(0)
        grid = O.Grid(level=l1.astype(np.int64), ixyz=x1.astype(np.int64), dim=3)
        u = O.alloc(grid, po)
        O.inicond_taylor_green(grid, po, u)
        amp = np.where(x1[:, 0] * 2 < 2 ** l1, 0.05, 1.0e-6)
        u += amp[:, None, None, None, None] * np.random.default_rng(2).standard_normal(u.shape)
        s1 = make(p, mb, wavelet)
        s1.set_forest(f1)
        h1 = np.zeros(s1.host_shape())
        h1[:grid.n] = u
        s1.upload(h1)
        f = s1.refine_tree(f1)
        nb1 = f.n_blocks
        _, _, dt1 = s1.timeStep_tree(0.0, 0)
        f, n0, n1 = s1.adapt_tree(f, eps=p.eps, Jmin=1)
        _, lf, xf, _ = f.active(0)
        ref = np.zeros(s1.host_shape())
        s1.download(ref, g_sync=0)
        s1.close()
        s = make(p, mb, wavelet)
        s.comm_init(rank, world)
        d = DistributedWabbit(s, fw, rank, world)
        assert d.in_library
        off = sum(fw.n_active(r) for r in range(rank))
        n = fw.n_active(rank)
        h = np.zeros(s.host_shape())
        h[:n] = u[off:off + n]
        s.upload(h)
        nb2 = d.refine_tree().n_blocks
        _, _, dt2 = d.timeStep_tree(0.0, 0)
        _, m0, m1 = d.adapt_tree(eps=p.eps, Jmin=1)
        _, l, x, _ = d.forest.active(rank)
        off2 = sum(d.forest.n_active(r) for r in range(rank))
        out = np.zeros(s.host_shape())
        s.download(out, g_sync=0)
        ok = (nb1 == nb2) and (dt1 == dt2) and (n0, n1) == (m0, m1) and np.array_equal(l, lf[off2:off2 + len(l)]) and \
            np.array_equal(x, xf[off2:off2 + len(l)]) and np.array_equal(interior(p, out[:len(l)]), interior(p, ref[off2:off2 + len(l)]))
        print(f"rank {rank}: cycle {wavelet} Bs={Bs} blocks {nb2} -> {m1} (single rank {nb1} -> {n1}) dt={dt2!r} ok={ok}", flush=True)
        s.close()
    elif what == "compression":
        # BASELINE config 5 (post_compression_unit_test.f90): adapt_tree with the full wavelet transformation from the equidistant grid, then
        # refineToEquidistant_tree, one component -- across ranks inside the library against the single-rank driver, bit for bit
        from wabbit_b200 import compression as CP
        wavelet, Jmax, eps = sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
        p = CP.compression_params(wavelet, 16, Jmax)
        mb = 2 * 8 ** Jmax
        f1 = Forest.uniform(3, Jmax, Jmax=Jmax, max_blocks=mb)
        fw = Forest.uniform(3, Jmax, Jmax=Jmax, n_ranks=world, max_blocks=mb)
        hvy, l1, x1, _ = f1.active(0)
        g = p.g
        s1 = make(p, mb, wavelet)
        s1.set_forest(f1)
        u = np.zeros(s1.host_shape())
        u[:len(hvy), 0, g:-g, g:-g, g:-g] = CP.set_block_testing_data(16, l1, x1)
        s1.upload(u)
        fa, n0, nb1 = s1.adapt_tree(f1, eps=eps, Jmin=1, full_tree=True)
        fr = CP.refineToEquidistant_tree(s1, fa, Jmax)
        ref = np.zeros(s1.host_shape())
        s1.download(ref, g_sync=0)
        s1.close()
        s = make(p, mb, wavelet)
        s.comm_init(rank, world)
        d = DistributedWabbit(s, fw, rank, world)
        assert d.in_library
        off = sum(fw.n_active(r) for r in range(rank))
        n = fw.n_active(rank)
        h = np.zeros(s.host_shape())
        h[:n] = u[off:off + n]
        s.upload(h)
        _, m0, nb2 = d.adapt_tree(eps=eps, Jmin=1, full_tree=True)
        CP.refineToEquidistant_tree(d, None, Jmax)
        out = np.zeros(s.host_shape())
        s.download(out, g_sync=0)
        off2 = sum(d.forest.n_active(r) for r in range(rank))
        n2 = d.forest.n_active(rank)
        ok = nb1 == nb2 and d.forest.n_blocks == 8 ** Jmax and np.array_equal(interior(p, out[:n2]), interior(p, ref[off2:off2 + n2]))
        print(f"rank {rank}: compression {wavelet} J={Jmax} eps={eps} Nb {nb2} (single rank {nb1}) ok={ok}", flush=True)
        s.close()
    else:
        raise SystemExit(f"unknown case {what}")
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
