import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np
import oracle as O
import test_multi_halo as T
from util import tg_params
from wabbit_b200.solver import HVY_TMP
world, wavelet, ignore_filter = 3, sys.argv[1] if len(sys.argv) > 1 else "CDF42", False
w = O.setup_wavelet(wavelet)
forest = T.graded_forest(world, seed=21)
p = tg_params(Bs=16, J=forest.Jmax, wavelet_g=w.g_default)
p.wavelet = wavelet
sols, grp = T._make_ranks(forest, world, p, wavelet)
po, grid, nbr = T._global_oracle(forest, world, p)
rng = np.random.default_rng(8)
u = O.alloc(grid, po); u[:] = rng.standard_normal(u.shape)
T._scatter(sols, forest, u)
for s in sols: s.set_ghost_filter(ignore_filter)
grp.exchange_array(0, 0)
ref = u.copy()
O.sync_ghosts_leaf(grid, po, ref, nbr, po.g, po.g, w.X, bool(w.lifted), ignore_filter=ignore_filter, w=w)
off = 0
for r, s in enumerate(sols):
    n = forest.n_active(r)
    ids = np.arange(1, n + 1, dtype=np.int32)
    got = np.zeros(s.host_shape()); got[:n] = u[off:off + n]
    s.download(got, g_sync=p.g, hvy_ids=ids)
    bad = np.argwhere(got[:n] != ref[off:off+n])
    hvy, lvl, ixyz, _ = forest.active(r)
    print("rank", r, "n", n, "mismatches", len(bad))
    if len(bad):
        blks = np.unique(bad[:,0])
        for b in blks[:6]:
            bb = bad[bad[:,0]==b]
            print("  blk", b, "level", lvl[b], "pos", ixyz[b], "n", len(bb), "z", bb[:,2].min(), bb[:,2].max(), "y", bb[:,3].min(), bb[:,3].max(), "x", bb[:,4].min(), bb[:,4].max())
        nb, wn, cnt, lists = s.topology_tables()
        print("  counts", cnt)
        for b in blks[:3]:
            print("  nbr row", nb[b].tolist())
            print("  wnbr row", wn[b].tolist())
    off += n
