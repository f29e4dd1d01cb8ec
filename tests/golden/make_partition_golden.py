"""Golden fixture for the space-filling-curve partition: per-rank block lists of files the reference wrote with several
MPI ranks (datasets `procs`, `level`, `block_treecode_num`; attribute max_level).  balanceLoad_tree sorts the leaves along the Hilbert
curve (treecode_to_hilbertcode_2D/3D) and hands out contiguous chunks, so `procs` pins the curve itself.

Run in the build container (where /root/reference exists):  python tests/golden/make_partition_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from h5lite import H5Lite  # noqa: E402

FILES = [
    "acm/acm_CDF40/ux_000000050000.h5", "acm/acm_CDF40/ux_000000100000.h5", "acm/acm_CDF44/ux_000000050000.h5",
    "acm/acm_CDF44/ux_000000100000.h5", "acm/acm_norm_CDF44/ux_000000200000.h5", "acm/acm_significant_CDF44/ux_000000200000.h5",
    "acm/taylorGreen/taylorGreenEqui_FD4_CDF40/ux_000010000000.h5", "acm/bumblebeeFlowEquiFD4_CDF40/ux_000002000000.h5",
    "conv/blob_equi_3D_CDF40/phi1_000000500000.h5", "conv/blob_adaptive_3D_CDF40/phi1_000000050000.h5",
    "conv/blob_adaptive_3D_CDF22/phi1_000000050000.h5", "conv/blob_adaptive_3D_CDF44/phi1_000000050000.h5",
]


def main():
    out = {"files": np.array(FILES)}
    for k, f in enumerate(FILES):
        h = H5Lite(os.path.join("/root/reference/TESTING", f))
        at = h.attrs("blocks")
        out[f"f{k}_dim"] = np.asarray(at["dim"], dtype=np.int32)
        out[f"f{k}_Jmax"] = np.asarray(at["max_level"], dtype=np.int32)
        out[f"f{k}_procs"] = h.read("procs").astype(np.int32)
        out[f"f{k}_level"] = h.read("level").astype(np.int32)
        out[f"f{k}_treecode"] = h.read("block_treecode_num").astype(np.int64)
    path = os.path.join(HERE, "partition_procs.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
