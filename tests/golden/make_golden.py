"""Generate the small golden fixtures under tests/golden/ from the reference checkout.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The GPU box has no /root/reference, so tests only read the .npz files written here.

taylor_green_{FD2_CDF20,FD4_CDF40,FD6_CDF60}.npz : strided samples (every 3rd interior point) of the reference's own
  regression fields TESTING/acm/taylorGreen/*/{ux,uy,uz,p}_000010000000.h5 (t = 10, written by the reference
  Fortran code), keyed by zero-based block coordinates, plus the iteration count stored in the file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
from h5lite import read_wabbit  # noqa: E402

REF = "/root/reference/TESTING/acm/taylorGreen"
STRIDE = 3


def main():
    for case in ("taylorGreenEqui_FD2_CDF20", "taylorGreenEqui_FD4_CDF40", "taylorGreenEqui_FD6_CDF60"):
        out = {}
        for tag, key in (("000000000000", "t0"), ("000010000000", "t1")):
            fields = []
            for name in ("ux", "uy", "uz", "p"):
                d = read_wabbit(os.path.join(REF, case, f"{name}_{tag}.h5"))
                Bs = int(d["attrs"]["block-size"][0])
                nb = d["blocks"].shape[0]
                # coords_origin is stored (z,y,x); block coordinate = origin / (Bs*dx)
                ixyz = np.rint(d["origin"][:, ::-1] / (d["spacing"][:, ::-1] * Bs)).astype(np.int32)
                order = np.lexsort((ixyz[:, 0], ixyz[:, 1], ixyz[:, 2]))
                blocks = d["blocks"][order][:, :Bs:STRIDE, :Bs:STRIDE, :Bs:STRIDE]
                fields.append(blocks)
                out[f"{key}_ixyz"] = ixyz[order]
                out[f"{key}_time"] = d["attrs"]["time"]
                out[f"{key}_iteration"] = d["attrs"]["iteration"]
                out["Bs"] = np.array([Bs])
            out[key] = np.stack(fields, axis=1)   # [block, comp, z, y, x] strided
        out["stride"] = np.array([STRIDE])
        path = os.path.join(HERE, case.replace("taylorGreenEqui", "taylor_green") + ".npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), out["t1"].shape, out["t1_iteration"])


if __name__ == "__main__":
    main()
