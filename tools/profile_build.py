"""Lean driver for ncu: two ISDF builds of a bench workload with device-resident tables (one warm, one to
capture) and nothing else -- no comparator GEMMs, no end-to-end arm.  `python tools/profile_build.py [workload]`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fft_isdf_scratch_b200 import fftisdf  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "diamond-standin-k333"
    nbuild = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cell, kpts, w = bench.make_workload(name)
    x0, f_all, _ = bench.ao_tables(cell, kpts, w["m0"], ops=fftisdf._get_ops(0))
    df = fftisdf.ISDF(cell, kpts, m0=w["m0"], c0=w["c0"], device=0)
    df.set_ao_tables(x0=torch.from_numpy(x0).to("cuda:0"), f_all=torch.from_numpy(f_all).to("cuda:0"))
    for _ in range(nbuild):
        df.build()
        print({k: round(v, 2) for k, v in df._stage_ms.items()}, "launches", df._ops.launches)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
