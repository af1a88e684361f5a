"""Time isdf_qrcp / the gelsy operator stages on random Hermitian PSD matrices:  python tools/qrcp_bench.py n batch"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fft_isdf_scratch_b200.fftisdf import _get_ops  # noqa: E402

n, batch = int(sys.argv[1]), int(sys.argv[2])
ops = _get_ops(0)
g = torch.Generator(device="cuda").manual_seed(1)
r = int(0.87 * n)
c = torch.randn((batch, r, n), dtype=torch.complex128, device="cuda", generator=g)
c *= (10.0 ** (-9.0 * torch.arange(r, device="cuda") / r))[None, :, None]
a = c.conj().transpose(1, 2) @ c


def timeit(f, reps=2):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


w = torch.empty_like(a)
def run_qr():
    ops.conj_copy(a, w)
    return ops.qrcp(w)
t, (vt, tau, piv, pos) = timeit(run_qr)
byt = batch * 16.0 / 3.0 * n ** 3
print(f"qrcp n={n} batch={batch}: {t:.1f} ms  ({byt / t / 1e6:.0f} GB/s of the 16/3 n^3 GEMV bytes)")
t2, rank = timeit(lambda: ops.gelsy_rank(w, piv, 2.2e-16))
print(f"gelsy_rank: {t2:.1f} ms ranks {rank.cpu().numpy()[:4]}")
st = dict(w=w, vt=vt, tau=tau, piv=piv, pos=pos, rank=rank)
rP = -(-int(rank.max()) // 64) * 64
ops.kernel_events = {}
t3, fac = timeit(lambda: ops.gelsy_operators(st, rP), reps=1)
print(f"gelsy_operators: {t3:.1f} ms (rP={rP})", {k: round(v, 1) for k, v in ops.kernel_ms().items()})
