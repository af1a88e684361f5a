"""Diagnostics of the device zgelsy operators at a bench-workload size (not a test; uses torch.linalg as comparator).
    python tools/gelsy_diag.py nio-afm-standin-k222 3"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from fft_isdf_scratch_b200 import fftisdf  # noqa: E402
from fft_isdf_scratch_b200.fftisdf import _get_ops  # noqa: E402

name, q = sys.argv[1], int(sys.argv[2])
cell, kpts, w = bench.make_workload(name)
ops = _get_ops(0)
x0, f_all, coord = bench.ao_tables(cell, kpts, w["m0"], ops=ops)
df = fftisdf.ISDF(cell, kpts, m0=w["m0"], c0=w["c0"])
df.set_ao_tables(x0=x0, f_all=f_all)
df.keep_metric = True
df.build()
s = df._qind.index(q)
a = df._a_q[s:s + 1].contiguous()
n = a.shape[1]
eps = float(np.finfo(np.float64).eps)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
st = ops.gelsy_qr(a.clone(), eps)
ev1.record()
torch.cuda.synchronize()
print("qrcp+rank ms", ev0.elapsed_time(ev1))
rank = int(st["rank"].cpu()[0])
rP = -(-rank // 64) * 64
ev0.record()
fac = ops.gelsy_operators(st, rP)
ev1.record()
torch.cuda.synchronize()
print("operators ms", ev0.elapsed_time(ev1), "rank", rank, "chol rank", int(fac["chol_rank"].cpu()[0]))
piv = st["piv"][0].long()
wm = st["w"][0]                      # w[c][k] = R[k][c]
r = torch.triu(wm[piv].T)            # pivoted order
d = r.diagonal().abs()
print("|R_kk| at 0, rank-1, rank:", d[0].item(), d[rank - 1].item(), d[min(rank, n - 1)].item())
q1 = fac["q1s"][0][:, :rank] * d[:rank][None, :]
I = torch.eye(rank, dtype=q1.dtype, device=q1.device)
print("Q1 orthonormality", (q1.conj().T @ q1 - I).abs().max().item())
ap = a[0][:, piv]
print("Q1^H A P - R1 (rel)", ((q1.conj().T @ ap) - r[:rank]).abs().max().item() / d[0].item())
eh = fac["eh"][0][:rank]
print("E orthonormality", (eh @ eh.conj().T - I).abs().max().item())
rhat = (r[:rank] / d[:rank, None])
inv = torch.empty_like(piv); inv[piv] = torch.arange(n, device=piv.device)
rhat_o = rhat[:, inv]                # original column order
sv = torch.linalg.svdvals(rhat)
print("cond(Rhat1)", (sv[0] / sv[-1]).item())
# row space: projector of eh vs projector from QR of rhat^H
qz, _ = torch.linalg.qr(rhat_o.conj().T)
print("row-space projector diff", ((eh.conj().T @ eh) - (qz @ qz.conj().T)).abs().max().item())
# full QR identity: A = Q R P^T  using the reflectors
vt, tau = st["vt"][0], st["tau"][0]
qfull = torch.eye(n, dtype=a.dtype, device=a.device)
for k in range(n - 1, -1, -1):
    v = vt[k]
    qfull -= tau[k] * torch.outer(v, v.conj() @ qfull)
print("Q orthonormality (all n)", (qfull.conj().T @ qfull - torch.eye(n, dtype=a.dtype, device=a.device)).abs().max().item())
print("Q R - A P (rel)", (qfull @ r - ap).abs().max().item() / a.abs().max().item())
# compare with LAPACK's R diagonal (host)
import scipy.linalg
qr_ref, jp, _, _, _ = scipy.linalg.lapack.zgeqp3(a[0].cpu().numpy())
dr = np.abs(np.diag(qr_ref))
dd = d.cpu().numpy()
same = int((jp - 1 == piv.cpu().numpy()).sum())
first_diff = int(np.argmax(jp - 1 != piv.cpu().numpy())) if same < n else n
print("pivots equal to LAPACK's:", same, "of", n, "first difference at", first_diff, "|R_kk| there / R_00", dd[min(first_diff, n - 1)] / dd[0])
print("max rel diff of |R_kk| over the first `rank`:", np.abs(dd[:rank] - dr[:rank]).max() / dr[0],
      " log-ratio spread", np.abs(np.log(dd[:rank] / dr[:rank])).max())
