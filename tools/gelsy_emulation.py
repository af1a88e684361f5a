"""Design study (CPU, numpy): what parity does a from-scratch restatement of LAPACK zgelsy reach on the
rank-deficient golden, compared with the truncated-Cholesky fit of round 1?

    python tools/gelsy_emulation.py [golden-name]

Prints, per variant, the per-q ranks and the relative deviation of K, J and the reconstructed ERIs from the
reference's own output (tests/golden/ref_<name>.npz).  Nothing here is on the product path.
"""
import os
import sys

import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import isdf_oracle as O  # noqa: E402
from oracle import pbc_helpers as H  # noqa: E402
from oracle import gelsy_port as GP  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def main(name="k222_sp"):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    a, kpts, kmesh, mesh = g["a"], g["kpts"], g["kmesh"].tolist(), g["mesh"].tolist()
    out = O.build(a, kpts, kmesh, mesh, g["x0"], g["f_all"], g["coord"], float(g["c0"]), keep_theta=True)
    ph = H.get_phase(a, kpts, kmesh)
    gv = H.get_Gv(a, mesh)
    vol = abs(np.linalg.det(a))
    ng = len(g["coord"])
    nk = len(kpts)
    x = out["x"]
    dms = g["dm"][None]
    vk_ref = g["vk"].reshape(g["dm"].shape)
    vj_ref = g["vj"].reshape(g["dm"].shape)
    print("oracle ranks", out["ranks"], " oracle-vs-golden W", rel(out["wq"], g["wq"]))

    def finish(thetas, label, ranks):
        wq = []
        for q in range(nk):
            fq = np.exp(-1j * g["coord"] @ kpts[q])
            b = H.fft(thetas[q] * fq, mesh) * np.sqrt(H.get_coulG(a, kpts[q], mesh, Gv=gv) * vol) / ng
            wq.append(b @ b.conj().T)
        wq = np.asarray(wq)
        vk = O.get_k_kpts(x, wq, dms, ph)[0]
        vj = O.get_j_kpts(x, wq[0], dms)[0]
        import itertools
        kidx = np.array(list(itertools.product(*[range(n) for n in kmesh])))
        find = lambda v: int(np.where((kidx == np.mod(v, kmesh)).all(1))[0][0])
        worst = 0.0
        for k1 in range(nk):
            for k2 in range(nk):
                q = find(kidx[k2] - kidx[k1])
                for k3 in range(0, nk, 3):
                    k4 = find(kidx[k3] - kidx[q])
                    e1 = O.eri_from_w(wq[q], x[k1], x[k2], x[k3], x[k4])
                    e0 = O.eri_from_w(g["wq"][q], x[k1], x[k2], x[k3], x[k4])
                    worst = max(worst, rel(e1, e0))
        print(f"{label:34s} ranks {list(ranks)}  K {rel(vk, vk_ref):.2e}  J {rel(vj, vj_ref):.2e}  ERI {worst:.2e}"
              f"  W {rel(wq, g['wq']):.2e}")

    rng = np.random.default_rng(0)
    # (0) the reference's own noise floor: eps-level perturbation of A_q before gelsy
    th, rk = [], []
    for q in range(nk):
        aq = out["x4_k"][q] * (1 + 1e-16 * rng.standard_normal(out["x4_k"][q].shape))
        r = scipy.linalg.lstsq(aq, out["y"][q].T, lapack_driver="gelsy")
        th.append(r[0]); rk.append(r[2])
    finish(th, "scipy gelsy, A*(1+1e-16 noise)", rk)
    # (1) from-scratch gelsy port (own Householder QRCP, ICE, RZ)
    th, rk = [], []
    for q in range(nk):
        t, r = GP.gelsy(out["x4_k"][q], out["y"][q].T)
        th.append(t); rk.append(r)
    finish(th, "gelsy port (numpy, unblocked)", rk)
    # (1b) the factored form the GPU path uses: Theta = E (T11^-1 (Q1^H Y^T)), W = E W~ E^H
    th, rk = [], []
    for q in range(nk):
        f = GP.gelsy_factor(out["x4_k"][q])
        c = f["q1"].conj().T @ out["y"][q].T
        c = scipy.linalg.solve_triangular(f["t11"], c)
        th.append(f["e"] @ c); rk.append(f["rank"])
    finish(th, "gelsy port, factored (E, T11, Q1)", rk)
    # (2) truncated pivoted Cholesky (round-1 GPU algorithm), pstrf rule
    for lab, rule in [("pchol basic, n*eps*max rule", "pstrf"), ("pchol basic, eps ratio rule", "eps")]:
        th, rk = [], []
        for q in range(nk):
            t, r = GP.pchol_basic(out["x4_k"][q], out["y"][q].T, rule)
            th.append(t); rk.append(r)
        finish(th, lab, rk)
    # (3) pivoted Cholesky at the eps rule + minimum-norm / least-squares finish
    for lab, mode in [("pchol eps, min-norm projection", "proj"), ("pchol eps, LS + min-norm (U^+ U^+H)", "ls")]:
        th, rk = [], []
        for q in range(nk):
            t, r = GP.pchol_minnorm(out["x4_k"][q], out["y"][q].T, mode)
            th.append(t); rk.append(r)
        finish(th, lab, rk)


if __name__ == "__main__":
    main(*sys.argv[1:])
