"""Device restatement of LAPACK zgelsy (the solver behind the reference's lstsq call, fftisdf.py:108) against
(a) the real LAPACK through scipy and (b) the numpy port oracle/gelsy_port.py, piece by piece through the C ABI:
QRCP pivots / R / reflectors, the incremental-condition-estimation rank, the three dense operators, and the
minimum-norm solution itself on rank-deficient Hermitian PSD matrices."""
import numpy as np
import pytest
import scipy.linalg

from oracle import gelsy_port as GP

pytestmark = pytest.mark.gpu
EPS = float(np.finfo(np.float64).eps)


def _ops():
    from fft_isdf_scratch_b200.fftisdf import _get_ops
    return _get_ops(0)


def _psd(n, r, seed, decay=0.0):
    """Hermitian PSD test matrix of rank r (r = n: full rank) with singular values spread over 10^-decay."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((r, n)) + 1j * rng.standard_normal((r, n))
    c *= 10.0 ** (-decay * np.arange(r) / max(r - 1, 1))[:, None]
    return c.conj().T @ c


def _dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _qrcp_dev(a_list):
    """Run isdf_qrcp on a list of equally sized matrices; returns numpy (R [b,n,n] in pivoted order, V [b,n,n]
    with reflector k as COLUMN k, tau, piv)."""
    ops = _ops()
    a = np.stack(a_list)
    w = _dev(a.transpose(0, 2, 1))                     # column-major working copy: w[c][i] = A[i][c]
    vt, tau, piv, pos = ops.qrcp(w)
    w, vt, tau, piv, pos = [t.cpu().numpy() for t in (w, vt, tau, piv, pos)]
    b, n, _ = a.shape
    r = np.zeros_like(a)
    for z in range(b):
        assert np.array_equal(np.sort(piv[z]), np.arange(n)) and np.array_equal(pos[z][piv[z]], np.arange(n))
        rr = w[z][piv[z]].T                             # rr[k][j] = w[piv[j]][k]
        r[z] = np.triu(rr)
    return r, vt.transpose(0, 2, 1), tau, piv


@pytest.mark.parametrize("n,seed", [(5, 1), (33, 2), (70, 3), (200, 4), (300, 5), (520, 6), (1100, 7)])
def test_qrcp_matches_lapack_on_tie_free_matrices(n, seed):
    """General complex matrices with well separated column norms: pivots identical to scipy's zgeqp3, R equal to
    rounding, and Q R = A P with Q rebuilt from the reflectors."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    a *= (1.0 + np.arange(n))[None, ::-1] ** 0.5        # distinct column norms
    r, v, tau, piv = _qrcp_dev([a, a.conj()])
    qr_ref, jp, tau_ref, _, _ = scipy.linalg.lapack.zgeqp3(a)
    assert np.array_equal(piv[0], jp - 1)
    scale = np.abs(qr_ref).max()
    assert np.abs(np.abs(r[0]) - np.abs(np.triu(qr_ref))).max() < 1e-12 * scale * n ** 0.5
    # Q from the reflectors: Q = H_0 H_1 ... H_{n-1}
    q = np.eye(n, dtype=complex)
    for k in range(n - 1, -1, -1):
        q -= tau[0][k] * np.outer(v[0][:, k], v[0][:, k].conj() @ q)
    assert np.abs(q.conj().T @ q - np.eye(n)).max() < 1e-12
    assert np.abs(q @ r[0] - a[:, piv[0]]).max() < 1e-12 * scale * n ** 0.5
    # second batch member (conjugated input): conjugated factors, same pivots
    assert np.array_equal(piv[1], piv[0])
    assert np.abs(r[1] - r[0].conj()).max() < 1e-12 * scale * n ** 0.5


@pytest.mark.parametrize("n,r,decay,seed", [(30, 30, 3.0, 1), (64, 40, 8.0, 2), (200, 120, 10.0, 3),
                                             (300, 300, 7.0, 4), (300, 300, 14.0, 4), (520, 410, 12.0, 5), (130, 7, 2.0, 6)])
def test_gelsy_rank_matches_lapack(n, r, decay, seed):
    ops = _ops()
    a = _psd(n, r, seed, decay)
    w = _dev(a.conj()[None])                           # Hermitian: A^T = conj(A)
    st = ops.gelsy_qr(_dev(a[None]), EPS)
    rank_dev = int(st["rank"].cpu()[0])
    res = scipy.linalg.lstsq(a, a[:, :2], lapack_driver="gelsy")
    # the port applied to the DEVICE's own R must give the device's rank exactly (same recurrence) ...
    piv = st["piv"].cpu().numpy()[0]
    rr = np.triu(st["w"].cpu().numpy()[0][piv].T)
    assert rank_dev == GP.gelsy_rank(rr)
    # ... and LAPACK's rank on its own R can differ only by where an eps-level plateau of |R_kk| is cut
    d = np.abs(np.diag(rr))
    lo, hi = min(rank_dev, res[2]), max(rank_dev, res[2])
    assert lo == hi or d[lo - 1] / d[0] < 1e2 * EPS, (rank_dev, res[2], d[lo - 1] / d[0])
    del w


@pytest.mark.parametrize("n,r,decay,seed", [(30, 30, 2.0, 11), (64, 40, 6.0, 12), (200, 150, 9.0, 13),
                                             (520, 380, 9.0, 14)])
def test_gelsy_operators_and_solution(n, r, decay, seed):
    """x = E U^-H (Q1 D^-1)^H b equals LAPACK's minimum-norm solution; the operators have the stated structure."""
    import torch
    ops = _ops()
    a = _psd(n, r, seed, decay)
    rng = np.random.default_rng(seed + 100)
    b = a @ (rng.standard_normal((n, 9)) + 1j * rng.standard_normal((n, 9)))      # consistent right-hand sides
    st = ops.gelsy_qr(_dev(np.stack([a, a])), EPS)
    rank = st["rank"].cpu().numpy()
    rP = max(64, -(-int(rank.max()) // 64) * 64)
    fac = ops.gelsy_operators(st, rP, debug=True)
    assert np.array_equal(fac["chol_rank"].cpu().numpy(), rank)
    rk = int(rank[0])
    q1s, eh = fac["q1s"].cpu().numpy()[0], fac["eh"].cpu().numpy()[0]
    piv = st["piv"].cpu().numpy()[0]
    rr = np.triu(st["w"].cpu().numpy()[0][piv].T)
    d = np.abs(np.diag(rr))[:rk]
    q1 = q1s[:, :rk] * d[None, :]
    assert np.abs(q1.conj().T @ q1 - np.eye(rk)).max() < 1e-12
    assert np.abs(q1s[:, rk:]).max() == 0.0 if rk < rP else True
    assert np.abs(q1.conj().T @ a[:, piv] - rr[:rk]).max() < 1e-12 * np.abs(a).max() * n ** 0.5
    assert np.abs(eh[:rk] @ eh[:rk].conj().T - np.eye(rk)).max() < 1e-12
    # device solve, two-step form:  x = E U^-H ((Q1 D^-1)^H b)
    bt = _dev(np.stack([b, b]))
    c = ops.gemm_hn(fac["q1s"], bt)                                        # [2, rP, 9]
    cpad = torch.zeros((2, rP, 64), dtype=torch.complex128, device=c.device)
    cpad[:, :, :9] = c
    for lf in fac["lfwd"]:                                                  # U^-H = U2^-H U1^-H
        ops.trsm_sweep(lf, cpad, nact=int(rank.max()), backward=False, ng=9)
    x2 = ops.gemm_hn(fac["eh"], cpad.contiguous())[:, :, :9].cpu().numpy()
    # ... and the fused operator the build uses:  x = E (G b),  G = U^-H D^-1 Q1^H
    x = ops.gemm_hn(fac["eh"], ops.gemm_nn(fac["gt"], bt)).cpu().numpy()
    d_all = np.abs(np.diag(rr))[:rk]
    assert np.abs(x - x2).max() < max(1e-11, 50 * d_all[0] / d_all[-1] * EPS) * np.abs(x2).max(), np.abs(x - x2).max() / np.abs(x2).max()
    gt = fac["gt"].cpu().numpy()[0]
    assert np.abs(gt[rk:]).max() == 0.0 if rk < rP else True
    ref = scipy.linalg.lstsq(a, b, lapack_driver="gelsy")
    xp, rkp = GP.gelsy(a, b)
    # residual and (when the ranks agree) the solution itself
    assert np.abs(a @ x[0] - b).max() < 1e-9 * np.abs(b).max()
    assert np.abs(x[0] - x[1]).max() == 0.0
    if rk == ref[2]:
        nrm = np.abs(ref[0]).max()
        cond_r = d[0] / d[-1]
        assert np.abs(x[0] - ref[0]).max() < max(1e-10, 50 * cond_r * EPS) * nrm, (np.abs(x[0] - ref[0]).max() / nrm, cond_r)


@pytest.mark.parametrize("n,sizes", [(300, ("4", "8", "16")), (1100, ("4", "8", "16")), (3120, ("8", "16"))])
def test_qrcp_is_bit_identical_for_every_cluster_size(n, sizes):
    """The cluster size follows the number of matrices per GPU (i.e. the number of GPUs); the factorisation must not
    depend on it, so that an N-GPU build takes the very same rank decisions as the single-GPU build.  n = 3120 (the
    NiO metric) is where the shared-memory budget would allow a wider panel for 16-CTA clusters than for 8: the panel
    width is pinned to the 8-CTA geometry (the 8-GPU NiO build once cut the eps-plateau at other ranks because of it)."""
    import os
    ops = _ops()
    a = _psd(n, int(0.8 * n), 77, 9.0)
    outs = []
    for cs in sizes:
        os.environ["ISDF_QR_CS"] = cs
        try:
            st = ops.gelsy_qr(_dev(a[None]), EPS)
        finally:
            del os.environ["ISDF_QR_CS"]
        outs.append([st[k].cpu().numpy() for k in ("w", "vt", "tau", "piv", "rank")])
    for o in outs[1:]:
        for x, y in zip(outs[0], o):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("n,r", [(64, 64), (100, 100), (200, 131), (448, 448), (130, 64)])
def test_blocked_unpivoted_cholesky(n, r):
    """isdf_chol_nopivot (the Cholesky of the Cholesky-QR step): U^H U = A for well-conditioned Hermitian matrices, and a
    clean stop at the first zero pivot when the trailing rows/columns are exactly zero (rank < n)."""
    ops = _ops()
    rng = np.random.default_rng(n + r)
    c = rng.standard_normal((2, r + 20, r)) + 1j * rng.standard_normal((2, r + 20, r))
    a = np.zeros((2, n, n), dtype=complex)
    a[:, :r, :r] = np.einsum("bki,bkj->bij", c.conj(), c) / (r + 20)
    u, piv, rank = ops.chol_nopivot(_dev(a))
    u, rank = u.cpu().numpy(), rank.cpu().numpy()
    assert rank.tolist() == [r, r] and np.array_equal(piv.cpu().numpy()[0], np.arange(n))
    for z in range(2):
        assert np.abs(np.tril(u[z], -1)).max() == 0.0 and (r == n or np.abs(u[z][r:]).max() == 0.0)
        assert np.abs(u[z].conj().T @ u[z] - a[z]).max() < 1e-13 * np.abs(a[z]).max()
        ref = np.linalg.cholesky(a[z][:r, :r]).conj().T
        assert np.abs(u[z][:r, :r] - ref).max() < 1e-12 * np.abs(ref).max()
