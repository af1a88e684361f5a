"""Parity at BASELINE.json sizes: the bench workloads themselves (diamond stand-in, Gamma and 3x3x3; NiO-AFM stand-in
2x2x2: n0 = 3375 candidates, nip = 520 / 3120, ng = 50 653 / 35 937) through ISDF(cell, kpts).build() against the
numpy/scipy oracle on the same AO tables.

The oracle's lstsq(gelsy) on all q is minutes of CPU at these sizes, so it runs on a few sampled q (all of them for
the Gamma case) and the comparison uses quantities that need only W_q contracted with AO pair products, exactly as
the reference's acceptance test does (fftdf-with-k-lstsq.py:219-258):
  * interpolation-point indices identical (scipy dpstrf on the 3375 x 3375 selection matrix);
  * zgelsy's rank for the sampled q: equal, or different only by where an eps-level plateau of |R_kk| is cut;
  * reconstructed ERIs (a fixed subset of AO pairs) device vs oracle within 10x of the reference solver's own
    reproducibility floor (scipy's gelsy re-run on the symmetrically permuted, i.e. mathematically identical,
    system: at NiO size LAPACK itself moves its rank by a few and the fitted densities by ~4e-6), and against the
    EXACT pair-density ERIs for both arms with err_device <= 2 err_oracle;
  * Gamma case: J, K and E_x as well.
Set ISDF_SKIP_HEAVY=1 to skip (the NiO case alone is ~1.5 min of host LAPACK).
"""
import os
import sys
import time

import numpy as np
import pytest
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import isdf_oracle as O   # noqa: E402
from oracle import pbc_helpers as H   # noqa: E402

pytestmark = pytest.mark.gpu
NPAIR = 6          # AO subset per index of the ERI check: NPAIR^2 pair densities on each side


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _oracle_y_q(f_all, xip, phase, q, blksize=8000):
    """fftisdf.py:72-85 for ONE q (the reference writes all nk to HDF5): y[q] [ng, nip]."""
    nk, ng, _ = f_all.shape
    out = np.empty((ng, xip.shape[1]), dtype=np.complex128)
    for g0 in range(0, ng, blksize):
        g1 = min(ng, g0 + blksize)
        fx_k = np.asarray([f.conj() @ x.T for f, x in zip(f_all[:, g0:g1], xip)])      # :76
        fx_s = (phase @ fx_k.reshape(nk, -1))                                          # :79
        assert abs(fx_s.imag).max() < 1e-10 * max(1.0, abs(fx_s.real).max())           # :81
        y_s = fx_s * fx_s                                                              # :83
        out[g0:g1] = (phase.T[q] @ y_s).reshape(g1 - g0, -1)                           # :84, row q only
    return out


def _pair_eri(theta_or_w, x, ks, q, a, kpts, mesh, coord, from_w):
    """ERIs (NPAIR^4 block) for the quadruple ks = (k1,k2,k3,k4) from W_q (device) or from Theta_q (oracle; the
    fitted pair densities are contracted with the Coulomb kernel exactly as fftisdf.py:113-121 does)."""
    k1, k2, k3, k4 = ks
    s = slice(0, NPAIR)
    lhs = np.einsum("Im,In->mnI", x[k1][:, s].conj(), x[k2][:, s]).reshape(NPAIR * NPAIR, -1)
    rhs = np.einsum("Jk,Jl->klJ", x[k3][:, s].conj(), x[k4][:, s]).reshape(NPAIR * NPAIR, -1)
    if from_w:
        return lhs @ theta_or_w @ rhs.T
    th = theta_or_w
    ng, vol = len(coord), abs(np.linalg.det(a))
    fq = np.exp(-1j * coord @ kpts[q])
    zl = lhs @ th                                                    # fitted pair densities [pairs, ng]
    zeta = H.ifft(H.fft(zl * fq, mesh) * H.get_coulG(a, kpts[q], mesh) * vol / ng, mesh) * fq.conj()
    return zeta @ (rhs.conj() @ th).conj().T                         # zeta_q @ z_q^H, contracted with the pair products


def _exact_eri(f_all, ks, q, a, kpts, mesh, coord):
    k1, k2, k3, k4 = ks
    s = slice(0, NPAIR)
    ng, vol = len(coord), abs(np.linalg.det(a))
    fq = np.exp(-1j * coord @ kpts[q])
    rho12 = np.einsum("gm,gn->mng", f_all[k1][:, s].conj(), f_all[k2][:, s]).reshape(-1, ng)
    zeta = H.ifft(H.fft(rho12 * fq, mesh) * H.get_coulG(a, kpts[q], mesh) * vol / ng, mesh) * fq.conj()
    rho34 = np.einsum("gk,gl->klg", f_all[k3][:, s].conj(), f_all[k4][:, s]).reshape(-1, ng)
    return zeta @ rho34.T


@pytest.mark.skipif(os.environ.get("ISDF_SKIP_HEAVY") == "1", reason="ISDF_SKIP_HEAVY=1")
@pytest.mark.parametrize("workload,qs,floor_q", [("diamond-standin-gamma", [0], True),
                                                 ("diamond-standin-k333", [0, 5, 13], True),
                                                 ("nio-afm-standin-k222", [3], True)])
def test_bench_workload_against_oracle(workload, qs, floor_q):
    import itertools
    import bench
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    from fft_isdf_scratch_b200.fftisdf import _get_ops
    t_start = time.time()
    cell, kpts, w = bench.make_workload(workload)
    x0, f_all, coord = bench.ao_tables(cell, kpts, w["m0"], ops=_get_ops(0))
    kmesh, mesh, a = w["kmesh"], cell.mesh, cell.a
    nk = len(kpts)
    df = fftisdf.ISDF(cell, kpts, m0=w["m0"], c0=w["c0"])
    df.set_ao_tables(x0=x0, f_all=f_all)
    df.keep_metric = True
    df.build()
    kpts = df.kpts
    nip = df._x.shape[1]
    # ---- selection (fftisdf.py:357-388) with scipy's dpstrf on the same x0
    xip, mask, rank0, chol, x4 = O.select_interpolation_points(x0, w["c0"], return_info=True)
    # The stand-in cells are symmetric: diag(x4) has exact ties, which LAPACK itself resolves by its summation order
    # (SURVEY 7-2).  Same point SET; any difference in ORDER only exchanges points with equal diagonal entries.
    assert np.array_equal(np.sort(df._mask), np.sort(mask)), "interpolation point set differs from dpstrf's"
    dg = np.diag(x4)
    assert np.abs(dg[df._mask] - dg[mask]).max() < 1e-10 * dg.max()
    n_swapped = int((df._mask != mask).sum())
    reorder = np.argsort(df._mask)[np.argsort(np.argsort(mask))]     # device row of every oracle row
    assert np.array_equal(df._x[:, reorder, :], xip)
    phase = H.get_phase(a, kpts, kmesh)
    x4_k = O.build_metric(xip, phase)
    tr = pk.pbc_tools.time_reversal_partner(kmesh)
    ranks_dev = np.zeros(nk, dtype=int)
    for s, q in enumerate(df._qind):
        ranks_dev[q] = ranks_dev[tr[q]] = df._ranks[s]
    kidx = np.array(list(itertools.product(*[range(n) for n in kmesh])))
    find = lambda v: int(np.where((kidx == np.mod(v, kmesh)).all(1))[0][0])
    wq = df._wq[:, reorder][:, :, reorder]                           # oracle row order
    report = [dict(points_in_tie_swapped_order=n_swapped)]
    floor0 = 0.0
    for iq, q in enumerate(qs):
        y_q = _oracle_y_q(f_all, xip, phase, q)
        res = scipy.linalg.lstsq(x4_k[q], y_q.T, lapack_driver="gelsy")                 # :108
        th, rank_ref = res[0], int(res[2])
        # rank: equal, or both cuts inside the eps plateau of |R_kk|
        rd = np.abs(np.diag(scipy.linalg.lapack.zgeqp3(x4_k[q])[0]))
        lo, hi = min(rank_ref, ranks_dev[q]), max(rank_ref, ranks_dev[q])
        assert lo == hi or rd[lo - 1] / rd[0] < 50 * 2.2e-16, (q, rank_ref, ranks_dev[q], rd[lo - 1] / rd[0])
        k1 = 0
        k2 = find(kidx[k1] + kidx[q])              # pair momentum k2 - k1 = q
        k3 = nk // 2
        k4 = find(kidx[k3] - kidx[q])
        ks = (k1, k2, k3, k4)
        e_dev = _pair_eri(wq[q], xip, ks, q, a, kpts, mesh, coord, True)
        e_ora = _pair_eri(th, xip, ks, q, a, kpts, mesh, coord, False)
        e_exact = _exact_eri(f_all, ks, q, a, kpts, mesh, coord)
        d_dev_ora = rel(e_dev, e_ora)
        err_dev, err_ora = rel(e_dev, e_exact), rel(e_ora, e_exact)
        floor, rank_alt = None, None
        if floor_q and iq == 0:
            # Reproducibility floor of the REFERENCE's own solver: scipy's lstsq(gelsy) on the same system with rows and
            # columns of A_q (and the rows of Y^T) permuted symmetrically -- the identical problem in exact arithmetic,
            # only the summation / tie order changes, which is what another BLAS or thread count does to the reference.
            perm = np.random.default_rng(9).permutation(nip)
            res2 = scipy.linalg.lstsq(x4_k[q][perm][:, perm], y_q.T[perm], lapack_driver="gelsy")
            th2 = np.empty_like(res2[0])
            th2[perm] = res2[0]
            floor = rel(_pair_eri(th2, xip, ks, q, a, kpts, mesh, coord, False), e_ora)
            rank_alt = int(res2[2])
            del res2, th2
            # (the device's A_q itself equals numpy's to rounding)
            sq = df._qind.index(q) if q in df._qind else None
            aq = df._a_q[sq].cpu().numpy() if sq is not None else df._a_q[df._qind.index(int(tr[q]))].cpu().numpy().conj()
            assert np.abs(aq[reorder][:, reorder] - x4_k[q]).max() < 1e-13 * np.abs(x4_k[q]).max()
        report.append(dict(q=q, rank_dev=int(ranks_dev[q]), rank_gelsy=rank_ref, nip=nip, eri_dev_vs_oracle=d_dev_ora,
                           eri_err_dev=err_dev, eri_err_oracle=err_ora, floor=floor,
                           rank_gelsy_permuted=(rank_alt if floor is not None else None)))
        print("\n", workload, report[-1], flush=True)
        assert err_dev < 1e-4 and err_dev <= 2 * err_ora + 1e-12, report[-1]           # reference's acceptance test
        # device vs oracle: within 10x of the reference solver's own reproducibility floor -- or, when zgelsy cut the
        # eps-plateau of |R_kk| elsewhere than LAPACK did on this box (rounding decides that; LAPACK's own cut moves
        # under the permutation above as well), at least 4x below the reference's own error against the exact ERIs
        if floor is not None:
            floor0 = floor
        bound = 10 * max(floor0, 1e-9)
        if floor is None or rank_ref != ranks_dev[q] or rank_alt != rank_ref:
            bound = max(bound, 0.5 * err_ora)
        assert d_dev_ora < bound, report[-1]
        if nk == 1:
            # Gamma: the whole consumer chain, against the oracle's W
            fq = np.exp(-1j * coord @ kpts[q])
            b = H.fft(th * fq, mesh) * np.sqrt(H.get_coulG(a, kpts[q], mesh) * abs(np.linalg.det(a))) / len(coord)
            w_ora = (b @ b.conj().T)[None]
            rng = np.random.default_rng(3)
            nao = xip.shape[-1]
            dm = rng.standard_normal((1, nao, nao))
            dm = dm + dm.transpose(0, 2, 1)
            vj, vk = df.get_jk(dm, kpts=kpts)
            vj_o = O.get_j_kpts(xip, w_ora[0], dm[None].astype(complex))[0]
            vk_o = O.get_k_kpts(xip, w_ora, dm[None].astype(complex), phase)[0]
            dj, dk = rel(vj, vj_o), rel(vk, vk_o)
            report[-1].update(J=dj, K=dk)
            tol = max(10 * max(floor, 1e-9), 1e-8)
            assert dj < tol and dk < tol, report[-1]
            ex = O.exchange_energy(np.asarray(vk)[None].astype(complex), dm[None].astype(complex))
            ex_o = O.exchange_energy(vk_o[None], dm[None].astype(complex))
            assert abs(ex - ex_o) < tol * abs(ex_o)
    print("\n%s (%.0f s): %s" % (workload, time.time() - t_start, report))
