"""TEST INFRASTRUCTURE (oracle) -- never imported by the shipped path.

numpy/scipy restatement of the reference's FFT-ISDF build, written against *tensors*
(AO tables, lattice, k-points, mesh) instead of a PySCF cell so it runs without PySCF.
Each block cites the reference lines it follows.  It issues the same library calls the
reference does (numpy `@` -> ZGEMM, scipy dpstrf, scipy lstsq(gelsy), numpy fftn/ifftn).

Pinning: the arithmetic of build()/select_interpolation_points() is pinned against the
reference's OWN code, executed in the build container by oracle/gen_golden.py (which imports
/root/reference/fftisdf.py with PySCF stubbed by oracle/pbc_helpers.py) -> tests/golden/*.npz.
The PySCF helper semantics themselves (get_coulG wrap-around, get_Gv ordering, fft sign)
remain *unpinned* because PySCF is not available: "parity unpinned at the PySCF boundary".
"""
import time

import numpy
import scipy.linalg

from . import pbc_helpers as H


def select_interpolation_points(x0, c0, tol=-1.0, return_info=False):
    """fftisdf.py:357-388.  x0: [nk, n0, nao] c128 AO values on the parent grid.
    Returns x0[:, mask, :] like the reference; with return_info also (mask, rank, chol)."""
    x0 = numpy.asarray(x0)
    nkpt, ng, nao = x0.shape                               # :373-374
    x2 = numpy.zeros((ng, ng), dtype=numpy.double)         # :376
    for q in range(nkpt):                                  # :377
        x2 += (x0[q].conj() @ x0[q].T).real                # :378
    x4 = (x2 * x2 / nkpt).real                             # :379
    chol, perm, rank = H.pivoted_cholesky(x4, tol=tol)     # :381-382
    nip = min(int(nao * c0), rank)                         # :383
    mask = perm[:nip]                                      # :384
    if return_info:
        return x0[:, mask, :], mask, rank, chol, x4
    return x0[:, mask, :]                                  # :388


def build_metric(xip, phase):
    """fftisdf.py:38-48 -> x4_k [nk, nip, nip]."""
    nkpt, nip, nao = xip.shape
    nimg = nkpt
    x2_k = numpy.asarray([xq.conj() @ xq.T for xq in xip])             # :38
    x2_s = phase @ x2_k.reshape(nkpt, -1)                              # :41
    x2_s = x2_s.reshape(nimg, nip, nip)                                # :42
    assert abs(x2_s.imag).max() < 1e-10                                # :43
    x4_s = x2_s * x2_s                                                 # :45
    x4_k = phase.conj().T @ x4_s.reshape(nimg, -1)                     # :46
    return x4_k.reshape(nkpt, nip, nip)                                # :47


def build_rhs_block(f_k, xip, phase):
    """fftisdf.py:73-85 for one grid block: f_k [nk, blk, nao] -> y_k [nk, blk, nip]."""
    nkpt, blk, nao = f_k.shape
    nip = xip.shape[1]
    fx_k = numpy.asarray([f.conj() @ x.T for f, x in zip(f_k, xip)])   # :76
    fx_s = phase @ fx_k.reshape(nkpt, -1)                              # :79
    fx_s = fx_s.reshape(nkpt, blk, nip)                                # :80
    assert abs(fx_s.imag).max() < 1e-10                                # :81
    y_s = fx_s * fx_s                                                  # :83
    y_k = phase.T @ y_s.reshape(nkpt, -1)                              # :84
    return y_k.reshape(nkpt, blk, nip)                                 # :85


def fit_and_coulomb_q(x4_q, y_q, fq, coulG_q, vol, mesh):
    """fftisdf.py:99-121 for one q.  y_q: [ngrid, nip].  Returns (w_q, z_q, rank)."""
    ngrid = y_q.shape[0]
    res = scipy.linalg.lstsq(x4_q, y_q.T, lapack_driver="gelsy")      # :108
    z_q = res[0]                                                       # :109
    rank = res[2]                                                      # :110
    zeta_q = H.fft(z_q * fq, mesh)                                     # :113
    zeta_q *= coulG_q                                                  # :114
    zeta_q *= vol / ngrid                                              # :115
    zeta_q = H.ifft(zeta_q, mesh)                                      # :118
    zeta_q *= fq.conj()                                                # :119
    w_q = zeta_q @ z_q.conj().T                                        # :121
    return w_q, z_q, rank


def build(a, kpts, kmesh, mesh, x0, f_all, coord, c0, blksize=8000, qlist=None,
          keep_theta=False, timers=None):
    """fftisdf.py:22-128 on tensors.

    a [3,3] lattice (bohr); kpts [nk,3]; mesh dense FFT mesh; x0 [nk,n0,nao] parent-grid AOs;
    f_all [nk,ng,nao] dense-grid AOs (what aoR_loop yields block by block); coord [ng,3].
    Returns dict(x, mask, rank, x4_k, wq, theta?, ranks).  `qlist` restricts the q loop
    (CPU-baseline sampling); the reference loops over all q.
    """
    a = numpy.asarray(a)
    nkpt = len(kpts)
    phase = H.get_phase(a, kpts, kmesh)                                # :28
    t0 = time.perf_counter()
    xip, mask, rank0, chol, _ = select_interpolation_points(x0, c0, return_info=True)  # :33
    t1 = time.perf_counter()
    nip = xip.shape[1]                                                 # :34
    x4_k = build_metric(xip, phase)                                    # :38-48
    ngrid = coord.shape[0]                                             # :54
    y = numpy.empty((nkpt, ngrid, nip), dtype=numpy.complex128)        # :62 (HDF5 scratch in the reference)
    for g0 in range(0, ngrid, blksize):                                # :72
        g1 = min(ngrid, g0 + blksize)
        y[:, g0:g1, :] = build_rhs_block(f_all[:, g0:g1, :], xip, phase)  # :73-85
    t2 = time.perf_counter()
    gv = H.get_Gv(a, mesh)                                             # :91
    vol = abs(numpy.linalg.det(a))
    wq, ranks, thetas = [], [], []
    qs = range(nkpt) if qlist is None else qlist
    for q in qs:                                                       # :97
        vq = kpts[q]
        fq = numpy.exp(-1j * numpy.dot(coord, vq))                     # :99
        coulG = H.get_coulG(a, vq, mesh, Gv=gv)                        # :114
        w_q, z_q, rank = fit_and_coulomb_q(x4_k[q], y[q], fq, coulG, vol, mesh)
        wq.append(w_q)
        ranks.append(rank)
        if keep_theta:
            thetas.append(z_q)
    t3 = time.perf_counter()
    if timers is not None:
        timers.update(select=t1 - t0, rhs=t2 - t1, fit_coulomb=t3 - t2)
    out = dict(x=xip, mask=mask, rank=rank0, chol_next=float(chol[nip, nip]) if nip < chol.shape[0] else 0.0,
               x4_k=x4_k, wq=numpy.asarray(wq), ranks=ranks, y=y)
    if keep_theta:
        out["theta"] = numpy.asarray(thetas)
    return out


def get_j_kpts(x, w0, dms):
    """fftisdf.py:155-166.  dms [nset,nk,nao,nao] -> vj [nset,nk,nao,nao]."""
    nkpt = x.shape[0]
    rho = numpy.einsum("kIm,kIn,xkmn->xI", x, x.conj(), dms, optimize=True)   # :155
    rho *= 1.0 / nkpt                                                          # :156
    v = numpy.einsum("IJ,xJ->xI", w0, rho, optimize=True)                      # :159
    return numpy.einsum("kIm,kIn,xI->xkmn", x.conj(), x, v, optimize=True)     # :166


def get_k_kpts(x, wq, dms, phase):
    """fftisdf.py:204-227."""
    nkpt, nip, nao = x.shape
    nset = dms.shape[0]
    ws = phase @ wq.reshape(nkpt, -1)                                          # :205
    ws = ws.reshape(nkpt, nip, nip)
    ws = ws.real * numpy.sqrt(nkpt)                                            # :207
    vk_kpts = []
    for dm in dms:                                                             # :210
        rhok = [xx @ d @ xx.conj().T for xx, d in zip(x, dm)]                  # :211
        rhok = numpy.asarray(rhok) / nkpt                                      # :212
        rhos = phase @ rhok.reshape(nkpt, -1)                                  # :215
        assert abs(rhos.imag).max() < 1e-10                                    # :216
        rhos = rhos.real.reshape(nkpt, nip, nip)                               # :217
        vs = ws * rhos.transpose(0, 2, 1)                                      # :219
        vk = phase.T @ vs.reshape(nkpt, -1)                                    # :222
        vk = vk.reshape(nkpt, nip, nip)
        vk_kpts.append([xx.conj().T @ v @ xx for xx, v in zip(x, vk)])         # :225
    return numpy.asarray(vk_kpts).reshape(nset, nkpt, nao, nao)                # :227


def exchange_energy(vk, dms):
    """E_x = -1/4 * sum_k Tr(K_k D_k) / nk (closed shell)."""
    nkpt = vk.shape[1]
    return -0.25 * numpy.einsum("xkmn,xknm->", vk, dms).real / nkpt


def eri_from_w(wq_q, x1, x2, x3, x4):
    """fftdf-with-k-lstsq.py:232: einsum("IJ,Im,In,Jk,Jl->mnkl", c[q], x1*, x2, x3*, x4)."""
    l = numpy.einsum("Im,In->Imn", x1.conj(), x2)
    r = numpy.einsum("Jk,Jl->Jkl", x3.conj(), x4)
    return numpy.einsum("IJ,Imn,Jkl->mnkl", wq_q, l, r, optimize=True)


def trans_2e_bruteforce(x, wq, kmesh, c_ao_emb):
    """Definition-level statement of the embedding ERI the stub at fftisdf.py:230-294 sets up (it stops before the
    contraction): sum over the momentum-conserving quadruples of eri_from_w with xmo = C_ao_emb[k].T @ x[k].T (:287).
    c_ao_emb [nk, nao, nemb] (already carrying the nkpts**-0.75 factor, :275-276).  Test oracle only."""
    import itertools
    nk = len(x)
    kidx = numpy.array(list(itertools.product(*[range(n) for n in kmesh])))
    find = lambda v: int(numpy.where((kidx == numpy.mod(v, kmesh)).all(1))[0][0])
    xe = [x[k] @ c_ao_emb[k] for k in range(nk)]                  # [nip, nemb] = (C^T x^T)^T
    nemb = c_ao_emb.shape[-1]
    eri = numpy.zeros((nemb,) * 4, dtype=numpy.complex128)
    for k1 in range(nk):
        for k2 in range(nk):
            q = find(kidx[k2] - kidx[k1])
            for k3 in range(nk):
                k4 = find(kidx[k3] - kidx[q])
                eri += eri_from_w(wq[q], xe[k1], xe[k2], xe[k3], xe[k4])
    return eri


def flop_model(nk, nao, n0, nip, ng, nq=None):
    """SURVEY.md section 8(d) / BASELINE.md section 3 algorithmic work (real flop; complex MAC = 8)."""
    import math
    nq = nk if nq is None else nq
    sel = 4.0 * nk * n0 * n0 * nao + 1.0 * n0 * nip * nip
    metric = 8.0 * nk * nip * nip * nao
    rhs = 8.0 * nk * ng * nip * nao
    fit = nq * ((4.0 / 3.0) * nip ** 3 + 16.0 * nip * nip * ng)
    fft = nq * 2 * 5.0 * nip * ng * math.log2(ng)
    kern = nq * 8.0 * nip * nip * ng
    return dict(select=sel, metric=metric, rhs=rhs, fit=fit, fft=fft, kernel=kern,
                total=sel + metric + rhs + fit + fft + kern)
