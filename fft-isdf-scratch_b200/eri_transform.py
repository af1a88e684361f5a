"""Embedding-space ERIs from the ISDF factors -- SURVEY.md section 8 row f-4.

The reference only has a commented-out, unfinished stub for this (`trans_2e`, /root/reference/fftisdf.py:230-294):
it sets up `C_ao_lo`, `C_lo_eo`, the `nkpts**0.75` normalisation (:275-276) and `xmo = C_ao_emb[s,k].T @ xk[k].T`
(:287-289), then stops at `eri = numpy.zeros(...)` (:292).  There is no reference behaviour to be identical to, so
parity is UNPINNED for this function; what is implemented is the contraction those lines set up,

    eri[s12][p,q,r,s] = sum_{k1,k2,k3} sum_{I,J} conj(xe1_{k1}[p,I]) xe1_{k2}[q,I] W_{k2-k1}[I,J] conj(xe2_{k3}[r,J]) xe2_{k4}[s,J]

with k4 = k3 - (k2 - k1) (momentum conservation, the convention of the reference's ERI check at
fftdf-with-k-lstsq.py:219-236) and xe_k = C_ao_emb[k]^T X_k^T.  `C_lo_eo` is given in R space (one block per cell of
the k2gamma supercell) and taken to k space with e^{-ik.R}; with Bloch AOs phi_k = sum_T e^{ik.T} chi(r-T) (PySCF's
`pbc_eval_gto`) and the nkpts**-0.75 factor this makes the default (`C_lo_eo = identity`: "k2gamma AO
transformation", stub :246) the ERIs of the supercell AOs -- pinned by `tests/test_oracle_cpu.py::
test_trans_2e_default_is_the_supercell_eri` against explicit supercell pair densities.

Two routes, same contraction: a host numpy statement (any object with numpy `_x` / `_wq`), and -- when `df_obj` is a
built ISDF object, i.e. holds `_x_dev` / `_wq_dev` and the kernel handle -- the same five steps on the device through
the library's own GEMM kernels (`isdf_gemm_nn/hn/tn`): Z_k = X_k C_k, the k-pair sums as nip small batched products,
W_q applied to all q in one batched call, and the sum over q folded into the contraction length of the last product.
O(nk nip nemb^2 + nk nip^2 nemb^2) work.
"""
import numpy

from . import pbc_tools

KPT_DIFF_TOL = 1e-5   # fftisdf.py:230


def _kmesh_index_table(kmesh):
    """idx[k] = integer mesh coordinates of k-point k (cartesian-product order of cell.get_kpts), and its inverse."""
    n1, n2, n3 = [int(n) for n in kmesh]
    idx = pbc_tools.cartesian_prod([numpy.arange(n1), numpy.arange(n2), numpy.arange(n3)]).astype(int)
    inv = numpy.arange(n1 * n2 * n3).reshape(n1, n2, n3)
    return idx, inv


def _add_spin_dim(c, spin):
    c = numpy.asarray(c)
    if c.shape[0] == spin:
        return c
    assert c.shape[0] == 1, "cannot broadcast the spin dimension"
    return numpy.concatenate([c] * spin, axis=0)


def _contract_device(ops, x_dev, wq_dev, C_ao_emb, kmesh):
    """eri[n][p,r,t,u] on the device; x_dev [nk,nip,nao], wq_dev [nk,nip,nip] (torch, complex128), C_ao_emb numpy
    [spin,nk,nao,nemb].  Same pairing conventions as the host statement below (rho: k -> k+q, sig: k -> k-q)."""
    import torch
    c128 = torch.complex128
    dev = x_dev.device
    spin, nk, nao, nemb = C_ao_emb.shape
    nip = x_dev.shape[1]
    idx, inv = _kmesh_index_table(kmesh)
    cdev = torch.from_numpy(numpy.ascontiguousarray(C_ao_emb)).to(dev)
    z = [ops.gemm_nn(x_dev.contiguous(), cdev[s].contiguous()) for s in range(spin)]          # Z[s][k] = X_k C_k  [nip, nemb]
    rho = torch.empty((spin, nk, nip, nemb, nemb), dtype=c128, device=dev)
    sig = torch.empty_like(rho)
    for q in range(nk):
        kp = torch.as_tensor([inv[tuple(numpy.mod(idx[k] + idx[q], kmesh))] for k in range(nk)], device=dev)
        km = torch.as_tensor([inv[tuple(numpy.mod(idx[k] - idx[q], kmesh))] for k in range(nk)], device=dev)
        for s in range(spin):
            a = z[s].permute(1, 0, 2)                                   # [I][k][p]: one nemb x nemb product per point I
            ops.gemm_hn_strided(a, z[s].index_select(0, kp).permute(1, 0, 2), rho[s, q])   # sum_k conj(Z_k[I,p]) Z_{k+q}[I,r]
            ops.gemm_hn_strided(a, z[s].index_select(0, km).permute(1, 0, 2), sig[s, q])
    pairs = [(0, 0)] if spin == 1 else [(0, 0), (1, 1), (0, 1)]
    n2 = nemb * nemb
    eri = torch.empty((len(pairs), n2, n2), dtype=c128, device=dev)
    half = torch.empty((n2, nk * nip), dtype=c128, device=dev)          # [p r][q, J]: the q sum becomes part of K
    for n, (s1, s2) in enumerate(pairs):
        r1 = rho[s1].reshape(nk, nip, n2)
        ops.gemm_tn_strided(r1, wq_dev, half.view(n2, nk, nip).permute(1, 0, 2))           # half[q] = rho_q^T W_q
        ops.gemm_nn_strided(half[None], sig[s2].reshape(1, nk * nip, n2), eri[n][None])
    return eri.reshape(len(pairs), nemb, nemb, nemb, nemb).cpu().numpy()


def trans_2e(df_obj, C_ao_lo=None, C_lo_eo=None, unit_eri=False, symmetry=1, t_reversal_symm=True, max_memory=None,
             kscaled_center=None, kconserv_tol=KPT_DIFF_TOL, fname=None, on_device=None):
    """Signature of the stub at fftisdf.py:231-234.

    C_ao_lo [nk, nao, nlo] or [spin, nk, nao, nlo] (k space; default identity); C_lo_eo [ncell, nlo, nemb] or
    [spin, ncell, nlo, nemb] (R space; default: every supercell orbital).  unit_eri: C_ao_emb = C_ao_lo / nk^{3/4}
    (:275-276).  Returns eri [spin(spin+1)/2, nemb, nemb, nemb, nemb] (aa, bb, ab) for symmetry = 1, or the
    pair-packed real array [.., npair, npair] for symmetry = 4.  `fname`: also saved with numpy.save.
    """
    if kscaled_center is not None:
        raise NotImplementedError("shifted k-meshes: the ISDF build assumes the Gamma-centred mesh (fftisdf.py:322)")
    assert symmetry in (1, 4)
    if on_device is None:   # a built ISDF object keeps its results on the GPU
        on_device = getattr(df_obj, "_wq_dev", None) is not None and getattr(df_obj, "_ops", None) is not None
    kmesh = [int(n) for n in df_obj.kmesh]
    if on_device:
        xk, wq = df_obj._x_dev, df_obj._wq_dev
    else:
        xk = numpy.asarray(df_obj._x)
        wq = numpy.asarray(df_obj._wq)
    nkpts, nip, nao = xk.shape
    assert nkpts == int(numpy.prod(kmesh)) and tuple(wq.shape) == (nkpts, nip, nip)   # :282-285

    if C_ao_lo is None:                                                               # :246-250
        C_ao_lo = numpy.asarray([numpy.eye(nao) for _ in range(nkpts)], dtype=numpy.complex128)
    C_ao_lo = numpy.asarray(C_ao_lo, dtype=numpy.complex128)
    if C_ao_lo.ndim == 3:                                                             # :253-254
        C_ao_lo = C_ao_lo[numpy.newaxis]
    nlo = C_ao_lo.shape[-1]
    assert C_ao_lo.shape[1:3] == (nkpts, nao)

    if unit_eri:                                                                      # :275-276
        C_ao_emb = C_ao_lo / (nkpts ** 0.75)
    else:
        if C_lo_eo is None:                                                           # :262-264
            C_lo_eo = numpy.eye(nlo * nkpts).reshape((1, nkpts, nlo, nlo * nkpts))
        C_lo_eo = numpy.asarray(C_lo_eo, dtype=numpy.complex128)
        if C_lo_eo.ndim == 3:
            C_lo_eo = C_lo_eo[numpy.newaxis]
        assert C_lo_eo.shape[1:3] == (nkpts, nlo)
        spin = max(C_ao_lo.shape[0], C_lo_eo.shape[0])                                # :266-271
        C_ao_lo = _add_spin_dim(C_ao_lo, spin)
        C_lo_eo = _add_spin_dim(C_lo_eo, spin)
        a = numpy.asarray(df_obj.cell.lattice_vectors())
        rvec = pbc_tools.translation_vectors_for_kmesh(a, kmesh)
        phase_rk = numpy.exp(-1j * (rvec @ numpy.asarray(df_obj.kpts).T))            # [R, k]
        C_lo_eo_k = numpy.einsum("Rk,sRln->skln", phase_rk, C_lo_eo)
        C_ao_emb = numpy.einsum("skal,skln->skan", C_ao_lo, C_lo_eo_k) / (nkpts ** 0.75)

    spin, _, _, nemb = C_ao_emb.shape                                                 # :278-279
    assert C_ao_emb.shape == (spin, nkpts, nao, nemb) and spin in (1, 2)

    if on_device:
        eri = _contract_device(df_obj._ops, xk, wq, C_ao_emb, kmesh)
        return _finish(eri, nemb, symmetry, t_reversal_symm, fname)

    # xmo[s, k] = C_ao_emb[s, k]^T X_k^T   [nemb, nip]                                  :287-289
    xmo = numpy.einsum("skan,kIa->sknI", C_ao_emb, xk)

    # pair densities at the interpolation points, summed over the k pairs with the same transfer momentum:
    #   rho[s][q][I, p, r] = sum_{k1} conj(xmo[s,k1][p,I]) xmo[s,k1+q][r,I]
    #   sig[s][q][J, p, r] = sum_{k3} conj(xmo[s,k3][p,J]) xmo[s,k3-q][r,J]
    idx, inv = _kmesh_index_table(kmesh)
    shift = lambda k, q, sgn: inv[tuple(numpy.mod(idx[k] + sgn * idx[q], kmesh))]
    rho = numpy.zeros((spin, nkpts, nip, nemb, nemb), dtype=numpy.complex128)
    sig = numpy.zeros_like(rho)
    for q in range(nkpts):
        for k in range(nkpts):
            kp, km = shift(k, q, +1), shift(k, q, -1)
            for s in range(spin):
                rho[s, q] += numpy.einsum("pI,rI->Ipr", xmo[s, k].conj(), xmo[s, kp])
                sig[s, q] += numpy.einsum("pI,rI->Ipr", xmo[s, k].conj(), xmo[s, km])

    pairs = [(0, 0)] if spin == 1 else [(0, 0), (1, 1), (0, 1)]                       # aa, bb, ab
    eri = numpy.zeros((len(pairs), nemb, nemb, nemb, nemb), dtype=numpy.complex128)   # :292
    for n, (s1, s2) in enumerate(pairs):
        for q in range(nkpts):
            half = numpy.einsum("Ipr,IJ->prJ", rho[s1, q], wq[q])
            eri[n] += numpy.einsum("prJ,Jtu->prtu", half, sig[s2, q])

    return _finish(eri, nemb, symmetry, t_reversal_symm, fname)


def _finish(eri, nemb, symmetry, t_reversal_symm, fname):
    if symmetry == 4:
        # real orbitals in a time-reversal-symmetric set: (pr|tu) is real and symmetric within each pair
        scale = max(numpy.abs(eri).max(), 1e-300)
        assert not t_reversal_symm or numpy.abs(eri.imag).max() < 1e-8 * scale, "ERIs are not real: symmetry=4 needs real embedding orbitals"
        tri = numpy.tril_indices(nemb)
        eri = numpy.ascontiguousarray(eri.real[:, tri[0], tri[1]][:, :, tri[0], tri[1]])
    if fname is not None:
        numpy.save(fname, eri)
    return eri
