"""TEST INFRASTRUCTURE (oracle) -- never imported by the shipped path.

CPU restatement of the PySCF helpers the reference hot path calls.  PySCF (version
unpinned in the reference; no requirements file) is NOT installed in the build image and
is absent from /root/reference, so these restate its published 2.x behaviour and are
*unpinned* at the PySCF boundary; every function cites the reference call site it serves.
"""
import numpy as np
import scipy.linalg


def cartesian_prod(arrays):
    # pyscf.lib.cartesian_prod: C order, last array fastest
    mg = np.meshgrid(*[np.asarray(x) for x in arrays], indexing="ij")
    return np.stack([m.ravel() for m in mg], axis=-1)


def reciprocal_vectors(a):
    return 2 * np.pi * np.linalg.inv(a).T


def get_kpts(a, kmesh):
    # cell.get_kpts(kmesh)  (fftisdf.py:322,434): make_kpts, wrap_around=False
    scaled = cartesian_prod([np.arange(n) / float(n) for n in kmesh])
    return scaled.dot(reciprocal_vectors(a))


def get_phase(a, kpts, kmesh):
    # pyscf.pbc.tools.k2gamma.get_phase(cell, kpts, kmesh, wrap_around=False)[1]  (fftisdf.py:28)
    r_rel = cartesian_prod([np.arange(n) for n in kmesh])
    r_abs = r_rel.dot(a)
    nr = len(r_abs)
    phase = np.exp(1j * np.einsum("Ru,ku->Rk", r_abs, kpts))
    phase /= np.sqrt(nr)
    return phase


def gen_uniform_grids(a, mesh, wrap_around=False):
    # cell.gen_uniform_grids(mesh)  (fftisdf.py:368)
    mesh = np.asarray(mesh)
    qv = cartesian_prod([np.arange(x) for x in mesh])
    if wrap_around:
        for i in range(3):
            qv[qv[:, i] >= (mesh[i] + 1) // 2, i] -= mesh[i]
    a_frac = np.einsum("i,ij->ij", 1.0 / mesh, a)
    return qv.dot(a_frac)


def get_Gv(a, mesh):
    # cell.get_Gv(mesh)  (fftisdf.py:91)
    gx = np.fft.fftfreq(mesh[0], 1.0 / mesh[0])
    gy = np.fft.fftfreq(mesh[1], 1.0 / mesh[1])
    gz = np.fft.fftfreq(mesh[2], 1.0 / mesh[2])
    return cartesian_prod((gx, gy, gz)).dot(reciprocal_vectors(a))


def get_coulG(a, k, mesh, Gv=None):
    # pbctools.get_coulG(cell, k=vq, mesh=mesh, Gv=gv), exxdiv=None, wrap_around=True  (fftisdf.py:114)
    mesh = np.asarray(mesh)
    if Gv is None:
        Gv = get_Gv(a, mesh)
    if abs(k).sum() > 1e-9:
        kG = k + Gv
    else:
        kG = Gv
    equal2boundary = None
    if abs(k).sum() > 1e-9:
        equal2boundary = np.zeros(Gv.shape[0], dtype=bool)
        b = reciprocal_vectors(a)
        box_edge = np.einsum("i,ij->ij", mesh // 2 + 0.5, b)
        assert all(np.linalg.solve(box_edge.T, k).round(9).astype(int) == 0)
        reduced_coords = np.linalg.solve(box_edge.T, kG.T).T.round(9)
        on_edge = reduced_coords.astype(int)
        for ax in range(3):
            equal2boundary |= reduced_coords[:, ax] == 1
            equal2boundary |= reduced_coords[:, ax] == -1
            kG[on_edge[:, ax] == 1] -= 2 * box_edge[ax]
            kG[on_edge[:, ax] == -1] += 2 * box_edge[ax]
    absG2 = np.einsum("gi,gi->g", kG, kG)
    G0_idx = np.where(absG2 == 0)[0]
    with np.errstate(divide="ignore"):
        coulG = 4 * np.pi / absG2
        coulG[G0_idx] = 0
    if equal2boundary is not None:
        coulG[equal2boundary] = 0
    return coulG


def fft(f, mesh):
    # pbctools.fft (fftisdf.py:113): unnormalised forward 3-D FFT over the last axis reshaped to mesh
    if f.size == 0:
        return np.zeros_like(f)
    f3d = f.reshape(-1, *mesh)
    g3d = np.fft.fftn(f3d, axes=(1, 2, 3))
    ngrids = np.prod(mesh)
    if f.ndim == 1 or (f.ndim == 3 and f.size == ngrids):
        return g3d.ravel()
    return g3d.reshape(-1, ngrids)


def ifft(g, mesh):
    # pbctools.ifft (fftisdf.py:118): inverse with 1/ng
    if g.size == 0:
        return np.zeros_like(g)
    g3d = g.reshape(-1, *mesh)
    f3d = np.fft.ifftn(g3d, axes=(1, 2, 3))
    ngrids = np.prod(mesh)
    if g.ndim == 1 or (g.ndim == 3 and g.size == ngrids):
        return f3d.ravel()
    return f3d.reshape(-1, ngrids)


def pivoted_cholesky(A, tol=-1.0, lower=False):
    # pyscf.lib.scipy_helper.pivoted_cholesky -> LAPACK dpstrf  (fftisdf.py:381-382)
    N = A.shape[0]
    assert A.shape == (N, N)
    L, piv, rank, info = scipy.linalg.lapack.dpstrf(A, tol=tol, lower=lower)
    if info < 0:
        raise RuntimeError("Pivoted Cholesky factorization failed.")
    if lower:
        L[np.triu_indices(N, k=1)] = 0
        L[:, rank:] = 0
    else:
        L[np.tril_indices(N, k=-1)] = 0
        L[rank:, :] = 0
    return L, piv - 1, rank


def pivoted_cholesky_steps(A, max_steps, tol=-1.0):
    """Plain-loop restatement of LAPACK dpstrf's pivot rule (unblocked dpstf2 order),
    stopped after `max_steps` pivots: returns (piv[:steps], steps, next_pivot_value).
    Pivot = first maximum of the running residual diagonal in *current position order*;
    stop when pivot <= n*eps*max(diag) (tol<0) -- used to pin the GPU selector's tie rule."""
    A = np.array(A, dtype=np.float64, copy=True)
    n = A.shape[0]
    piv = np.arange(n)
    d0 = np.diag(A).copy()
    ajj = d0.max()
    if not (ajj > 0):
        return piv[:0], 0, float(ajj)
    dstop = n * np.finfo(np.float64).eps * ajj if tol < 0 else tol
    L = np.zeros((max_steps, n))
    d = d0.copy()
    steps = 0
    for j in range(max_steps):
        sub = d[piv[j:]]
        p = j + int(np.argmax(sub))
        ajj = d[piv[p]]
        if ajj <= dstop or np.isnan(ajj):
            break
        piv[[j, p]] = piv[[p, j]]
        o = piv[j]
        row = (A[o, :] - L[:j, o].dot(L[:j, :])) / np.sqrt(ajj)
        row[piv[: j + 1]] = 0.0
        row[o] = np.sqrt(ajj)
        L[j] = row
        rest = piv[j + 1:]
        d[rest] -= row[rest] ** 2
        steps += 1
    nxt = float(d[piv[steps:]].max()) if steps < n else 0.0
    return piv[:steps].copy(), steps, nxt
