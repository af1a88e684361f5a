"""Host-side periodic-cell helpers for the ISDF build (numpy, tiny arrays only).

These restate the handful of PySCF helpers the reference hot path calls
(/root/reference/fftisdf.py:28,91,99,114,317-322,367-370) so that the B200 path
can run without PySCF.  They only ever touch O(nk), O(mesh) or O(ng) *tables*
(k-points, phase matrices, Coulomb weights) that are then uploaded once; every
O(ng*nip) operation is done by the CUDA library.

PySCF is not installed in the build image, so the PySCF-side semantics below are
restated from PySCF 2.x and are re-validated against PySCF whenever it imports
(tests/test_pyscf_optional.py is skipped otherwise).
"""
from __future__ import annotations

import numpy as np


def cartesian_prod(arrays):
    """pyscf.lib.cartesian_prod: C-ordered cartesian product, last axis fastest."""
    arrays = [np.asarray(a) for a in arrays]
    grids = np.meshgrid(*arrays, indexing="ij")
    return np.stack([g.reshape(-1) for g in grids], axis=1)


def reciprocal_vectors(a):
    """b such that a_i . b_j = 2 pi delta_ij (rows are vectors)."""
    return 2.0 * np.pi * np.linalg.inv(np.asarray(a, dtype=np.float64)).T


def make_kpts(a, kmesh):
    """cell.make_kpts(kmesh) / cell.get_kpts(kmesh): Gamma-centred, no wrap-around
    (used at fftisdf.py:322,434)."""
    ks = [np.arange(n, dtype=np.float64) / n for n in kmesh]
    scaled = cartesian_prod(ks)
    return scaled @ reciprocal_vectors(a)


def get_scaled_kpts(a, kpts):
    return np.asarray(kpts) @ np.asarray(a).T / (2.0 * np.pi)


def kpts_to_kmesh(a, kpts):
    """pyscf.pbc.tools.k2gamma.kpts_to_kmesh (fftisdf.py:317-318)."""
    sk = get_scaled_kpts(a, kpts).round(8)
    return [len(np.unique(sk[:, i])) for i in range(3)]


def translation_vectors_for_kmesh(a, kmesh):
    """R vectors of the k2gamma supercell, wrap_around=False (fftisdf.py:28)."""
    rel = cartesian_prod([np.arange(n) for n in kmesh])
    return rel @ np.asarray(a, dtype=np.float64)


def get_phase(a, kpts, kmesh):
    """k2gamma.get_phase(...)[1]: P[R,k] = exp(i k.R)/sqrt(nk) (fftisdf.py:28)."""
    rvec = translation_vectors_for_kmesh(a, kmesh)
    phase = np.exp(1j * (rvec @ np.asarray(kpts).T))
    return phase / np.sqrt(len(rvec))


def get_phase_axes(kmesh):
    """Per-axis factors of the Bloch phase matrix for a Gamma-centred regular mesh:
    P[(m1,m2,m3),(j1,j2,j3)] = prod_a U_a[m_a, j_a],  U_a[m,j]=exp(2 pi i m j/N_a)/sqrt(N_a).
    The CUDA k-transform kernels take these three small matrices."""
    mats = []
    for n in kmesh:
        m = np.arange(n)
        mats.append(np.exp(2j * np.pi * np.outer(m, m) / n) / np.sqrt(n))
    return mats


def phase_is_separable(phase, kmesh, tol=1e-12):
    u1, u2, u3 = get_phase_axes(kmesh)
    kron = np.einsum("ai,bj,ck->abcijk", u1, u2, u3).reshape(phase.shape)
    return bool(np.abs(kron - phase).max() < tol)


def gen_uniform_grids(a, mesh, wrap_around=False):
    """cell.gen_uniform_grids(mesh) (fftisdf.py:368): fractional grid i/mesh, C order."""
    mesh = np.asarray(mesh)
    frac = cartesian_prod([np.arange(n) / n for n in mesh])
    if wrap_around:
        frac[frac >= 0.5] -= 1.0
    return frac @ np.asarray(a, dtype=np.float64)


def get_Gv(a, mesh):
    """cell.get_Gv(mesh) (fftisdf.py:91): fftfreq-ordered integer triples @ b."""
    rx = [np.fft.fftfreq(n, 1.0 / n) for n in mesh]
    return cartesian_prod(rx) @ reciprocal_vectors(a)


def get_coulG(a, k, mesh, Gv=None, wrap_around=True):
    """pbctools.get_coulG(cell, k=vq, mesh=mesh, Gv=gv) with exxdiv=None
    (fftisdf.py:114): 4 pi / |k+G|^2, zero at k+G=0; k+G wrapped into the first
    zone and box-boundary terms zeroed when k != 0 (PySCF >= 2.1 semantics)."""
    a = np.asarray(a, dtype=np.float64)
    mesh = np.asarray(mesh)
    k = np.asarray(k, dtype=np.float64).reshape(3)
    if Gv is None:
        Gv = get_Gv(a, mesh)
    if np.abs(k).sum() > 1e-9:
        kG = k + Gv
    else:
        kG = Gv.copy()
    equal2boundary = None
    if wrap_around and np.abs(k).sum() > 1e-9:
        equal2boundary = np.zeros(Gv.shape[0], dtype=bool)
        b = reciprocal_vectors(a)
        box_edge = np.einsum("i,ij->ij", mesh // 2 + 0.5, b)
        reduced = np.linalg.solve(box_edge.T, kG.T).T.round(9)
        on_edge = reduced.astype(int)
        for ax in range(3):
            equal2boundary |= reduced[:, ax] == 1
            equal2boundary |= reduced[:, ax] == -1
            kG[on_edge[:, ax] == 1] -= 2 * box_edge[ax]
            kG[on_edge[:, ax] == -1] += 2 * box_edge[ax]
    absG2 = np.einsum("gi,gi->g", kG, kG)
    with np.errstate(divide="ignore"):
        coulG = 4.0 * np.pi / absG2
    coulG[absG2 == 0] = 0.0
    if equal2boundary is not None:
        coulG[equal2boundary] = 0.0
    return coulG


def cutoff_to_mesh(a, ke_cutoff):
    """pbctools.cutoff_to_mesh, orthogonal-lattice form: n_i = 2*ceil(sqrt(2 ke)/|b_i|)+1."""
    b = reciprocal_vectors(a)
    gmax = np.sqrt(2.0 * ke_cutoff)
    n = np.ceil(gmax / np.linalg.norm(b, axis=1)).astype(int)
    return (2 * n + 1).tolist()


def time_reversal_partner(kmesh):
    """For a Gamma-centred regular mesh, index of -q (mod G) for every q."""
    n1, n2, n3 = kmesh
    idx = np.arange(n1 * n2 * n3).reshape(n1, n2, n3)
    j1, j2, j3 = np.meshgrid(np.arange(n1), np.arange(n2), np.arange(n3), indexing="ij")
    return idx[(-j1) % n1, (-j2) % n2, (-j3) % n3].reshape(-1)
