"""CPU suite: the oracle against the reference-generated golden vectors, host helpers, and the
C-ABI library's exported symbols (no compute calls -- there is no GPU here)."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from oracle import isdf_oracle as O
from oracle import pbc_helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz")))


def test_golden_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_matches_reference_golden(path):
    """oracle.build == the reference's own build() (fftisdf.py:22-128) on the stored inputs.
    Both issue the same LAPACK/FFT calls, so agreement is to rounding (1e-12 relative)."""
    g = np.load(path)
    out = O.build(g["a"], g["kpts"], g["kmesh"].tolist(), g["mesh"].tolist(), g["x0"], g["f_all"], g["coord"],
                  float(g["c0"]), blksize=int(g["blksize"]))
    assert np.array_equal(out["mask"], g["mask"])
    assert np.abs(out["x"] - g["x"]).max() == 0.0
    assert np.abs(out["wq"] - g["wq"]).max() <= 1e-12 * np.abs(g["wq"]).max()
    ph = H.get_phase(g["a"], g["kpts"], g["kmesh"].tolist())
    dms = g["dm"][None]
    vj = O.get_j_kpts(out["x"], out["wq"][0], dms)[0]
    vk = O.get_k_kpts(out["x"], out["wq"], dms, ph)[0]
    assert np.abs(vj - g["vj"]).max() <= 1e-12 * np.abs(g["vj"]).max()
    assert np.abs(vk - g["vk"]).max() <= 1e-12 * np.abs(g["vk"]).max()


def test_phase_is_unitary_dft():
    """P @ . is the unitary inverse DFT over the k-mesh axes; P is symmetric."""
    a = np.eye(3) * 5.0 + 0.2
    kmesh = [3, 2, 4]
    ph = H.get_phase(a, H.get_kpts(a, kmesh), kmesh)
    x = np.random.default_rng(0).standard_normal((24, 7)) + 0j
    ref = np.fft.ifftn(x.reshape(3, 2, 4, 7), axes=(0, 1, 2), norm="ortho").reshape(24, 7)
    assert np.abs(ph @ x - ref).max() < 1e-13
    assert np.abs(ph - ph.T).max() < 1e-13


def test_parseval_form_of_w():
    """zeta @ Theta^H == B B^H with B = FFT[Theta fq] sqrt(v vol)/ng."""
    g = np.load(GOLDEN[1])
    a, kpts, mesh = g["a"], g["kpts"], g["mesh"].tolist()
    rng = np.random.default_rng(1)
    ng = len(g["coord"])
    th = rng.standard_normal((7, ng)) + 1j * rng.standard_normal((7, ng))
    q = 3
    fq = np.exp(-1j * g["coord"] @ kpts[q])
    cg = H.get_coulG(a, kpts[q], mesh)
    vol = abs(np.linalg.det(a))
    zeta = H.ifft(H.fft(th * fq, mesh) * cg * vol / ng, mesh) * fq.conj()
    w_ref = zeta @ th.conj().T
    b = H.fft(th * fq, mesh) * np.sqrt(cg * vol) / ng
    assert np.abs(b @ b.conj().T - w_ref).max() < 1e-13 * np.abs(w_ref).max()


def test_time_reversal_symmetry_of_w():
    """W_{-q} = conj(W_q) in the reference's own output -- for odd FFT meshes with
    numerically full-rank A_q; the host check that gates the shortcut must agree."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200.fftisdf import _time_reversal_valid
    T = pk.pbc_tools
    for name, expect in [("k231_odd", True), ("k321_spd", False)]:
        g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
        kmesh, mesh = g["kmesh"].tolist(), g["mesh"].tolist()
        part = T.time_reversal_partner(kmesh)
        coulg = [T.get_coulG(g["a"], k, mesh) for k in g["kpts"]]
        ok = _time_reversal_valid(kmesh, mesh, coulg, part)
        wq = g["wq"]
        for q in range(len(part)):
            if part[q] == q:
                continue
            assert bool(ok[q]) == expect
            if ok[q]:
                assert np.abs(wq[part[q]] - wq[q].conj()).max() < 1e-11 * np.abs(wq[q]).max()


def test_pivot_rule_restatement_matches_dpstrf():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((150, 30))
    x4 = (x @ x.T) ** 2
    _, piv, rank = H.pivoted_cholesky(x4.copy())
    p2, steps, _ = H.pivoted_cholesky_steps(x4, 100)
    assert steps == 100 and np.array_equal(p2, piv[:100])


def test_product_helpers_match_oracle_helpers():
    import fft_isdf_scratch_b200 as pk
    T = pk.pbc_tools
    a = np.array([[6.0, 0.4, 0.0], [0.1, 5.0, 0.3], [0.0, -0.2, 7.0]])
    kmesh, mesh = [2, 3, 2], [7, 8, 9]
    kp = T.make_kpts(a, kmesh)
    assert np.abs(kp - H.get_kpts(a, kmesh)).max() < 1e-14
    assert np.abs(T.get_phase(a, kp, kmesh) - H.get_phase(a, kp, kmesh)).max() < 1e-14
    assert T.phase_is_separable(T.get_phase(a, kp, kmesh), kmesh)
    assert np.abs(T.get_Gv(a, mesh) - H.get_Gv(a, mesh)).max() < 1e-13
    assert np.abs(T.gen_uniform_grids(a, mesh) - H.gen_uniform_grids(a, mesh)).max() < 1e-13
    for k in kp:
        assert np.abs(T.get_coulG(a, k, mesh) - H.get_coulG(a, k, mesh)).max() < 1e-12
    assert T.kpts_to_kmesh(a, kp) == kmesh


def test_cabi_library_exports_every_declared_symbol():
    import fft_isdf_scratch_b200._cabi as cabi
    hdr = open(os.path.join(ROOT, "include", "isdf_b200.h")).read()
    declared = set(re.findall(r"\b(isdf_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    if not os.path.exists(cabi.lib_path()):          # a fresh checkout: nvcc cross-compiles without a GPU
        import __graft_entry__
        __graft_entry__.build()
    assert os.path.exists(cabi.lib_path()), "libisdf_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(cabi.lib_path())
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/isdf_b200.h but not exported"
    assert declared == set(cabi.SIGNATURES), (declared ^ set(cabi.SIGNATURES))
    assert lib.isdf_abi_version() == 1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    cell = pk.random_cubic_cell(8, 3, seed=0, L=6.0)
    with pytest.raises(Exception):
        fftisdf.ISDF(cell, cell.get_kpts([1, 1, 1]))


def test_time_reversal_rule_odd_mesh():
    """build() exploits W_{-q} = W_q^* iff every FFT mesh dimension is odd; check that rule against the
    numeric criterion on the Coulomb tables for several lattices / meshes / k-meshes."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200.fftisdf import _time_reversal_valid
    T = pk.pbc_tools
    rng = np.random.default_rng(3)
    for mesh, kmesh in [([9, 11, 7], [2, 3, 1]), ([15, 15, 15], [4, 4, 4]), ([7, 7, 9], [3, 3, 3]), ([5, 9, 11], [5, 2, 4]),
                        ([37, 37, 37], [3, 3, 3])]:
        a = np.eye(3) * 6.0 + rng.uniform(-1, 1, (3, 3))
        kp = T.make_kpts(a, kmesh)
        part = T.time_reversal_partner(kmesh)
        cg = [T.get_coulG(a, k, mesh) for k in kp]
        ok = _time_reversal_valid(kmesh, mesh, cg, part)
        assert all(ok[q] for q in range(len(kp)) if part[q] != q), (mesh, kmesh)
    a = np.eye(3) * 6.0 + rng.uniform(-1, 1, (3, 3))
    kp = T.make_kpts(a, [3, 2, 1])
    cg = [T.get_coulG(a, k, [8, 9, 10]) for k in kp]
    ok = _time_reversal_valid([3, 2, 1], [8, 9, 10], cg, T.time_reversal_partner([3, 2, 1]))
    assert not ok.any()


def _exact_isdf_factors(side, kmesh, nao=2, seed=81):
    """Tiny cell in the exact-ISDF regime (every grid point a candidate, nip = full rank), built by the oracle."""
    import types
    import fft_isdf_scratch_b200 as pk
    cell = pk.random_cubic_cell(side, nao, seed=seed, L=5.0, ltypes="s")
    kpts = cell.get_kpts(kmesh)
    coord = cell.gen_uniform_grids(cell.mesh)
    phi = cell.eval_ao_kpts(coord, kpts)
    out = O.build(cell.a, kpts, kmesh, list(cell.mesh), phi, phi, coord, 50.0)
    return cell, types.SimpleNamespace(_x=out["x"], _wq=out["wq"], kmesh=kmesh, kpts=kpts, cell=cell)


@pytest.mark.parametrize("side,kmesh", [(6, [2, 1, 1]), (6, [2, 2, 1]), (5, [3, 1, 1])])
def test_trans_2e_default_is_the_supercell_eri(side, kmesh):
    """with C_ao_lo = C_lo_eo = identity ("k2gamma AO transformation", fftisdf.py:246) the embedding
    ERIs are the ERIs of the supercell AOs.  Pins the R->k phase e^{-ik.R}, the nkpts**-0.75 normalisation (:275) and
    the momentum convention against EXPLICIT supercell pair densities (exact-ISDF regime, so no fitting error)."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import eri_transform as E
    cell, df = _exact_isdf_factors(side, kmesh)
    eri = E.trans_2e(df)[0]
    rvec = pk.pbc_tools.translation_vectors_for_kmesh(cell.a, kmesh)
    shells = [(cell._cen[i], "s", cell._alp[i]) for i in range(cell.nao_nr())]
    mesh_sc = [int(m * k) for m, k in zip(cell.mesh, kmesh)]
    sup = pk.SyntheticCell(np.diag(kmesh) @ cell.a, [(c + r, l, al) for r in rvec for (c, l, al) in shells], mesh_sc)
    csc = sup.gen_uniform_grids(mesh_sc)
    chi = sup.eval_ao_kpts(csc, np.zeros((1, 3)))[0].real
    n, ng = chi.shape[1], len(csc)
    rho = np.einsum("gm,gn->mng", chi, chi).reshape(n * n, ng)
    zeta = H.ifft(H.fft(rho, mesh_sc) * H.get_coulG(sup.a, np.zeros(3), mesh_sc) * sup.vol / ng, mesh_sc)
    ref = (zeta @ rho.T).reshape(n, n, n, n)
    assert np.abs(eri - ref).max() < 1e-10 * np.abs(ref).max()
    assert np.abs(eri.imag).max() < 1e-12 * np.abs(ref).max()
    packed = E.trans_2e(df, symmetry=4)[0]
    tri = np.tril_indices(n)
    assert np.abs(packed - ref[tri[0], tri[1]][:, tri[0], tri[1]]).max() < 1e-10 * np.abs(ref).max()


def test_trans_2e_matches_definition_for_general_orbitals():
    """Random k-space C_ao_lo, R-space C_lo_eo, two spins: the factored contraction == the quadruple sum."""
    from fft_isdf_scratch_b200 import eri_transform as E
    import types
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_k231_odd.npz"))
    kmesh = g["kmesh"].tolist()
    nk, nip, nao = g["x"].shape
    cell = types.SimpleNamespace(lattice_vectors=lambda: g["a"])
    df = types.SimpleNamespace(_x=g["x"], _wq=g["wq"], kmesh=kmesh, kpts=g["kpts"], cell=cell)
    rng = np.random.default_rng(5)
    nlo, nemb = 5, 3
    c_ao_lo = rng.standard_normal((2, nk, nao, nlo)) + 1j * rng.standard_normal((2, nk, nao, nlo))
    c_lo_eo = rng.standard_normal((1, nk, nlo, nemb))
    eri = E.trans_2e(df, C_ao_lo=c_ao_lo, C_lo_eo=c_lo_eo)
    assert eri.shape == (3, nemb, nemb, nemb, nemb)
    import itertools
    rvec = np.array(list(itertools.product(*[range(n) for n in kmesh]))) @ g["a"]
    ph = np.exp(-1j * rvec @ g["kpts"].T)
    c_emb = [np.einsum("kal,kln->kan", c_ao_lo[s], np.einsum("Rk,Rln->kln", ph, c_lo_eo[0])) / nk ** 0.75 for s in range(2)]
    ref_aa = O.trans_2e_bruteforce(g["x"], g["wq"], kmesh, c_emb[0])
    assert np.abs(eri[0] - ref_aa).max() < 1e-12 * np.abs(ref_aa).max()
    ref_bb = O.trans_2e_bruteforce(g["x"], g["wq"], kmesh, c_emb[1])
    assert np.abs(eri[1] - ref_bb).max() < 1e-12 * np.abs(ref_bb).max()
    # unit_eri: C_ao_emb = C_ao_lo / nk^{3/4} (fftisdf.py:275-276)
    eri_u = E.trans_2e(df, C_ao_lo=c_ao_lo[0], unit_eri=True)
    ref_u = O.trans_2e_bruteforce(g["x"], g["wq"], kmesh, c_ao_lo[0] / nk ** 0.75)
    assert eri_u.shape == (1, nlo, nlo, nlo, nlo) and np.abs(eri_u[0] - ref_u).max() < 1e-12 * np.abs(ref_u).max()
    with pytest.raises(NotImplementedError):
        E.trans_2e(df, kscaled_center=[0.5, 0.0, 0.0])


# ---- oracle/gelsy_port.py (numpy restatement of LAPACK zgelsy) pinned against the real LAPACK through scipy ----------
def _graded_psd(n, r, seed, decay):
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((r, n)) + 1j * rng.standard_normal((r, n))
    c *= 10.0 ** (-decay * np.arange(r) / max(r - 1, 1))[:, None]
    return c.conj().T @ c


@pytest.mark.parametrize("n,seed", [(7, 1), (40, 2), (90, 3)])
def test_gelsy_port_qrcp_matches_zgeqp3(n, seed):
    import scipy.linalg
    from oracle import gelsy_port as GP
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    a *= (1.0 + np.arange(n))[None, ::-1] ** 0.5
    qr, tau, jp = GP.qrcp(a)
    qr2, jp2, tau2, _, _ = scipy.linalg.lapack.zgeqp3(a)
    assert np.array_equal(jp, jp2 - 1)
    assert np.abs(np.triu(qr) - np.triu(qr2)).max() < 1e-12 * np.abs(qr2).max()
    assert np.abs(tau - tau2).max() < 1e-12


@pytest.mark.parametrize("n,r,decay,seed", [(12, 12, 1.0, 1), (40, 25, 5.0, 2), (60, 60, 7.5, 3), (80, 50, 9.0, 4),
                                             (57, 6, 1.0, 5)])
def test_gelsy_port_rank_and_solution_match_scipy_gelsy(n, r, decay, seed):
    import scipy.linalg
    from oracle import gelsy_port as GP
    a = _graded_psd(n, r, seed, decay)
    rng = np.random.default_rng(seed + 50)
    b = a @ (rng.standard_normal((n, 4)) + 1j * rng.standard_normal((n, 4)))
    ref = scipy.linalg.lstsq(a, b, lapack_driver="gelsy")
    x, rank = GP.gelsy(a, b)
    assert rank == ref[2]
    f = GP.gelsy_factor(a)
    assert np.abs(f["q1"].conj().T @ f["q1"] - np.eye(rank)).max() < 1e-13
    assert np.abs(f["e"].conj().T @ f["e"] - np.eye(rank)).max() < 1e-13
    # same rank + same algorithm: the minimum-norm solutions agree to the conditioning of the kept block
    d = np.abs(np.diag(f["r"]))[:rank]
    assert np.abs(x - ref[0]).max() < max(1e-11, 100 * d[0] / d[-1] * 2.2e-16) * np.abs(ref[0]).max()
    assert np.abs(a @ x - b).max() < 1e-9 * np.abs(b).max()


@pytest.mark.parametrize("name", ["k222_sp", "k221_rd"])
def test_gelsy_port_reproduces_the_reference_ranks(name):
    """The per-q ranks the REFERENCE logged (fftisdf.py:122, stored in the golden by gen_golden.py) come out of the
    port's QRCP + incremental condition estimation on the same A_q."""
    from oracle import gelsy_port as GP
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    ph = H.get_phase(g["a"], g["kpts"], g["kmesh"].tolist())
    x4_k = O.build_metric(g["x"], ph)
    ranks = [GP.gelsy_factor(x4_k[q])["rank"] for q in range(len(x4_k))]
    assert ranks == g["ranks"].tolist()
