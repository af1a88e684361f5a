"""End-to-end parity of the B200 ISDF build (through the reference-shaped Python surface and the C ABI)
against (a) golden vectors produced by the reference's own code and (b) the numpy/scipy oracle.

Tolerances (FP64, relative to the largest reference entry):
  * interpolation-point indices: identical.
  * A_q numerically full rank (cond <~ 1e6): Theta, W_q, J, K, E_x within 1e-10 (north-star bar).
  * cond ~ 1e8 (gamma_s): 1e-8 on Theta/W (forward error ~ cond*eps of either solver), 1e-10 on J/K.
  * rank-deficient A_q (k222_sp, k221_rd: the regime the reference's c0 default lands in): the device zgelsy takes the
    reference's rank decision q by q; K, J, E_x and the reconstructed ERIs within 10x of the reference's own
    reproducibility floor (its result under a 1e-16 perturbation of A_q: ~6e-7 on K), and the reference's ERI
    acceptance test (vs exact pair densities) for both arms.
"""
import os

import numpy as np
import pytest
import scipy.linalg

from oracle import isdf_oracle as O
from oracle import pbc_helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def run_golden(name, **attrs):
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    g = np.load(os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz"))
    cell = pk.TableCell(g["a"], g["mesh"], g["x0"].shape[-1])
    df = fftisdf.ISDF(cell, g["kpts"], m0=g["m0"].tolist(), c0=float(g["c0"]))
    df.blksize = int(g["blksize"])
    df.set_ao_tables(x0=g["x0"], f_all=g["f_all"])
    for k, v in attrs.items():
        setattr(df, k, v)
    df.build()
    return g, df


@pytest.mark.parametrize("name,tol_w,tol_jk", [("k321_spd", 1e-10, 1e-10), ("k231_odd", 1e-10, 1e-10),
                                               ("k333_odd", 1e-10, 1e-10), ("k434_odd", 1e-10, 1e-10),
                                               ("gamma_s", 1e-8, 1e-10)])
def test_full_rank_cases_match_reference(name, tol_w, tol_jk):
    g, df = run_golden(name, keep_theta=True)
    assert np.array_equal(df._mask, g["mask"])                 # indices identical
    assert np.array_equal(df._x, g["x"])                       # same AO rows, bit for bit
    assert df._wq.shape == g["wq"].shape and df._w0.shape == g["w0"].shape
    assert all(int(r) == df._x.shape[1] for r in df._ranks)
    assert rel(df._wq, g["wq"]) < tol_w
    assert rel(df._w0, g["w0"]) < tol_w
    vj, vk = df.get_jk(g["dm"], kpts=g["kpts"])
    assert rel(vj, g["vj"].reshape(vj.shape)) < tol_jk
    assert rel(vk, g["vk"].reshape(vk.shape)) < tol_jk
    shp = (1,) + g["dm"].shape
    ex_ref = O.exchange_energy(g["vk"].reshape(shp), g["dm"].reshape(shp))
    ex = O.exchange_energy(np.asarray(vk).reshape(shp), g["dm"].reshape(shp))
    assert abs(ex - ex_ref) < tol_jk * abs(ex_ref)
    # reconstructed ERIs (fftdf-with-k-lstsq.py:232) from the GPU W_q vs from the reference's W_q
    x = g["x"]
    nk = len(x)
    for (k1, k2, k3) in [(0, 0, 0), (0, nk - 1, 0), (nk // 2, 0, nk - 1)]:
        for q in range(nk):
            e_gpu = O.eri_from_w(df._wq[q], x[k1], x[k2], x[k3], x[k1])
            e_ref = O.eri_from_w(g["wq"][q], x[k1], x[k2], x[k3], x[k1])
            assert rel(e_gpu, e_ref) < tol_w
    # Theta against the oracle's gelsy solution
    out = O.build(g["a"], g["kpts"], g["kmesh"].tolist(), g["mesh"].tolist(), g["x0"], g["f_all"], g["coord"],
                  float(g["c0"]), keep_theta=True)
    th = df._theta_dev.cpu().numpy()
    for s, q in enumerate(df._qind):
        assert rel(th[s], out["theta"][q]) < tol_w


def test_time_reversal_shortcut_equals_full_computation():
    g, df1 = run_golden("k231_odd", use_time_reversal=True)
    _, df2 = run_golden("k231_odd", use_time_reversal=False)
    assert len(df1._qind) < len(df2._qind) == len(g["kpts"])
    assert rel(df1._wq, df2._wq) < 1e-11
    # even mesh + skewed lattice: the host check must refuse the shortcut
    _, df3 = run_golden("k321_spd")
    assert len(df3._qind) == len(g["kpts"])


def _momentum_quadruples(kmesh, nk, stride3=1):
    import itertools
    kidx = np.array(list(itertools.product(*[range(n) for n in kmesh])))
    find = lambda v: int(np.where((kidx == np.mod(v, kmesh)).all(1))[0][0])
    for k1 in range(nk):
        for k2 in range(nk):
            q = find(kidx[k2] - kidx[k1])
            for k3 in range(0, nk, stride3):
                yield k1, k2, k3, find(kidx[k3] - kidx[q]), q


def _eri_rel_to_reference(x, wq, wq_ref, kmesh):
    """max relative deviation of the reconstructed ERIs (fftdf-with-k-lstsq.py:232) from those of the reference's W."""
    worst = 0.0
    for k1, k2, k3, k4, q in _momentum_quadruples(kmesh, len(x)):
        e1 = O.eri_from_w(wq[q], x[k1], x[k2], x[k3], x[k4])
        e0 = O.eri_from_w(wq_ref[q], x[k1], x[k2], x[k3], x[k4])
        worst = max(worst, rel(e1, e0))
    return worst


def _eri_error_vs_exact(x, wq, g, kmesh):
    """The reference's acceptance test (fftdf-with-k-lstsq.py:219-258): reconstructed ERIs against the exact ERIs of the
    same pair densities on the dense grid (PySCF-free stand-in for FFTDF.get_eri), worst relative error."""
    a, kpts, mesh, coord, phi = g["a"], g["kpts"], g["mesh"].tolist(), g["coord"], g["f_all"]
    ng, vol = len(coord), abs(np.linalg.det(g["a"]))
    worst, cache = 0.0, {}
    for k1, k2, k3, k4, q in _momentum_quadruples(kmesh, len(x)):
        if (k1, k2) not in cache:
            cache.clear()
            fq = np.exp(-1j * coord @ kpts[q])
            rho12 = np.einsum("gm,gn->mng", phi[k1].conj(), phi[k2]).reshape(-1, ng)
            cache[(k1, k2)] = H.ifft(H.fft(rho12 * fq, mesh) * H.get_coulG(a, kpts[q], mesh) * vol / ng, mesh) * fq.conj()
        rho34 = np.einsum("gk,gl->klg", phi[k3].conj(), phi[k4]).reshape(-1, ng)
        eri_ref = cache[(k1, k2)] @ rho34.T
        eri = O.eri_from_w(wq[q], x[k1], x[k2], x[k3], x[k4]).reshape(eri_ref.shape)
        worst = max(worst, float(np.abs(eri - eri_ref).max() / np.abs(eri_ref).max()))
    return worst


def _reference_noise_floor(g, out, nrep=3):
    """How far the REFERENCE's own result (scipy lstsq(gelsy), fftisdf.py:108) moves under changes that are exact
    no-ops mathematically: A_q perturbed at the 1e-16 level, and the system permuted symmetrically (rows and columns of
    A_q, rows of Y^T: only the summation / tie order changes, which is what another BLAS does to the reference).
    Max over nrep variants of the relative change of K, J and the reconstructed ERIs: the reproducibility floor of
    the reference in the rank-deficient regime."""
    a, kpts, kmesh, mesh = g["a"], g["kpts"], g["kmesh"].tolist(), g["mesh"].tolist()
    ph, gv, vol, ng = H.get_phase(a, kpts, kmesh), H.get_Gv(a, mesh), abs(np.linalg.det(a)), len(g["coord"])
    dms = g["dm"][None]
    vk_ref, vj_ref = g["vk"].reshape(g["dm"].shape), g["vj"].reshape(g["dm"].shape)
    fk = fj = fe = 0.0
    for rep in range(nrep):
        rng = np.random.default_rng(1000 + rep)
        wq2 = []
        for q in range(len(kpts)):
            aq, yt = out["x4_k"][q], out["y"][q].T
            if rep % 2 == 0:
                th = scipy.linalg.lstsq(aq * (1.0 + 1e-16 * rng.standard_normal(aq.shape)), yt, lapack_driver="gelsy")[0]
            else:
                perm = rng.permutation(len(aq))
                th = np.empty_like(yt)
                th[perm] = scipy.linalg.lstsq(aq[perm][:, perm], yt[perm], lapack_driver="gelsy")[0]
            fq = np.exp(-1j * g["coord"] @ kpts[q])
            b = H.fft(th * fq, mesh) * np.sqrt(H.get_coulG(a, kpts[q], mesh, Gv=gv) * vol) / ng
            wq2.append(b @ b.conj().T)
        wq2 = np.asarray(wq2)
        fk = max(fk, rel(O.get_k_kpts(out["x"], wq2, dms, ph)[0], vk_ref))
        fj = max(fj, rel(O.get_j_kpts(out["x"], wq2[0], dms)[0], vj_ref))
        fe = max(fe, _eri_rel_to_reference(out["x"], wq2, g["wq"], kmesh))
    return fk, fj, fe


@pytest.mark.parametrize("name", ["k222_sp", "k221_rd"])
def test_rank_deficient_cases_at_the_reference_noise_floor(name):
    """Rank-deficient A_q -- the regime the reference's defaults land in (nip = nao*c0 beyond the local pair rank; the
    author logs `rank / nip`, fftisdf.py:122).  The device zgelsy must take the reference's rank decision, and K, J,
    E_x and the reconstructed ERIs must sit within 10x of the reference's own eps-perturbation floor."""
    g, df = run_golden(name)
    nip = df._x.shape[1]
    kmesh = g["kmesh"].tolist()
    assert np.array_equal(df._mask, g["mask"]) and np.array_equal(df._x, g["x"])
    assert all(int(r) < nip for r in g["ranks"])
    out = O.build(g["a"], g["kpts"], kmesh, g["mesh"].tolist(), g["x0"], g["f_all"], g["coord"], float(g["c0"]))
    assert list(out["ranks"]) == list(g["ranks"])
    ranks = np.zeros(len(g["kpts"]), dtype=int)
    tr = __import__("fft_isdf_scratch_b200").pbc_tools.time_reversal_partner(kmesh)
    for s, q in enumerate(df._qind):
        ranks[q] = ranks[tr[q]] = df._ranks[s]
    # zgelsy's rank, q by q: the reference's, or cut elsewhere INSIDE the eps-plateau of |R_kk| (rounding decides where;
    # LAPACK's own cut moves when the system is merely permuted).  In the latter case the build is repeated with the
    # reference's ranks imposed, so that everything else in the solver is still held to the noise floor.
    differ = [q for q in range(len(ranks)) if ranks[q] != g["ranks"][q]]
    for q in differ:
        rd = np.abs(np.diag(scipy.linalg.lapack.zgeqp3(out["x4_k"][q])[0]))
        lo = min(ranks[q], g["ranks"][q])
        assert abs(int(ranks[q]) - int(g["ranks"][q])) <= 2 and rd[lo - 1] / rd[0] < 50 * 2.3e-16, (q, ranks[q], g["ranks"][q])
    print(f"\n{name}: device ranks {ranks.tolist()} reference {g['ranks'].tolist()}")
    if differ:
        _, df = run_golden(name, gelsy_rank_override={q: int(g["ranks"][q]) for q in df._qind})
    w = df._wq
    for q in range(len(w)):
        assert np.abs(w[q] - w[q].conj().T).max() == 0.0                  # exactly Hermitian
    fk, fj, fe = _reference_noise_floor(g, out)
    vj, vk = df.get_jk(g["dm"], kpts=g["kpts"])
    dk, dj = rel(vk, g["vk"].reshape(vk.shape)), rel(vj, g["vj"].reshape(vj.shape))
    de = _eri_rel_to_reference(g["x"], w, g["wq"], kmesh)
    print(f"\n{name}: K {dk:.2e} (floor {fk:.2e})  J {dj:.2e} (floor {fj:.2e})  ERI {de:.2e} (floor {fe:.2e})")
    assert dk < 10 * fk and dj < 10 * fj and de < 10 * fe, (dk, fk, dj, fj, de, fe)
    shp = (1,) + g["dm"].shape
    ex_ref = O.exchange_energy(g["vk"].reshape(shp), g["dm"].reshape(shp))
    ex = O.exchange_energy(np.asarray(vk).reshape(shp), g["dm"].reshape(shp))
    assert abs(ex - ex_ref) < 10 * fk * abs(ex_ref)
    # the reference's own acceptance test, both arms: ERIs against the exact pair densities
    e_gpu, e_ref = _eri_error_vs_exact(g["x"], w, g, kmesh), _eri_error_vs_exact(g["x"], g["wq"], g, kmesh)
    print(f"{name}: ERI vs exact  device {e_gpu:.3e}  reference {e_ref:.3e}")
    assert e_gpu <= 2 * e_ref and e_gpu < 1e-3      # (this tiny cell's ISDF error itself is 3e-4 / 2e-5)
    # the cheaper rank-revealing Cholesky route (fit = "cholesky") is NOT held to this bar: it truncates differently
    _, dfc = run_golden(name, fit="cholesky")
    vjc, vkc = dfc.get_jk(g["dm"], kpts=g["kpts"])
    assert rel(vkc, g["vk"].reshape(vkc.shape)) < 1e-3 and _eri_error_vs_exact(g["x"], dfc._wq, g, kmesh) < 1e-3


def test_synthetic_cell_end_to_end_vs_oracle():
    """Cell -> AO evaluation -> build, the path a user takes (no precomputed tables)."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    cell = pk.random_cubic_cell(12, 10, seed=21, L=8.0, ltypes="spd")
    kmesh = [2, 1, 2]
    kpts = cell.get_kpts(kmesh)
    df = fftisdf.ISDF(cell, kpts, m0=[7, 7, 7], c0=3.0)
    df.blksize = 700          # ragged last block (1728 = 2*700 + 328)
    df.ao_on_device = False   # host AO evaluation: the very same AO values the oracle gets
    df.build()
    x0 = cell.eval_ao_kpts(cell.gen_uniform_grids([7, 7, 7]), df.kpts)
    coord = cell.gen_uniform_grids(cell.mesh)
    f_all = cell.eval_ao_kpts(coord, df.kpts)
    out = O.build(cell.a, df.kpts, kmesh, cell.mesh, x0, f_all, coord, 3.0)
    assert np.array_equal(df._mask, out["mask"])
    assert np.array_equal(df._x, out["x"])
    conds = [np.linalg.cond(a) for a in out["x4_k"]]
    tol = max(1e-10, 50 * max(conds) * 2.2e-16)
    assert rel(df._wq, out["wq"]) < tol, (rel(df._wq, out["wq"]), max(conds))
    # AO values from the library's own AO kernel (differ from numpy's at the 1e-16 level)
    df2 = fftisdf.ISDF(cell, kpts, m0=[7, 7, 7], c0=3.0)
    df2.blksize = 700
    df2.build()
    assert np.array_equal(df2._mask, out["mask"])
    assert rel(df2._x, out["x"]) < 1e-13
    assert rel(df2._wq, out["wq"]) < tol
    # function-form twin (fftdf-with-k-lstsq.py:189)
    coul_q, x_k = fftisdf.get_coul(df, kmesh=kmesh, c0=3.0, m0=[7, 7, 7], blksize=700)
    assert rel(x_k, df._x) < 1e-13 and rel(coul_q, df._wq) < tol   # device AOs differ at 1e-16, amplified by cond(A_q)


def test_reference_error_conventions():
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    cell = pk.random_cubic_cell(8, 4, seed=3, L=6.0)
    df = fftisdf.ISDF(cell, cell.get_kpts([1, 1, 2]), m0=[5, 5, 5], c0=2.0)
    df.build()
    dm = np.zeros((2, 4, 4), complex)
    with pytest.raises(NotImplementedError):
        df.get_jk(dm, omega=0.1)                      # fftisdf.py:392-393
    with pytest.raises(NotImplementedError):
        df.get_jk(dm, exxdiv="ewald")                 # fftisdf.py:395-396
    assert df._x.shape[0] == 2 and df._wq.shape[0] == 2 and df._w0.shape == df._wq.shape[1:]


def _eri_worst_error(df, cell, kmesh):
    """max relative error of einsum("IJ,Im,In,Jk,Jl->mnkl", W_q, x1*, x2, x3*, x4) (fftdf-with-k-lstsq.py:232) over all
    momentum-conserving quadruples, against the exact ERI of the same pair densities evaluated on the dense grid
    with the same Coulomb kernel (PySCF-free stand-in for FFTDF.get_eri)."""
    import itertools
    x, wq = df._x, df._wq
    coord = cell.gen_uniform_grids(cell.mesh)
    phi = cell.eval_ao_kpts(coord, df.kpts)              # [nk, ng, nao]
    ng, mesh, nk = len(coord), cell.mesh, len(df.kpts)
    kidx = np.array(list(itertools.product(*[range(n) for n in kmesh])))
    find = lambda v: int(np.where((kidx == np.mod(v, kmesh)).all(1))[0][0])
    worst = 0.0
    for k1 in range(nk):
        for k2 in range(nk):
            q = find(kidx[k2] - kidx[k1])                  # pair momentum k2 - k1
            fq = np.exp(-1j * coord @ df.kpts[q])
            cg = H.get_coulG(cell.a, df.kpts[q], mesh)
            rho12 = np.einsum("gm,gn->mng", phi[k1].conj(), phi[k2]).reshape(-1, ng)
            zeta = H.ifft(H.fft(rho12 * fq, mesh) * cg * cell.vol / ng, mesh) * fq.conj()
            for k3 in range(nk):
                k4 = find(kidx[k3] - kidx[q])
                rho34 = np.einsum("gk,gl->klg", phi[k3].conj(), phi[k4]).reshape(-1, ng)
                eri_ref = zeta @ rho34.T
                eri = O.eri_from_w(wq[q], x[k1], x[k2], x[k3], x[k4]).reshape(eri_ref.shape)
                worst = max(worst, float(np.abs(eri - eri_ref).max() / np.abs(eri_ref).max()))
    return worst


def test_eri_reconstruction_against_exact_pair_densities():
    """the reference's de-facto acceptance test (fftdf-with-k-lstsq.py:219-258, abort above 1e-4)."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    cell = pk.random_cubic_cell(14, 6, seed=41, L=7.0, ltypes="sp")
    kmesh = [2, 1, 1]
    df = fftisdf.ISDF(cell, cell.get_kpts(kmesh), m0=[9, 9, 9], c0=12.0)
    df.ao_on_device = False
    df.build()
    worst = _eri_worst_error(df, cell, kmesh)
    assert worst < 1e-4, worst


@pytest.mark.parametrize("kmesh", [[2, 1, 1], [2, 2, 1]])
def test_exact_isdf_reproduces_eris(kmesh):
    """The known-answer test of the reference's isdf.py (:44-52 full-rank selection, :160-185 ERIs within 1e-10):
    when every grid point is a candidate and nip is the full numerical rank of the pair-density Gram, ISDF is
    exact, so the reconstructed ERIs equal the exact ones even though every A_q is singular (the fit is a
    consistent system; the oracle's zgelsy route gives 2e-12 / 6e-12 here)."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    cell = pk.random_cubic_cell(6, 2, seed=81, L=5.0, ltypes="s")
    df = fftisdf.ISDF(cell, cell.get_kpts(kmesh), m0=list(cell.mesh), c0=50.0)
    df.ao_on_device = False
    df.build()
    assert df._x.shape[1] < 2 * 50                           # nip is rank-limited (fftisdf.py:383)
    assert any(int(r) < df._x.shape[1] for r in df._ranks)   # ... and the per-q metrics are singular
    worst = _eri_worst_error(df, cell, kmesh)
    assert worst < 1e-9, worst


@pytest.mark.parametrize("name", ["k321_spd", "k231_odd", "k333_odd", "k434_odd", "gamma_s"])
def test_device_jk_equals_host_jk_and_reference(name):
    """get_j_kpts / get_k_kpts on the device == the oracle's numpy statement of fftisdf.py:133-228 applied
    to the device's own (X, W_q) == golden."""
    g, df = run_golden(name)
    dm2 = np.stack([g["dm"], g["dm"].conj().transpose(0, 2, 1) * 0.5])      # two density-matrix sets
    vj_d, vk_d = df.get_jk(dm2, kpts=g["kpts"])
    ph = H.get_phase(g["a"], g["kpts"], g["kmesh"].tolist())
    vj_h = O.get_j_kpts(df._x, df._w0, dm2)
    vk_h = O.get_k_kpts(df._x, df._wq, dm2, ph)
    assert rel(vj_d, vj_h) < 1e-12 and rel(vk_d, vk_h) < 1e-12
    assert rel(vj_d[0], g["vj"].reshape(vj_d[0].shape)) < 1e-10
    assert rel(vk_d[0], g["vk"].reshape(vk_d[0].shape)) < 1e-10


@pytest.mark.parametrize("kmesh", [[1, 2, 4], [4, 3, 1], [2, 3, 4], [3, 4, 2], [1, 5, 2], [6, 1, 1]])
def test_jk_on_every_supported_kmesh_shape(kmesh):
    """k-meshes whose register k-transform instantiations were missing in round 1 (permutations of (1,2,4), (1,3,4),
    (2,3,4): get_jk hit an assert) and meshes with an axis > 4 (shared-memory kernel, exchange modes added)."""
    import fft_isdf_scratch_b200 as pk
    cell = pk.random_cubic_cell(8, 5, seed=90 + sum(kmesh), L=6.0, ltypes="sp")
    df, out = _run_synth(cell, kmesh, [5, 5, 5], 2.0)
    nk, nao = len(df.kpts), cell.nao_nr()
    rng = np.random.default_rng(5)
    dm = rng.standard_normal((nk, nao, nao)) + 1j * rng.standard_normal((nk, nao, nao))
    dm = dm + dm.conj().transpose(0, 2, 1)
    tr = pk.pbc_tools.time_reversal_partner(kmesh)
    dm = 0.5 * (dm + dm[tr].conj())
    vj, vk = df.get_jk(dm, kpts=df.kpts)
    ph = H.get_phase(cell.a, df.kpts, kmesh)
    assert rel(vj, O.get_j_kpts(df._x, df._w0, dm[None])[0]) < 1e-12
    assert rel(vk, O.get_k_kpts(df._x, df._wq, dm[None], ph)[0]) < 1e-12


def _run_synth(cell, kmesh, m0, c0, **attrs):
    from fft_isdf_scratch_b200 import fftisdf
    kpts = cell.get_kpts(kmesh)
    df = fftisdf.ISDF(cell, kpts, m0=m0, c0=c0)
    df.ao_on_device = False
    for k, v in attrs.items():
        setattr(df, k, v)
    df.build()
    x0 = cell.eval_ao_kpts(cell.gen_uniform_grids(m0), df.kpts)
    coord = cell.gen_uniform_grids(cell.mesh)
    f_all = cell.eval_ao_kpts(coord, df.kpts)
    out = O.build(cell.a, df.kpts, kmesh, cell.mesh, x0, f_all, coord, c0)
    return df, out


def test_fallback_kernels_large_kmesh_axis_and_long_fft_axis():
    """k-mesh axis 5 (shared-memory k-transform, also for J/K) and a 50-point FFT axis (Stockham kernel
    instead of the tensor-core DFT) through the same build(): same parity bar as the fast paths."""
    import fft_isdf_scratch_b200 as pk
    cell = pk.random_cubic_cell(12, 10, seed=52, L=8.0, ltypes="spd")
    cell.mesh = [50, 8, 9]                      # cond(A_q) ~ 3e2 at c0 = 2: strict parity applies
    df, out = _run_synth(cell, [5, 1, 1], [7, 7, 7], 2.0, blksize=1000)
    assert np.array_equal(df._mask, out["mask"]) and np.array_equal(df._x, out["x"])
    conds = [np.linalg.cond(m) for m in out["x4_k"]]
    tol = max(1e-10, 50 * max(conds) * 2.2e-16)
    assert rel(df._wq, out["wq"]) < tol, (rel(df._wq, out["wq"]), max(conds))
    nk, nao = len(df.kpts), cell.nao_nr()
    rng = np.random.default_rng(5)
    dm = rng.standard_normal((nk, nao, nao)) + 1j * rng.standard_normal((nk, nao, nao))
    dm = dm + dm.conj().transpose(0, 2, 1)
    import fft_isdf_scratch_b200 as pk2
    tr = pk2.pbc_tools.time_reversal_partner([5, 1, 1])
    dm = 0.5 * (dm + dm[tr].conj())
    vj, vk = df.get_jk(dm, kpts=df.kpts)
    ph = H.get_phase(cell.a, df.kpts, [5, 1, 1])
    vj_ref = O.get_j_kpts(out["x"], out["wq"][0], dm[None])[0]
    vk_ref = O.get_k_kpts(out["x"], out["wq"], dm[None], ph)[0]
    assert rel(vj, vj_ref) < tol and rel(vk, vk_ref) < tol


def test_all_parent_points_selected_and_single_block():
    """Edge sizes: c0 so large that nip is limited by the pivoted-Cholesky rank / the parent grid (fftisdf.py:383),
    one aoR block covering the whole grid, Gamma point only."""
    import fft_isdf_scratch_b200 as pk
    cell = pk.random_cubic_cell(10, 3, seed=61, L=6.0, ltypes="s")
    df, out = _run_synth(cell, [1, 1, 1], [3, 3, 3], 50.0, blksize=100000)
    assert df._x.shape[1] == out["x"].shape[1] <= 27
    assert np.array_equal(df._mask, out["mask"])
    # nip is rank-limited here (A_q at the edge of numerical rank): check structure, not digits
    w = df._wq[0]
    assert np.isfinite(w).all() and np.abs(w - w.conj().T).max() == 0.0
    vj, vk = df.get_jk(np.eye(3)[None], kpts=df.kpts)          # real-dtype density matrix (common at Gamma)
    dmc = np.eye(3)[None, None].astype(complex)
    vj_h = O.get_j_kpts(df._x, df._w0, dmc)[0]
    vk_h = O.get_k_kpts(df._x, df._wq, dmc, H.get_phase(cell.a, df.kpts, [1, 1, 1]))[0]
    assert np.isfinite(vj).all() and np.isfinite(vk).all()
    assert rel(vj, vj_h) < 1e-10 and rel(vk, vk_h) < 1e-10


def test_kmesh_metric_equals_supercell_pair_gram():
    """BASELINE configs[2] / fftisdf-supercell-4.py:83-119: the k-mesh metric, unfolded with the Bloch phases, is the
    Hadamard square of the pair Gram of the explicit k2gamma SUPERCELL (real-space, Gamma point) at the same points:
    (P @ A_q)[R][I][J] == nk * ( sum_{S,mu} chi^sup_{S mu}(r_I) chi^sup_{S mu}(r_J + R) )^2."""
    import fft_isdf_scratch_b200 as pk
    from fft_isdf_scratch_b200 import fftisdf
    T = pk.pbc_tools
    cell = pk.random_cubic_cell(9, 5, seed=71, L=5.5, ltypes="sp")
    kmesh = [3, 1, 2]
    nk = 6
    df = fftisdf.ISDF(cell, cell.get_kpts(kmesh), m0=[5, 5, 5], c0=2.0)
    df.ao_on_device = False
    df.keep_metric = True
    df.use_time_reversal = False
    df.build()
    a_q = df._a_q.cpu().numpy()                                   # [nk, nip, nip]
    phase = H.get_phase(cell.a, df.kpts, kmesh)
    x4_s = (phase @ a_q.reshape(nk, -1)).reshape(a_q.shape)       # A = P^H x4_s  =>  x4_s = P A
    assert np.abs(x4_s.imag).max() < 1e-12 * np.abs(x4_s).max()
    # explicit supercell: lattice diag(kmesh) a, the primitive shells replicated on every image
    rvec = T.translation_vectors_for_kmesh(cell.a, kmesh)
    shells, i = [], 0
    while i < cell.nao_nr():
        l = {0: "s", 1: "p", 2: "d"}[sum(cell._terms[i][0][1])]
        shells.append((cell._cen[i], l, cell._alp[i]))
        i += {"s": 1, "p": 3, "d": 5}[l]
    sup = pk.SyntheticCell(np.diag(kmesh) @ cell.a, [(c + r, l, al) for r in rvec for (c, l, al) in shells], [9, 9, 9])
    pts = cell.gen_uniform_grids([5, 5, 5])[df._mask]
    gam = np.zeros((1, 3))
    phi0 = sup.eval_ao_kpts(pts, gam)[0].real
    scale = nk * np.abs(phi0 @ phi0.T).max() ** 2          # the R = 0 block sets the scale; far blocks are tiny
    for ir, r in enumerate(rvec):
        gram = phi0 @ sup.eval_ao_kpts(pts + r, gam)[0].real.T
        ref = nk * gram ** 2
        assert np.abs(x4_s[ir].real - ref).max() < 1e-12 * scale, ir


@pytest.mark.parametrize("name", ["k231_odd", "k222_sp"])
def test_trans_2e_device_route_equals_host_statement(name):
    """trans_2e (the reference's stub fftisdf.py:230-294): the device contraction through the
    library's GEMM kernels == the host numpy statement (itself pinned against the brute-force quadruple sum and the
    explicit supercell ERIs in tests/test_oracle_cpu.py), for general two-spin orbitals, unit_eri and symmetry = 4."""
    from fft_isdf_scratch_b200 import eri_transform as E
    g, df = run_golden(name)
    nk, nip, nao = df._x.shape
    assert df._wq_dev is not None
    rng = np.random.default_rng(5)
    nlo, nemb = 5, 3
    c_ao_lo = rng.standard_normal((2, nk, nao, nlo)) + 1j * rng.standard_normal((2, nk, nao, nlo))
    c_lo_eo = rng.standard_normal((1, nk, nlo, nemb))
    dev = E.trans_2e(df, C_ao_lo=c_ao_lo, C_lo_eo=c_lo_eo)                       # auto: on the device
    host = E.trans_2e(df, C_ao_lo=c_ao_lo, C_lo_eo=c_lo_eo, on_device=False)
    assert dev.shape == host.shape == (3, nemb, nemb, nemb, nemb)
    tol = 1e-10        # normwise bound of the 3M products over a W_q spanning many decades (k231_odd: < 1e-12)
    assert rel(dev, host) < tol
    dev_u = E.trans_2e(df, C_ao_lo=c_ao_lo[0], unit_eri=True)
    assert rel(dev_u, E.trans_2e(df, C_ao_lo=c_ao_lo[0], unit_eri=True, on_device=False)) < tol
    if name == "k231_odd":      # default orbitals (supercell AOs): real ERIs where W_{-q} = conj(W_q) holds to rounding
        p4 = E.trans_2e(df, symmetry=4)
        assert p4.dtype == np.float64 and rel(p4, E.trans_2e(df, symmetry=4, on_device=False)) < tol
