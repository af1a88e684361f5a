"""Kernel-level parity: every C-ABI entry point against numpy/scipy on the same seeded inputs.
Tolerances are stated per test (FP64; relative to the largest reference entry)."""
import numpy as np
import pytest
import scipy.linalg
import torch

from oracle import pbc_helpers as H
from oracle import isdf_oracle as O

pytestmark = pytest.mark.gpu


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("shape", [(3, 150, 70, 37), (1, 128, 64, 8), (2, 5, 3, 1), (1, 257, 129, 26)])
def test_gram_conja(ops, shape):
    batch, m, n, k = shape
    rng = np.random.default_rng(1)
    a, b = crand(rng, batch, m, k), crand(rng, batch, n, k)
    c = ops.gram_conja(dev(a), dev(b)).cpu().numpy()
    ref = np.einsum("zik,zjk->zij", a.conj(), b)
    assert relerr(c, ref) < 1e-13


def test_select_gram(ops):
    rng = np.random.default_rng(2)
    nk, n0, nao = 3, 203, 7
    x0 = crand(rng, nk, n0, nao)
    x4 = ops.select_gram(dev(x0)).cpu().numpy()
    x2 = sum((x0[q].conj() @ x0[q].T).real for q in range(nk))
    ref = x2 * x2 / nk
    assert relerr(x4.real, ref) < 1e-13
    assert np.abs(x4.imag).max() == 0.0
    assert np.array_equal(x4.real, x4.real.T)  # exactly symmetric (mirrored tiles)


@pytest.mark.parametrize("shape", [(2, 64, 300, 130), (1, 64, 128, 64), (3, 40, 77, 19)])
def test_gemm_nn(ops, shape):
    batch, m, n, k = shape
    rng = np.random.default_rng(3)
    a, b = crand(rng, batch, m, k), crand(rng, batch, k, n)
    c = ops.gemm_nn(dev(a), dev(b)).cpu().numpy()
    assert relerr(c, a @ b) < 1e-13


def test_herk_scatter(ops):
    rng = np.random.default_rng(4)
    batch, n, k = 2, 150, 1003
    b = crand(rng, batch, n, k)
    perm = np.stack([rng.permutation(n) for _ in range(batch)]).astype(np.int32)
    w = ops.herk(dev(b), alpha=0.5, perm=dev(perm)).cpu().numpy()
    for z in range(batch):
        ref = np.zeros((n, n), complex)
        ref[np.ix_(perm[z], perm[z])] = 0.5 * b[z] @ b[z].conj().T
        assert relerr(w[z], ref) < 1e-13
        assert np.array_equal(w[z], w[z].conj().T)  # exactly Hermitian


def _psd(rng, n, r, complex_=True):
    b = crand(rng, n, r) if complex_ else rng.standard_normal((n, r)) + 0j
    return b @ b.conj().T


@pytest.mark.parametrize("n,r,nb", [(200, 200, 32), (130, 40, 16), (300, 300, 64), (65, 65, 7), (1100, 300, 32),
                                    (2051, 2051, 32)])
def test_pchol_complex(ops, n, r, nb):
    rng = np.random.default_rng(5)
    a = np.stack([_psd(rng, n, r), _psd(rng, n, r)])
    u, piv, rank, nxt = ops.pchol(dev(a.copy()), max_steps=n, tol=-1.0, nb=nb)
    u, piv, rank = u.cpu().numpy(), piv.cpu().numpy(), rank.cpu().numpy()
    for z in range(2):
        rk = int(rank[z])
        assert (rk == n) if r == n else (r <= rk <= r + 2)
        assert sorted(piv[z].tolist()) == list(range(n))
        rec = u[z, :rk].conj().T @ u[z, :rk]
        assert relerr(rec, a[z]) < 1e-11
        up = u[z, :rk][:, piv[z][:rk]]
        assert np.abs(np.tril(up, -1)).max() == 0.0  # upper triangular in pivot order
        d = np.diag(up).real
        assert np.all(d[:-1] >= d[1:] * (1 - 1e-12))  # non-increasing pivots


def test_pchol_matches_dpstrf_real(ops):
    """Real symmetric input carried as complex: pivots identical to LAPACK dpstrf (tie-free input)."""
    rng = np.random.default_rng(6)
    n = 257
    x = rng.standard_normal((n, 60))
    x2 = x @ x.T
    x4 = x2 * x2
    _, piv_ref, rank_ref = H.pivoted_cholesky(x4.copy())
    nsteps = 120
    u, piv, rank, nxt = ops.pchol(dev(x4[None].astype(complex)), max_steps=nsteps, tol=-1.0, nb=32)
    piv = piv.cpu().numpy()[0]
    assert int(rank.cpu()[0]) == nsteps
    assert np.array_equal(piv[:nsteps], piv_ref[:nsteps])
    p2, steps, nx = H.pivoted_cholesky_steps(x4, nsteps)
    assert np.array_equal(piv[:nsteps], p2)
    assert abs(float(nxt.cpu()[0]) - nx) <= 1e-9 * abs(nx)


def test_pchol_early_stop_and_zero_steps(ops):
    rng = np.random.default_rng(7)
    a = _psd(rng, 90, 90)[None]
    u, piv, rank, nxt = ops.pchol(dev(a.copy()), max_steps=0)
    assert int(rank.cpu()[0]) == 0
    assert abs(float(nxt.cpu()[0]) - np.diag(a[0]).real.max()) < 1e-12 * np.diag(a[0]).real.max()
    u, piv, rank, nxt = ops.pchol(dev(a.copy()), max_steps=90, tol=1e300)
    assert int(rank.cpu()[0]) == 0


@pytest.mark.parametrize("trim", [False, True])
@pytest.mark.parametrize("n,ng", [(100, 517), (64, 128), (200, 1000), (157, 333)])
def test_trsm_solve(ops, n, ng, trim):
    rng = np.random.default_rng(8)
    batch = 2
    a = np.stack([_psd(rng, n, n + 20) + 0.1 * np.eye(n) for _ in range(batch)])
    y = crand(rng, batch, n, ng)
    u, piv, rank, _ = ops.pchol(dev(a.copy()), max_steps=n)
    nP = -(-n // 64) * 64
    lfwd, ubwd = ops.trsm_prepare(u, piv, rank, nP)
    pivh = piv.cpu().numpy()
    t = np.zeros((batch, nP, ng), complex)
    for z in range(batch):
        t[z, :n] = y[z][pivh[z]]
    td = dev(t)
    # trim: only the first n rows (>= every rank) are swept; the padding rows are neither read nor written
    ops.trsm_sweeps(lfwd, ubwd, td, nact=n if trim else None)
    sol = td.cpu().numpy()
    for z in range(batch):
        ref = np.linalg.solve(a[z], y[z])
        got = np.zeros_like(ref)
        got[pivh[z]] = sol[z, :n]
        assert relerr(got, ref) < 1e-10
        assert nP == n or np.abs(sol[z, n:]).max() == 0.0


@pytest.mark.parametrize("kmesh", [[1, 1, 1], [2, 2, 2], [3, 2, 1], [4, 4, 4], [1, 5, 3]])
def test_ktransform_metric_and_rhs(ops, kmesh):
    rng = np.random.default_rng(9)
    nk = int(np.prod(kmesh))
    a = np.eye(3) * 6.0 + 0.3 * rng.standard_normal((3, 3))
    kpts = H.get_kpts(a, kmesh)
    phase = H.get_phase(a, kpts, kmesh)
    nip, blk = 21, 37
    # time-reversal symmetric real-space tables -> k-space (so that phase @ v is real)
    vs = rng.standard_normal((nk, blk * nip))
    vk = (phase.conj().T @ vs).reshape(nk, blk, nip)
    uax = ops.pack_uaxes(kmesh)
    diag = torch.zeros(2, dtype=torch.float64, device="cuda")
    # rhs form: out[q][i][g] (transposed), phase.T second transform
    out = torch.zeros((nk, nip, 50), dtype=torch.complex128, device="cuda")
    ops.ktransform_square(dev(vk), blk * nip, nip, out, nip * 50, 1, 50, 5, blk, nip, kmesh, uax, 0, 1, diag=diag)
    ys = (phase @ vk.reshape(nk, -1))
    ref = (phase.T @ (ys * ys)).reshape(nk, blk, nip)
    got = out.cpu().numpy()[:, :, 5:5 + blk].transpose(0, 2, 1)
    assert relerr(got, ref) < 1e-13
    d = diag.cpu().numpy()
    assert d[0] < 1e-12 and abs(d[1] - np.abs(ys.real).max()) < 1e-12
    # metric form: out[q][g][i], phase^H second transform
    out2 = torch.zeros((nk, blk, nip), dtype=torch.complex128, device="cuda")
    ops.ktransform_square(dev(vk), blk * nip, nip, out2, blk * nip, nip, 1, 0, blk, nip, kmesh, uax, 1, 0)
    ref2 = (phase.conj().T @ (ys * ys)).reshape(nk, blk, nip)
    assert relerr(out2.cpu().numpy(), ref2) < 1e-13


def test_ktransform_qslot_rowmap(ops):
    rng = np.random.default_rng(10)
    kmesh = [2, 2, 1]
    nk = 4
    a = np.eye(3) * 5.0
    phase = H.get_phase(a, H.get_kpts(a, kmesh), kmesh)
    nip, blk = 10, 9
    vs = rng.standard_normal((nk, blk * nip))
    vk = (phase.conj().T @ vs).reshape(nk, blk, nip)
    qslot = np.array([0, -1, 1, -1], dtype=np.int32)
    rowmap = np.stack([rng.permutation(nip), rng.permutation(nip)]).astype(np.int32)
    rowmap[0, 3] = -1
    out = torch.zeros((2, nip, blk), dtype=torch.complex128, device="cuda")
    ops.ktransform_square(dev(vk), blk * nip, nip, out, nip * blk, 1, blk, 0, blk, nip, kmesh, ops.pack_uaxes(kmesh),
                          0, 1, qslot=dev(qslot), rowmap=dev(rowmap), rowmap_sq=nip)
    ys = phase @ vk.reshape(nk, -1)
    full = (phase.T @ (ys * ys)).reshape(nk, blk, nip)
    got = out.cpu().numpy()
    for slot, q in enumerate([0, 2]):
        ref = np.zeros((nip, blk), complex)
        for i in range(nip):
            if rowmap[slot, i] >= 0:
                ref[rowmap[slot, i]] = full[q][:, i]
        assert relerr(got[slot], ref) < 1e-13


@pytest.mark.parametrize("mesh", [[9, 10, 12], [15, 15, 15], [37, 5, 4], [32, 32, 32], [1, 1, 7], [33, 31, 36]])
def test_fft3d(ops, mesh):
    rng = np.random.default_rng(11)
    ng = int(np.prod(mesh))
    nvec = 5
    x = crand(rng, nvec, ng)
    pre = np.exp(1j * rng.uniform(0, 6.28, ng))
    post = rng.uniform(0.1, 2.0, ng)
    d = dev(x.copy())
    ops.fft3d(d, mesh, pre=dev(pre), post=dev(post), group_vecs=2)
    ref = np.fft.fftn((x * pre).reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng) * post
    assert relerr(d.cpu().numpy(), ref) < 1e-13
    d = dev(x.copy())
    ops.fft3d(d, mesh)
    ref = np.fft.fftn(x.reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng)
    assert relerr(d.cpu().numpy(), ref) < 1e-13


@pytest.mark.parametrize("mesh", [[33, 33, 33], [15, 15, 15], [64, 64, 16], [8, 96, 96], [48, 40, 27], [17, 19, 23],
                                  [37, 37, 3], [2, 3, 1], [1, 35, 1], [5, 1, 26], [13, 11, 7], [100, 4, 6]])
def test_fft3d_stockham_kernels(ops, mesh):
    """The shared-memory Stockham path on its own: every hard-coded radix (2,3,4,5,7,8,11,13), the direct-DFT stage for
    larger primes, fused planes and the unfused fallback (96 x 96 plane), degenerate axes, a padded vector pitch."""
    rng = np.random.default_rng(21)
    ng = int(np.prod(mesh))
    nvec, ldv = 3, ng + 3
    x = crand(rng, nvec, ldv)
    pre = np.exp(1j * rng.uniform(0, 6.28, ng))
    post = rng.uniform(0.1, 2.0, ng)
    d = dev(x.copy())
    ops.fft3d(d, mesh, pre=dev(pre), post=dev(post), nvec=nvec, ldv=ldv, mode="stockham", group_vecs=2)
    got = d.cpu().numpy()
    ref = np.fft.fftn((x[:, :ng] * pre).reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng) * post
    assert relerr(got[:, :ng], ref) < 1e-13
    assert np.array_equal(got[:, ng:], x[:, ng:])          # padding untouched


@pytest.mark.parametrize("mesh", [[32, 32, 32], [33, 33, 33], [37, 37, 37], [48, 48, 48], [64, 64, 64], [15, 96, 96],
                                  [1, 45, 45], [61, 17, 17], [108, 25, 25], [16, 105, 105], [27, 59, 59]])
def test_fft3d_register_kernels(ops, mesh):
    """Register-resident FFT (fft_reg.cu): two-factor lengths (in-place plane buffer, one exchange per axis), direct
    symmetric DFT for primes, mixed x / plane lengths, a 2-D mesh, a padded vector pitch, ragged x tiles, groups."""
    assert ops.fft3d_reg_supported(mesh)
    rng = np.random.default_rng(31)
    ng = int(np.prod(mesh))
    nvec, ldv = 3, ng + 3
    x = crand(rng, nvec, ldv)
    pre = np.exp(1j * rng.uniform(0, 6.28, ng))
    post = rng.uniform(0.1, 2.0, ng)
    d = dev(x.copy())
    ops.fft3d(d, mesh, pre=dev(pre), post=dev(post), nvec=nvec, ldv=ldv, mode="reg", group_vecs=2)
    got = d.cpu().numpy()
    ref = np.fft.fftn((x[:, :ng] * pre).reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng) * post
    assert relerr(got[:, :ng], ref) < 1e-13
    assert np.array_equal(got[:, ng:], x[:, ng:])          # padding untouched
    d = dev(x.copy())
    ops.fft3d(d, mesh, nvec=nvec, ldv=ldv)                 # auto routing picks the same kernels; no phase / weight
    ref = np.fft.fftn(x[:, :ng].reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng)
    assert relerr(d.cpu().numpy()[:, :ng], ref) < 1e-13
    if mesh[0] == mesh[1] == mesh[2]:                      # single persistent launch with dependency counters
        nvec = 70 if ng < 40000 else 9                     # several groups of vectors
        x = crand(rng, nvec, ng)
        d = dev(x.copy())
        ops.fft3d(d, mesh, pre=dev(pre), post=dev(post), mode="reg-fused")
        ref = np.fft.fftn((x * pre).reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng) * post
        assert relerr(d.cpu().numpy(), ref) < 1e-13


def test_fft3d_register_kernels_every_length(ops):
    """Every instantiated axis length once (n x n planes, x pass of the same length on a thin mesh)."""
    rng = np.random.default_rng(32)
    for n in range(15, 109):
        if not ops.fft3d_reg_supported([1, n, n]):
            continue
        for mesh in ([1, n, n], [n, 15, 15]):
            ng = int(np.prod(mesh))
            x = crand(rng, 2, ng)
            d = dev(x.copy())
            ops.fft3d(d, mesh, mode="reg")
            ref = np.fft.fftn(x.reshape(2, *mesh), axes=(1, 2, 3)).reshape(2, ng)
            assert relerr(d.cpu().numpy(), ref) < 1e-13, mesh
    assert not ops.fft3d_reg_supported([8, 64, 32]) and not ops.fft3d_reg_supported([4, 127, 127])


def test_gemm_hn_herm(ops):
    """a^H (H a) with Hermitian H: the lower-tile product equals the full product, exactly Hermitian, real diagonal."""
    rng = np.random.default_rng(41)
    for k, n in [(70, 157), (130, 64), (33, 200)]:
        a = crand(rng, 2, k, n)
        hm = crand(rng, 2, k, k)
        hm = hm + hm.conj().transpose(0, 2, 1)
        b = hm @ a
        out = ops.gemm_hn_herm(dev(a), dev(b)).cpu().numpy()
        ref = a.conj().transpose(0, 2, 1) @ b
        assert relerr(out, ref) < 1e-13
        assert np.array_equal(out, out.conj().transpose(0, 2, 1))
        assert np.all(out[:, np.arange(n), np.arange(n)].imag == 0.0)


def test_gather_and_conj(ops):
    rng = np.random.default_rng(12)
    src = crand(rng, 2, 11, 301)
    idx = np.array([[3, -1, 0, 10], [1, 1, -1, 5]], dtype=np.int32)
    out = ops.gather_rows(dev(src), dev(idx)).cpu().numpy()
    for z in range(2):
        for i in range(4):
            ref = src[z, idx[z, i]] if idx[z, i] >= 0 else np.zeros(301)
            assert np.array_equal(out[z, i], ref)
    s = dev(src)
    d = torch.empty_like(s)
    ops.conj_copy(s, d)
    assert np.array_equal(d.cpu().numpy(), src.conj())


@pytest.mark.parametrize("mesh,kmesh", [([9, 11, 7], [2, 3, 1]), ([8, 9, 10], [3, 2, 1]), ([12, 12, 12], [2, 2, 2]),
                                        ([15, 15, 15], [4, 4, 4])])
def test_device_coulomb_weights_and_phase(ops, mesh, kmesh):
    """Device tables == host restatement of pbctools.get_coulG (wrap-around, boundary zeroing) and exp(-iq.r)."""
    import fft_isdf_scratch_b200 as pk
    T = pk.pbc_tools
    a = np.array([[6.1, 0.9, 0.0], [0.0, 5.3, -0.7], [0.5, 0.0, 7.2]])
    kpts = T.make_kpts(a, kmesh)
    ks = T.get_scaled_kpts(a, kpts)
    b = T.reciprocal_vectors(a)
    ng = int(np.prod(mesh))
    vol = abs(np.linalg.det(a))
    coords = T.gen_uniform_grids(a, mesh)
    cd = dev(coords)
    w = torch.empty(ng, dtype=torch.float64, device="cuda")
    f = torch.empty(ng, dtype=torch.complex128, device="cuda")
    for q in range(len(kpts)):
        ops.coulomb_weights(b, ks[q], mesh, vol, w)
        ref = np.sqrt(T.get_coulG(a, kpts[q], mesh) * vol) / ng
        got = w.cpu().numpy()
        assert np.array_equal(got == 0.0, ref == 0.0)          # same zeroed entries (G=0, box boundary)
        assert relerr(got, ref) < 1e-13
        ops.phase_table(cd, kpts[q], f)
        assert relerr(f.cpu().numpy(), np.exp(-1j * coords @ kpts[q])) < 1e-13


@pytest.mark.parametrize("kmesh", [[1, 1, 1], [2, 2, 2], [3, 3, 3], [3, 2, 1], [1, 4, 4], [2, 4, 4], [4, 4, 4], [5, 1, 1],
                                   [3, 4, 4], [4, 3, 3], [3, 3, 4], [3, 4, 3]])
def test_ktransform_rows_register_path(ops, kmesh):
    """Register-resident k-transform == dense phase-matrix arithmetic of the reference (or a clean
    'unsupported' for meshes that must use the shared-memory kernel)."""
    rng = np.random.default_rng(13)
    nk = int(np.prod(kmesh))
    a = np.eye(3) * 6.0 + 0.3 * rng.standard_normal((3, 3))
    phase = H.get_phase(a, H.get_kpts(a, kmesh), kmesh)
    nrows, ncols = 11, 203
    vs = rng.standard_normal((nk, nrows * ncols))
    vk = (phase.conj().T @ vs).reshape(nk, nrows, ncols)
    qslot = np.arange(nk, dtype=np.int32)
    qslot[nk // 2] = -1 if nk > 1 else 0
    rowmap = np.stack([rng.permutation(nrows) for _ in range(nk)]).astype(np.int32)
    rowmap[0, 2] = -1
    out = torch.zeros((nk, nrows, 300), dtype=torch.complex128, device="cuda")
    diag = torch.zeros(2, dtype=torch.float64, device="cuda")
    ok = ops.ktransform_rows(dev(vk), nrows * ncols, ncols, out, nrows * 300, 300, 7, nrows, ncols, kmesh,
                             ops.pack_uaxes_host(kmesh), conj2=0, qslot=dev(qslot), rowmap=dev(rowmap), rowmap_sq=nrows,
                             diag=diag)
    if max(kmesh) > 4:
        assert not ok
        return
    assert ok
    ys = phase @ vk.reshape(nk, -1)
    full = (phase.T @ (ys * ys)).reshape(nk, nrows, ncols)
    got = out.cpu().numpy()
    for q in range(nk):
        ref = np.zeros((nrows, 300), complex)
        if qslot[q] >= 0:
            for r in range(nrows):
                if rowmap[qslot[q], r] >= 0:
                    ref[rowmap[qslot[q], r], 7:7 + ncols] = full[q, r]
            assert relerr(got[qslot[q]], ref) < 1e-13
    assert diag.cpu().numpy()[0] < 1e-12
    out2 = torch.zeros((nk, nrows, ncols), dtype=torch.complex128, device="cuda")
    assert ops.ktransform_rows(dev(vk), nrows * ncols, ncols, out2, nrows * ncols, ncols, 0, nrows, ncols, kmesh,
                               ops.pack_uaxes_host(kmesh), conj2=1)
    ref2 = (phase.conj().T @ (ys * ys)).reshape(nk, nrows, ncols)
    assert relerr(out2.cpu().numpy(), ref2) < 1e-13


def test_gram_conjb(ops):
    rng = np.random.default_rng(14)
    a, b = crand(rng, 3, 150, 26), crand(rng, 3, 333, 26)
    c = ops.gram_conjb(dev(a), dev(b)).cpu().numpy()
    assert relerr(c, np.einsum("zik,zjk->zij", a, b.conj())) < 1e-13


def test_pchol_cluster_matches_dpstrf_large(ops):
    """n >= 256 runs on the 8-CTA cluster panel kernel (DSMEM): same pivots as LAPACK dpstrf."""
    rng = np.random.default_rng(15)
    n = 1500
    x = rng.standard_normal((n, 40))
    x4 = (x @ x.T) ** 2
    _, piv_ref, _ = H.pivoted_cholesky(x4.copy())
    nsteps = 333
    u, piv, rank, nxt = ops.pchol(dev(x4[None].astype(complex)), max_steps=nsteps, tol=-1.0, nb=32)
    assert int(rank.cpu()[0]) == nsteps
    assert np.array_equal(piv.cpu().numpy()[0][:nsteps], piv_ref[:nsteps])


def test_device_ao_evaluation_matches_host(ops):
    import fft_isdf_scratch_b200 as pk
    cell = pk.random_cubic_cell(10, 14, seed=31, L=7.5, ltypes="spd")
    a = cell.a.copy(); a[0, 1] = 0.8; a[2, 0] = -0.5
    cell = pk.SyntheticCell(a, [(cell._cen[i], "s", cell._alp[i]) for i in range(3)] +
                            [(cell._cen[4], "p", 0.7), (cell._cen[8], "d", 0.9)], [10, 10, 10])
    kpts = cell.get_kpts([2, 3, 2])
    coords = cell.gen_uniform_grids([7, 6, 5]) + 0.123
    ref = cell.eval_ao_kpts(coords, kpts)
    got = cell.eval_ao_device(ops, coords, kpts).cpu().numpy()
    assert got.shape == ref.shape
    assert relerr(got, ref) < 1e-13


@pytest.mark.parametrize("mesh", [[37, 37, 37], [31, 33, 37], [9, 10, 12], [48, 47, 41], [15, 15, 15], [2, 3, 5]])
def test_dft3d_tensor_core_path(ops, mesh):
    """Dense-DFT-on-DMMA path == numpy fftn (with fused phase and weight), incl. a padded vector pitch."""
    rng = np.random.default_rng(16)
    ng = int(np.prod(mesh))
    nvec, ldv = 7, ng + 5
    x = crand(rng, nvec, ldv)
    pre = np.exp(1j * rng.uniform(0, 6.28, ng))
    post = rng.uniform(0.1, 2.0, ng)
    d = dev(x.copy())
    ops.fft3d(d, mesh, pre=dev(pre), post=dev(post), nvec=nvec, ldv=ldv, mode="dmma")
    got = d.cpu().numpy()
    ref = np.fft.fftn((x[:, :ng] * pre).reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng) * post
    assert relerr(got[:, :ng], ref) < 1e-13
    assert np.array_equal(got[:, ng:], x[:, ng:])          # padding untouched
    d = dev(x.copy())
    ops.fft3d(d, mesh, nvec=nvec, ldv=ldv, mode="dmma")
    ref = np.fft.fftn(x[:, :ng].reshape(nvec, *mesh), axes=(1, 2, 3)).reshape(nvec, ng)
    assert relerr(d.cpu().numpy()[:, :ng], ref) < 1e-13


def test_cabi_error_conventions(ops):
    """Status codes, never exceptions, across the ABI: < 0 argument error with a message in isdf_last_error."""
    import ctypes as C
    lib, h = ops.lib, ops.h
    a = torch.zeros((1, 8, 8), dtype=torch.complex128, device="cuda")
    rc = lib.isdf_pchol(h, C.c_void_p(a.data_ptr()), 8, 1, 8, C.c_double(-1.0), 999, C.c_void_p(a.data_ptr()), 8,
                        C.c_void_p(a.data_ptr()), C.c_void_p(a.data_ptr()), None, C.c_void_p(a.data_ptr()), None)
    assert rc < 0 and b"nb <= 64" in lib.isdf_last_error(h)
    rc = lib.isdf_select_gram(h, None, 1, 8, 8, C.c_void_p(a.data_ptr()), None)
    assert rc < 0 and b"null pointer" in lib.isdf_last_error(h)
    m = (C.c_int * 3)(64, 5, 5)
    rc = lib.isdf_dft3d_dmma(h, C.c_void_p(a.data_ptr()), 1, 1600, m, None, None, None)
    assert rc == -2                                   # axis > 48: "use the Stockham entry point", nothing launched
    with pytest.raises(Exception):
        ops.fft3d(a.reshape(-1), [64, 5, 5], mode="dmma")
    torch.cuda.synchronize()
