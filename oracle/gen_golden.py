"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the REFERENCE'S OWN CODE.

    python -m oracle.gen_golden            (build container only: needs /root/reference)

/root/reference/fftisdf.py is imported unmodified.  PySCF/opt_einsum/h5py are not installed,
so the modules it imports are stubbed with thin adapters around oracle/pbc_helpers.py
(get_phase, fft/ifft, get_coulG, pivoted_cholesky=scipy dpstrf, an in-memory H5TmpFile, a
block_loop over precomputed AO tables).  The reference's `build(df_obj)`,
`select_interpolation_points`, `get_j_kpts`, `get_k_kpts` then run line by line on a small
synthetic cell and their inputs/outputs are stored.  This pins oracle/isdf_oracle.py's
restatement of fftisdf.py:22-228,357-388; it cannot pin PySCF's own helpers (unavailable).
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("ISDF_REFERENCE_DIR", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import pbc_helpers as H  # noqa: E402

GELSY_RANKS = {}   # q -> rank, filled from the reference's own log line during build()


def _install_stubs():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    oe = mod("opt_einsum")
    oe.contract = np.einsum
    pyscf = mod("pyscf")
    lib = mod("pyscf.lib")
    logger = mod("pyscf.lib.logger")
    scipy_helper = mod("pyscf.lib.scipy_helper")
    pbc = mod("pyscf.pbc")
    df = mod("pyscf.pbc.df")
    dffft = mod("pyscf.pbc.df.fft")
    dfjk = mod("pyscf.pbc.df.df_jk")
    aft = mod("pyscf.pbc.df.aft")
    tools = mod("pyscf.pbc.tools")
    k2gamma = mod("pyscf.pbc.tools.k2gamma")
    pbclib = mod("pyscf.pbc.lib")
    kpts_helper = mod("pyscf.pbc.lib.kpts_helper")
    scf = mod("pyscf.pbc.scf")
    pyscf.lib, pyscf.pbc = lib, pbc
    pbc.df, pbc.tools, pbc.lib, pbc.scf = df, tools, pbclib, scf
    df.fft, df.df_jk, df.aft = dffft, dfjk, aft
    tools.k2gamma = k2gamma
    pbclib.kpts_helper = kpts_helper
    lib.logger, lib.scipy_helper = logger, scipy_helper

    import time

    logger.process_clock = time.process_time
    logger.perf_counter = time.perf_counter

    class _Log:
        def info(self, *a):
            pass

        debug = info

        def timer(self, msg, *t0):
            # the reference reports zgelsy's rank only through this message (fftisdf.py:122)
            m = re.match(r"w\[\s*(\d+)\], rank =\s*(\d+) /", msg)
            if m:
                GELSY_RANKS[int(m.group(1))] = int(m.group(2))
            return (time.process_time(), time.perf_counter())

    logger.new_logger = lambda obj=None, verbose=None: _Log()
    lib.current_memory = lambda: (0.0, 0.0)
    lib.asarray = lambda x, order=None: np.asarray(x, order=order)

    class H5TmpFile(dict):
        filename = "<memory>"

        def create_dataset(self, name, shape=None, dtype=None):
            self[name] = np.zeros(shape, dtype=dtype)
            return self[name]

    lib.H5TmpFile = H5TmpFile
    scipy_helper.pivoted_cholesky = H.pivoted_cholesky

    tools.fft = H.fft
    tools.ifft = H.ifft
    tools.get_coulG = lambda cell, k=np.zeros(3), mesh=None, Gv=None, **kw: H.get_coulG(
        cell.lattice_vectors(), np.asarray(k), mesh, Gv=Gv)
    k2gamma.get_phase = lambda cell, kpts, kmesh=None, wrap_around=False: (
        None, H.get_phase(cell.lattice_vectors(), np.asarray(kpts), kmesh))

    def kpts_to_kmesh(cell, kpts):
        sk = (np.asarray(kpts) @ cell.lattice_vectors().T / (2 * np.pi)).round(8)
        return [len(np.unique(sk[:, i])) for i in range(3)]

    k2gamma.kpts_to_kmesh = kpts_to_kmesh
    kpts_helper.is_zero = lambda k: bool(abs(np.asarray(k)).max() < 1e-9)

    def _format_dms(dm_kpts, kpts):
        nkpts = len(kpts)
        nao = dm_kpts.shape[-1]
        return dm_kpts.reshape(-1, nkpts, nao, nao)

    def _format_kpts_band(kpts_band, kpts):
        if kpts_band is None:
            kpts_band = kpts
        return np.reshape(kpts_band, (-1, 3))

    def _format_jks(v_kpts, dm_kpts, kpts_band, kpts):
        assert kpts_band is None or kpts_band is kpts
        return v_kpts.reshape(dm_kpts.shape)

    dfjk._format_dms, dfjk._format_kpts_band, dfjk._format_jks = _format_dms, _format_kpts_band, _format_jks
    aft._check_kpts = lambda mydf, kpts: (np.asarray(mydf.kpts if kpts is None else kpts), False)

    class _Grids:
        non0tab = True

        def __init__(self, cell):
            self.cell = cell
            self.coords = cell.gen_uniform_grids(cell.mesh)

    class _NumInt:
        def block_loop(self, cell, grids, nao, deriv, kpts, max_memory=None, blksize=None):
            coords = grids.coords
            for p0 in range(0, len(coords), blksize):
                c = coords[p0:p0 + blksize]
                ao = cell.pbc_eval_gto("GTOval", c, kpts=kpts)
                yield ao, ao, None, None, c

    class FFTDF:
        def __init__(self, cell, kpts=np.zeros((1, 3))):
            self.cell = cell
            self.kpts = np.asarray(kpts)
            self.mesh = cell.mesh
            self.grids = _Grids(cell)
            self._numint = _NumInt()
            self.verbose = 0
            self.max_memory = 4000
            self.stdout = sys.stdout

    dffft.FFTDF = FFTDF


def load_reference():
    _install_stubs()
    spec = importlib.util.spec_from_file_location("ref_fftisdf", os.path.join(REF, "fftisdf.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


CASES = {
    # name: (kind, mesh_side, nao, seed, kmesh, m0, c0, ltypes, blksize)
    "gamma_s": dict(mesh=[10, 10, 10], nao=5, seed=11, kmesh=[1, 1, 1], m0=[6, 6, 6], c0=4.0, ltypes="s", blksize=300),
    "k222_sp": dict(mesh=[9, 9, 9], nao=6, seed=12, kmesh=[2, 2, 2], m0=[6, 6, 6], c0=5.0, ltypes="sp", blksize=250),
    "k321_spd": dict(mesh=[8, 9, 10], nao=8, seed=13, kmesh=[3, 2, 1], m0=[5, 6, 7], c0=4.0, ltypes="spd", blksize=8000,
                     skew=True),
    # odd mesh (what PySCF's cutoff_to_mesh produces), skewed lattice, nip below the local pair rank:
    # numerically full-rank A_q and exact time-reversal symmetry W_{-q} = W_q^*
    "k231_odd": dict(mesh=[9, 11, 7], nao=8, seed=14, kmesh=[2, 3, 1], m0=[6, 7, 5], c0=3.0, ltypes="spd", blksize=500,
                     skew=True),
    # the headline k-mesh (3x3x3: lane-split register k-transform, 13 time-reversal pairs) and a 48-point mesh with
    # two axes of length 4 (four-lane split), both on odd FFT meshes with full-rank, well-conditioned A_q
    # (cond ~ 2e4; seeds 15/16 or c0 = 3 give cond(A_q) ~ 1e9-1e12 on these coarse parent grids, where the
    # reference's own zgelsy output only reproduces to ~1e-8)
    "k333_odd": dict(mesh=[7, 7, 9], nao=8, seed=17, kmesh=[3, 3, 3], m0=[5, 5, 5], c0=2.5, ltypes="spd", blksize=200,
                     skew=True),
    # mid-size RANK-DEFICIENT case (the regime of the reference's defaults: nip = nao*c0 beyond the local pair rank,
    # zgelsy truncates every A_q), odd FFT mesh; the reference's per-q ranks are stored next to its outputs
    "k221_rd": dict(mesh=[11, 11, 13], nao=10, seed=31, kmesh=[2, 2, 1], m0=[8, 8, 8], c0=12.0, ltypes="sp", blksize=600),
    "k434_odd": dict(mesh=[5, 7, 5], nao=8, seed=17, kmesh=[4, 3, 4], m0=[4, 5, 4], c0=2.5, ltypes="spd", blksize=8000,
                     skew=True),
}


def make_cell(spec):
    import fft_isdf_scratch_b200 as pk
    cell = pk.random_cubic_cell(spec["mesh"][0], spec["nao"], spec["seed"], L=7.0, ltypes=spec["ltypes"])
    if spec.get("skew"):
        a = cell.a.copy()
        a[0, 1] = 0.9
        a[1, 2] = -0.7
        a[2, 0] = 0.5
        cell = pk.SyntheticCell(a, _shells_of(cell), spec["mesh"])
    cell.mesh = list(spec["mesh"])
    return cell


def _shells_of(cell):
    # rebuild shell list from the expanded AO list (s:1, p:3, d:5 consecutive entries share centre/alpha)
    shells, i, n = [], 0, cell.nao_nr()
    while i < n:
        nterm = len(cell._terms[i])
        lsum = sum(cell._terms[i][0][1])
        l = {0: "s", 1: "p", 2: "d"}[lsum]
        shells.append((cell._cen[i], l, cell._alp[i]))
        i += {"s": 1, "p": 3, "d": 5}[l]
    return shells


def run_case(ref, name, spec):
    cell = make_cell(spec)
    kmesh = spec["kmesh"]
    kpts = cell.get_kpts(kmesh)
    ref.cell = cell  # fftisdf.py:322 reads the *global* `cell`
    df = ref.ISDF(cell, kpts, m0=spec["m0"], c0=spec["c0"])
    df.blksize = spec["blksize"]
    GELSY_RANKS.clear()
    df.build()
    ranks = np.asarray([GELSY_RANKS[q] for q in range(len(kpts))])
    x0 = np.asarray(cell.pbc_eval_gto("GTOval", cell.gen_uniform_grids(spec["m0"]), kpts=df.kpts))
    coord = df.grids.coords
    f_all = np.asarray(cell.pbc_eval_gto("GTOval", coord, kpts=df.kpts))
    nk, nip, nao = df._x.shape
    # recover the mask (the reference returns AO values, not indices: fftisdf.py:388)
    mask = []
    for i in range(nip):
        d = abs(x0[:, :, :] - df._x[:, i:i + 1, :]).max(axis=(0, 2))
        mask.append(int(np.argmin(d)))
        assert d[mask[-1]] == 0.0
    rng = np.random.default_rng(spec["seed"] + 100)
    dm = rng.standard_normal((nk, nao, nao)) + 1j * rng.standard_normal((nk, nao, nao))
    dm = dm + dm.conj().transpose(0, 2, 1)
    tr = _time_reversal(kmesh)
    dm = 0.5 * (dm + dm[tr].conj())  # D(-k) = D(k)^*  (real-space density matrix real)
    vj = ref.get_j_kpts(df, dm, 1, df.kpts, None)
    vk = ref.get_k_kpts(df, dm, 1, df.kpts, None)
    out = dict(a=cell.a, kpts=df.kpts, kmesh=np.asarray(kmesh), mesh=np.asarray(spec["mesh"]),
               m0=np.asarray(spec["m0"]), c0=spec["c0"], blksize=spec["blksize"], x0=x0, f_all=f_all,
               coord=coord, x=df._x, wq=df._wq, w0=df._w0, mask=np.asarray(mask), dm=dm, vj=vj, vk=vk, ranks=ranks)
    path = os.path.join(ROOT, "tests", "golden", f"ref_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: nk={nk} nip={nip} nao={nao} ng={len(coord)} ranks={ranks.tolist()} -> {path} ({os.path.getsize(path)/1e6:.2f} MB)")


def _time_reversal(kmesh):
    n1, n2, n3 = kmesh
    idx = np.arange(n1 * n2 * n3).reshape(n1, n2, n3)
    j = np.meshgrid(np.arange(n1), np.arange(n2), np.arange(n3), indexing="ij")
    return idx[(-j[0]) % n1, (-j[1]) % n2, (-j[2]) % n3].ravel()


def main():
    ref = load_reference()
    only = sys.argv[1:]          # optional case names; default: regenerate every fixture
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(ref, name, spec)


if __name__ == "__main__":
    main()
