"""PySCF-free periodic cells for the ISDF build.

`SyntheticCell` duck-types the handful of `pyscf.pbc.gto.Cell` members the reference
hot path touches (/root/reference/fftisdf.py:24-31,91,115,322,367-370):
`lattice_vectors()`, `vol`, `nao_nr()`, `get_kpts()/make_kpts()`, `gen_uniform_grids()`,
`get_Gv()`, `pbc_eval_gto("GTOval", coords, kpts=...)`, `mesh`.

The atomic orbitals are real cartesian Gaussians (s, p, five d combinations) summed over
lattice images with Bloch phases, phi_k(r) = sum_T e^{ik.T} chi(r - T), which is PySCF's
`pbc_eval_gto` convention.  AO evaluation is the *input producer* of the ISDF build
(SURVEY.md section 8, row a5/f-3), not one of the accelerated kernels, so it stays on the host here.

Named stand-ins (`diamond_standin`, `nio_afm_standin`) reproduce the lattice, atom
positions, AO count and symmetry of the reference's cells
(/root/reference/fftdf-with-k-lstsq.py:192-202, /root/reference/nio-afm.vasp) with
single-primitive Gaussians in place of the GTH contractions, which need PySCF's basis
files.  They are labelled "stand-in" wherever results are reported.
"""
from __future__ import annotations

import numpy as np

from . import pbc_tools

BOHR = 0.52917721092

# real angular factors as lists of (coef, (px, py, pz))
_ANG = {
    "s": [[(1.0, (0, 0, 0))]],
    "p": [[(1.0, (1, 0, 0))], [(1.0, (0, 1, 0))], [(1.0, (0, 0, 1))]],
    "d": [
        [(1.0, (1, 1, 0))],
        [(1.0, (0, 1, 1))],
        [(1.0, (1, 0, 1))],
        [(0.5, (2, 0, 0)), (-0.5, (0, 2, 0))],
        [(1.0, (0, 0, 2)), (-0.5, (2, 0, 0)), (-0.5, (0, 2, 0))],
    ],
}


class SyntheticCell:
    """Periodic cell with analytic Gaussian AOs.

    Parameters
    ----------
    a : (3,3) lattice vectors in bohr (rows).
    shells : list of (center[3] bohr, l in "spd", alpha) tuples.
    mesh : FFT mesh [n1,n2,n3] (dense grid).
    """

    def __init__(self, a, shells, mesh, name="synthetic"):
        self.a = np.asarray(a, dtype=np.float64).reshape(3, 3)
        self.mesh = [int(m) for m in mesh]
        self.name = name
        self.verbose = 0
        self.max_memory = 160000
        self.dimension = 3
        self.low_dim_ft_type = None
        cen, alp, terms = [], [], []
        for c, l, al in shells:
            for t in _ANG[l]:
                cen.append(np.asarray(c, dtype=np.float64))
                alp.append(float(al))
                terms.append(t)
        self._cen = np.asarray(cen).reshape(-1, 3)
        self._alp = np.asarray(alp)
        self._terms = terms
        # normalise each AO roughly to unit self-overlap of the isolated Gaussian so
        # the Gram matrices are well scaled: N = (2 alpha/pi)^(3/4) * (4 alpha)^(l/2)
        lsum = np.array([sum(t[0][1]) for t in terms])
        self._norm = (2 * self._alp / np.pi) ** 0.75 * (4 * self._alp) ** (lsum / 2.0)
        self._images = self._make_images()

    # ---- Cell-like surface -------------------------------------------------
    def lattice_vectors(self):
        return self.a

    def reciprocal_vectors(self):
        return pbc_tools.reciprocal_vectors(self.a)

    @property
    def vol(self):
        return float(abs(np.linalg.det(self.a)))

    def nao_nr(self):
        return len(self._alp)

    def make_kpts(self, kmesh):
        return pbc_tools.make_kpts(self.a, kmesh)

    get_kpts = make_kpts

    def get_scaled_kpts(self, kpts):
        return pbc_tools.get_scaled_kpts(self.a, kpts)

    def gen_uniform_grids(self, mesh=None, wrap_around=False):
        return pbc_tools.gen_uniform_grids(self.a, self.mesh if mesh is None else mesh, wrap_around)

    get_uniform_grids = gen_uniform_grids

    def get_Gv(self, mesh=None):
        return pbc_tools.get_Gv(self.a, self.mesh if mesh is None else mesh)

    # ---- AO evaluation -----------------------------------------------------
    def _make_images(self):
        amin = float(self._alp.min())
        rcut = np.sqrt(36.0 / amin) + 1e-9
        b = pbc_tools.reciprocal_vectors(self.a)
        heights = 2 * np.pi / np.linalg.norm(b, axis=1)
        nimg = np.ceil(rcut / heights).astype(int) + 1
        rel = pbc_tools.cartesian_prod([np.arange(-n, n + 1) for n in nimg])
        return rel @ self.a

    def _eval_images(self, coords):
        """chi[g, mu, T] = chi_mu(r_g - T) (real)."""
        d = coords[:, None, None, :] - self._cen[None, :, None, :] - self._images[None, None, :, :]
        r2 = np.einsum("gmtx,gmtx->gmt", d, d)
        rad = np.exp(-self._alp[None, :, None] * r2)
        ang = np.zeros_like(rad)
        for mu, terms in enumerate(self._terms):
            acc = 0.0
            for coef, (px, py, pz) in terms:
                t = coef
                if px:
                    t = t * d[:, mu, :, 0] ** px
                if py:
                    t = t * d[:, mu, :, 1] ** py
                if pz:
                    t = t * d[:, mu, :, 2] ** pz
                acc = acc + t
            ang[:, mu, :] = acc
        return rad * ang * self._norm[None, :, None]

    def pbc_eval_gto(self, eval_name, coords, kpts=None, kpt=None, chunk=None):
        """List over k of [npts, nao] complex128 arrays (PySCF `GTOval` convention)."""
        assert eval_name in ("GTOval", "GTOval_sph", "GTOval_cart")
        coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
        single = kpts is None
        if single:
            kpts = np.zeros((1, 3)) if kpt is None else np.asarray(kpt).reshape(1, 3)
        kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
        nk, npts, nao, nt = len(kpts), len(coords), self.nao_nr(), len(self._images)
        eikt = np.exp(1j * (self._images @ kpts.T))  # [T, nk]
        out = np.empty((nk, npts, nao), dtype=np.complex128)
        if chunk is None:
            chunk = max(1, int(2.5e7 // (nao * nt * 4)))
        for p0 in range(0, npts, chunk):
            p1 = min(npts, p0 + chunk)
            chi = self._eval_images(coords[p0:p1])  # [g, mu, T]
            out[:, p0:p1, :] = np.einsum("gmt,tk->kgm", chi, eikt, optimize=True)
        if single:
            return out[0]
        return [out[k] for k in range(nk)]

    def eval_ao_kpts(self, coords, kpts):
        return np.asarray(self.pbc_eval_gto("GTOval", coords, kpts=kpts))

    # ---- device evaluation (same functions, libisdf_b200's ao_eval kernel) ---------------------
    def _ao_desc_bytes(self):
        """Pack the AO list in the record layout of csrc/ao_eval.cu::AoDesc."""
        dt = np.dtype([("c", "<f8", 3), ("alpha", "<f8"), ("norm", "<f8"), ("coef", "<f8", 3), ("pw", "<i4", (3, 3)),
                       ("nterm", "<i4")], align=True)
        rec = np.zeros(self.nao_nr(), dtype=dt)
        for mu, terms in enumerate(self._terms):
            rec["c"][mu] = self._cen[mu]
            rec["alpha"][mu] = self._alp[mu]
            rec["norm"][mu] = self._norm[mu]
            rec["nterm"][mu] = len(terms)
            for i, (coef, pw) in enumerate(terms):
                rec["coef"][mu][i] = coef
                rec["pw"][mu][i] = pw
        return rec

    def eval_ao_device(self, ops, coords, kpts, out=None):
        """[nk, npts, nao] complex128 CUDA tensor; `coords` numpy [npts,3] or CUDA tensor."""
        import torch
        key = id(ops)
        cache = self.__dict__.setdefault("_dev_cache", {})
        if key not in cache:
            rec = self._ao_desc_bytes()
            assert rec.dtype.itemsize == ops.lib.isdf_ao_desc_bytes(), (rec.dtype.itemsize, ops.lib.isdf_ao_desc_bytes())
            cache[key] = (torch.from_numpy(rec.view(np.uint8).reshape(-1)).to(ops.device),
                          torch.from_numpy(np.ascontiguousarray(self._images)).to(ops.device))
        desc, images = cache[key]
        kpts = np.asarray(kpts, dtype=np.float64).reshape(-1, 3)
        kphase = torch.from_numpy(np.ascontiguousarray(np.exp(1j * (kpts @ self._images.T)))).to(ops.device)
        if not torch.is_tensor(coords):
            coords = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.float64)).to(ops.device)
        return ops.eval_ao(coords.contiguous(), desc, self.nao_nr(), images, kphase, out=out)


def random_cubic_cell(ng_side, nao, seed, L=None, alpha_range=(0.3, 3.0), ltypes="s"):
    """SURVEY.md section 8(d) synthetic family: cubic cell, side L = 10 bohr * (ng/32^3)^(1/3), `nao`
    Gaussians at generic positions (tie-free pivots), log-uniform exponents."""
    rng = np.random.default_rng(seed)
    if L is None:
        L = 10.0 * (ng_side / 32.0)
    a = np.eye(3) * L
    shells = []
    n = 0
    while n < nao:
        l = ltypes[rng.integers(len(ltypes))]
        k = {"s": 1, "p": 3, "d": 5}[l]
        if n + k > nao:
            l, k = "s", 1
        c = rng.uniform(0, L, size=3)
        al = float(np.exp(rng.uniform(np.log(alpha_range[0]), np.log(alpha_range[1]))))
        shells.append((c, l, al))
        n += k
    return SyntheticCell(a, shells, [ng_side] * 3, name=f"cubic{ng_side}-nao{nao}-seed{seed}")


def diamond_standin(mesh=None, ke_cutoff=100.0):
    """Stand-in for the reference's diamond C2 cell (fftdf-with-k-lstsq.py:192-202):
    same lattice `ones*3.5668 - eye*3.5668` Angstrom, same two atoms, 13 AOs per atom
    (2s 2p 1d like gth-dzvp) -> nao = 26."""
    a = (np.ones((3, 3)) * 3.5668 - np.eye(3) * 3.5668) / BOHR
    atoms = [np.zeros(3), np.ones(3) * 0.8917 / BOHR]
    shells = []
    for c in atoms:
        shells += [(c, "s", 0.32), (c, "s", 1.10), (c, "p", 0.28), (c, "p", 0.95), (c, "d", 0.60)]
    if mesh is None:
        mesh = pbc_tools.cutoff_to_mesh(a, ke_cutoff)
    return SyntheticCell(a, shells, mesh, name="diamond-standin")


def nio_afm_standin(mesh=None, ke_cutoff=200.0):
    """Stand-in for the NiO AFM cell of /root/reference/nio-afm.vasp (2 Ni + 2 O,
    rhombohedral): Ni 4s 4p 2d = 26 AOs, O 2s 2p 1d = 13 AOs -> nao = 78."""
    a = np.array([[4.17, 2.085, 2.085], [2.085, 4.17, 2.085], [2.085, 2.085, 4.17]]) / BOHR
    frac = {"Ni": [[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]], "O": [[0.25, 0.25, 0.25], [0.75, 0.75, 0.75]]}
    shells = []
    for f in frac["Ni"]:
        c = np.asarray(f) @ a
        shells += [(c, "s", 0.35), (c, "s", 1.0), (c, "s", 2.6), (c, "s", 0.18), (c, "p", 0.4), (c, "p", 1.3),
                   (c, "p", 0.20), (c, "d", 0.7), (c, "d", 2.0), (c, "p", 0.8)]
    for f in frac["O"]:
        c = np.asarray(f) @ a
        shells += [(c, "s", 0.35), (c, "s", 1.2), (c, "p", 0.3), (c, "p", 1.1), (c, "d", 0.8)]
    if mesh is None:
        mesh = pbc_tools.cutoff_to_mesh(a, ke_cutoff)
    cell = SyntheticCell(a, shells, mesh, name="nio-afm-standin")
    return cell


class TableCell:
    """Geometry-only cell whose AO values are supplied as precomputed tables (e.g. dumped from PySCF's
    `pbc_eval_gto`, or the golden fixtures): use with `ISDF.set_ao_tables(x0=..., f_all=...)`."""

    def __init__(self, a, mesh, nao, name="table-cell"):
        self.a = np.asarray(a, dtype=np.float64).reshape(3, 3)
        self.mesh = [int(m) for m in mesh]
        self._nao = int(nao)
        self.name = name
        self.verbose = 0
        self.max_memory = 160000
        self.dimension = 3
        self.low_dim_ft_type = None

    def lattice_vectors(self):
        return self.a

    @property
    def vol(self):
        return float(abs(np.linalg.det(self.a)))

    def nao_nr(self):
        return self._nao

    def make_kpts(self, kmesh):
        return pbc_tools.make_kpts(self.a, kmesh)

    get_kpts = make_kpts

    def gen_uniform_grids(self, mesh=None, wrap_around=False):
        return pbc_tools.gen_uniform_grids(self.a, self.mesh if mesh is None else mesh, wrap_around)

    get_uniform_grids = gen_uniform_grids

    def get_Gv(self, mesh=None):
        return pbc_tools.get_Gv(self.a, self.mesh if mesh is None else mesh)

    def pbc_eval_gto(self, *args, **kwargs):
        raise RuntimeError("TableCell has no basis: supply AO tables with ISDF.set_ao_tables()")
