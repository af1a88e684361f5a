"""The branches that only run when PySCF imports (`_HAVE_PYSCF`: the class subclasses pyscf.pbc.df.fft.FFTDF and
aoR_loop pulls AOs through NumInt.block_loop, fftisdf.py:327-355).  PySCF is not in the image, so the test installs the
SAME thin stubs the golden generator uses to run the reference's own code (oracle/gen_golden.py: an FFTDF with `grids`
and `_numint.block_loop`), imports the package in a fresh interpreter and checks the build against the reference's
golden outputs and against the table-fed build."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
import numpy as np
sys.path.insert(0, ROOT)
from oracle import gen_golden as GG
GG._install_stubs()                                   # pyscf.pbc.df.fft.FFTDF etc. (stubs), BEFORE the package import
import fft_isdf_scratch_b200 as pk
from fft_isdf_scratch_b200 import fftisdf
assert fftisdf._HAVE_PYSCF
assert issubclass(fftisdf.ISDF, sys.modules["pyscf.pbc.df.fft"].FFTDF)

class PlainCell:                                      # a cell without the SyntheticCell device hooks
    def __init__(self, c):
        self._c = c
        self.mesh = list(c.mesh)
        self.a = c.a
        self.vol = c.vol
        self.dimension = 3
    def lattice_vectors(self): return self._c.lattice_vectors()
    def nao_nr(self): return self._c.nao_nr()
    def get_kpts(self, kmesh): return self._c.get_kpts(kmesh)
    def gen_uniform_grids(self, mesh): return self._c.gen_uniform_grids(mesh)
    def get_Gv(self, mesh): return self._c.get_Gv(mesh)
    def pbc_eval_gto(self, name, coords, kpts=None): return self._c.pbc_eval_gto(name, coords, kpts=kpts)

for name in ["k231_odd", "k222_sp"]:
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_%s.npz" % name))
    spec = GG.CASES[name]
    cell = PlainCell(GG.make_cell(spec))
    df = fftisdf.ISDF(cell, g["kpts"])
    df.m0, df.c0, df.blksize = g["m0"].tolist(), float(g["c0"]), int(spec["blksize"])
    calls = []
    orig = df._numint.block_loop
    def counted(*a, **k):
        for blk in orig(*a, **k):
            calls.append(blk[4].shape[0])
            yield blk
    df._numint.block_loop = counted
    df.build()
    assert calls and sum(calls) >= len(g["coord"]), "NumInt.block_loop was not used"
    assert np.array_equal(df._mask, g["mask"]) and np.array_equal(df._x, g["x"])
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    if name == "k231_odd":                            # full-rank A_q: the 1e-10 bar
        assert rel(df._wq, g["wq"]) < 1e-10, rel(df._wq, g["wq"])
        vj, vk = df.get_jk(g["dm"], kpts=g["kpts"])
        assert rel(vj, g["vj"].reshape(vj.shape)) < 1e-10 and rel(vk, g["vk"].reshape(vk.shape)) < 1e-10
    # same build fed from AO tables (the route every other test takes): identical device results
    t = pk.TableCell(g["a"], g["mesh"], g["x0"].shape[-1])
    d2 = fftisdf.ISDF(t, g["kpts"], m0=g["m0"].tolist(), c0=float(g["c0"]))
    d2.blksize = int(spec["blksize"])
    d2.set_ao_tables(x0=g["x0"], f_all=g["f_all"])
    d2.build()
    assert rel(df._wq, d2._wq) < 1e-12, rel(df._wq, d2._wq)
    print(name, "ok", len(calls), "blocks")
print("PYSCF_BRANCH_OK")
'''


def test_pyscf_class_branch_with_stubbed_pyscf():
    p = subprocess.run([sys.executable, "-c", "ROOT = %r\n" % ROOT + SCRIPT], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "PYSCF_BRANCH_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
