"""Register-resident FFT on the CPU: tools/fft_reg_host_check.cu compiles the SAME phase functions the kernels run
(fft_reg_core.cuh: in-register butterflies, the in-place plane slot maps, the x-pass exchange, the direct symmetric DFT
for primes) as host code and executes them thread by thread against a naive DFT, for every instantiated length.  Also
checks that the committed size tables are what tools/gen_fft_sizes.py generates."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fft-isdf-scratch_b200", "csrc")


def _nvcc():
    return os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(_nvcc()), reason="nvcc not found")
def test_fft_phases_against_naive_dft_on_the_host(tmp_path):
    exe = str(tmp_path / "fft_reg_host_check")
    p = subprocess.run([_nvcc(), "-Wno-deprecated-gpu-targets", "--expt-relaxed-constexpr", "-std=c++17", "-O1", "-I", CSRC,
                        os.path.join(ROOT, "tools", "fft_reg_host_check.cu"), "-o", exe], capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    lines = r.stdout.strip().splitlines()
    worst = float(lines[-1].split()[-1])
    assert worst < 1e-14
    lengths = {int(m.group(1)) for m in (re.match(r"(?:plane|direct) N=(\d+)", ln) for ln in lines) if m}
    assert {32, 33, 37, 48, 64, 96}.issubset(lengths) and len(lengths) >= 50


def test_size_tables_match_their_generator(tmp_path):
    before = {}
    for k in range(6):
        for name in ("fft_reg_sizes_p%d.inc" % k, "fft_reg_part%d.cu" % k):
            before[name] = open(os.path.join(CSRC, name)).read()
    before["fft_reg_parts.inc"] = open(os.path.join(CSRC, "fft_reg_parts.inc")).read()
    # the generator writes in place: run it on a copy of the tree layout
    work = tmp_path / "repo"
    (work / "tools").mkdir(parents=True)
    (work / "fft-isdf-scratch_b200" / "csrc").mkdir(parents=True)
    shutil.copy(os.path.join(ROOT, "tools", "gen_fft_sizes.py"), work / "tools" / "gen_fft_sizes.py")
    subprocess.run([sys.executable, str(work / "tools" / "gen_fft_sizes.py")], check=True, capture_output=True)
    for name, text in before.items():
        assert open(work / "fft-isdf-scratch_b200" / "csrc" / name).read() == text, name
