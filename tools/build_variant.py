"""Build a tuning variant of libisdf_b200.so with extra -D flags (A/B measurements on the GPU box):
    python tools/build_variant.py 4m -DISDF_GEMM_3M=0 -DISDF_DFT_3M=0
writes fft-isdf-scratch_b200/lib/variants/libisdf_b200_4m.so; select it with ISDF_B200_LIB=<path>."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as G  # noqa: E402


def main():
    name, defs = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(G.PKG, "lib", "variants")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libisdf_b200_{name}.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + G.NVCC_FLAGS + defs + [os.path.join(G.CSRC, s) for s in G.SOURCES] + ["-o", out]
    subprocess.run(cmd, check=True, cwd=G.PKG)
    print(out)


if __name__ == "__main__":
    main()
