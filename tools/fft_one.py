"""One 3-D FFT batch for ncu:  python tools/fft_one.py 64 256 [mode]   (mode: reg | reg-fused | stockham | dmma)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_isdf_scratch_b200.kernels as K
n, nvec = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "reg"
ops = K.IsdfOps(0)
mesh = [n] * 3
ng = n ** 3
x = torch.randn(nvec, ng, dtype=torch.complex128, device="cuda")
pre = torch.randn(ng, dtype=torch.complex128, device="cuda")
post = torch.rand(ng, dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.fft3d(x, mesh, pre=pre, post=post, mode=mode)
torch.cuda.synchronize()
