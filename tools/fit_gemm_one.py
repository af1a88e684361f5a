"""The dominant launch of the NiO-AFM 4x4x4 build in isolation (for ncu): Theta~ = G Y^T for one grid block,
G [36, 2752, 3120], Y^T [36, 3120, 8000] complex128.   python tools/fit_gemm_one.py [nq rP n blk]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_isdf_scratch_b200.kernels as K
nq, rP, n, blk = [int(x) for x in sys.argv[1:5]] if len(sys.argv) > 4 else (36, 2752, 3120, 8000)
ops = K.IsdfOps(0)
g = torch.randn(nq, rP, n, dtype=torch.complex128, device="cuda")
y = torch.randn(nq, n, blk, dtype=torch.complex128, device="cuda")
out = torch.empty(nq, rP, blk, dtype=torch.complex128, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm_nn_strided(g, y, out); e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"{t*1e3:.1f} ms  executed {6.0*nq*rP*n*blk/t/1e12:.2f} TFLOP/s (3M: 6 flop per complex MAC), zgemm-equivalent {8.0*nq*rP*n*blk/t/1e12:.2f}")
