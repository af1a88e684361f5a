import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_isdf_scratch_b200.kernels as K
ops = K.IsdfOps(0)
for n, nvec in [(64, 256), (33, 2176), (96, 96), (48, 600)]:
    mesh = [n] * 3; ng = n ** 3
    x = torch.randn(nvec, ng, dtype=torch.complex128, device="cuda")
    pre = torch.randn(ng, dtype=torch.complex128, device="cuda")
    post = torch.rand(ng, dtype=torch.float64, device="cuda")
    for mb in [24, 48, 96, 192, 100000]:
        gv = max(1, int(mb * 1024 * 1024 / (ng * 16)))
        for _ in range(2):
            ops.fft3d(x, mesh, pre=pre, post=post, mode="stockham", group_vecs=gv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.fft3d(x, mesh, pre=pre, post=post, mode="stockham", group_vecs=gv)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 3 * 1e-3
        print(f"mesh {n}^3 nvec {nvec} group {mb:6d} MB ({gv} vecs): {t*1e3:8.3f} ms -> {2*16*nvec*ng/t/1e9:8.1f} GB/s of min", flush=True)
