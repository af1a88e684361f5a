"""Micro-benchmark of isdf_fft3d_batched (kernel-tuning aid): vectors of mesh^3 with phase + weight."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_isdf_scratch_b200.kernels as K
ops = K.IsdfOps(0)
for mesh, nvec in [([37] * 3, 448 * 4), ([33] * 3, 544 * 4), ([32] * 3, 2048), ([64] * 3, 256), ([96] * 3, 96), ([15] * 3, 8192), ([45] * 3, 700), ([48] * 3, 600), ([25] * 3, 4000)]:
    ng = mesh[0] * mesh[1] * mesh[2]
    x = torch.randn(nvec, ng, dtype=torch.complex128, device="cuda")
    pre = torch.randn(ng, dtype=torch.complex128, device="cuda")
    post = torch.rand(ng, dtype=torch.float64, device="cuda")
    for mode in (sys.argv[1:] or ["reg", "stockham", "dmma"]):
        if mode == "dmma" and max(mesh) > 48:
            continue
        if mode == "reg" and not ops.fft3d_reg_supported(mesh):
            continue
        for _ in range(2):
            ops.fft3d(x, mesh, pre=pre, post=post, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.fft3d(x, mesh, pre=pre, post=post, mode=mode)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 3 * 1e-3
        print(f"mesh {mesh[0]}^3 nvec {nvec} {mode:9s}: {t*1e3:8.3f} ms  -> {2*16*nvec*ng/t/1e9:8.1f} GB/s of the minimum traffic", flush=True)
