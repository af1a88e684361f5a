"""FP64 roofline calibration on the GPU box (comparators only -- never on the shipped path):
cuBLAS DGEMM/ZGEMM through torch.matmul, an HBM copy, and this library's GEMM engine on the
shapes of the dominant stages.  Writes gpurun_out/calib.json."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), sum(ts) / len(ts)


def main():
    out = {}
    dev = torch.device("cuda", 0)
    if "--engine-only" in sys.argv:      # A/B of library builds (ISDF_B200_LIB=...): skip the cuBLAS / HBM comparators
        return engine(out, dev)
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    tmin, tavg = timeit(lambda: torch.matmul(a, b))
    out["cublas_dgemm_tflops_best"] = 2 * n ** 3 / tmin / 1e12
    out["cublas_dgemm_tflops_avg"] = 2 * n ** 3 / tavg / 1e12
    t0 = time.time()
    k = 0
    torch.cuda.synchronize()
    while time.time() - t0 < 3.0:
        torch.matmul(a, b)
        k += 1
        if k % 4 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    out["cublas_dgemm_tflops_sustained"] = 2 * n ** 3 * k / (time.time() - t0) / 1e12
    del a, b
    n = 4096
    a = torch.randn(n, n, dtype=torch.complex128, device=dev)
    b = torch.randn(n, n, dtype=torch.complex128, device=dev)
    tmin, tavg = timeit(lambda: torch.matmul(a, b))
    out["cublas_zgemm_tflops_best"] = 8 * n ** 3 / tmin / 1e12
    del a, b
    x = torch.empty(1 << 28, dtype=torch.float64, device=dev)
    y = torch.empty_like(x)
    tmin, _ = timeit(lambda: y.copy_(x))
    out["hbm_copy_gbs"] = 2 * x.numel() * 8 / tmin / 1e9
    del x, y

    engine(out, dev)


def engine(out, dev):
    import fft_isdf_scratch_b200.kernels as K
    ops = K.IsdfOps(0)
    # HERK shape: nip=2048, ng=32768
    nip, ng = 2048, 32768
    bmat = torch.randn(1, nip, ng, dtype=torch.complex128, device=dev)
    w = torch.empty(1, nip, nip, dtype=torch.complex128, device=dev)
    tmin, tavg = timeit(lambda: ops.herk(bmat, out=w))
    out["herk_2048x32768_zgemm_equiv_tflops"] = 8 * nip * nip * ng / tmin / 1e12
    # plain NN gemm (the sweep kernel): M=64 rows, K=2048, N=32768, batch 8
    am = torch.randn(8, 64, 2048, dtype=torch.complex128, device=dev)
    bm = torch.randn(8, 2048, ng, dtype=torch.complex128, device=dev)
    tmin, tavg = timeit(lambda: ops.gemm_nn(am, bm))
    out["gemm_nn_64x32768x2048_b8_tflops"] = 8 * 8 * 64 * ng * 2048 / tmin / 1e12
    # conj(A) B^T gram, M=N=4096, K=4096
    am = torch.randn(1, 4096, 4096, dtype=torch.complex128, device=dev)
    tmin, tavg = timeit(lambda: ops.gram_conja(am, am))
    out["gram_conja_4096_tflops"] = 8 * 4096 ** 3 / tmin / 1e12
    print(json.dumps(out, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/calib.json", "w"), indent=1)


if __name__ == "__main__":
    main()
