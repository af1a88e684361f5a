"""torchrun worker: multi-GPU ISDF build == single-GPU build (run by tests/test_multigpu.py and by hand:
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_build_check.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fft_isdf_scratch_b200 as pk  # noqa: E402
from fft_isdf_scratch_b200 import fftisdf  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    worst = 0.0
    for name in ["k231_odd", "k321_spd", "gamma_s", "k222_sp", "k221_rd"]:
        g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"ref_{name}.npz"))
        cell = pk.TableCell(g["a"], g["mesh"], g["x0"].shape[-1])

        def run(comm, exchange="p2p"):
            df = fftisdf.ISDF(cell, g["kpts"], m0=g["m0"].tolist(), c0=float(g["c0"]), device=local)
            df.blksize = 97
            df.set_ao_tables(x0=g["x0"], f_all=g["f_all"])
            df.comm = comm
            df.exchange = exchange
            df.keep_theta = True
            df.build()
            return df

        d1 = run(None)
        dn = run(dist.group.WORLD)                      # exchange fused into the DFT kernels over NVLink peer memory
        dc = run(dist.group.WORLD, exchange="nccl")     # two NCCL all-to-alls (the route of meshes with an axis > 48)
        assert float(np.abs(dc._wq - d1._wq).max() / np.abs(d1._wq).max()) < 1e-11
        assert np.array_equal(d1._mask, dn._mask)
        assert np.array_equal(d1._ranks, dn._ranks)
        err = float(np.abs(dn._wq - d1._wq).max() / np.abs(d1._wq).max())
        # theta shard == the matching columns of the single-GPU theta
        from fft_isdf_scratch_b200 import sharding
        lo, hi, c = sharding.col_shard(len(g["coord"]), world, rank)
        t1 = d1._theta_dev.cpu().numpy()[:, :, lo:hi]
        tn = dn._theta_dev.cpu().numpy()
        terr = float(np.abs(tn - t1).max() / max(np.abs(t1).max(), 1e-300)) if hi > lo else 0.0
        worst = max(worst, err, terr)
        print(f"rank {rank}/{world} {name}: |W_N - W_1|/|W_1| = {err:.2e}  theta shard {terr:.2e}", flush=True)
        # vs the reference golden where it is well conditioned
        if name in ("k231_odd", "k321_spd"):
            e2 = float(np.abs(dn._wq - g["wq"]).max() / np.abs(g["wq"]).max())
            assert e2 < 1e-10, e2
    assert worst < 1e-11, worst
    # kernel level: the NVLink gather/scatter variants of both transform families == the local transform of the
    # gathered vectors (register-resident FFT: two-factor 33^3, prime 37 x two-factor 15^2; tensor-core DFT 9^3)
    from fft_isdf_scratch_b200 import kernels, sharding
    ops = kernels.IsdfOps(local)
    for mesh, mode in [([33, 33, 33], "reg"), ([37, 15, 15], "reg"), ([9, 9, 9], "dmma"), ([33, 33, 33], "dmma")]:
        ng = int(np.prod(mesh))
        lo, hi, ncol = sharding.col_shard(ng, world, rank)
        rows = 6 * world
        gen = torch.Generator(device="cpu").manual_seed(7)
        full = torch.randn(rows, ncol * world, dtype=torch.complex128, generator=gen)
        full[:, ng:] = 0
        pre = torch.randn(ng, dtype=torch.complex128, generator=gen).cuda()
        post = torch.rand(ng, dtype=torch.float64, generator=gen).cuda()
        buf = sharding.PeerBuffer((rows, ncol), torch.device("cuda", local), dist.group.WORLD)
        buf.tensor.copy_(full[:, rank * ncol:(rank + 1) * ncol])
        v_lo, v_cnt = sharding.vector_shard(rows, world, rank)
        work = torch.empty((v_cnt, ng), dtype=torch.complex128, device="cuda")
        buf.barrier()
        ops.dft3d_p2p(buf.ptrs, ncol, v_lo, work, v_cnt, mesh, pre=pre, post=post, mode=mode)
        buf.barrier()
        ref = full[:, :ng].cuda().clone()
        ops.fft3d(ref, mesh, pre=pre, post=post)
        got = buf.tensor[:, : min(hi, ng) - lo] if hi > lo else buf.tensor[:, :0]
        e3 = float((got - ref[:, lo:min(hi, ng)]).abs().max() / ref.abs().max()) if hi > lo else 0.0
        print(f"rank {rank}/{world} p2p {mode} {mesh}: {e3:.2e}", flush=True)
        assert e3 < 1e-13, e3
        del buf
    dist.barrier()
    if rank == 0:
        print("DIST_OK", worst, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
