"""world_size-2 gloo tests (CPU) of the multi-GPU layouts: grid-column <-> vector all-to-alls,
slot all-gather and the W all-reduce used by build() when df.comm is set."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from fft_isdf_scratch_b200 import sharding as S
    dist.init_process_group("gloo")
    w, r = dist.get_world_size(), dist.get_rank()
    nq, nipP, ng = 3, 8, 11                       # ng not divisible by world -> padded last shard
    lo, hi, c = S.col_shard(ng, w, r)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(nq, nipP, ng, 2, generator=g, dtype=torch.float64)
    full = torch.view_as_complex(full)
    cols = torch.zeros(nq, nipP, c, dtype=torch.complex128)
    cols[:, :, : hi - lo] = full[:, :, lo:hi]
    v = S.to_vector_layout(cols, dist.group.WORLD)
    nv = nipP // w
    assert v.shape == (nq, nv, w * c)
    assert torch.equal(v[:, :, :ng], full[:, r * nv:(r + 1) * nv, :]), "vector layout wrong"
    assert float(v[:, :, ng:].abs().max() if w * c > ng else 0.0) == 0.0
    back = S.to_column_layout(v, dist.group.WORLD)
    assert torch.equal(back, cols), "round trip failed"
    # slot all-gather (round robin)
    nslot = 5
    mine = S.slot_shard(nslot, w, r)
    local = torch.stack([torch.full((4,), float(s), dtype=torch.float64) for s in mine]) if mine else torch.zeros(0, 4, dtype=torch.float64)
    allv = S.allgather_slots(local, nslot, dist.group.WORLD)
    assert torch.equal(allv[:, 0], torch.arange(nslot, dtype=torch.float64))
    lc = torch.stack([torch.full((2, 2), complex(s, -s), dtype=torch.complex128) for s in mine]) if mine else torch.zeros(0, 2, 2, dtype=torch.complex128)
    allc = S.allgather_slots(lc, nslot, dist.group.WORLD)
    assert torch.equal(allc[:, 0, 0].real, torch.arange(nslot, dtype=torch.float64))
    agc = S.AsyncSlotGather(lc, nslot, dist.group.WORLD).result()
    assert torch.equal(agc, allc)
    # partial W all-reduce == full contraction
    part = cols[:, :, : hi - lo] @ cols[:, :, : hi - lo].conj().transpose(1, 2)
    S.allreduce_sum_(part, dist.group.WORLD)
    ref = full @ full.conj().transpose(1, 2)
    assert (part - ref).abs().max() < 1e-12
    t = torch.full((3,), float(r + 1), dtype=torch.complex128)
    S.broadcast_(t, 0, dist.group.WORLD)
    assert torch.equal(t.real, torch.ones(3, dtype=torch.float64))
    dist.destroy_process_group()
    open(sys.argv[1] + f"/ok{r}", "w").write("ok")
''')


def test_layouts_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), str(tmp_path)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_vector_shard_covers_live_rows_once():
    from fft_isdf_scratch_b200 import sharding as S
    for nlive in (0, 1, 7, 52, 413, 2140):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, cnt = S.vector_shard(nlive, w, r)
                assert cnt >= 0 and (cnt == 0 or lo + cnt <= nlive)
                seen += list(range(lo, lo + cnt))
            assert seen == list(range(nlive))


def test_col_shard_covers_grid():
    from fft_isdf_scratch_b200 import sharding as S
    for ng in (1, 7, 50653, 32768):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                lo, hi, c = S.col_shard(ng, w, r)
                assert hi - lo <= c
                seen += list(range(lo, hi))
            assert seen == list(range(ng))


def test_slab_destinations_cover_every_owner_slot_once():
    """Fused reduce-scatter of W~: the (owner, offset) every rank writes q-slot s to must be the owner's local index of
    that slot (slot_shard) and slab = writer rank -- all world x nslot slabs distinct and inside the owner's buffer."""
    from fft_isdf_scratch_b200 import sharding
    for nslot, world in [(36, 8), (14, 2), (3, 4), (1, 2), (36, 1)]:
        slab = 7
        n_own = -(-nslot // world)
        seen = set()
        for rank in range(world):
            owner, off = sharding.slab_destinations(nslot, world, rank, slab)
            for s in range(nslot):
                mine = sharding.slot_shard(nslot, world, owner[s])
                assert s in mine
                assert off[s] == (mine.index(s) * world + rank) * slab
                assert 0 <= off[s] and off[s] + slab <= n_own * world * slab
                seen.add((owner[s], off[s]))
            ptrs = sharding.slab_destinations(nslot, world, rank, slab, base_ptrs=[1000 * (r + 1) for r in range(world)])
            assert ptrs == [1000 * (o + 1) + 16 * f for o, f in zip(owner, off)]
        assert len(seen) == nslot * world
