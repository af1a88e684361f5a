"""Multi-GPU data layouts of the ISDF build (one process per GPU, torch.distributed).

The path shards along two axes with one exchange between them:

  grid-column layout   T_p [nq][nipP][c]          rank p owns dense-grid points [p*c, (p+1)*c)
      right-hand side + triangular sweeps (columns are independent given the replicated A_q factors)
  vector layout        V_p [nq][nipP/P][P*c]       rank p owns interpolation vectors [p*nipP/P, ...)
      3-D FFT (needs the whole grid of a vector)

`to_vector_layout` / `to_column_layout` are the two all-to-alls; the Coulomb-kernel contraction runs
in the column layout (K = local grid points) followed by one all-reduce of W_q.  Everything here is
device-agnostic torch (NCCL on GPUs, gloo in the CPU tests); complex tensors travel as float64 views.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def col_shard(ng, world, rank):
    """Contiguous shard of the dense grid: returns (lo, hi, c) with c = ceil(ng/world) the padded width."""
    c = -(-ng // world)
    lo = min(ng, rank * c)
    hi = min(ng, lo + c)
    return lo, hi, c


def vector_shard(nlive, world, rank):
    """Contiguous shard of the live interpolation vectors (rows below the maximum rank) for the FFT stage:
    returns (lo, count); the counts of all ranks add up to nlive, trailing ranks may get none."""
    per = -(-nlive // world)
    lo = rank * per
    return lo, max(0, min(per, nlive - lo))


def slot_shard(nslot, world, rank):
    """Round-robin assignment of q-slots to ranks (factorisation of A_q)."""
    return list(range(rank, nslot, world))


def slab_destinations(nslot, world, rank, slab_elems, base_ptrs=None, itemsize=16):
    """Fused reduce-scatter of the partial W~ (isdf_herk_to_peers): q-slot s is owned by rank s % world (slot_shard) and
    is that rank's local slot s // world; rank `rank` writes its partial product into slab `rank` of that local slot.
    Returns (owner[s], offset[s]) with the offset in ELEMENTS inside the owner's [n_own][world][slab_elems] buffer, or
    the absolute byte addresses when the ranks' base pointers are given."""
    owner = [s % world for s in range(nslot)]
    off = [((s // world) * world + rank) * slab_elems for s in range(nslot)]
    if base_ptrs is None:
        return owner, off
    return [int(base_ptrs[o]) + itemsize * f for o, f in zip(owner, off)]


def _a2a(out, inp, group):
    dist.all_to_all_single(torch.view_as_real(out), torch.view_as_real(inp), group=group)


def to_vector_layout(t_cols, group=None):
    """[nq][nipP][c] (grid-column shard) -> [nq][nipP/P][P*c] (vector shard, full padded grid)."""
    world, _ = world_info(group)
    if world == 1:
        return t_cols
    nq, nipP, c = t_cols.shape
    assert nipP % world == 0
    nv = nipP // world
    send = t_cols.reshape(nq, world, nv, c).permute(1, 0, 2, 3).contiguous()   # [dest][nq][nv][c]
    recv = torch.empty_like(send)                                               # [src][nq][nv][c]
    _a2a(recv, send, group)
    return recv.permute(1, 2, 0, 3).reshape(nq, nv, world * c).contiguous()


def to_column_layout(t_vecs, group=None):
    """Inverse of `to_vector_layout`."""
    world, _ = world_info(group)
    if world == 1:
        return t_vecs
    nq, nv, pc = t_vecs.shape
    c = pc // world
    send = t_vecs.reshape(nq, nv, world, c).permute(2, 0, 1, 3).contiguous()   # [dest][nq][nv][c]
    recv = torch.empty_like(send)                                               # [src][nq][nv][c]
    _a2a(recv, send, group)
    return recv.permute(1, 0, 2, 3).reshape(nq, world * nv, c).contiguous()


def allreduce_sum_(t, group=None):
    world, _ = world_info(group)
    if world > 1:
        dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_slots(local, nslot, group=None):
    """Each rank holds the slots `slot_shard(nslot, world, rank)` stacked on axis 0 (zero-padded to the
    common count); returns the full [nslot, ...] tensor on every rank."""
    world, rank = world_info(group)
    if world == 1:
        return local
    per = -(-nslot // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    flat = torch.view_as_real(pad) if pad.is_complex() else pad
    out = torch.empty((world * flat.shape[0],) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat.contiguous(), group=group)
    out = out.reshape((world,) + tuple(flat.shape))
    if local.is_complex():
        out = torch.view_as_complex(out)
    full = torch.empty((nslot,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = slot_shard(nslot, world, r)
        if idx:
            full[idx] = out[r, : len(idx)]
    return full


class AsyncSlotGather:
    """`allgather_slots` split in two so that the (large) gather of the sweep operators overlaps the right-hand
    side stage: start() issues the NCCL all-gather asynchronously, result() waits and reorders the slots."""

    def __init__(self, local, nslot, group=None):
        self.local, self.nslot, self.group = local, nslot, group
        self.world, self.rank = world_info(group)
        self.work = None
        if self.world == 1:
            return
        per = -(-nslot // self.world)
        pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        self.flat = (torch.view_as_real(pad) if pad.is_complex() else pad).contiguous()
        self.out = torch.empty((self.world * self.flat.shape[0],) + tuple(self.flat.shape[1:]), dtype=self.flat.dtype,
                               device=self.flat.device)
        self.work = dist.all_gather_into_tensor(self.out, self.flat, group=group, async_op=True)

    def result(self):
        if self.world == 1:
            return self.local
        self.work.wait()
        out = self.out.reshape((self.world,) + tuple(self.flat.shape))
        if self.local.is_complex():
            out = torch.view_as_complex(out)
        full = torch.empty((self.nslot,) + tuple(self.local.shape[1:]), dtype=self.local.dtype,
                           device=self.local.device)
        for r in range(self.world):
            idx = slot_shard(self.nslot, self.world, r)
            if idx:
                full[idx] = out[r, : len(idx)]
        return full


def broadcast_(t, src=0, group=None):
    world, _ = world_info(group)
    if world > 1:
        dist.broadcast(torch.view_as_real(t) if t.is_complex() else t, src=src, group=group)
    return t


class PeerBuffer:
    """A complex128 tensor in NVLink peer-mapped (symmetric) memory: every rank can load/store every other
    rank's copy from inside a kernel.  Thin wrapper over torch.distributed._symmetric_memory."""

    def __init__(self, shape, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        numel = 1
        for d in shape:
            numel *= int(d)
        self.raw = symm_mem.empty(numel * 2, dtype=torch.float64, device=device)
        self.handle = symm_mem.rendezvous(self.raw, group.group_name if hasattr(group, "group_name") else group)
        self.tensor = torch.view_as_complex(self.raw.view(-1, 2)).view(*shape)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]

    def barrier(self):
        """All ranks' prior work on the current stream is visible to every peer afterwards."""
        self.handle.barrier()
