"""Multi-GPU partitioning of the hot path.

Every (polynomial, residue) unit is independent in the forward / inverse transform, the pointwise functors and the
fused product, so the path shards with NO data-path collective: one process per GPU, each owning a slab of units
and a context (`nflgpu_ctx_create(..., first_modulus, ...)`) over just the residues it holds.  A collective is only
needed when a caller wants the complete RNS vector of every polynomial on one device (e.g. before a CRT lift,
include/nfl/gmp.hpp:183-209): `gather_residues` does that with one all_gather over NCCL (NVLink / NVSwitch).

`gather_residues_peer` does it without a collective library: every rank exports its slab (CUDA IPC through the C ABI), maps
its partners' and pulls them over NVLink with one strided copy per slab, straight into [npolys][nmoduli][degree]
(nflgpu_gather_residues) -- the form a C++ caller of nflgpu_poly2mpz uses.

Host-side logic only (pure Python + torch.distributed for the handle exchange); the partitioning and the exchange protocol
are tested on CPU with gloo, world_size 2."""
from dataclasses import dataclass


@dataclass(frozen=True)
class Shard:
    """What one rank owns: polys [poly0, poly0 + npolys) x residues [res0, res0 + nres)."""
    poly0: int
    npolys: int
    res0: int
    nres: int


def split_even(total, parts, index):
    """Contiguous split of `total` items into `parts` ranges whose sizes differ by at most one."""
    base, extra = divmod(total, parts)
    start = index * base + min(index, extra)
    return start, base + (1 if index < extra else 0)


def shard_batch(batch, nmoduli, world, rank):
    """Primary strategy: contiguous polynomial ranges; every rank keeps all residues (tiny twiddle tables)."""
    p0, n = split_even(batch, world, rank)
    return Shard(p0, n, 0, nmoduli)


def residue_groups(nmoduli, world):
    """Largest divisor of `world` that also divides `nmoduli` (14 moduli on 8 GPUs -> 2 residue groups x 4 batch groups)."""
    best = 1
    for g in range(1, world + 1):
        if world % g == 0 and nmoduli % g == 0:
            best = g
    return best


def shard_residues(batch, nmoduli, world, rank):
    """BASELINE.json configs[3] ("residues sharded over 8xB200"): residue groups x batch groups so that every rank
    gets the same number of units even when nmoduli does not divide by world."""
    rg = residue_groups(nmoduli, world)
    bg = world // rg
    r, b = rank % rg, rank // rg
    res0, nres = split_even(nmoduli, rg, r)
    p0, n = split_even(batch, bg, b)
    return Shard(p0, n, res0, nres)


def local_view(full, shard):
    """Slice of a host array [batch][nmoduli][degree] owned by `shard` (copy, contiguous)."""
    import numpy as np
    return np.ascontiguousarray(full[shard.poly0:shard.poly0 + shard.npolys, shard.res0:shard.res0 + shard.nres, :])


def gather_residues(local, shard, batch, nmoduli, world, group=None):
    """All ranks end with the full [batch][nmoduli][degree] tensor.  `local` is this rank's [npolys][nres][degree]
    torch tensor (CUDA with the nccl backend, CPU with gloo).  Shards may be uneven: slabs are padded to the
    largest shard for the all_gather and trimmed when placed."""
    import torch
    import torch.distributed as dist
    degree = local.shape[-1]
    shards = [None] * world
    dist.all_gather_object(shards, (shard.poly0, shard.npolys, shard.res0, shard.nres), group=group)
    max_elems = max(s[1] * s[3] for s in shards) * degree
    flat = torch.zeros(max_elems, dtype=local.dtype, device=local.device)
    flat[:local.numel()] = local.reshape(-1)
    slabs = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(slabs, flat, group=group)
    full = torch.empty((batch, nmoduli, degree), dtype=local.dtype, device=local.device)
    for (p0, n, r0, nr), slab in zip(shards, slabs):
        full[p0:p0 + n, r0:r0 + nr, :] = slab[:n * nr * degree].reshape(n, nr, degree)
    return full


def residue_partners(shards, rank):
    """Ranks that hold the OTHER residue ranges of this rank's polynomial range (`shards` = every rank's Shard)."""
    me = shards[rank]
    return [r for r, s in enumerate(shards) if r != rank and s.poly0 == me.poly0 and s.npolys == me.npolys]


def gather_residues_peer(full_ctx, slab_ctx, slab_ptr, shard, world, rank, stream=0, group=None):
    """Every residue of this rank's polynomials on this rank's device, through peer memory.

    slab_ptr: this rank's transformed slab [npolys][nres][degree], an allocation of its own made by slab_ctx.alloc() (its CUDA
    IPC handle must name exactly the slab); full_ctx: a context over all residues on the same device.  Returns
    (dst pointer from full_ctx.alloc, close) -- call close() once every rank is done reading (it unmaps the peers).
    The caller has synchronised the stream that produced slab_ptr; the handle exchange below is the cross-rank barrier."""
    import torch.distributed as dist
    info = [None] * world
    dist.all_gather_object(info, (slab_ctx.ipc_export(slab_ptr), shard.poly0, shard.npolys, shard.res0, shard.nres), group=group)
    shards = [Shard(i[1], i[2], i[3], i[4]) for i in info]
    mapped = [(full_ctx.ipc_open(info[r][0]), shards[r].res0, shards[r].nres) for r in residue_partners(shards, rank)]
    slabs = [(slab_ptr, shard.res0, shard.nres)] + mapped
    covered = sorted((r0, n) for _, r0, n in slabs)
    assert sum(n for _, n in covered) == full_ctx.nmoduli and covered[0][0] == 0, "residue ranges of the partners do not tile the full context"
    dst = full_ctx.alloc(shard.npolys)
    full_ctx.gather_residues(dst, slabs, shard.npolys, stream)

    def close():
        for p, _, _ in mapped:
            full_ctx.ipc_close(p)
    return dst, slabs, close
