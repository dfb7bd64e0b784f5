"""Host-side multi-GPU logic (nfllib_b200/sharding.py) on CPU: partition arithmetic, and a world_size-2 gloo run in
which each rank transforms only its shard (the CPU oracle stands in for the kernel — test infrastructure only)
and the gathered result must equal the unsharded transform."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nfllib_b200 import sharding as sh
from oracle_lib import Oracle, golden_params, random_polys


def test_split_even_covers_everything():
    for total in (0, 1, 7, 14, 4096):
        for parts in (1, 2, 3, 8):
            ranges = [sh.split_even(total, parts, i) for i in range(parts)]
            assert sum(n for _, n in ranges) == total
            assert all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(parts - 1))
            assert max(n for _, n in ranges) - min(n for _, n in ranges) <= 1


def test_residue_sharding_of_config_c4_is_balanced():
    # N=4096, uint32, 14 moduli, batch 8192 over 8 GPUs -> 2 residue groups x 4 batch groups, 14336 units each
    shards = [sh.shard_residues(8192, 14, 8, r) for r in range(8)]
    assert sh.residue_groups(14, 8) == 2
    assert all(s.npolys * s.nres == 14336 for s in shards)
    cover = np.zeros((8192, 14), np.int32)
    for s in shards:
        cover[s.poly0:s.poly0 + s.npolys, s.res0:s.res0 + s.nres] += 1
    assert (cover == 1).all()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bits, N, M, batch = 32, 64, 4, 10
    a = random_polys(bits, N, M, batch, 321)
    shard = (sh.shard_batch if mode == "batch" else sh.shard_residues)(batch, M, world, rank)
    g = golden_params(bits)
    sub = {k: (g[k][shard.res0:shard.res0 + shard.nres] if isinstance(g[k], list) else g[k]) for k in g}
    o = Oracle(bits, N, shard.nres, params=sub)  # a context over this rank's residues only (first_modulus = res0)
    local = o.run("fwd", sh.local_view(a, shard))
    full = sh.gather_residues(torch.from_numpy(local.astype(np.int64)), shard, batch, M, world)
    exp = Oracle(bits, N, M).run("fwd", a).astype(np.int64)
    q.put((rank, bool(np.array_equal(full.numpy(), exp))))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["batch", "residue"])
def test_two_rank_gloo_shard_transform_gather(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_residue_partners_of_config_c4():
    shards = [sh.shard_residues(8192, 14, 8, r) for r in range(8)]
    for r in range(8):
        partners = sh.residue_partners(shards, r)
        assert len(partners) == 1 and partners[0] == (r ^ 1)  # the rank with the other 7 residues of the same 2048 polynomials
        assert shards[partners[0]].res0 != shards[r].res0


class _FileCtx:
    """CPU stand-in for nfllib_b200.Context in the peer-gather protocol test: 'device memory' is a file-backed numpy array, an
    'IPC handle' is its path (padded to 64 bytes), gather_residues is the same strided placement the C ABI does."""
    registry = {}

    def __init__(self, tmp, N, nmoduli, dtype):
        self.tmp, self.N, self.nmoduli, self.dtype = tmp, N, nmoduli, dtype

    def alloc(self, batch):
        path = os.path.join(self.tmp, f"buf_{os.getpid()}_{len(_FileCtx.registry)}.bin")
        arr = np.lib.format.open_memmap(path, mode="w+", dtype=self.dtype, shape=(batch, self.nmoduli, self.N))
        _FileCtx.registry[path] = arr
        return path

    def ipc_export(self, ptr):
        return ptr.encode().ljust(256, b"\0")

    def ipc_open(self, handle):
        path = handle.rstrip(b"\0").decode()
        _FileCtx.registry[path + "#peer"] = np.load(path, mmap_mode="r")
        return path + "#peer"

    def ipc_close(self, ptr):
        _FileCtx.registry.pop(ptr)

    def gather_residues(self, dst, slabs, batch, stream=0):
        out = _FileCtx.registry[dst]
        for ptr, r0, n in slabs:
            out[:batch, r0:r0 + n, :] = _FileCtx.registry[ptr][:batch]


def _peer_worker(rank, world, port, tmp, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bits, N, M, batch = 32, 64, 14, 12
    a = random_polys(bits, N, M, batch, 555)
    shard = sh.shard_residues(batch, M, world, rank)
    g = golden_params(bits)
    sub = {k: (g[k][shard.res0:shard.res0 + shard.nres] if isinstance(g[k], list) else g[k]) for k in g}
    local = Oracle(bits, N, shard.nres, params=sub).run("fwd", sh.local_view(a, shard))
    slab_ctx, full_ctx = _FileCtx(tmp, N, shard.nres, np.uint32), _FileCtx(tmp, N, M, np.uint32)
    mine = slab_ctx.alloc(shard.npolys)
    _FileCtx.registry[mine][...] = local
    _FileCtx.registry[mine].flush()
    dst, slabs, close = sh.gather_residues_peer(full_ctx, slab_ctx, mine, shard, world, rank)
    exp = Oracle(bits, N, M).run("fwd", a)[shard.poly0:shard.poly0 + shard.npolys]
    ok = bool(np.array_equal(np.asarray(_FileCtx.registry[dst]), exp)) and len(slabs) == 2
    dist.barrier()
    close()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_rank_gloo_peer_gather_protocol(tmp_path):
    """gather_residues_peer's host logic (handle exchange, partner selection, residue tiling, strided placement) with two gloo
    ranks and file-backed stand-ins for device memory; the CUDA IPC version of the same exchange runs in tests/test_round2.py."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
