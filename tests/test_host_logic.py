"""Host-side logic that needs no GPU: input construction, shard arithmetic, and the multi-rank gather
plumbing over gloo (world_size 2)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle.pyapi import LZ4, ZSTD, OraclePort, have_reference

needs_ref = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 131072)])
def test_tile_and_replicate_equals_writer_on_tiled_input(codec, level, frame):
    """frames are compressed independently and deterministically, so replicating the
    compressed frames of a tile k times == running the reference writer over the tiled input."""
    from datagen import refwriter, zsyn
    tile = zsyn.gen(4 * frame)
    one = refwriter.write(tile, codec, level, frame)
    assert refwriter.replicate(one, 3) == refwriter.write(tile * 3, codec, level, frame)
    assert refwriter.write_parallel(tile * 2, codec, level, frame, piece_frames=2, workers=2) == refwriter.write(tile * 2, codec, level, frame)


def test_zsyn_is_deterministic():
    from datagen import zsyn
    import hashlib
    assert hashlib.sha256(zsyn.gen(1 << 20)).hexdigest() == hashlib.sha256(zsyn.gen(1 << 20)).hexdigest()
    assert zsyn.gen(100000) == zsyn.gen(200000)[:100000]


def test_shard_ranges_partition_the_frames():
    from libzseek_b200.sharding import shard_byte_ranges, shard_range
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    d_off = np.array([0, 10, 30, 35, 100], dtype=np.uint64)
    assert shard_byte_ranges(d_off, 2) == [(0, 30), (30, 100)]


def _gather_worker(rank, world, port, image_path, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from libzseek_b200.sharding import gather_to, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    image = open(image_path, "rb").read()
    with OraclePort(image) as op:  # stands in for the per-rank GPU decode of its own shard
        lo, hi = shard_range(op.frames, rank, world)
        full = op.decode_all()
        local = torch.from_numpy(full[int(op.d_off[lo]):int(op.d_off[hi])].copy())
        out = gather_to(local, op.d_off, dst_rank=0)
        if rank == 0:
            q.put(bool((out.numpy() == full).all()))
        else:
            assert out is None
    dist.barrier()
    dist.destroy_process_group()


def _pipeline_worker(rank, world, port, image_path, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from libzseek_b200.sharding import chunk_plan, decode_and_gather, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    image = open(image_path, "rb").read()
    with OraclePort(image) as op:  # stands in for the per-rank GPU decode of chunk [f0, f1)
        full = op.decode_all()
        d_off = op.d_off
        calls = []

        def decode_chunk(f0, f1, view):
            calls.append((f0, f1))
            view.copy_(torch.from_numpy(full[int(d_off[f0]):int(d_off[f1])].copy()))

        lo, hi = shard_range(op.frames, rank, world)
        out = torch.zeros(op.size, dtype=torch.uint8) if rank == 0 else None
        local = torch.zeros(int(d_off[hi] - d_off[lo]), dtype=torch.uint8) if rank != 0 else None
        got = decode_and_gather(decode_chunk, d_off, out=out, local=local, dst_rank=0, chunk_bytes=50000)
        assert calls == chunk_plan(d_off, lo, hi, 50000) and len(calls) > 1
        if rank == 0:
            q.put(bool((got.numpy() == full).all()))
        else:
            assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_chunk_plan_covers_the_range():
    from libzseek_b200.sharding import chunk_plan
    d_off = np.cumsum([0] + [100, 100, 250, 40, 1000, 5, 5, 5]).astype(np.uint64)
    for lo, hi in ((0, 8), (2, 7), (3, 3)):
        plan = chunk_plan(d_off, lo, hi, 300)
        assert [a for a, _ in plan] == [lo] * bool(plan) + [b for _, b in plan][:-1]
        assert (not plan and lo == hi) or plan[-1][1] == hi
        for a, b in plan:
            assert b > a and (b - a == 1 or int(d_off[b] - d_off[a]) <= 300)


def test_decode_and_gather_pipeline_over_gloo_world2(tmp_path):
    """The chunked 'decode chunk j -> send chunk j' pipeline on CPU: rank 0 ends up with the whole file, byte-exact."""
    import torch.multiprocessing as mp
    from conftest import GOLDEN
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, os.path.join(GOLDEN, "mix_zstd3.zsk"), q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok


def test_gather_to_over_gloo_world2(tmp_path):
    """N>1 path on CPU: each rank holds the decoded bytes of its frame shard, rank 0 ends up with the
    whole file, byte-exact."""
    import torch.multiprocessing as mp
    from conftest import GOLDEN
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, os.path.join(GOLDEN, "mix_zstd3.zsk"), q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok
