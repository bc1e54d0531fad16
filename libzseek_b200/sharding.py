"""Frame-range sharding across the GPUs of one box.

Every frame decodes independently, so rank g of G takes the contiguous frame range
[N*g/G, N*(g+1)/G) and touches only its slice of the compressed file: no data-path collective.
A collective (NCCL over NVLink via torch.distributed) is used ONLY when the caller asks for the
decoded bytes on a single rank (`gather_to`).  torch is plumbing here: device memory, process group.
"""
import numpy as np


def shard_range(n_frames: int, rank: int, world: int):
    """Same arithmetic as zseek_b200_set_shard (libzseek_b200/csrc/reader.c)."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def shard_byte_ranges(d_off: np.ndarray, world: int):
    """[(byte_lo, byte_hi)] of every rank's decoded slice."""
    n = len(d_off) - 1
    return [(int(d_off[shard_range(n, r, world)[0]]), int(d_off[shard_range(n, r, world)[1]])) for r in range(world)]


def gather_to(local, d_off: np.ndarray, dst_rank: int = 0, group=None, out=None):
    """Collects every rank's decoded shard on `dst_rank` as one contiguous uint8 tensor (None elsewhere).

    `local` is this rank's decoded slice (uint8 tensor on the rank's device, or on CPU under gloo).
    Shards are ragged by at most one frame, so the exchange is point-to-point sends of exact byte
    ranges into the right offsets of the destination buffer (no padding, no reduction)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ranges = shard_byte_ranges(d_off, world)
    assert local.numel() == ranges[rank][1] - ranges[rank][0]
    if rank == dst_rank:
        if out is None:
            out = torch.empty(int(d_off[-1]), dtype=torch.uint8, device=local.device)
        lo, hi = ranges[rank]
        out[lo:hi].copy_(local)
        reqs = [dist.irecv(out[ranges[r][0]:ranges[r][1]], src=r, group=group) for r in range(world)
                if r != dst_rank and ranges[r][1] > ranges[r][0]]
        for q in reqs:
            q.wait()
        return out[:int(d_off[-1])]
    if local.numel():
        dist.send(local.contiguous(), dst=dst_rank, group=group)
    return None


def chunk_plan(d_off: np.ndarray, frame_lo: int, frame_hi: int, chunk_bytes: int):
    """[(f0, f1)]: consecutive frame ranges of [frame_lo, frame_hi) of at most chunk_bytes decoded bytes each (at least
    one frame).  A pure function of the seek table, so sender and receiver derive the same plan."""
    d = np.asarray(d_off, dtype=np.uint64)
    out, a = [], int(frame_lo)
    while a < frame_hi:
        b = int(np.searchsorted(d, d[a] + np.uint64(chunk_bytes), side="right")) - 1
        b = min(max(b, a + 1), int(frame_hi))
        out.append((a, b))
        a = b
    return out


def decode_and_gather(decode_chunk, d_off: np.ndarray, out=None, local=None, dst_rank: int = 0, chunk_bytes: int = 256 << 20, group=None):
    """'decode chunk j -> send chunk j' pipeline: every rank decodes its frame shard in chunks and ships
    each chunk to `dst_rank` while the next one decodes; `dst_rank` posts all its receives up front, straight into the
    final positions of `out` (the whole decoded file, uint8 tensor on its device), and decodes its own shard in place.

    decode_chunk(f0, f1, view) must fill `view` (d_off[f1] - d_off[f0] bytes) and have COMPLETED when it returns
    (zseek_b200_decode_frames does).  `local` is the sender's staging tensor for its shard.  No reduction, no padding:
    point-to-point NCCL sends over NVLink of exact byte ranges."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    d = np.asarray(d_off, dtype=np.uint64)
    n = len(d) - 1
    reqs = []
    if rank == dst_rank:
        assert out is not None and out.numel() >= int(d[-1])
        for r in range(world):
            if r == dst_rank:
                continue
            lo, hi = shard_range(n, r, world)
            for f0, f1 in chunk_plan(d, lo, hi, chunk_bytes):
                reqs.append(dist.irecv(out[int(d[f0]):int(d[f1])], src=r, group=group))
        lo, hi = shard_range(n, rank, world)
        for f0, f1 in chunk_plan(d, lo, hi, chunk_bytes):
            decode_chunk(f0, f1, out[int(d[f0]):int(d[f1])])
    else:
        lo, hi = shard_range(n, rank, world)
        base = int(d[lo])
        assert local is not None and local.numel() >= int(d[hi]) - base
        for f0, f1 in chunk_plan(d, lo, hi, chunk_bytes):
            view = local[int(d[f0]) - base:int(d[f1]) - base]
            decode_chunk(f0, f1, view)
            reqs.append(dist.isend(view, dst=dst_rank, group=group))
    for q in reqs:
        q.wait()
    return out if rank == dst_rank else None
