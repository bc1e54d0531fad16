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


def gather_to(local, d_off: np.ndarray, dst_rank: int = 0, group=None):
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
        out = torch.empty(int(d_off[-1]), dtype=torch.uint8, device=local.device)
        lo, hi = ranges[rank]
        out[lo:hi].copy_(local)
        reqs = [dist.irecv(out[ranges[r][0]:ranges[r][1]], src=r, group=group) for r in range(world)
                if r != dst_rank and ranges[r][1] > ranges[r][0]]
        for q in reqs:
            q.wait()
        return out
    if local.numel():
        dist.send(local.contiguous(), dst=dst_rank, group=group)
    return None
