"""Worker of tests/test_multi_gpu.py (one process per GPU under torch.distributed.run, backend nccl): every rank decodes
its frame shard on its own B200, rank 0 collects the decoded file over NVLink — plain gather_to and the chunked
'decode chunk j -> send chunk j' pipeline — and compares it with the writer's input, byte for byte."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import libzseek_b200 as z
    from datagen import refwriter, zsyn
    from libzseek_b200.sharding import decode_and_gather, gather_to, shard_range

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ["ZSEEK_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    data = zsyn.gen(24 << 20, seed=77) + b"tail" * 1000          # ragged last frame
    for codec, level, frame in ((1, 0, 65536), (0, 3, 262144)):
        image = refwriter.write(data, codec, level, frame)           # deterministic: every rank writes the same file
        with z.Reader(image=image, cache_size=0) as rd:
            rd.set_shard(rank, world)
            lo, hi = shard_range(rd.frames, rank, world)
            d_off = rd.d_off
            nbytes = int(d_off[hi] - d_off[lo])
            # (1) decode the shard, then gather
            local_buf = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
            assert rd.decode_frames(lo, hi, local_buf) == nbytes
            whole = gather_to(local_buf[:nbytes], d_off, dst_rank=0)
            if rank == 0:
                assert whole.cpu().numpy().tobytes() == data, "gather_to over NCCL differs from the writer's input"
            # (2) chunked pipeline, chunks of 1 MiB
            out = torch.zeros(rd.size + 64, dtype=torch.uint8, device="cuda") if rank == 0 else None
            local_buf.zero_()

            def decode_chunk(f0, f1, view):
                assert rd.decode_frames(f0, f1, view) == view.numel()

            got = decode_and_gather(decode_chunk, d_off, out=out, local=local_buf, dst_rank=0, chunk_bytes=1 << 20)
            torch.cuda.synchronize()
            if rank == 0:
                assert got[:rd.size].cpu().numpy().tobytes() == data, "decode_and_gather over NCCL differs from the writer's input"
            # a frame outside the shard is refused
            other = (hi % rd.frames) if world > 1 else None
            if other is not None and not (lo <= other < hi):
                try:
                    rd.pread(10, int(d_off[other]))
                    raise AssertionError("read outside the shard succeeded")
                except z.ZseekError:
                    pass
        dist.barrier()
    if rank == 0:
        print("MGPU_GATHER_OK", world, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
