"""BASELINE configs[4]: 1 MiB frames, zstd level 19 and LZ4, frame-range sharded over the GPUs of one box, no collective.

    python tools/config5.py [--gib-per-gpu 8] [--steps 3]                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config5.py ...

Every rank decodes its own contiguous frame range (here: its own G GiB shard, built by tile-and-replicate from a tile the
reference writer compressed once; N ranks x G GiB = the corpus).  Timing: CUDA events on the launching stream, barrier
on both sides, max over ranks.  Rank 0 prints one JSON line per codec.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import libzseek_b200 as z
    from datagen import refwriter, zsyn
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib-per-gpu", type=float, default=8.0)
    ap.add_argument("--tile-mib", type=int, default=256)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["ZSEEK_B200_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    base = f"/dev/shm/zsk_cfg5_{os.environ.get('MASTER_PORT', 'solo')}"
    kinds = {"lz4_1m": (1, 0), "zstd19_1m": (0, 19)}
    if rank == 0:
        tile = zsyn.gen_parallel(args.tile_mib << 20)
        for k, (codec, level) in kinds.items():
            t0 = time.time()
            img = refwriter.write_parallel(tile, codec, level, 1 << 20, strategy=0, piece_frames=4)
            open(f"{base}_{k}.zsk", "wb").write(img)
            print(f"[config5] reference writer {k}: ratio {len(tile) / len(img):.3f} in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    if world > 1:
        dist.barrier()
    reps = max(1, int(args.gib_per_gpu * (1 << 30)) // (args.tile_mib << 20))
    for k in kinds:
        one = open(f"{base}_{k}.zsk", "rb").read()
        image = np.frombuffer(refwriter.replicate(one, reps), dtype=np.uint8)
        with z.Reader(image=torch.from_numpy(image.copy()).pin_memory(), cache_size=0) as rd:
            out = torch.empty(rd.size + 64, dtype=torch.uint8, device="cuda")
            rd.load(0, rd.frames)
            rd.decode_frames(0, rd.frames, out)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            rd.timer_start()
            for _ in range(args.steps):
                rd.decode_frames(0, rd.frames, out)
            ms = rd.timer_stop()
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            ok = bool((out[:1 << 20].cpu().numpy() == np.frombuffer(zsyn.gen_parallel(args.tile_mib << 20), dtype=np.uint8)[:1 << 20]).all()) if rank == 0 else True
            if rank == 0:
                C, D = int(rd.c_off[-1]), rd.size
                print(json.dumps({"workload": f"BASELINE configs[4]: {k}, 1 MiB frames, {args.gib_per_gpu:g} GiB per GPU x {world} GPUs, frame-range shards, no collective",
                                  "n_gpus": world, "frames_per_gpu": rd.frames, "decompressed_GBps": round(D * world * args.steps / (ms / 1e3) / 1e9, 2),
                                  "ms_per_step": round(ms / args.steps, 3), "algorithmic_GBps_per_gpu": round((C + D) * args.steps / (ms / 1e3) / 1e9, 1),
                                  "kernel": rd.last_decode_kernel, "first_mib_matches_source": ok}), flush=True)
            del out
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        for k in kinds:
            os.remove(f"{base}_{k}.zsk")


if __name__ == "__main__":
    main()
