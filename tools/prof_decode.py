"""Small driver for profiling one decode kernel under ncu (inputs: reference writer over a zsyn-v1 tile).

    python tools/prof_decode.py lz4|zstd3|zstd19|lz4_1m|lz4near [size_mib] [iters] [tile_mib]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import libzseek_b200 as z
    from datagen import refwriter, zsyn
    kind = sys.argv[1] if len(sys.argv) > 1 else "lz4"
    size = (int(sys.argv[2]) if len(sys.argv) > 2 else 256) << 20
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    tile_mib = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    if kind.endswith("near"):
        # diagnostic corpus: same token density as text, but every match source lies within ~600 bytes
        rng = np.random.Generator(np.random.PCG64(5))
        base = rng.integers(97, 123, 600, dtype=np.uint8)
        reps = (64 << 20) // 600 + 1
        arr = np.tile(base, reps)[:64 << 20].copy()
        idx = np.arange(0, arr.size, 11) + rng.integers(0, 5, (arr.size + 10) // 11)
        idx = idx[idx < arr.size]
        arr[idx] = rng.integers(97, 123, idx.size, dtype=np.uint8)
        tile = arr.tobytes()
        kind = kind[:-4]
    else:
        tile = zsyn.gen_parallel(min(size, tile_mib << 20))
    codec, level, frame = {"lz4": (1, 0, 65536), "zstd3": (0, 3, 262144), "zstd19": (0, 19, 1 << 20), "lz4_1m": (1, 0, 1 << 20)}[kind]
    one = refwriter.write_parallel(tile, codec, level, frame, piece_frames=max(1, (4 << 20) // frame))
    image = refwriter.replicate(one, max(1, size // len(tile)))
    with z.Reader(image=image) as rd:
        out = torch.empty(rd.size + 64, dtype=torch.uint8, device="cuda")
        rd.load(0, rd.frames)
        for i in range(iters):
            rd.decode_frames(0, rd.frames, out)
            ms = rd.last_decode_ms
            print(f"{kind}: {rd.size / 2**20:.0f} MiB, {rd.frames} frames, C={int(rd.c_off[-1])} kernel {ms:.3f} ms -> {rd.size / ms / 1e6:.1f} GB/s decompressed, "
                  f"{(rd.size + int(rd.c_off[-1])) / ms / 1e6:.1f} GB/s algorithmic", flush=True)
        ref = np.frombuffer(tile, dtype=np.uint8)
        got = out[:len(tile)].cpu().numpy()
        assert (got == ref).all(), "decode mismatch"


if __name__ == "__main__":
    main()
