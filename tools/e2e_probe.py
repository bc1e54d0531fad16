"""Host-buffer range read (zseek_b200_read_range: H2D + decode + D2H) under different pipeline settings.

    python tools/e2e_probe.py [size_mib] [chunk_mb:ramp_mb ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import libzseek_b200 as z
    from datagen import refwriter, zsyn
    size = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096) << 20
    settings = sys.argv[2:] or ["512:512", "512:16", "256:16", "128:16", "128:8", "64:8"]
    tile = zsyn.gen_parallel(64 << 20)
    one = refwriter.write_parallel(tile, 1, 0, 65536, piece_frames=64)
    image = np.frombuffer(refwriter.replicate(one, size // len(tile)), dtype=np.uint8)
    pinned = torch.from_numpy(image.copy()).pin_memory()
    host_out = torch.empty(size, dtype=torch.uint8).pin_memory()
    for s in settings:
        chunk, ramp = s.split(":")
        os.environ["ZSEEK_B200_CHUNK_MB"], os.environ["ZSEEK_B200_RAMP_MB"] = chunk, ramp
        rd = z.Reader(image=pinned, cache_size=0)
        best = None
        for i in range(5):
            rd.unload()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            assert rd.read_range_into(host_out, size, 0) == size
            t = time.perf_counter() - t0
            if i >= 1:
                best = t if best is None else min(best, t)
        ok = bool((host_out[:len(tile)].numpy() == np.frombuffer(tile, dtype=np.uint8)).all())
        print(f"chunk {chunk} MiB ramp {ramp} MiB: best {best * 1e3:.2f} ms -> {size / best / 1e9:.2f} GB/s  ok={ok}", flush=True)
        rd.close()


if __name__ == "__main__":
    main()
