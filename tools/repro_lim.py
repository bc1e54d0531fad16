"""Cold batch over many frames with per-job limits (debug helper)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def main():
    import torch
    import libzseek_b200 as z
    from datagen import refwriter, zsyn
    size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
    nreq = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    tile = zsyn.gen_parallel(min(size, 256 << 20))
    one = refwriter.write_parallel(tile, 0, 3, 262144, piece_frames=64)
    image = refwriter.replicate(one, max(1, size // len(tile)))
    raw = torch.from_numpy(np.frombuffer(tile, dtype=np.uint8).copy()).cuda()
    rng = np.random.Generator(np.random.PCG64(3))
    with z.Reader(image=image, cache_size=1 << 30) as rd:
        rd.load(0, rd.frames)
        offs = rng.integers(0, rd.size - 4096, nreq, dtype=np.uint64)
        out = torch.zeros(nreq * 4096, dtype=torch.uint8, device="cuda")
        for attempt in range(3):
            rd.cache_clear()
            try:
                res = rd.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)
            except z.ZseekError as e:
                print("attempt", attempt, "FAILED:", e, flush=True)
                continue
            ar = torch.arange(4096, device="cuda")
            so = torch.from_numpy(offs.astype(np.int64)).cuda()
            ok = 0
            for o in range(0, nreq, 8192):
                k = min(8192, nreq - o)
                exp = raw[(so[o:o + k, None] + ar[None, :]) % len(tile)]
                ln = torch.from_numpy(res[o:o + k]).cuda()
                ok += int(((exp == out[o * 4096:(o + k) * 4096].view(k, 4096)) | (ar[None, :] >= ln[:, None])).all(dim=1).sum())
            print("attempt", attempt, "ok requests", ok, "of", nreq, flush=True)
if __name__ == "__main__":
    main()
