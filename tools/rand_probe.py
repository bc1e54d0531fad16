import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import libzseek_b200 as z
from datagen import refwriter, zsyn
tile = zsyn.gen_parallel(256 << 20)
one = refwriter.write_parallel(tile, 0, 3, 262144, strategy=0, piece_frames=256)
image = np.frombuffer(refwriter.replicate(one, 16), dtype=np.uint8)
total = len(tile) * 16
rd = z.Reader(image=torch.from_numpy(image.copy()).pin_memory(), cache_size=16384)
rd.load(0, rd.frames)
out = torch.empty(1000000 * 4096, dtype=torch.uint8, device="cuda")
def offs(n, seed):
    return np.random.Generator(np.random.PCG64(seed)).integers(0, total - 4096, n, dtype=np.uint64)
rd.pread_batch(offs(1000, 0), fixed_count=4096, dst=out, dst_stride=4096)
for n in (10000, 1000000, 1000000, 1000000):
    rd.cache_clear(); torch.cuda.synchronize(); t0 = time.perf_counter()
    rd.pread_batch(offs(n, 100), fixed_count=4096, dst=out, dst_stride=4096); torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"n={n}: cold {t*1e3:.2f} ms kernel {rd.last_decode_ms:.2f} ms ({rd.last_decode_kernel})", flush=True)
