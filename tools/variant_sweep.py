"""A/B of kernel build variants on the same inputs: writes each corpus once (reference writer over unique zsyn-v1 data),
then times whole-file decodes with every library given (a fresh process per library; ZSEEK_B200_LIB selects it).

    python tools/variant_sweep.py zstd3:2048,zstd19:1024,lz4:4096 libzseek_b200/libzseek_b200.so libzseek_b200/libzsk_alt_*.so
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KINDS = {"lz4": (1, 0, 65536), "zstd3": (0, 3, 262144), "zstd19": (0, 19, 1 << 20), "lz4_1m": (1, 0, 1 << 20), "zstd3_64k": (0, 3, 65536)}


def child(path, rawpath, iters):
    import numpy as np
    import torch
    import libzseek_b200 as z
    image = np.fromfile(path, dtype=np.uint8)
    with z.Reader(image=image) as rd:
        out = torch.empty(rd.size + 64, dtype=torch.uint8, device="cuda")
        rd.load(0, rd.frames)
        ms = []
        err = ""
        for _ in range(iters):
            try:
                rd.decode_frames(0, rd.frames, out)
            except Exception as e:  # noqa: BLE001 -- experiment builds may decode garbage on purpose; the kernel time still counts
                err = f"  [{e}]"
            ms.append(rd.last_decode_ms)
        raw = torch.from_numpy(np.fromfile(rawpath, dtype=np.uint8)).cuda()
        ok = bool(torch.equal(out[:raw.numel()], raw))
        best = min(ms)
        print(f"  {os.path.basename(os.environ.get('ZSEEK_B200_LIB', 'default')):28s} {rd.frames} frames  best {best:8.3f} ms  {rd.size / best / 1e6:7.1f} GB/s  "
              f"all {[round(m, 2) for m in ms]}  verified {ok}{err}", flush=True)


def main():
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], int(sys.argv[4]))
        return
    from datagen import refwriter, zsyn
    specs = [k.split(":") for k in sys.argv[1].split(",")]
    libs = sys.argv[2:]
    for kind, mib in specs:
        size = int(mib) << 20
        codec, level, frame = KINDS[kind]
        tile = zsyn.gen_parallel(size)
        path, rawpath = f"/dev/shm/zsk_sweep_{kind}.zsk", f"/dev/shm/zsk_sweep_{kind}.raw"
        open(rawpath, "wb").write(tile)
        open(path, "wb").write(refwriter.write_parallel(tile, codec, level, frame, piece_frames=max(1, (16 << 20) // frame)))
        del tile
        print(f"{kind}: {mib} MiB", flush=True)
        for lib in libs:
            env = dict(os.environ, ZSEEK_B200_LIB=os.path.abspath(lib))
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path, rawpath, "4"], env=env, capture_output=True, text=True)
            sys.stdout.write(p.stdout if p.returncode == 0 else f"  {lib}: FAILED\n{p.stderr[-1500:]}\n")
            sys.stdout.flush()
        os.remove(path)
        os.remove(rawpath)


if __name__ == "__main__":
    main()
