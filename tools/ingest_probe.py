"""File (page cache) -> HBM ingest rate of a reader opened BY PATH (zseek_reader_open over a FILE*), with 1 and with N
pread(2) worker threads (ZSEEK_B200_IO_THREADS), next to the same load from a pinned memory image.

    python tools/ingest_probe.py [size_mib] [threads ...]
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import libzseek_b200 as z
    from datagen import refwriter, zsyn
    size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
    threads = [int(t) for t in sys.argv[2:]] or [1, 2, 4, 8, 16]
    data = zsyn.gen_parallel(size)
    image = refwriter.write_parallel(data, 1, 0, 65536, piece_frames=1024)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    path = os.path.join(d, "zsk_ingest_probe.zsk")
    open(path, "wb").write(image)
    C = len(image)
    try:
        for t in threads:
            os.environ["ZSEEK_B200_IO_THREADS"] = str(t)
            with z.Reader(path=path) as rd:
                best = 1e9
                for _ in range(3):
                    rd.unload()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    rd.load(0, rd.frames)
                    best = min(best, time.perf_counter() - t0)
                print(f"by path, {t:2d} I/O threads: {C / best / 1e9:6.2f} GB/s file -> HBM ({C >> 20} MiB compressed)", flush=True)
        pinned = torch.frombuffer(bytearray(image), dtype=torch.uint8).pin_memory()
        with z.Reader(image=pinned) as rd:
            best = 1e9
            for _ in range(3):
                rd.unload()
                t0 = time.perf_counter()
                rd.load(0, rd.frames)
                best = min(best, time.perf_counter() - t0)
            print(f"pinned memory image      : {C / best / 1e9:6.2f} GB/s (one DMA)", flush=True)
    finally:
        os.remove(path)


if __name__ == "__main__":
    main()
