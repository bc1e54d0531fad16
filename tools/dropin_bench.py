"""The six-symbol drop-in path under the caller's own threads: sequential scan GB/s and random 4 KiB
ops/s through plain zseek_pread — no additive call — at T = 1 and T = all cores, one reader per thread and one reader
shared by all threads, for the reference build and for libzseek_b200.so.  The same C harness (oracle/refdrive.c) drives
both libraries; each library runs in its own process.  Prints one JSON object.

    python tools/dropin_bench.py [size_mib] [threads]
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib, path, threads):
    import ctypes as C
    import struct
    import numpy as np
    L = C.CDLL(os.path.join(ROOT, "oracle", "librefdrive.so"))
    L.refdrive_init.argtypes = [C.c_char_p]
    L.refdrive_scan.restype = C.c_double
    L.refdrive_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    L.refdrive_random.restype = C.c_double
    L.refdrive_random.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    L.refdrive_set_shared.argtypes = [C.c_int]
    assert L.refdrive_init(lib.encode()) == 0
    image = np.fromfile(path, dtype=np.uint8)
    n = struct.unpack("<I", image[-9:-5].tobytes())[0]
    ent = np.frombuffer(image, dtype="<u4", count=2 * n, offset=len(image) - (8 + 8 * n + 9) + 8).reshape(n, 2)
    total = int(ent[:, 1].astype(np.uint64).sum())
    rng = np.random.Generator(np.random.PCG64(1))
    out = {}
    nbytes, ops = C.c_uint64(), C.c_uint64()
    for shared in (0, 1):
        L.refdrive_set_shared(shared)
        for T in sorted({1, threads}):
            if shared and T == 1:
                continue
            key = f"T{T}_{'shared_reader' if shared else 'reader_per_thread'}"
            # best of three passes for both libraries; every pass opens fresh readers (libzseek_b200 hands a new reader the
            # device state of a closed one, so its first pass is the cold-process figure and the later ones the steady state)
            times = []
            for _ in range(3):
                t = L.refdrive_scan(image.ctypes.data, image.size, total, T, 4096, 1, 1, None, C.byref(nbytes))
                assert t > 0 and nbytes.value == total, t
                times.append(t)
            out.setdefault("scan_4k_GBps", {})[key] = round(total / min(times) / 1e9, 3)
            out.setdefault("scan_4k_GBps_first_pass", {})[key] = round(total / times[0] / 1e9, 3)
            offs = rng.integers(0, total - 4096, 200000 if T > 1 else 20000, dtype=np.uint64)
            rates = []
            for _ in range(2):
                t = L.refdrive_random(image.ctypes.data, image.size, offs.ctypes.data, offs.size, 4096, T, 0, 1, None, C.byref(ops))
                assert t > 0, t
                rates.append(ops.value / t)
            out.setdefault("random_4k_ops_per_s", {})[key] = round(max(rates))
            out.setdefault("random_4k_ops_per_s_first_pass", {})[key] = round(rates[0])
    print(json.dumps(out))


def main():
    if sys.argv[1:2] == ["--child"]:
        child(sys.argv[2], sys.argv[3], int(sys.argv[4]))
        return
    from datagen import refwriter, zsyn
    size = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024) << 20
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 1)
    tile = zsyn.gen_parallel(size)
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    res = {"size_mib": size >> 20, "threads": threads, "request": "4096-byte zseek_pread, cache_size 1 (scan) / 0 (random)"}
    for kind, (codec, level, frame) in {"zstd3": (0, 3, 262144), "lz4": (1, 0, 65536)}.items():
        path = os.path.join(d, f"zsk_dropin_{kind}.zsk")
        open(path, "wb").write(refwriter.write_parallel(tile, codec, level, frame, piece_frames=max(1, (16 << 20) // frame)))
        try:
            for name, lib in (("reference", os.path.join(ROOT, "oracle", "_ref", "libzseek_ref.so")),
                              ("b200", os.path.join(ROOT, "libzseek_b200", "libzseek_b200.so"))):
                p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", lib, path, str(threads)], capture_output=True, text=True, timeout=1200)
                if p.returncode != 0:
                    res.setdefault(kind, {})[name] = {"error": p.stderr[-400:]}
                else:
                    res.setdefault(kind, {})[name] = json.loads(p.stdout.strip().splitlines()[-1])
        finally:
            os.remove(path)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
