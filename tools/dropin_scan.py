"""Sequential zseek_pread scan through the plain reference API (no additive calls), timed the same way for the
reference build and for libzseek_b200.so: oracle/refdrive.c's scan harness is pointed at either library.

    python tools/dropin_scan.py lz4|zstd3 [size_mib] [request_bytes] [threads]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(lib, path, req, threads):
    import ctypes as C
    import numpy as np
    L = C.CDLL(os.path.join(ROOT, "oracle", "librefdrive.so"))
    L.refdrive_init.argtypes = [C.c_char_p]
    L.refdrive_scan.restype = C.c_double
    L.refdrive_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    assert L.refdrive_init(lib.encode()) == 0
    image = np.fromfile(path, dtype=np.uint8)
    import struct
    n = struct.unpack("<I", image[-9:-5].tobytes())[0]
    ent = np.frombuffer(image, dtype="<u4", count=2 * n, offset=len(image) - (8 + 8 * n + 9) + 8).reshape(n, 2)
    total = int(ent[:, 1].astype(np.uint64).sum())
    out = np.empty(total, dtype=np.uint8)
    nbytes = C.c_uint64()
    best = None
    for _ in range(3):
        t = L.refdrive_scan(image.ctypes.data, image.size, total, threads, req, 1, 0, out.ctypes.data, C.byref(nbytes))
        assert t > 0 and nbytes.value == total, t
        best = t if best is None else min(best, t)
    import hashlib
    print(f"{os.path.basename(lib)}: {total / best / 1e9:.3f} GB/s ({threads} thread(s), {req}-byte zseek_pread, cache_size 1) sha {hashlib.sha256(out).hexdigest()[:12]}")


def main():
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
        return
    from datagen import refwriter, zsyn
    kind = sys.argv[1]
    size = (int(sys.argv[2]) if len(sys.argv) > 2 else 1024) << 20
    req = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    threads = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    codec, level, frame = {"lz4": (1, 0, 65536), "zstd3": (0, 3, 262144)}[kind]
    tile = zsyn.gen_parallel(min(size, 128 << 20))
    image = refwriter.replicate(refwriter.write_parallel(tile, codec, level, frame, piece_frames=256), max(1, size // len(tile)))
    path = f"/dev/shm/dropin_{kind}.zsk"
    open(path, "wb").write(image)
    for lib in (os.path.join(ROOT, "oracle", "_ref", "libzseek_ref.so"), os.path.join(ROOT, "libzseek_b200", "libzseek_b200.so")):
        subprocess.run([sys.executable, __file__, "--child", lib, path, str(req), str(threads)], check=True)
    os.remove(path)


if __name__ == "__main__":
    main()
