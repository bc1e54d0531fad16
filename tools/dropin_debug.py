"""T-thread reader-per-thread scan through the C harness with ZSEEK_B200_DEBUG timings (debug helper)."""
import os, sys, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
def child(lib, path, T):
    import ctypes as C, struct, time
    import numpy as np
    L = C.CDLL(os.path.join(ROOT, "oracle", "librefdrive.so"))
    L.refdrive_init.argtypes = [C.c_char_p]
    L.refdrive_scan.restype = C.c_double
    L.refdrive_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    assert L.refdrive_init(lib.encode()) == 0
    image = np.fromfile(path, dtype=np.uint8)
    n = struct.unpack("<I", image[-9:-5].tobytes())[0]
    ent = np.frombuffer(image, dtype="<u4", count=2 * n, offset=len(image) - (8 + 8 * n + 9) + 8).reshape(n, 2)
    total = int(ent[:, 1].astype(np.uint64).sum())
    nb = C.c_uint64()
    for rep in range(int(os.environ.get("REPS", "3"))):
        t0 = time.time()
        t = L.refdrive_scan(image.ctypes.data, image.size, total, T, 4096, 1, 1, None, C.byref(nb))
        print(f"rep {rep}: harness {t:.3f} s ({total / t / 1e9:.2f} GB/s), wall incl. open/close {time.time() - t0:.3f} s", flush=True)
def main():
    if sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], int(sys.argv[4])); return
    from datagen import refwriter, zsyn
    kind, size, T = sys.argv[1], int(sys.argv[2]) << 20, int(sys.argv[3])
    codec, level, frame = {"lz4": (1, 0, 65536), "zstd3": (0, 3, 262144)}[kind]
    tile = zsyn.gen_parallel(size)
    path = os.path.join("/dev/shm", f"zsk_dbg_{kind}.zsk")
    open(path, "wb").write(refwriter.write_parallel(tile, codec, level, frame, piece_frames=max(1, (16 << 20) // frame)))
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child", os.path.join(ROOT, "libzseek_b200", "libzseek_b200.so"), path, str(T)])
    os.remove(path)
if __name__ == "__main__":
    main()
