"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

RefDrive  -> oracle/librefdrive.so driving the unmodified reference build oracle/_ref/libzseek_ref.so
             (kind "reference"): writer over memory, reader over memory, threaded CPU timing harness.
OraclePort-> oracle/libzsk_oracle.so, the plain-C restatement of the read path (kind "port").
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libzseek_ref.so")
DRIVE_SO = os.path.join(HERE, "librefdrive.so")
PORT_SO = os.path.join(HERE, "libzsk_oracle.so")

ZSTD, LZ4 = 0, 1
ERRBUF = 80


def build():
    """Compile the checkers (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def have_reference() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(DRIVE_SO)


class ReaderStats(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in ("seek_table_memory", "frames", "decompressed_size",
                                           "cache_memory", "cached_frames", "buffer_size")]


def _as_u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a)


class RefDrive:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not have_reference():
                raise RuntimeError("oracle/_ref/libzseek_ref.so or oracle/librefdrive.so missing; run `make -C oracle`")
            L = C.CDLL(DRIVE_SO)
            L.refdrive_init.argtypes = [C.c_char_p]
            L.refdrive_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                            C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_char_p]
            L.refdrive_free.argtypes = [C.c_void_p]
            L.refdrive_reader_open.restype = C.c_void_p
            L.refdrive_reader_open.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p]
            L.refdrive_pread.restype = C.c_ssize_t
            L.refdrive_pread.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p]
            L.refdrive_pread_full.restype = C.c_ssize_t
            L.refdrive_pread_full.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p]
            L.refdrive_read.restype = C.c_ssize_t
            L.refdrive_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_char_p]
            L.refdrive_stats.argtypes = [C.c_void_p, C.POINTER(ReaderStats), C.c_char_p]
            L.refdrive_reader_close.argtypes = [C.c_void_p, C.c_char_p]
            L.refdrive_scan.restype = C.c_double
            L.refdrive_scan.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                        C.c_void_p, C.POINTER(C.c_uint64)]
            L.refdrive_random.restype = C.c_double
            L.refdrive_random.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                          C.c_size_t, C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
            if L.refdrive_init(REF_SO.encode()) != 0:
                raise RuntimeError("refdrive_init failed")
            cls._lib = L
        return cls._lib

    # ---- reference writer
    @classmethod
    def compress(cls, data, codec: int, level: int, min_frame_size: int, chunk: int, strategy: int = 0,
                 nb_workers: int = 0) -> bytes:
        L = cls.lib()
        src = _as_u8(data)
        out = C.c_void_p()
        outlen = C.c_size_t()
        err = C.create_string_buffer(ERRBUF)
        rc = L.refdrive_compress(src.ctypes.data, src.size, codec, level, strategy, nb_workers, min_frame_size, chunk,
                                 C.byref(out), C.byref(outlen), err)
        if rc != 0:
            raise RuntimeError(f"reference writer failed rc={rc}: {err.value.decode()}")
        try:
            return C.string_at(out.value, outlen.value)
        finally:
            L.refdrive_free(out)

    # ---- CPU baseline harness
    @classmethod
    def scan(cls, image: np.ndarray, dsize: int, threads: int, req: int = 1 << 20, cache_size: int = 0,
             pin: bool = True, dst: np.ndarray | None = None):
        L = cls.lib()
        nbytes = C.c_uint64()
        t = L.refdrive_scan(image.ctypes.data, image.size, dsize, threads, req, cache_size, int(pin),
                            dst.ctypes.data if dst is not None else None, C.byref(nbytes))
        if t < 0:
            raise RuntimeError(f"reference scan failed ({t})")
        return t, nbytes.value

    @classmethod
    def random(cls, image: np.ndarray, offsets: np.ndarray, count: int, threads: int, cache_size: int = 0,
               pin: bool = True, dst: np.ndarray | None = None):
        L = cls.lib()
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ops = C.c_uint64()
        t = L.refdrive_random(image.ctypes.data, image.size, offsets.ctypes.data, offsets.size, count, threads,
                              cache_size, int(pin), dst.ctypes.data if dst is not None else None, C.byref(ops))
        if t < 0:
            raise RuntimeError(f"reference random failed ({t})")
        return t, ops.value


class RefReader:
    """The reference reader (zseek_reader_open_full … zseek_reader_close) over a memory image."""

    def __init__(self, image, cache_size: int = 0):
        self.L = RefDrive.lib()
        self.image = _as_u8(image)  # keep alive
        self.err = C.create_string_buffer(ERRBUF)
        self.h = self.L.refdrive_reader_open(self.image.ctypes.data, self.image.size, cache_size, self.err)
        if not self.h:
            raise OSError(self.err.value.decode())

    def pread(self, count: int, offset: int):
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.L.refdrive_pread(self.h, buf.ctypes.data, count, offset, self.err)
        if r < 0:
            raise OSError(self.err.value.decode())
        return r, buf[:r].tobytes()

    def pread_full(self, count: int, offset: int) -> bytes:
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.L.refdrive_pread_full(self.h, buf.ctypes.data, count, offset, self.err)
        if r < 0:
            raise OSError(self.err.value.decode())
        return buf[:r].tobytes()

    def read(self, count: int):
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.L.refdrive_read(self.h, buf.ctypes.data, count, self.err)
        if r < 0:
            raise OSError(self.err.value.decode())
        return r, buf[:r].tobytes()

    def stats(self) -> ReaderStats:
        st = ReaderStats()
        if not self.L.refdrive_stats(self.h, C.byref(st), self.err):
            raise OSError(self.err.value.decode())
        return st

    def close(self):
        if self.h:
            self.L.refdrive_reader_close(self.h, self.err)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class OraclePort:
    """The plain-C restatement (oracle/zsk_oracle.c) over a memory image."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(PORT_SO):
                raise RuntimeError("oracle/libzsk_oracle.so missing; run `make -C oracle`")
            L = C.CDLL(PORT_SO)
            L.zo_reader_open.restype = C.c_void_p
            L.zo_reader_open.argtypes = [C.c_void_p, C.c_size_t]
            L.zo_reader_close.argtypes = [C.c_void_p]
            L.zo_reader_type.argtypes = [C.c_void_p]
            L.zo_reader_seek_table.restype = C.c_void_p
            L.zo_reader_seek_table.argtypes = [C.c_void_p]
            L.zo_seek_table_frames.restype = C.c_uint64
            L.zo_seek_table_frames.argtypes = [C.c_void_p]
            L.zo_seek_table_coff.restype = C.POINTER(C.c_uint64)
            L.zo_seek_table_coff.argtypes = [C.c_void_p]
            L.zo_seek_table_doff.restype = C.POINTER(C.c_uint64)
            L.zo_seek_table_doff.argtypes = [C.c_void_p]
            L.zo_offset_to_frame.restype = C.c_int64
            L.zo_offset_to_frame.argtypes = [C.c_void_p, C.c_uint64]
            L.zo_pread.restype = C.c_ssize_t
            L.zo_pread.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
            L.zo_decode_all.restype = C.c_ssize_t
            L.zo_decode_all.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
            L.zo_reader_decode_frame.restype = C.c_ssize_t
            L.zo_reader_decode_frame.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t]
            L.zo_zstd_frame_decode.restype = C.c_ssize_t
            L.zo_zstd_frame_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            L.zo_lz4_frame_decode.restype = C.c_ssize_t
            L.zo_lz4_frame_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
            cls._lib = L
        return cls._lib

    def __init__(self, image):
        self.L = self.lib()
        self.image = _as_u8(image)
        self.h = self.L.zo_reader_open(self.image.ctypes.data, self.image.size)
        if not self.h:
            raise OSError("oracle: open failed")
        st = self.L.zo_reader_seek_table(self.h)
        self.frames = int(self.L.zo_seek_table_frames(st))
        self.c_off = np.ctypeslib.as_array(self.L.zo_seek_table_coff(st), shape=(self.frames + 1,)).copy()
        self.d_off = np.ctypeslib.as_array(self.L.zo_seek_table_doff(st), shape=(self.frames + 1,)).copy()
        self.codec = self.L.zo_reader_type(self.h)
        self.size = int(self.d_off[-1])

    def offset_to_frame(self, offset: int) -> int:
        return int(self.L.zo_offset_to_frame(self.L.zo_reader_seek_table(self.h), offset))

    def pread(self, count: int, offset: int):
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.L.zo_pread(self.h, buf.ctypes.data, count, offset)
        if r < 0:
            raise OSError("oracle: pread failed")
        return r, buf[:r].tobytes()

    def decode_all(self) -> np.ndarray:
        out = np.empty(max(self.size, 1), dtype=np.uint8)
        r = self.L.zo_decode_all(self.h, out.ctypes.data, self.size)
        if r < 0:
            raise OSError(f"oracle: decode_all failed ({r})")
        return out[:r]

    def decode_frame(self, idx: int) -> np.ndarray:
        dsize = int(self.d_off[idx + 1] - self.d_off[idx])
        out = np.empty(max(dsize, 1), dtype=np.uint8)
        r = self.L.zo_reader_decode_frame(self.h, idx, out.ctypes.data, dsize)
        if r < 0:
            raise OSError(f"oracle: decode_frame failed ({r})")
        return out[:r]

    def close(self):
        if self.h:
            self.L.zo_reader_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
