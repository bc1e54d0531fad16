"""ctypes mirror of include/zseek.h + include/zseek_b200.h.

`Reader` keeps the reference's call shapes — open (FILE*, callbacks or memory image), pread, read,
stats, close — and adds the batched / multi-frame / sharded entry points.  Buffers are passed as
plain addresses: numpy arrays for host memory, ``tensor.data_ptr()`` for device memory.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("ZSEEK_B200_LIB") or os.path.join(ROOT, "libzseek_b200", "libzseek_b200.so")
ERRBUF = 80
ZSTD, LZ4 = 0, 1

PREAD_CB = C.CFUNCTYPE(C.c_ssize_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p)
FSIZE_CB = C.CFUNCTYPE(C.c_ssize_t, C.c_void_p, C.c_void_p)


class ReadFile(C.Structure):  # zseek_read_file_t, reference src/zseek.h:109-116
    _fields_ = [("user_data", C.c_void_p), ("pread", PREAD_CB), ("fsize", FSIZE_CB)]


class ReaderStats(C.Structure):  # zseek_reader_stats_t, reference src/zseek.h:190-203
    _fields_ = [(n, C.c_size_t) for n in ("seek_table_memory", "frames", "decompressed_size", "cache_memory",
                                           "cached_frames", "buffer_size")]


class ZseekError(OSError):
    pass


def build():
    """Compile libzseek_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", ROOT, "-s"], check=True)


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZseekError(f"{LIB_PATH} is missing: run `make` (there is no fallback implementation)")
    L = C.CDLL(LIB_PATH)
    vp, sz, cp = C.c_void_p, C.c_size_t, C.c_char_p
    L.zseek_reader_open_full.restype = vp
    L.zseek_reader_open_full.argtypes = [ReadFile, sz, vp, cp]
    L.zseek_reader_open.restype = vp
    L.zseek_reader_open.argtypes = [vp, sz, vp, cp]
    L.zseek_reader_close.restype = C.c_bool
    L.zseek_reader_close.argtypes = [vp, vp, cp]
    L.zseek_pread.restype = C.c_ssize_t
    L.zseek_pread.argtypes = [vp, vp, sz, sz, vp, cp]
    L.zseek_read.restype = C.c_ssize_t
    L.zseek_read.argtypes = [vp, vp, sz, vp, cp]
    L.zseek_reader_stats.restype = C.c_bool
    L.zseek_reader_stats.argtypes = [vp, C.POINTER(ReaderStats), cp]
    L.zseek_b200_reader_open_mem.restype = vp
    L.zseek_b200_reader_open_mem.argtypes = [vp, sz, sz, cp]
    L.zseek_b200_set_shard.restype = C.c_bool
    L.zseek_b200_set_shard.argtypes = [vp, C.c_uint, C.c_uint, cp]
    L.zseek_b200_get_shard.restype = C.c_bool
    L.zseek_b200_get_shard.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
    L.zseek_b200_seek_table.restype = C.c_bool
    L.zseek_b200_seek_table.argtypes = [vp, C.POINTER(sz), C.POINTER(C.POINTER(C.c_uint64)),
                                        C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_int)]
    L.zseek_b200_load.restype = C.c_bool
    L.zseek_b200_load.argtypes = [vp, sz, sz, vp, cp]
    L.zseek_b200_decode_frames.restype = C.c_ssize_t
    L.zseek_b200_decode_frames.argtypes = [vp, sz, sz, vp, vp, cp]
    L.zseek_b200_read_range.restype = C.c_ssize_t
    L.zseek_b200_read_range.argtypes = [vp, vp, sz, sz, vp, cp]
    L.zseek_b200_pread_batch.restype = C.c_ssize_t
    L.zseek_b200_pread_batch.argtypes = [vp, sz, vp, vp, C.c_uint64, vp, vp, C.c_uint64, vp, vp, cp]
    L.zseek_b200_pread_batch_async.restype = C.c_ssize_t
    L.zseek_b200_pread_batch_async.argtypes = [vp, sz, vp, vp, C.c_uint64, vp, vp, C.c_uint64, vp, vp, cp]
    L.zseek_b200_batch_wait.restype = C.c_int
    L.zseek_b200_batch_wait.argtypes = [vp, cp]
    L.zseek_b200_cache_clear.restype = None
    L.zseek_b200_cache_clear.argtypes = [vp]
    L.zseek_b200_unload.restype = None
    L.zseek_b200_unload.argtypes = [vp]
    L.zseek_b200_timer_start.restype = C.c_bool
    L.zseek_b200_timer_start.argtypes = [vp]
    L.zseek_b200_timer_stop.restype = C.c_double
    L.zseek_b200_timer_stop.argtypes = [vp]
    L.zseek_b200_launch_count.restype = C.c_ulonglong
    L.zseek_b200_launch_count.argtypes = [vp]
    L.zseek_b200_last_decode_ms.restype = C.c_double
    L.zseek_b200_last_decode_ms.argtypes = [vp]
    L.zseek_b200_last_decode_kernel.restype = C.c_char_p
    L.zseek_b200_last_decode_kernel.argtypes = [vp]
    L.zseek_b200_device.restype = C.c_int
    L.zseek_b200_device.argtypes = [vp]
    _lib = L
    return L


def _addr(buf):
    """Address of a destination/source buffer: int (device pointer), numpy array, or torch tensor."""
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    raise TypeError(f"unsupported buffer type {type(buf)}")


class Reader:
    """zseek_reader_t.  Construct with exactly one of: `path` (zseek_reader_open over a FILE*),
    `image` (host bytes/ndarray -> zseek_b200_reader_open_mem), or `pread`/`fsize` Python callbacks
    (zseek_reader_open_full)."""

    def __init__(self, path=None, image=None, pread=None, fsize=None, cache_size=0):
        self.L = load_library()
        self.err = C.create_string_buffer(ERRBUF)
        self._keep = []
        self._file = None
        self.h = None
        if path is not None:
            libc = C.CDLL(None)
            libc.fopen.restype = C.c_void_p
            libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
            libc.fclose.argtypes = [C.c_void_p]
            self._libc = libc
            self._file = libc.fopen(os.fsencode(path), b"rb")
            if not self._file:
                raise ZseekError(f"cannot open {path}")
            h = self.L.zseek_reader_open(self._file, cache_size, None, self.err)
        elif image is not None:
            if hasattr(image, "data_ptr"):  # torch CPU tensor (e.g. pinned memory)
                self._keep.append(image)
                h = self.L.zseek_b200_reader_open_mem(image.data_ptr(), image.numel(), cache_size, self.err)
            else:
                arr = np.frombuffer(image, dtype=np.uint8) if not isinstance(image, np.ndarray) else image
                arr = np.ascontiguousarray(arr)
                self._keep.append(arr)
                h = self.L.zseek_b200_reader_open_mem(arr.ctypes.data, arr.size, cache_size, self.err)
        elif pread is not None and fsize is not None:
            def _pread(data, size, offset, user, call):
                try:
                    b = pread(size, offset)
                    if b is None:
                        return -1
                    C.memmove(data, b, len(b))
                    return len(b)
                except Exception:
                    return -1

            def _fsize(user, call):
                try:
                    return int(fsize())
                except Exception:
                    return -1
            cb1, cb2 = PREAD_CB(_pread), FSIZE_CB(_fsize)
            self._keep += [cb1, cb2]
            h = self.L.zseek_reader_open_full(ReadFile(None, cb1, cb2), cache_size, None, self.err)
        else:
            raise TypeError("Reader needs path=, image= or pread=/fsize=")
        if not h:
            self._close_file()
            raise ZseekError(self.err.value.decode())
        self.h = h
        n, co, do, codec = C.c_size_t(), C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)(), C.c_int()
        self.L.zseek_b200_seek_table(h, C.byref(n), C.byref(co), C.byref(do), C.byref(codec))
        self.frames = n.value
        self.codec = codec.value
        self.c_off = np.ctypeslib.as_array(co, shape=(self.frames + 1,)).copy()
        self.d_off = np.ctypeslib.as_array(do, shape=(self.frames + 1,)).copy()
        self.size = int(self.d_off[-1])

    # ---- reference API
    def pread_into(self, buf, count: int, offset: int) -> int:
        """zseek_pread into `buf` (numpy array, torch tensor or raw address); returns its result."""
        r = self.L.zseek_pread(self.h, _addr(buf), count, offset, None, self.err)
        if r < 0:
            raise ZseekError(self.err.value.decode())
        return r

    def pread(self, count: int, offset: int):
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.pread_into(buf, count, offset)
        return r, buf[:r].tobytes()

    def read(self, count: int):
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.L.zseek_read(self.h, buf.ctypes.data, count, None, self.err)
        if r < 0:
            raise ZseekError(self.err.value.decode())
        return r, buf[:r].tobytes()

    def stats(self) -> ReaderStats:
        st = ReaderStats()
        if not self.L.zseek_reader_stats(self.h, C.byref(st), self.err):
            raise ZseekError(self.err.value.decode())
        return st

    def close(self):
        if self.h:
            self.L.zseek_reader_close(self.h, None, self.err)
            self.h = None
        self._close_file()

    def _close_file(self):
        if self._file:
            self._libc.fclose(self._file)  # the library never closes the caller's FILE*
            self._file = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- additive API (include/zseek_b200.h)
    def set_shard(self, rank: int, world: int):
        if not self.L.zseek_b200_set_shard(self.h, rank, world, self.err):
            raise ZseekError(self.err.value.decode())
        lo, hi = C.c_size_t(), C.c_size_t()
        self.L.zseek_b200_get_shard(self.h, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def load(self, frame_lo: int = 0, frame_hi: int | None = None):
        if not self.L.zseek_b200_load(self.h, frame_lo, self.frames if frame_hi is None else frame_hi, None, self.err):
            raise ZseekError(self.err.value.decode())

    def decode_frames(self, frame_lo: int, frame_hi: int, dev_dst) -> int:
        r = self.L.zseek_b200_decode_frames(self.h, frame_lo, frame_hi, _addr(dev_dst), None, self.err)
        if r < 0:
            raise ZseekError(self.err.value.decode())
        return r

    def read_range_into(self, buf, count: int, offset: int) -> int:
        r = self.L.zseek_b200_read_range(self.h, _addr(buf), count, offset, None, self.err)
        if r < 0:
            raise ZseekError(self.err.value.decode())
        return r

    def read_range(self, count: int, offset: int) -> bytes:
        buf = np.empty(max(count, 1), dtype=np.uint8)
        r = self.read_range_into(buf, count, offset)
        return buf[:r].tobytes()

    def pread_batch(self, offsets, counts=None, fixed_count: int = 0, dst=None, dst_offs=None, dst_stride: int = 0):
        """-> int64 results array (what zseek_pread would return per request)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = offsets.size
        if counts is not None:
            counts = np.ascontiguousarray(counts, dtype=np.uint64)
        if dst_offs is not None:
            dst_offs = np.ascontiguousarray(dst_offs, dtype=np.uint64)
        results = np.zeros(n, dtype=np.int64)
        r = self.L.zseek_b200_pread_batch(self.h, n, offsets.ctypes.data, counts.ctypes.data if counts is not None else None,
                                          fixed_count, _addr(dst), dst_offs.ctypes.data if dst_offs is not None else None,
                                          dst_stride, results.ctypes.data, None, self.err)
        if r < 0:
            raise ZseekError(self.err.value.decode())
        return results

    def pread_batch_async(self, dev_offsets, dev_dst, fixed_count: int = 0, dev_counts=None, dev_dst_offs=None, dst_stride: int = 0,
                          dev_results=None, stream=None):
        """zseek_b200_pread_batch_async: every array is a device tensor (uint64 offsets/counts/dst_offs viewed as int64 is
        fine, int64 results); queued on `stream` (a torch.cuda.Stream, an int cudaStream_t, or None = the reader's own
        stream); returns at once.  batch_wait() reports frame errors."""
        n = dev_offsets.numel()
        sp = stream.cuda_stream if hasattr(stream, "cuda_stream") else (stream or 0)
        r = self.L.zseek_b200_pread_batch_async(self.h, n, _addr(dev_offsets), _addr(dev_counts) if dev_counts is not None else None,
                                                fixed_count, _addr(dev_dst), _addr(dev_dst_offs) if dev_dst_offs is not None else None,
                                                dst_stride, _addr(dev_results) if dev_results is not None else None, sp or None, self.err)
        if r < 0:
            raise ZseekError(self.err.value.decode())
        return r

    def batch_wait(self):
        if self.L.zseek_b200_batch_wait(self.h, self.err) != 0:
            raise ZseekError(self.err.value.decode())

    def cache_clear(self):
        self.L.zseek_b200_cache_clear(self.h)

    def unload(self):
        self.L.zseek_b200_unload(self.h)

    def timer_start(self):
        if not self.L.zseek_b200_timer_start(self.h):
            raise ZseekError("timer_start failed")

    def timer_stop(self) -> float:
        ms = float(self.L.zseek_b200_timer_stop(self.h))
        if ms < 0:
            raise ZseekError("timer_stop failed")
        return ms

    @property
    def launch_count(self) -> int:
        return int(self.L.zseek_b200_launch_count(self.h))

    @property
    def last_decode_ms(self) -> float:
        return float(self.L.zseek_b200_last_decode_ms(self.h))

    @property
    def last_decode_kernel(self) -> str:
        return (self.L.zseek_b200_last_decode_kernel(self.h) or b"").decode()

    @property
    def device(self) -> int:
        return int(self.L.zseek_b200_device(self.h))


def pread_full(reader, count: int, offset: int) -> bytes:
    """Loop zseek_pread over the short reads at frame boundaries (reference test/example.c:64-80)."""
    out = bytearray()
    while len(out) < count:
        r, b = reader.pread(count - len(out), offset + len(out))
        if r == 0:
            break
        out += b
    return bytes(out)
