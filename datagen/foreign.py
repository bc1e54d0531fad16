"""Seekable files that the reference WRITER never emits but the reference READER accepts (it hands every frame
to libzstd / liblz4): frames compressed directly with the system codec runtimes through ctypes — LZ4 block sizes
256 KiB / 1 MiB / 4 MiB, LZ4 block and content checksums, independent blocks, zstd content checksums, zstd
frames without content size — wrapped in a seek table (optionally with per-entry checksum fields, which the
reference parses over and never verifies, src/seek_table.c:95-98).  Used to widen parity beyond the writer's
output.  Test/bench input production only.
"""
import ctypes as C
import struct

import numpy as np

from .refwriter import SEEK_MAGIC_FOOTER, SEEK_MAGIC_SKIPPABLE

_zstd = C.CDLL("libzstd.so.1")
_lz4 = C.CDLL("liblz4.so.1")
_zstd.ZSTD_createCCtx.restype = C.c_void_p
_zstd.ZSTD_CCtx_setParameter.argtypes = [C.c_void_p, C.c_int, C.c_int]
_zstd.ZSTD_CCtx_setParameter.restype = C.c_size_t
_zstd.ZSTD_compress2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
_zstd.ZSTD_compress2.restype = C.c_size_t
_zstd.ZSTD_compressBound.argtypes = [C.c_size_t]
_zstd.ZSTD_compressBound.restype = C.c_size_t
_zstd.ZSTD_freeCCtx.argtypes = [C.c_void_p]
_zstd.ZSTD_isError.argtypes = [C.c_size_t]


class _FrameInfo(C.Structure):
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int), ("frameType", C.c_int),
                ("contentSize", C.c_ulonglong), ("dictID", C.c_uint), ("blockChecksumFlag", C.c_int)]


class _Prefs(C.Structure):
    _fields_ = [("frameInfo", _FrameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint), ("favorDecSpeed", C.c_uint),
                ("reserved", C.c_uint * 3)]


_lz4.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.POINTER(_Prefs)]
_lz4.LZ4F_compressFrameBound.restype = C.c_size_t
_lz4.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(_Prefs)]
_lz4.LZ4F_compressFrame.restype = C.c_size_t
_lz4.LZ4F_isError.argtypes = [C.c_size_t]


def zstd_frame(data: bytes, level=3, checksum=False, content_size=True, window_log=0) -> bytes:
    cctx = _zstd.ZSTD_createCCtx()
    try:
        for param, val in ((100, level), (201, int(checksum)), (200, int(content_size))):
            assert not _zstd.ZSTD_isError(_zstd.ZSTD_CCtx_setParameter(cctx, param, val))
        if window_log:
            assert not _zstd.ZSTD_isError(_zstd.ZSTD_CCtx_setParameter(cctx, 101, window_log))
        cap = _zstd.ZSTD_compressBound(len(data))
        dst = C.create_string_buffer(cap)
        n = _zstd.ZSTD_compress2(cctx, dst, cap, data, len(data))
        assert not _zstd.ZSTD_isError(n)
        return dst.raw[:n]
    finally:
        _zstd.ZSTD_freeCCtx(cctx)


def lz4_frame(data: bytes, block_size_id=4, independent=False, block_checksum=False, content_checksum=False,
              content_size=True, level=0) -> bytes:
    p = _Prefs()
    p.frameInfo.blockSizeID = block_size_id
    p.frameInfo.blockMode = int(independent)
    p.frameInfo.contentChecksumFlag = int(content_checksum)
    p.frameInfo.blockChecksumFlag = int(block_checksum)
    p.frameInfo.contentSize = len(data) if content_size else 0
    p.compressionLevel = level
    cap = _lz4.LZ4F_compressFrameBound(len(data), C.byref(p))
    dst = C.create_string_buffer(cap)
    n = _lz4.LZ4F_compressFrame(dst, cap, data, len(data), C.byref(p))
    assert not _lz4.LZ4F_isError(n)
    return dst.raw[:n]


def seekable(frames, dsizes, entry_checksums=False) -> bytes:
    """Concatenated frames + zstd-seekable seek table (reference src/seek_table.c:15-23,112-176)."""
    n = len(frames)
    es = 12 if entry_checksums else 8
    body = b"".join(struct.pack("<II", len(f), d) + (struct.pack("<I", 0xDEADBEEF) if entry_checksums else b"")
                    for f, d in zip(frames, dsizes))
    return (b"".join(frames) + struct.pack("<II", SEEK_MAGIC_SKIPPABLE, n * es + 9) + body
            + struct.pack("<IBI", n, 0x80 if entry_checksums else 0, SEEK_MAGIC_FOOTER))


def build(data: bytes, frame_size: int, codec: str, **kw) -> bytes:
    pieces = [data[o:o + frame_size] for o in range(0, len(data), frame_size)]
    ent = kw.pop("entry_checksums", False)
    frames = [zstd_frame(p, **kw) if codec == "zstd" else lz4_frame(p, **kw) for p in pieces]
    return seekable(frames, [len(p) for p in pieces], ent)
