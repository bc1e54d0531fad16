"""Seekable-file production with the REFERENCE CPU writer (oracle/_ref/libzseek_ref.so).

north_star: "The write/compress path (compress.c) stays the reference CPU writer and is used only to
produce the inputs."  Everything here funnels through reference zseek_writer_open_full / zseek_write /
zseek_writer_close (reference src/compress.c:247,815,578) via oracle/refdrive.c.

Large files are built by the tile-and-replicate construction of SURVEY.md §8d: because the writer
compresses every frame independently and deterministically, writing a tile whose length is a
multiple of the frame size and replicating its compressed frames k times (then emitting a fresh seek
table, format reference src/seek_table.c:365-419 / SURVEY Appendix A.0) is byte-identical to running
the writer over the tiled input.
"""
import os
import struct
from concurrent.futures import ProcessPoolExecutor

import numpy as np

ZSTD, LZ4 = 0, 1
SEEK_MAGIC_SKIPPABLE = 0x184D2A5E
SEEK_MAGIC_FOOTER = 0x8F92EAB1


def write(data, codec: int, level: int, min_frame_size: int, chunk: int | None = None, strategy: int = 0,
          nb_workers: int = 0) -> bytes:
    """One pass of the reference writer, constant `chunk`-byte zseek_write calls."""
    from oracle.pyapi import RefDrive
    return RefDrive.compress(data, codec, level, min_frame_size, chunk or min_frame_size, strategy, nb_workers)


def split(image: bytes):
    """-> (payload bytes, list of (cSize, dSize)) of a seekable file image (no checksums)."""
    n, desc, magic = struct.unpack("<IBI", image[-9:])
    assert magic == SEEK_MAGIC_FOOTER and desc == 0
    tbl = len(image) - (8 + 8 * n + 9)
    ent = np.frombuffer(image, dtype="<u4", count=2 * n, offset=tbl + 8).reshape(n, 2)
    return image[:tbl], ent


def seek_table(entries: np.ndarray) -> bytes:
    n = len(entries)
    return (struct.pack("<II", SEEK_MAGIC_SKIPPABLE, 8 * n + 9) + np.ascontiguousarray(entries, dtype="<u4").tobytes()
            + struct.pack("<IBI", n, 0, SEEK_MAGIC_FOOTER))


def _write_piece(args):
    data, codec, level, min_frame_size, chunk, strategy = args
    return write(data, codec, level, min_frame_size, chunk, strategy)


def write_parallel(data: bytes, codec: int, level: int, frame_size: int, strategy: int = 0,
                   piece_frames: int = 64, workers: int | None = None) -> bytes:
    """Reference writer over `data` in worker processes, one run per `piece_frames` frames, frames and
    seek-table entries concatenated.  len(data) must be a multiple of frame_size (so that every piece
    sees exactly the frames a single run would produce)."""
    assert len(data) % frame_size == 0
    piece = piece_frames * frame_size
    jobs = [(data[o:o + piece], codec, level, frame_size, frame_size, strategy) for o in range(0, len(data), piece)]
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    if workers <= 1 or len(jobs) == 1:
        images = [_write_piece(j) for j in jobs]
    else:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            images = list(ex.map(_write_piece, jobs))
    payloads, entries = zip(*(split(im) for im in images))
    return b"".join(payloads) + seek_table(np.concatenate(entries))


def replicate(image: bytes, times: int) -> bytes:
    """File image equal to the writer's output on the decompressed content tiled `times` times."""
    payload, ent = split(image)
    return payload * times + seek_table(np.tile(ent, (times, 1)))
