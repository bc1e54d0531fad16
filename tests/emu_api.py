"""ctypes wrapper of the kernel-logic emulator (tests/emu, TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_SO = os.path.join(EMU_DIR, "libzsk_emu.so")
FRONT_PAD, BACK_PAD = 16, 64


def lib():
    subprocess.run(["make", "-C", EMU_DIR, "-s"], check=True)
    L = C.CDLL(EMU_SO)
    L.emu_decode.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
    L.emu_lookup.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p,
                             C.c_void_p, C.c_void_p, C.c_void_p]
    L.emu_gather.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_uint64, C.c_uint32]
    L.emu_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                            C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_int, C.c_uint32]
    L.emu_last_deferred.restype = C.c_uint32
    L.emu_set_addr_bias.argtypes = [C.c_uint32]
    return L


def decode_all(L, image: bytes, codec: int, c_off: np.ndarray, d_off: np.ndarray, ctas: int = 3, misalign: int = 0, limits=None, wrap_at=None):
    """Whole-file decode with the emulated K2/K3; returns (decoded bytes, status array)."""
    n = len(c_off) - 1
    payload = int(c_off[-1])
    comp = np.zeros(FRONT_PAD + payload + BACK_PAD, dtype=np.uint8)
    comp[FRONT_PAD:FRONT_PAD + payload] = np.frombuffer(image, dtype=np.uint8, count=payload)
    total = int(d_off[-1])
    raw = np.full(total + 128, 0xEE, dtype=np.uint8)
    lead = (-raw.ctypes.data) % 64 + misalign      # dst starts `misalign` bytes past a 64-byte boundary
    dst = raw[lead:lead + total + 32]
    status = np.full(max(n, 1), -1, dtype=np.int32)
    c_off = np.ascontiguousarray(c_off, dtype=np.uint64)
    d_off = np.ascontiguousarray(d_off, dtype=np.uint64)
    if wrap_at is not None:   # low 32 address bits wrap around `wrap_at` bytes into the compressed image
        L.emu_set_addr_bias((-(comp.ctypes.data + FRONT_PAD + wrap_at)) & 0xFFFFFFF0)
    L.emu_decode(codec, comp.ctypes.data + FRONT_PAD, 0, c_off.ctypes.data, d_off.ctypes.data, None, None,
                 dst.ctypes.data, 0, 0, n, status.ctypes.data, ctas, limits.ctypes.data if limits is not None else None)
    L.emu_set_addr_bias(0)
    assert (dst[total:] == 0xEE).all(), "emulated kernel wrote past the end of the output"
    assert (raw[:lead] == 0xEE).all(), "emulated kernel wrote before the start of the output"
    return dst[:total], status[:n]


def batch(L, image: bytes, codec: int, c_off, d_off, offsets, counts, stride: int, shard=None, max_jobs=None, limits=True, ctas: int = 2):
    """The device side of a stream-ordered batch on the emulator (lookup -> compaction -> decode with the job count in
    "device" memory -> gather): returns (dst, results, job statuses, job count, compaction error flag)."""
    n_frames = len(c_off) - 1
    payload = int(c_off[-1])
    comp = np.zeros(FRONT_PAD + payload + BACK_PAD, dtype=np.uint8)
    comp[FRONT_PAD:FRONT_PAD + payload] = np.frombuffer(image, dtype=np.uint8, count=payload)
    c_off = np.ascontiguousarray(c_off, dtype=np.uint64)
    d_off = np.ascontiguousarray(d_off, dtype=np.uint64)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    counts = np.ascontiguousarray(counts, dtype=np.uint64)
    n = offsets.size
    slot = (int((d_off[1:] - d_off[:-1]).max()) + 255) & ~255
    lo, hi = shard if shard is not None else (0, n_frames)
    max_jobs = min(n, hi - lo) if max_jobs is None else max_jobs
    slab = np.full(FRONT_PAD + max(max_jobs, 1) * slot + BACK_PAD, 0xEE, dtype=np.uint8)
    dst = np.full(n * stride + 64, 0x5A, dtype=np.uint8)
    results = np.full(n, -7, dtype=np.int64)
    status = np.full(max(max_jobs, 1), -1, dtype=np.int32)
    ctl = np.zeros(2, dtype=np.uint32)
    L.emu_batch(codec, comp.ctypes.data + FRONT_PAD, c_off.ctypes.data, d_off.ctypes.data, n_frames, lo, hi, offsets.ctypes.data,
                counts.ctypes.data, 0, n, slot, max_jobs, slab.ctypes.data + FRONT_PAD, dst.ctypes.data, stride, results.ctypes.data,
                status.ctypes.data, ctl.ctypes.data, int(limits), ctas)
    assert (dst[n * stride:] == 0x5A).all()
    return dst[:n * stride], results, status, int(ctl[0]), int(ctl[1])
