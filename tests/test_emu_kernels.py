"""Kernel LOGIC check without a GPU: the product's kernel sources (libzseek_b200/csrc/zsk_*.cuh) are
compiled for the host against tests/emu/cuda_emu.h (every CUDA thread a fiber, warp collectives and
barriers as rendezvous) and compared with the oracle.  Small inputs only — the emulator is slow.
This is test infrastructure; the shipped library has no host decode path."""
import hashlib

import numpy as np
import pytest

import emu_api
from oracle.pyapi import OraclePort

EMU_CASES = ["tiny_zstd", "tiny_lz4", "zsyn_lz4_64k", "zsyn_lz4_256k_linked", "zsyn_zstd3_128k", "zsyn_zstd19_256k",
             "zsyn_zstd3_mt", "mix_lz4", "mix_zstd3", "mix_zstd19"]


@pytest.fixture(scope="module")
def emu():
    return emu_api.lib()


@pytest.mark.parametrize("name", EMU_CASES)
def test_emulated_decode_matches_reference_output(emu, golden, name):
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        out, status = emu_api.decode_all(emu, c["image"], op.codec, op.c_off, op.d_off, ctas=2)
    assert (status == 0).all(), status
    assert hashlib.sha256(out.tobytes()).hexdigest() == c["input_sha256"]


ZSTD_CASES = ["tiny_zstd", "zsyn_zstd3_128k", "zsyn_zstd19_256k", "zsyn_zstd3_mt", "mix_zstd3", "mix_zstd19"]


@pytest.mark.parametrize("name,kernel", [(n, 201) for n in ZSTD_CASES] + [("tiny_zstd", 200), ("mix_zstd19", 200)])   # 201 hands most frames to the kernel that 200 runs alone
def test_emulated_zstd_deferred_frames(emu, golden, name, kernel):
    """200: the one-CTA-per-frame kernel alone (what deferred frames get).  201: the pipeline with scratch pools far too
    small, so P0 hands most frames over to that kernel inside the same launch."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        out, status = emu_api.decode_all(emu, c["image"], kernel, op.c_off, op.d_off, ctas=2)
        if kernel == 201 and op.frames > 1 and name != "tiny_zstd":
            assert emu.emu_last_deferred() > 0
    assert (status == 0).all(), status
    assert hashlib.sha256(out.tobytes()).hexdigest() == c["input_sha256"]


@pytest.mark.parametrize("name", ["zsyn_zstd3_128k", "mix_zstd19"])
def test_emulated_zstd_bitstream_ring_across_a_4gib_address_boundary(emu, golden, name):
    """The FSE stage stages each block's bitstream in a shared-memory ring addressed by the LOW 32 bits of the global address.
    A compressed image of a few GiB straddles a 4 GiB boundary somewhere, where those bits wrap (seen on the B200: one
    frame of 16,384 failed).  The emulator moves the wrap to many places inside the image."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        payload = int(op.c_off[-1])
        for wrap_at in list(range(16, payload, max(16, payload // 4))) + [payload - 7]:   # ~5 full decodes per file keep the CPU suite short
            out, status = emu_api.decode_all(emu, c["image"], op.codec, op.c_off, op.d_off, ctas=2, wrap_at=wrap_at)
            assert (status == 0).all(), (wrap_at, status)
            assert hashlib.sha256(out.tobytes()).hexdigest() == c["input_sha256"], wrap_at


@pytest.mark.parametrize("misalign", [1, 15, 17])
@pytest.mark.parametrize("name", ["zsyn_zstd3_128k", "mix_zstd19"])
def test_emulated_zstd_pipeline_unaligned_output(emu, golden, name, misalign):
    """The executor's ring is flushed in 16-byte vectors aligned on the GLOBAL address; heads and tails go bytewise."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        out, status = emu_api.decode_all(emu, c["image"], op.codec, op.c_off, op.d_off, ctas=2, misalign=misalign)
    assert (status == 0).all(), status
    assert hashlib.sha256(out.tobytes()).hexdigest() == c["input_sha256"]


LZ4_CASES = ["tiny_lz4", "zsyn_lz4_64k", "zsyn_lz4_256k_linked", "mix_lz4"]


LANE_KERNELS = [102]   # the lane-per-frame kernel (launches with many frames)


@pytest.mark.parametrize("kernel", LANE_KERNELS)
@pytest.mark.parametrize("name", LZ4_CASES)
def test_emulated_lane_lz4_kernel(emu, golden, name, kernel):
    """The lane-per-frame LZ4 kernels (taken for launches with many frames) decode every golden LZ4 file."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        out, status = emu_api.decode_all(emu, c["image"], kernel, op.c_off, op.d_off, ctas=1)
    assert (status == 0).all(), status
    assert hashlib.sha256(out.tobytes()).hexdigest() == c["input_sha256"]


@pytest.mark.parametrize("kernel", LANE_KERNELS)
@pytest.mark.parametrize("misalign", [1, 7, 15, 17, 31])
def test_emulated_lane_lz4_kernel_unaligned_output(emu, golden, misalign, kernel):
    """Frames whose output does not start on a 16-byte boundary: the head and tail go bytewise and neighbours stay intact."""
    cases, _ = golden
    c = cases["zsyn_lz4_4k_chunks"]
    with OraclePort(c["image"]) as op:
        out, status = emu_api.decode_all(emu, c["image"], kernel, op.c_off, op.d_off, ctas=1, misalign=misalign)
    assert (status == 0).all(), status
    assert hashlib.sha256(out.tobytes()).hexdigest() == c["input_sha256"]


@pytest.mark.parametrize("name,codec", [("zsyn_lz4_64k", None), ("zsyn_lz4_64k", 102), ("zsyn_zstd3_128k", None),
                                        ("zsyn_zstd3_128k", 200), ("zsyn_zstd3_128k", 201)])
def test_emulated_decode_flags_corrupt_frames(emu, golden, name, codec):
    """Truncation and bit flips must end in a non-zero status, never a hang or an out-of-bounds write."""
    cases, _ = golden
    if True:
        c = cases[name]
        with OraclePort(c["image"]) as op:
            img = bytearray(c["image"])
            c0, c1 = int(op.c_off[1]), int(op.c_off[2])
            for k in range(c0 + 20, c1, 997):  # sprinkle corruption over frame 1 only
                img[k] ^= 0x55
            out, status = emu_api.decode_all(emu, bytes(img), codec or op.codec, op.c_off, op.d_off, ctas=2)
            good = op.decode_all()
        assert status[0] == 0 and (status[2:] == 0).all()
        d0, d1 = int(op.d_off[1]), int(op.d_off[2])
        assert (out[:d0] == good[:d0]).all() and (out[d1:] == good[d1:]).all()
        # either the frame was rejected or (vanishingly unlikely) it decoded to other bytes of the right size
        assert status[1] != 0 or not (out[d0:d1] == good[d0:d1]).all()


def test_emulated_lookup_and_gather(emu, golden):
    cases, _ = golden
    c = cases["mix_zstd3"]
    with OraclePort(c["image"]) as op:
        decoded = op.decode_all()
        d_off = np.ascontiguousarray(op.d_off, dtype=np.uint64)
        n_frames = op.frames
        rng = np.random.Generator(np.random.PCG64(3))
        offsets = np.concatenate([rng.integers(0, op.size + 100, 500), d_off.astype(np.int64), d_off.astype(np.int64)[1:] - 1,
                                  [op.size, op.size + 10 ** 9]]).astype(np.uint64)
        counts = rng.choice([0, 1, 100, 4096, 70000], offsets.size).astype(np.uint64)
        n = offsets.size
        frame = np.zeros(n, np.int32)
        inframe = np.zeros(n, np.uint32)
        nbytes = np.zeros(n, np.uint32)
        touched = np.zeros(n_frames + 1, np.uint32)
        emu.emu_lookup(d_off.ctypes.data, n_frames, offsets.ctypes.data, counts.ctypes.data, 0, n, frame.ctypes.data,
                       inframe.ctypes.data, nbytes.ctypes.data, touched.ctypes.data)
        for i in range(n):
            assert frame[i] == op.offset_to_frame(int(offsets[i]))
            r, b = op.pread(int(counts[i]), int(offsets[i]))
            assert nbytes[i] == r
        # gather out of a "cache" that simply holds the whole decoded file
        stride = 70000
        src = np.zeros(16 + decoded.size + 64, np.uint8)
        src[16:16 + decoded.size] = decoded
        frame_src = d_off[:-1].astype(np.int64)
        dst = np.zeros(n * stride + 64, np.uint8)
        emu.emu_gather(frame.ctypes.data, inframe.ctypes.data, nbytes.ctypes.data, frame_src.ctypes.data,
                       src.ctypes.data + 16, dst.ctypes.data, None, stride, n)
        for i in range(n):
            o, k = int(offsets[i]), int(nbytes[i])
            assert (dst[i * stride:i * stride + k] == decoded[o:o + k]).all()


@pytest.mark.parametrize("name,codec,limits", [("zsyn_lz4_64k", 1, True), ("mix_lz4", 102, True), ("zsyn_lz4_256k_linked", 102, False),
                                               ("zsyn_zstd3_128k", None, True), ("mix_zstd19", None, False)])
def test_emulated_stream_ordered_batch_chain(emu, golden, name, codec, limits):
    """n1 on the CPU: the kernels of zseek_b200_pread_batch_async in the order the product queues them — lookup, compaction
    of the touched frames into a job list whose COUNT stays in device memory, a decode launch that reads that count
    (njobs is only an upper bound) and stops every frame at the last byte the batch needs of it, gather + results.
    Every request must come out as the oracle's zseek_pread returns it; nothing outside the requested bytes is written."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        rng = np.random.Generator(np.random.PCG64(11))
        d_off = np.asarray(op.d_off, dtype=np.int64)
        offsets = np.concatenate([rng.integers(0, op.size + 50, 220), d_off[1:] - rng.integers(1, 300, op.frames), [op.size, op.size + 10 ** 9]]).astype(np.uint64)
        counts = rng.choice([0, 1, 100, 4096], offsets.size).astype(np.uint64)
        stride = 4096
        dst, res, status, njobs, err = emu_api.batch(emu, c["image"], codec if codec is not None else op.codec, op.c_off, op.d_off,
                                                     offsets, counts, stride, limits=limits)
        assert err == 0
        touched = {op.offset_to_frame(int(o)) for o, k in zip(offsets, counts) if op.pread(int(k), int(o))[0] > 0 or op.offset_to_frame(int(o)) >= 0}
        touched.discard(-1)
        assert njobs == len(touched)
        assert (status[:njobs] == 0).all(), status[:njobs]
        for i in range(offsets.size):
            r, b = op.pread(int(counts[i]), int(offsets[i]))
            assert res[i] == r, i
            assert dst[i * stride:i * stride + r].tobytes() == b, i
            assert (dst[i * stride + r:(i + 1) * stride] == 0x5A).all(), i
        # a request outside the shard and a slab that is too small are flagged by the compaction, not silently dropped
        assert emu_api.batch(emu, c["image"], codec if codec is not None else op.codec, op.c_off, op.d_off, offsets, counts, stride,
                             shard=(0, max(1, op.frames // 2)), max_jobs=0)[4] == (1 if op.frames > 1 else 0)
        if len(touched) > 1:
            assert emu_api.batch(emu, c["image"], codec if codec is not None else op.codec, op.c_off, op.d_off, offsets, counts, stride,
                                 max_jobs=1)[4] == 2


def _shapes_corpus():
    """Every LZ4 sequence shape the lane kernel spreads over several trips: periodic runs of period 1..40 (overlapping
    matches), long zero runs, incompressible stretches (long literal runs, raw blocks), long far matches."""
    rng = np.random.Generator(np.random.PCG64(42))
    parts = []
    text = rng.integers(97, 123, 3000, dtype=np.uint8).tobytes()
    for period in list(range(1, 41)) + [63, 64, 65, 191, 192, 193, 255, 256, 257]:
        pat = rng.integers(0, 256, period, dtype=np.uint8).tobytes()
        parts.append(pat * (rng.integers(2, 300) // 1) + text[:int(rng.integers(0, 40))])
    parts.append(bytes(70000))
    parts.append(rng.integers(0, 256, 80000, dtype=np.uint8).tobytes())
    parts.append(text * 20)
    for _ in range(200):   # far matches of assorted lengths separated by short literal runs
        o = int(rng.integers(0, len(text) - 600))
        parts.append(text[o:o + int(rng.integers(4, 600))] + rng.integers(0, 256, int(rng.integers(0, 30)), dtype=np.uint8).tobytes())
    return b"".join(parts)


@pytest.mark.parametrize("kw", [dict(), dict(block_checksum=True, content_checksum=True), dict(block_size_id=5, level=9),
                                dict(content_size=False, independent=True)])
def test_emulated_lane_lz4_kernel_sequence_shapes(emu, kw):
    from datagen import foreign
    from oracle.pyapi import have_reference
    if not have_reference():
        pytest.skip("liblz4 frames are built through oracle/_ref's codec libraries")
    data = _shapes_corpus()
    image = foreign.build(data, 100000, "lz4", **kw)
    with OraclePort(image) as op:
        assert op.decode_all().tobytes() == data
        for codec, mis in ((102, 0), (102, 3), (1, 0), (1, 5)):
            out, status = emu_api.decode_all(emu, image, codec, op.c_off, op.d_off, ctas=1, misalign=mis)
            assert (status == 0).all(), status
            assert out.tobytes() == data


@pytest.mark.parametrize("name,codec", [("zsyn_lz4_64k", 1), ("zsyn_lz4_64k", 102), ("zsyn_lz4_256k_linked", 102), ("mix_lz4", 1),
                                        ("zsyn_zstd3_128k", None), ("zsyn_zstd19_256k", None), ("mix_zstd3", None),
                                        ("zsyn_zstd3_128k", 200), ("mix_zstd3", 201)])
def test_emulated_decode_stops_at_job_limits(emu, golden, name, codec):
    """Batches of small reads decode a frame only up to the last byte they need (zsk_decode_args.limits): the prefix
    must be exact, the status OK, and nothing may be written outside the frame."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        good = op.decode_all()
        sizes = np.diff(op.d_off.astype(np.int64))
        rng = np.random.Generator(np.random.PCG64(9))
        limits = np.array([int(rng.integers(1, max(2, s + 1))) for s in sizes], dtype=np.uint32)
        limits[::5] = 1
        limits[1::7] = sizes[1::7]          # exactly the frame size
        limits[2::9] = 0xFFFFFFFF           # no limit
        out, status = emu_api.decode_all(emu, c["image"], codec or op.codec, op.c_off, op.d_off, ctas=2, limits=limits)
        assert (status == 0).all(), status
        for f, s in enumerate(sizes):
            d0, k = int(op.d_off[f]), int(min(int(limits[f]), s))
            assert (out[d0:d0 + k] == good[d0:d0 + k]).all(), f


def _checksum_cases():
    from datagen import foreign, zsyn
    rng = np.random.Generator(np.random.PCG64(21))
    data = zsyn.gen(90000, seed=3) + bytes(3000) + rng.integers(0, 256, 30000, dtype=np.uint8).tobytes() + b"xyz"
    yield "lz4", data, foreign.build(data, 40000, "lz4", block_checksum=True, content_checksum=True)
    yield "lz4_no_content_size", data, foreign.build(data, 40000, "lz4", block_checksum=True, content_checksum=True, content_size=False)
    yield "zstd", data, foreign.build(data, 40000, "zstd", checksum=True)


def test_emulated_kernels_verify_checksums_like_the_reference(emu):
    """liblz4 verifies header, block and content checksums and libzstd the content checksum, so the reference's cached
    zseek_pread fails on a mismatch; the kernels must accept the intact files and reject exactly the frames the
    reference rejects (probe and expectations come from the reference itself)."""
    from oracle.pyapi import RefReader, have_reference
    if not have_reference():
        pytest.skip("needs oracle/_ref (the reference build) for the expected verdicts")
    for name, data, good in _checksum_cases():
        with OraclePort(good) as op:
            c_off, d_off, codec = op.c_off.copy(), op.d_off.copy(), op.codec
        kernels = (1, 102) if name.startswith("lz4") else (codec,)
        c0, c1, c2 = int(c_off[0]), int(c_off[1]), int(c_off[2])
        flips = {"intact": None, "content checksum of frame 0": c1 - 1, "payload of frame 1": c1 + (c2 - c1) // 2,
                 "frame header of frame 1": c1 + 5}
        for what, pos in flips.items():
            img = bytearray(good)
            if pos is not None:
                img[pos] ^= 0x04
            want_fail = []
            for f in range(len(c_off) - 1):
                # a fresh reader per frame: after a failed frame the reference's LZ4F context stays poisoned and every later
                # frame fails too (it is not reset on the cached path); the cached path decodes whole frames, like the kernels
                with RefReader(bytes(img), cache_size=2) as rr:
                    try:
                        rr.pread(10, int(d_off[f]))
                        want_fail.append(False)
                    except OSError:
                        want_fail.append(True)
            for k in kernels:
                out, status = emu_api.decode_all(emu, bytes(img), k, c_off, d_off, ctas=1)
                assert [bool(s) for s in status] == want_fail, (name, what, k, status)
                if pos is None:
                    assert out.tobytes() == data


@pytest.mark.parametrize("seed", [1, 2])
def test_emulated_lz4_kernels_on_random_mixtures(emu, seed):
    """Property check (decode == the writer's input) on randomly assembled data: runs of random length drawn from text,
    zeros, noise, short periods and far repeats, compressed by the reference writer at random frame sizes and levels; every
    shipped LZ4 kernel, aligned and unaligned destinations."""
    from datagen import refwriter, zsyn
    from oracle.pyapi import LZ4, have_reference
    if not have_reference():
        pytest.skip("inputs come from the reference writer (oracle/_ref)")
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    text = zsyn.gen(200000, seed=seed)
    parts, total = [], 0
    while total < 260000:
        kind, n = int(rng.integers(0, 6)), int(rng.integers(1, 9000))
        if kind == 0:
            o = int(rng.integers(0, len(text) - n)); p = text[o:o + n]
        elif kind == 1:
            p = bytes(n)
        elif kind == 2:
            p = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 3:
            per = rng.integers(0, 256, int(rng.integers(1, 24)), dtype=np.uint8).tobytes(); p = (per * (n // len(per) + 1))[:n]
        elif kind == 4 and parts:
            src = b"".join(parts)[-70000:]; o = int(rng.integers(0, max(1, len(src) - 1))); p = src[o:o + n]
        else:
            p = rng.integers(97, 101, n, dtype=np.uint8).tobytes()
        parts.append(p); total += len(p)
    data = b"".join(parts)
    frame = int(rng.choice([4096, 30000, 65536, 150000]))
    image = refwriter.write(data, LZ4, int(rng.choice([0, 3, 9])), frame, int(rng.choice([frame, 4093])))
    with OraclePort(image) as op:
        assert op.decode_all().tobytes() == data
        for codec, mis in ((1, 0), (102, 0), (102, 13), (1, 7)):
            out, status = emu_api.decode_all(emu, image, codec, op.c_off, op.d_off, ctas=1, misalign=mis)
            assert (status == 0).all(), (codec, status)
            assert out.tobytes() == data, codec


@pytest.mark.parametrize("seed,level", [(1, 1), (2, 3), (3, 19), (4, -3)])
def test_emulated_zstd_kernel_on_random_mixtures(emu, seed, level):
    """Same property for the zstd kernel (raw / RLE / compressed blocks, predefined / RLE / compressed / repeat tables, 1- and
    4-stream and treeless literals all occur in such mixtures), with and without per-job limits."""
    from datagen import refwriter, zsyn
    from oracle.pyapi import ZSTD, have_reference
    if not have_reference():
        pytest.skip("inputs come from the reference writer (oracle/_ref)")
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    text = zsyn.gen(120000, seed=seed)
    parts, total = [], 0
    while total < 180000:
        kind, n = int(rng.integers(0, 5)), int(rng.integers(1, 12000))
        if kind == 0:
            o = int(rng.integers(0, len(text) - n)); p = text[o:o + n]
        elif kind == 1:
            p = bytes([int(rng.integers(0, 256))]) * n
        elif kind == 2:
            p = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        elif kind == 3:
            per = rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes(); p = (per * (n // len(per) + 1))[:n]
        else:
            p = rng.integers(97, 100, n, dtype=np.uint8).tobytes()
        parts.append(p); total += len(p)
    data = b"".join(parts)
    frame = int(rng.choice([20000, 70000, 140000]))
    image = refwriter.write(data, ZSTD, level, frame, int(rng.choice([frame, 4093])))
    with OraclePort(image) as op:
        good = op.decode_all()
        assert good.tobytes() == data
        out, status = emu_api.decode_all(emu, image, op.codec, op.c_off, op.d_off, ctas=2)
        assert (status == 0).all(), status
        assert out.tobytes() == data
        sizes = np.diff(op.d_off.astype(np.int64))
        limits = np.array([int(rng.integers(1, s + 1)) for s in sizes], dtype=np.uint32)
        out, status = emu_api.decode_all(emu, image, op.codec, op.c_off, op.d_off, ctas=2, limits=limits)
        assert (status == 0).all(), status
        for f, s in enumerate(sizes):
            d0 = int(op.d_off[f])
            assert (out[d0:d0 + int(limits[f])] == good[d0:d0 + int(limits[f])]).all(), f
