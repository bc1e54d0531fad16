"""Host logic of the reader (libzseek_b200/csrc/reader.c) without a GPU.

tests/emu/libzsk_hostemu.so = the UNMODIFIED reader.c linked against a stand-in launch layer (tests/emu/hostemu.cpp) that
keeps "device" memory on the host and runs the product's kernel sources on the lock-step emulator.  TEST INFRASTRUCTURE
ONLY — not a fallback, never named by the product (test_abi.py::test_product_never_touches_the_oracle).  The scenarios
are the GPU parity tests themselves (tests/test_gpu_parity.py, the ones that only use host buffers), called here with the
emulated library on the small golden files, plus cache / parking checks through "device" buffers of the stand-in.  What
this adds to the kernel-level emulator tests: seek table, LRU cache, read-ahead windows and their failure handling, batch
bookkeeping, callbacks and I/O errors — the reference components src/decompress.c, src/cache.c, src/buffer.c replaced by
reader.c — are checked against the reference's golden vectors in the no-GPU container."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import emu_api
import test_gpu_parity as G
from oracle.pyapi import ZSTD, OraclePort, have_reference

HOSTEMU = os.path.join(emu_api.EMU_DIR, "libzsk_hostemu.so")


@pytest.fixture(scope="module")
def hostlib():
    """The libzseek_b200 Python mirror bound to the emulated library for the duration of this module."""
    subprocess.run(["make", "-C", emu_api.EMU_DIR, "-s", "libzsk_hostemu.so"], check=True, stderr=subprocess.DEVNULL)
    import libzseek_b200 as z
    from libzseek_b200 import reader as R
    saved = (R._lib, R.LIB_PATH)
    R._lib, R.LIB_PATH = None, HOSTEMU
    try:
        L = z.load_library()
        L.hostemu_device_alloc.restype = C.c_void_p
        L.hostemu_device_alloc.argtypes = [C.c_size_t]
        L.hostemu_device_free.argtypes = [C.c_void_p]
        L.hostemu_device_allocs.restype = C.c_size_t
        yield z
    finally:
        R._lib, R.LIB_PATH = saved


GOLDEN_RUNS = [("tiny_zstd", 0), ("tiny_zstd", 1), ("tiny_lz4", 0), ("tiny_lz4", 4), ("zsyn_zstd3_128k", 0), ("zsyn_zstd19_256k", 4)]


@pytest.mark.parametrize("name,cache_size", GOLDEN_RUNS)
def test_golden_preads_through_the_host_code(hostlib, golden, name, cache_size):
    """every golden zseek_pread call (return value + bytes of the reference reader), stats, whole-file scan"""
    G.test_pread_matches_reference_golden(hostlib, golden, name, cache_size, None)


def test_read_cursor_callbacks_and_io_errors(hostlib, golden):
    G.test_python_callbacks_reader(hostlib, golden, None)
    G.test_io_errors_surface_like_the_reference(hostlib, golden, None)
    # zseek_read: the cursor advances by what each call returned (reference src/decompress.c:826-835)
    cases, _ = golden
    c = cases["zsyn_zstd3_256k_chunks"]
    with OraclePort(c["image"]) as op:
        want = op.decode_all().tobytes()
    with hostlib.Reader(image=c["image"], cache_size=2) as rd:
        got = bytearray()
        while True:
            r, b = rd.read(70001)
            if r == 0:
                break
            got += b
    assert bytes(got) == want


def test_scan_over_a_file_handle(hostlib, golden):
    """reference test/example.c:36-87 over a FILE* (default I/O callbacks), 4 KiB preads, cache_size 1"""
    G.test_example_scan_4k_cache1(hostlib, golden, "mix_zstd19", None)


def test_host_batch_semantics(hostlib, golden):
    """zseek_b200_pread_batch into host memory stores exactly what a loop of zseek_pread stores (sentinel bytes survive)"""
    G.test_host_batch_leaves_unproduced_bytes_untouched(hostlib, golden, None, names=("zsyn_zstd3_128k",))


def test_corrupt_frames(hostlib, golden):
    G.test_corrupt_frame_fails_instead_of_hanging(hostlib, golden, None)


@pytest.mark.skipif(not have_reference(), reason="verdicts come from oracle/_ref")
def test_read_ahead_window_with_a_bad_frame(hostlib, golden):
    """a bad frame inside a read-ahead window fails only the reads of that frame (reference src/decompress.c:700-790)"""
    G.test_sequential_scan_across_a_corrupt_frame(hostlib, golden, None, names=("mix_zstd3",), cache_sizes=(0,))


def test_shard_limits(hostlib, golden):
    G.test_shard_restricts_frames(hostlib, golden, None)


def test_no_device_means_no_reader(hostlib):
    """the open fails loudly when the launch layer finds no device — there is nothing to fall back to (a fresh process:
    readers parked by the tests above would bring their context along)"""
    import sys
    code = ("import libzseek_b200 as z, sys\n"
            "img = open(sys.argv[1], 'rb').read()\n"
            "try:\n    z.Reader(image=img)\n    print('opened')\n"
            "except z.ZseekError as e:\n    print(e)\n")
    env = dict(os.environ, ZSEEK_B200_LIB=HOSTEMU, ZSK_HOSTEMU_NO_DEVICE="1", PYTHONPATH=G.ROOT if hasattr(G, "ROOT") else os.path.dirname(emu_api.HERE))
    out = subprocess.run([sys.executable, "-c", code, os.path.join(G.GOLDEN, "tiny_lz4.zsk")], env=env, capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "context creation failed: no CUDA device: emulated absence", (out.stdout, out.stderr[-300:])


@pytest.mark.skipif(not have_reference(), reason="inputs come from the reference writer (oracle/_ref)")
def test_cache_capacity_and_lru_order(hostlib):
    """reference test/test_cache.c:135-159 restated for the HBM cache: capacity honoured, least recently used evicted,
    a hit promotes (the same scenario as the GPU test, through a "device" buffer of the stand-in)."""
    from datagen import refwriter, zsyn
    data = zsyn.gen(140 * 8192, seed=5)
    image = refwriter.write(data, ZSTD, 1, 8192)
    cap = 66
    L = hostlib.load_library()
    buf = L.hostemu_device_alloc(8192)
    try:
        with hostlib.Reader(image=image, cache_size=cap) as rd:
            assert rd.frames == 140
            rd.load(0, rd.frames)

            def read(f):
                l0 = rd.launch_count
                assert rd.pread_into(buf, 100, f * 8192 + 7) == 100
                assert C.string_at(buf, 100) == data[f * 8192 + 7:f * 8192 + 107]
                return rd.launch_count > l0       # True = a kernel ran = miss

            order = list(range(2 * cap, -1, -2))[:cap + 1]        # cap + 1 distinct frames, never sequential
            assert all(read(f) for f in order[:cap])
            assert rd.stats().cached_frames == cap
            assert not read(order[0])                             # hit: promoted to most recently used
            assert read(order[cap])                               # one more frame evicts order[1]
            assert rd.stats().cached_frames == cap
            assert not read(order[0]) and not read(order[cap])
            for f in order[2:cap]:
                assert not read(f), f
            assert read(order[1])                                 # the victim is gone
            assert rd.stats().cached_frames == cap
    finally:
        L.hostemu_device_free(buf)


def test_closed_readers_are_parked_and_reused(hostlib, golden):
    """zseek_reader_close keeps the device side for the next open: after the first open / read / close cycle, later cycles
    over other files allocate (almost) nothing new, and every reader still returns the right bytes."""
    cases, _ = golden
    L = hostlib.load_library()
    names = ["tiny_zstd", "tiny_lz4", "zsyn_zstd3_256k_chunks", "tiny_zstd", "tiny_lz4"]
    allocs = []
    for name in names:
        c = cases[name]
        with OraclePort(c["image"]) as op:
            want = op.decode_all().tobytes()
        with hostlib.Reader(image=c["image"], cache_size=0) as rd:
            n = min(5000, len(want))
            assert rd.pread(n, 0)[1] == want[:n]
            assert rd.read_range(len(want), 0) == want
        allocs.append(int(L.hostemu_device_allocs()))
    first = allocs[2]                      # every kind of file has been seen once
    assert allocs[-1] - first <= 4, allocs  # re-opening them costs next to nothing


class _DevBuf:
    """a "device" buffer of the stand-in with the few tensor methods the GPU tests use"""

    def __init__(self, L, n, keep=None, addr=None):
        self.L, self.n = L, n
        self.addr = addr if addr is not None else L.hostemu_device_alloc(max(n, 1))
        self.keep = keep       # a view keeps its parent alive

    def data_ptr(self):
        return self.addr

    def numel(self):
        return self.n

    def __getitem__(self, s):
        lo, hi, _ = s.indices(self.n)
        return _DevBuf(self.L, hi - lo, keep=self, addr=self.addr + lo)

    def cpu(self):
        return self

    def numpy(self):
        return np.frombuffer(C.string_at(self.addr, self.n), dtype=np.uint8)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.n
        C.memmove(self.addr, arr.ctypes.data, arr.nbytes)
        return self

    def __del__(self):
        if self.keep is None and self.addr:
            self.L.hostemu_device_free(self.addr)


class _FakeTorch:
    uint8 = "uint8"

    def __init__(self, L):
        self.L = L

    def empty(self, n, dtype=None, device=None):
        return _DevBuf(self.L, n)


@pytest.mark.parametrize("mode", ["pinned", "hbm"])
def test_random_host_readers_go_resident(hostlib, golden, mode, monkeypatch):
    """isolated misses -> the whole shard is decoded once, into a pinned window or (pinned budget spent) into HBM; reads from
    six threads, a device destination, and the way back through cache_clear — the GPU test's scenario on the stand-in"""
    G.test_random_host_readers_go_resident(hostlib, golden, "zsyn_zstd3_128k", mode, monkeypatch, _FakeTorch(hostlib.load_library()))


@pytest.mark.parametrize("name", ["mix_zstd19"])
def test_stream_ordered_batch_through_the_host_code(hostlib, golden, name):
    """zseek_b200_pread_batch_async with every array in "device" memory: results and bytes are the oracle's zseek_pread,
    a second batch reuses the buffers, other entry points may run before batch_wait, a corrupt frame surfaces in
    batch_wait, and a shard that is not resident is refused."""
    cases, _ = golden
    c = cases[name]
    L = hostlib.load_library()
    with OraclePort(c["image"]) as op:
        want = op.decode_all()
        rng = np.random.Generator(np.random.PCG64(8))
        with hostlib.Reader(image=c["image"], cache_size=0) as rd:
            n, stride = 300, 4096
            d_offs, d_counts = _DevBuf(L, 8 * n), _DevBuf(L, 8 * n)
            d_res, dst = _DevBuf(L, 8 * n), _DevBuf(L, n * stride)
            with pytest.raises(hostlib.ZseekError) as e:       # nothing resident yet
                rd.pread_batch_async(d_offs, dst, fixed_count=100, dst_stride=stride)
            assert "zseek_b200_load" in str(e.value)
            rd.load(0, rd.frames)
            for rnd in range(2):
                offs = rng.integers(0, op.size + 100, n).astype(np.uint64)
                counts = rng.choice([0, 1, 700, 4096], n).astype(np.uint64)
                d_offs.upload(offs), d_counts.upload(counts)
                dst.upload(np.full(n * stride, 0x5A, dtype=np.uint8))
                d_res.upload(np.full(n, -7, dtype=np.int64))
                # numel() of the offsets buffer is the request count for the Python mirror: hand it a view of n "elements"
                view = _DevBuf(L, n, keep=d_offs, addr=d_offs.addr)
                rd.pread_batch_async(view, dst, dev_counts=d_counts, dst_stride=stride, dev_results=d_res)
                if rnd == 1:                                     # another entry point before the wait
                    assert rd.pread(100, 5)[1] == want[5:105].tobytes()
                rd.batch_wait()
                res = np.frombuffer(C.string_at(d_res.addr, 8 * n), dtype=np.int64)
                out = dst.numpy()
                for i in range(n):
                    r, b = op.pread(int(counts[i]), int(offs[i]))
                    assert res[i] == r, (rnd, i)
                    assert out[i * stride:i * stride + r].tobytes() == b, (rnd, i)
                    assert (out[i * stride + r:(i + 1) * stride] == 0x5A).all(), (rnd, i)
        # corrupt frame 1: the call itself succeeds, batch_wait reports it, good frames are served
        img = bytearray(c["image"])
        for k in range(int(op.c_off[1]) + 12, int(op.c_off[2]), 5):
            img[k] ^= 0x3C
        with hostlib.Reader(image=bytes(img), cache_size=0) as rd:
            rd.load(0, rd.frames)
            three = (op.d_off[:3].astype(np.int64) + 100).astype(np.uint64)
            d3, dst3 = _DevBuf(L, 24).upload(three), _DevBuf(L, 3 * 4096)
            rd.pread_batch_async(_DevBuf(L, 3, keep=d3, addr=d3.addr), dst3, fixed_count=4096, dst_stride=4096)
            with pytest.raises(hostlib.ZseekError) as e:
                rd.batch_wait()
            assert str(e.value).startswith("decompress frame")
            k = min(4096, int(op.d_off[1]) - 100)
            assert dst3.numpy()[:k].tobytes() == want[100:100 + k].tobytes()


@pytest.mark.skipif(not have_reference(), reason="inputs come from the reference writer (oracle/_ref)")
def test_scans_over_many_small_frames(hostlib):
    """A forward scan whose read-ahead window grows (1, 8, 64 frames ...) while the next window is decoded behind the
    caller's back, then reads that jump backwards (the window collapses, cached frames are served from HBM): every byte as
    written, short reads exactly at the frame ends (B1)."""
    from datagen import refwriter, zsyn
    data = zsyn.gen(120 * 8192 + 1234, seed=9)
    image = refwriter.write(data, ZSTD, 1, 8192)
    with hostlib.Reader(image=image, cache_size=1) as rd:
        assert rd.frames == 121
        off, calls = 0, 0
        while off < len(data):
            r, b = rd.pread(3000, off)
            want_r = min(3000, (off // 8192 + 1) * 8192 - off, len(data) - off)
            assert r == want_r and b == data[off:off + r], off
            off += r
            calls += 1
        assert rd.pread(3000, off) == (0, b"")
        launches_forward = rd.launch_count
        assert launches_forward < 40, launches_forward           # windows, not one launch per frame
        for f in range(118, 60, -7):                             # backwards, never sequential
            r, b = rd.pread(8192, f * 8192 + 100)
            assert r == 8092 and b == data[f * 8192 + 100:(f + 1) * 8192]


@pytest.mark.skipif(not have_reference(), reason="inputs come from the reference writer (oracle/_ref)")
def test_host_range_read_in_many_pipeline_stages(hostlib, monkeypatch):
    """zseek_b200_read_range into host memory is a pipeline of H2D / decode / D2H stages over chunks of frames; with the
    stage size forced down to one frame the chunk arithmetic (staging halves, offsets, the ragged tail, a start in the
    middle of a frame) is walked ~40 times instead of once."""
    from datagen import refwriter, zsyn
    data = zsyn.gen(40 * 8192 + 777, seed=3)
    image = refwriter.write(data, ZSTD, 1, 8192)
    monkeypatch.setenv("ZSEEK_B200_CHUNK_MB", "0")
    monkeypatch.setenv("ZSEEK_B200_RAMP_MB", "0")
    with hostlib.Reader(image=image, cache_size=0) as rd:
        assert rd.read_range(len(data) + 50, 0) == data                      # short only at EOF
        assert rd.read_range(100000, 12345) == data[12345:112345]
        assert rd.read_range(10, len(data)) == b""


FUZZ_STEPS = int(os.environ.get("ZSK_HOSTEMU_FUZZ_STEPS", "80"))          # soak runs: more steps, more seeds
FUZZ_RUNS = [(1, 0), (2, 3)] + [(s, [0, 1, 5, 70][s % 4]) for s in range(3, 3 + int(os.environ.get("ZSK_HOSTEMU_FUZZ_EXTRA_SEEDS", "0")))]


@pytest.mark.skipif(not have_reference(), reason="inputs come from the reference writer (oracle/_ref)")
@pytest.mark.parametrize("seed,cache_size", FUZZ_RUNS)
def test_random_sequences_of_calls_against_the_model(hostlib, monkeypatch, seed, cache_size):
    """Differential state-machine test of reader.c: a sequence of randomly chosen calls on ONE reader — zseek_pread /
    zseek_read with host and device buffers, multi-frame ranges, host and device batches, stream-ordered batches (with
    other calls before the wait), cache_clear, unload / load, shard changes, with residency kicking in on the way — each
    checked against the model the reference defines (B1-B4 over the writer's input).  What is tested is the interplay of
    the cache (prefix-valid entries of partially decoded frames), windows, residency and shards."""
    from datagen import refwriter, zsyn
    monkeypatch.setenv("ZSEEK_B200_RESIDENT_AFTER", "9")
    data = zsyn.gen(36 * 8192 + 4321, seed=40 + seed)
    image = refwriter.write(data, ZSTD, 1, 8192) if os.environ.get("ZSK_HOSTEMU_FUZZ_CODEC", "zstd") == "zstd" else refwriter.write(data, 1, 0, 8192)
    total, F = len(data), 8192
    L = hostlib.load_library()
    rng = np.random.Generator(np.random.PCG64(seed))

    def model(off, cnt, lo, hi):
        """(result, bytes) of zseek_pread, or None when the frame is outside the shard [lo, hi)"""
        if off >= total or cnt == 0:
            return 0, b""
        f = off // F
        if not lo <= f < hi:
            return None
        n = min(cnt, min((f + 1) * F, total) - off)
        return n, data[off:off + n]

    with hostlib.Reader(image=image, cache_size=cache_size) as rd:
        nfr = rd.frames
        lo, hi, cursor = 0, nfr, 0
        dev = _DevBuf(L, 1 << 20)
        log = []
        for step in range(FUZZ_STEPS):
            b_lo, b_hi = lo * F, min(hi * F, total)
            op = rng.choice(["pread", "pread", "pread", "pread_dev", "read", "range", "range_dev", "batch", "batch_dev", "async",
                             "clear", "unload", "load", "shard", "stats", "frames"])
            log.append(op)
            try:
                if op in ("pread", "pread_dev"):
                    off = int(rng.integers(0, total + 100)) if rng.random() < 0.3 else int(rng.integers(b_lo, b_hi))
                    cnt = int(rng.choice([0, 1, 100, 3000, 8192, 20000]))
                    want = model(off, cnt, lo, hi)
                    if want is None:
                        with pytest.raises(hostlib.ZseekError) as e:
                            rd.pread(cnt, off)
                        assert "shard" in str(e.value)
                    elif op == "pread":
                        assert rd.pread(cnt, off) == want
                    else:
                        assert rd.pread_into(dev, cnt, off) == want[0] and dev.numpy()[:want[0]].tobytes() == want[1]
                elif op == "read":
                    cnt = int(rng.choice([1, 500, 9000]))
                    want = model(cursor, cnt, lo, hi)
                    if want is None:
                        with pytest.raises(hostlib.ZseekError):
                            rd.read(cnt)
                    else:
                        assert rd.read(cnt) == want
                        cursor += want[0]
                elif op in ("range", "range_dev"):
                    off = int(rng.integers(b_lo, b_hi))
                    cnt = int(rng.integers(0, min(60000, b_hi - off) + 1))
                    if op == "range":
                        assert rd.read_range(cnt, off) == data[off:off + cnt]
                    else:
                        assert rd.read_range_into(dev, cnt, off) == cnt and dev.numpy()[:cnt].tobytes() == data[off:off + cnt]
                elif op in ("batch", "batch_dev", "async"):
                    n = int(rng.integers(1, 60))
                    offs = rng.integers(b_lo, b_hi, n).astype(np.uint64)
                    if rng.random() < 0.3:
                        offs[0] = total + 5                                                   # EOF inside a batch
                    counts = rng.choice([0, 1, 700, 4096], n).astype(np.uint64)
                    stride = 4200
                    if op == "batch":
                        dst = np.full(n * stride, 0x5A, dtype=np.uint8)
                        res = rd.pread_batch(offs, counts=counts, dst=dst, dst_stride=stride)
                        out = dst
                    else:
                        dev.upload(np.full(n * stride, 0x5A, dtype=np.uint8))
                        if op == "batch_dev":
                            res = rd.pread_batch(offs, counts=counts, dst=dev, dst_stride=stride)
                        else:
                            if not (lo == 0 and hi == nfr) or rng.random() < 0.5:
                                rd.load(lo, hi)                                               # the shard must be resident
                            else:
                                rd.load(0, nfr)
                            d_offs, d_counts, d_res = _DevBuf(L, 8 * n).upload(offs), _DevBuf(L, 8 * n).upload(counts), _DevBuf(L, 8 * n)
                            rd.pread_batch_async(_DevBuf(L, n, keep=d_offs, addr=d_offs.addr), dev, dev_counts=d_counts, dst_stride=stride,
                                                 dev_results=d_res)
                            if rng.random() < 0.5:                                            # another call before the wait
                                o2 = int(rng.integers(b_lo, b_hi))
                                assert rd.pread(50, o2) == model(o2, 50, lo, hi)
                            rd.batch_wait()
                            res = np.frombuffer(C.string_at(d_res.addr, 8 * n), dtype=np.int64)
                        out = dev.numpy()
                    for i in range(n):
                        r, b = model(int(offs[i]), int(counts[i]), lo, hi)
                        assert res[i] == r, i
                        assert out[i * stride:i * stride + r].tobytes() == b, i
                        assert (out[i * stride + r:(i + 1) * stride] == 0x5A).all(), i
                elif op == "clear":
                    rd.cache_clear()
                elif op == "unload":
                    rd.unload()
                elif op == "load":
                    a = int(rng.integers(lo, hi))
                    rd.load(a, int(rng.integers(a, hi)) + 1)
                elif op == "shard":
                    world = int(rng.choice([1, 1, 2, 3]))
                    lo, hi = rd.set_shard(int(rng.integers(0, world)), world)
                elif op == "stats":
                    st = rd.stats()
                    assert st.frames == nfr and st.decompressed_size == total and st.cached_frames <= max(cache_size, 64)
                elif op == "frames":
                    a = int(rng.integers(lo, hi))
                    b = min(hi, a + int(rng.integers(1, 12)))
                    nbytes = min(b * F, total) - a * F
                    assert rd.decode_frames(a, b, dev) == nbytes and dev.numpy()[:nbytes].tobytes() == data[a * F:a * F + nbytes]
            except AssertionError:
                print("failing step", step, op, "after", log[-12:])
                raise
        assert rd.pread(10, lo * F)[1] == data[lo * F:lo * F + 10]


def test_concurrent_callers(hostlib, golden):
    """twelve caller threads — scans and random reads on one SHARED reader plus readers of their own — as in the GPU test;
    the stand-in serialises kernel launches, the locking of reader.c (reader mutex, residency read lock, parked readers,
    process-wide budgets) runs as it does in the product (reference src/decompress.c:387,499: B10)"""
    G.test_eight_concurrent_callers_on_one_reader(hostlib, golden, "zsyn_zstd3_128k", None)
