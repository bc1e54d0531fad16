"""Parity tests proper: the CUDA path, called through the C-ABI library exactly as a libzseek caller
would call it, against (a) the committed golden vectors = outputs of the reference itself, (b) the
CPU oracle on freshly written files, (c) size-independent properties at larger sizes.
Bar: bit-exact (integer/byte work)."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_case_names, sha16
from oracle.pyapi import LZ4, ZSTD, OraclePort, RefReader, have_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


# --------------------------------------------------------------------------- golden vectors
@pytest.mark.parametrize("cache_size", [0, 1, 4])
@pytest.mark.parametrize("name", golden_case_names())
def test_pread_matches_reference_golden(lib, golden, name, cache_size, torch_cuda):
    cases, _ = golden
    c = cases[name]
    with lib.Reader(image=c["image"], cache_size=cache_size) as rd:
        st = rd.stats()
        assert (st.frames, st.decompressed_size, st.seek_table_memory) == (c["frames"], c["input_len"], c["seek_table_memory"])
        for off, cnt, ret, digest in c["reads"]:
            r, b = rd.pread(cnt, off)
            assert r == ret, (off, cnt)
            assert sha16(b) == digest, (off, cnt)
        whole = lib.pread_full(rd, c["input_len"] + 10, 0)
        assert hashlib.sha256(whole).hexdigest() == c["input_sha256"]


@pytest.mark.parametrize("name", ["zsyn_lz4_4k_chunks", "mix_zstd19"])
def test_example_scan_4k_cache1(lib, golden, name, torch_cuda):
    """reference test/example.c:36-87: open with cache_size 1 over a FILE*, scan with 4 KiB preads looping on
    short reads, compare every chunk."""
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        want = op.decode_all().tobytes()
    with lib.Reader(path=os.path.join(GOLDEN, name + ".zsk"), cache_size=1) as rd:
        off = 0
        while off < len(want):
            chunk = lib.pread_full(rd, min(4096, len(want) - off), off)
            assert chunk == want[off:off + 4096]
            off += len(chunk)
        assert rd.pread(4096, off) == (0, b"")


def test_zseek_read_cursor(lib, golden, torch_cuda):
    cases, _ = golden
    c = cases["mix_lz4"]
    with OraclePort(c["image"]) as op:
        want = op.decode_all().tobytes()
    with lib.Reader(image=c["image"], cache_size=2) as rd:
        got = bytearray()
        while True:
            r, b = rd.read(7001)
            if r == 0:
                break
            got += b
        assert bytes(got) == want


def test_python_callbacks_reader(lib, golden, torch_cuda):
    """zseek_reader_open_full with user callbacks (reference src/zseek.h:88-116), short read only at EOF."""
    cases, _ = golden
    img = cases["zsyn_zstd3_128k"]["image"]
    calls = []

    def pread(size, offset):
        calls.append((size, offset))
        return img[offset:offset + size]
    with lib.Reader(pread=pread, fsize=lambda: len(img), cache_size=0) as rd:
        r, b = rd.pread(1000, 200000)
        with OraclePort(img) as op:
            assert (r, b) == op.pread(1000, 200000)
    assert calls[0] == (4, 0)  # magic sniff first, like the reference


def test_io_errors_surface_like_the_reference(lib, golden, torch_cuda):
    cases, _ = golden
    img = cases["zsyn_lz4_64k"]["image"]
    state = {"fail": False}

    def pread(size, offset):
        if state["fail"] and offset < 100000:
            return None  # <0
        return img[offset:offset + size]
    with lib.Reader(pread=pread, fsize=lambda: len(img), cache_size=0) as rd:
        state["fail"] = True
        with pytest.raises(lib.ZseekError) as e:
            rd.pread(10, 5)
        assert str(e.value) == "read file failed"
    short = img[:50000] + bytes(len(img) - 50000)

    def pread2(size, offset):
        if offset + size <= 60000 and offset < 50000 and size > 16:
            return img[offset:offset + size][:size // 2]  # short in the middle of the file
        return img[offset:offset + size]
    with lib.Reader(pread=pread2, fsize=lambda: len(img), cache_size=0) as rd:
        with pytest.raises(lib.ZseekError) as e:
            rd.pread(10, 5)
        assert str(e.value) == "unexpected EOF"


def test_corrupt_frame_fails_instead_of_hanging(lib, golden, torch_cuda):
    """the reference can spin forever on a frame that yields fewer bytes than the seek table
    claims; the replacement must return -1."""
    cases, _ = golden
    for name in ("zsyn_lz4_64k", "zsyn_zstd3_128k", "mix_zstd19"):
        c = cases[name]
        with OraclePort(c["image"]) as op:
            good = op.decode_all()
            img = bytearray(c["image"])
            f = int(np.argmax(np.diff(op.c_off.astype(np.int64))[1:-1])) + 1  # a big frame that is neither first nor last
            c0, c1 = int(op.c_off[f]), int(op.c_off[f + 1])
            for k in range(c0 + 20, c1, 97):
                img[k] ^= 0x55
            d0, d1 = int(op.d_off[f]), int(op.d_off[f + 1])
        with lib.Reader(image=bytes(img), cache_size=0) as rd:
            assert rd.pread(100, 0)[1] == good[:100].tobytes()
            try:
                b = rd.read_range(d1 - d0, d0)   # either rejected, or decoded to different bytes of the claimed size
                assert len(b) == d1 - d0 and b != good[d0:d1].tobytes()
            except lib.ZseekError as e:
                assert str(e).startswith("decompress frame")
            assert rd.pread(100, d1 + 5)[1] == good[d1 + 5:d1 + 105].tobytes()


def test_sequential_scan_across_a_corrupt_frame(lib, golden, torch_cuda, names=("zsyn_lz4_4k_chunks", "mix_zstd3"), cache_sizes=(0, 1)):
    """The read-ahead window of a sequential host scan may contain a bad frame far ahead of the one a call asks for.  The
    reference decodes only the requested frame (src/decompress.c:700-790): every good frame must still be served, only
    reads of the bad frame fail, and a retry gives the same verdict."""
    cases, _ = golden
    for name in names:
        c = cases[name]
        with OraclePort(c["image"]) as op:
            good = op.decode_all()
            nfr = op.frames
            bad = nfr // 2
            img = bytearray(c["image"])
            c0, c1 = int(op.c_off[bad]), int(op.c_off[bad + 1])
            for k in range(c0 + 8, c1, 7):
                img[k] ^= 0xA5
            d_off = [int(x) for x in op.d_off]
        # the reference's own verdict per frame, a fresh reader each (a failed LZ4F context stays poisoned)
        ref_ok = []
        for f in range(nfr):
            with RefReader(bytes(img), cache_size=2) as rr:
                try:
                    ref_ok.append(rr.pread(16, d_off[f])[1] == good[d_off[f]:d_off[f] + 16].tobytes())
                except OSError:
                    ref_ok.append(False)
        assert not ref_ok[bad] and all(ref_ok[:bad]) and all(ref_ok[bad + 1:])
        for cache_size in cache_sizes:
            with lib.Reader(image=bytes(img), cache_size=cache_size) as rd:
                off, failures = 0, 0
                while off < len(good):
                    try:
                        r, b = rd.pread(4096, off)
                        f = max(i for i in range(nfr) if d_off[i] <= off)
                        assert f != bad or b != good[off:off + r].tobytes()
                        if f != bad:
                            assert b == good[off:off + r].tobytes(), (name, off)
                        off += r
                    except lib.ZseekError as e:
                        assert d_off[bad] <= off < d_off[bad + 1], (name, off, str(e))
                        failures += 1
                        with pytest.raises(lib.ZseekError):      # sticky state must not flip the verdict
                            rd.pread(4096, off)
                        off = d_off[bad + 1]
                assert failures <= 1


def test_host_batch_leaves_unproduced_bytes_untouched(lib, golden, torch_cuda, names=("mix_lz4", "zsyn_zstd3_128k")):
    """zseek_b200_pread_batch with a HOST destination stores exactly the bytes a loop of zseek_pread would store: stride
    gaps, the tail of reads clipped at a frame boundary, requests at/after EOF and count = 0 stay as they were."""
    cases, _ = golden
    for name in names:
        c = cases[name]
        with OraclePort(c["image"]) as op:
            good = op.decode_all()
            ends = op.d_off[1:].astype(np.int64)
            total = op.size
        rng = np.random.Generator(np.random.PCG64(12))
        n = 400
        offs = rng.integers(0, total + 2000, n).astype(np.uint64)
        offs[:40] = (ends[rng.integers(0, len(ends), 40)] - rng.integers(1, 300, 40)).astype(np.uint64)   # clipped at a boundary
        offs[40] = total
        offs[41] = total + 10 ** 9
        counts = rng.choice([0, 1, 700, 4096], n).astype(np.uint64)
        stride = 5000
        with lib.Reader(image=c["image"], cache_size=8) as rd:
            for variant in ("stride", "offs"):
                dst = np.full(n * stride + 64, 0x5A, dtype=np.uint8)
                if variant == "stride":
                    res = rd.pread_batch(offs, counts=counts, dst=dst, dst_stride=stride)
                    where = np.arange(n, dtype=np.int64) * stride
                else:
                    where = rng.permutation(n).astype(np.int64) * stride + 7
                    res = rd.pread_batch(offs, counts=counts, dst=dst, dst_offs=where.astype(np.uint64))
                expect = np.full_like(dst, 0x5A)
                for i in range(n):
                    o, cnt = int(offs[i]), int(counts[i])
                    if o >= total or cnt == 0:
                        want = 0
                    else:
                        want = min(cnt, int(ends[np.searchsorted(ends, o, side="right")]) - o)
                    assert res[i] == want, (name, variant, i)
                    expect[where[i]:where[i] + want] = good[o:o + want]
                assert (dst == expect).all(), (name, variant)
        # a count that would overflow offset arithmetic is clipped like any other ("rest of the frame")
        with lib.Reader(image=c["image"], cache_size=8) as rd:
            dst = np.full(1 << 20, 0x5A, dtype=np.uint8)
            res = rd.pread_batch(np.array([5], dtype=np.uint64), counts=np.array([2 ** 64 - 1], dtype=np.uint64), dst=dst, dst_stride=0)
            k = int(ends[0]) - 5
            assert res[0] == k and (dst[:k] == good[5:5 + k]).all() and (dst[k:] == 0x5A).all()


# --------------------------------------------------------------------------- oracle on fresh files
def fresh_files():
    from datagen import refwriter, zsyn
    data = zsyn.gen(6 << 20, seed=99)
    yield "lz4_64k", data, refwriter.write(data, LZ4, 0, 65536)
    yield "lz4_1m_linked", data, refwriter.write(data, LZ4, 0, 1 << 20, 4096)
    yield "zstd3_256k", data, refwriter.write(data, ZSTD, 3, 262144)
    yield "zstd19_1m", data[:3 << 20], refwriter.write(data[:3 << 20], ZSTD, 19, 1 << 20)
    yield "zstd1_ragged", data, refwriter.write(data, ZSTD, 1, 100000, 4093)
    yield "zstd_mt", data, refwriter.write(data, ZSTD, 3, 262144, 262144, 0, 2)


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
def test_fresh_files_all_entry_points(lib, torch_cuda):
    torch = torch_cuda
    for name, data, image in fresh_files():
        total = len(data)
        ref = np.frombuffer(data, dtype=np.uint8)
        with lib.Reader(image=image, cache_size=8) as rd, OraclePort(image) as op, RefReader(image) as rr:
            assert op.decode_all().tobytes() == data
            # (1) whole-file decode straight into device memory
            dev = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
            dev.fill_(0xEE)
            assert rd.decode_frames(0, rd.frames, dev) == total
            assert torch.equal(dev[:total].cpu(), torch.from_numpy(ref.copy())), name
            assert bool((dev[total:] == 0xEE).all()), "wrote past the end"
            assert rd.launch_count >= 1 and rd.last_decode_ms > 0
            # (2) zseek_pread with a DEVICE destination
            rng = np.random.Generator(np.random.PCG64(5))
            for _ in range(20):
                off, cnt = int(rng.integers(0, total)), int(rng.choice([1, 4096, 300000]))
                d = torch.zeros(cnt, dtype=torch.uint8, device="cuda")
                r = rd.pread_into(d, cnt, off)
                want = rr.pread(cnt, off)
                assert r == want[0] and d[:r].cpu().numpy().tobytes() == want[1]
            # (3) multi-frame range reads, host and device destinations
            for off, cnt in [(0, total), (1, total), (70001, 1500000), (total - 5, 100), (total, 10), (123, 0)]:
                want = data[off:off + cnt]
                assert rd.read_range(cnt, off) == want, (name, off, cnt)
                d = torch.zeros(max(cnt, 1), dtype=torch.uint8, device="cuda")
                r = rd.read_range_into(d, cnt, off)
                assert r == len(want) and d[:r].cpu().numpy().tobytes() == want
            # (4) batch == loop of zseek_pread (reference), host and device destinations
            n = 3000
            offs = rng.integers(0, total + 1000, n).astype(np.uint64)
            offs[:op.frames] = op.d_off[1:] - 1  # boundary straddlers
            cnts = rng.choice([0, 1, 100, 4096, 8192], n).astype(np.uint64)
            stride = 8192
            hdst = np.zeros(n * stride, dtype=np.uint8)
            res = rd.pread_batch(offs, counts=cnts, dst=hdst, dst_stride=stride)
            ddst = torch.zeros(n * stride, dtype=torch.uint8, device="cuda")
            rd.cache_clear()
            res2 = rd.pread_batch(offs, fixed_count=4096, dst=ddst, dst_stride=stride)
            dd = ddst.cpu().numpy()
            for i in range(n):
                r, b = rr.pread(int(cnts[i]), int(offs[i]))
                assert res[i] == r and hdst[i * stride:i * stride + r].tobytes() == b, (name, i)
                r, b = rr.pread(4096, int(offs[i]))
                assert res2[i] == r and dd[i * stride:i * stride + r].tobytes() == b, (name, i)


def test_shard_restricts_frames(lib, golden, torch_cuda):
    cases, _ = golden
    c = cases["mix_zstd3"]
    with lib.Reader(image=c["image"]) as rd, OraclePort(c["image"]) as op:
        lo, hi = rd.set_shard(1, 2)
        assert (lo, hi) == (op.frames // 2, op.frames)
        good = op.decode_all()
        o = int(op.d_off[lo])
        assert rd.pread(100, o)[1] == good[o:o + 100].tobytes()
        with pytest.raises(lib.ZseekError) as e:
            rd.pread(100, 0)
        assert "shard" in str(e.value)


# --------------------------------------------------------------------------- properties at size
@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 262144)])
def test_round_trip_256mib(lib, codec, level, frame, torch_cuda):
    """Size-independent property (the reference's own check, test/example.c:82-86): decode(write(x)) == x,
    at 256 MiB built by tile-and-replicate; compared on the device, plus random 4 KiB probes vs the source."""
    torch = torch_cuda
    from datagen import refwriter, zsyn
    tile = zsyn.gen_parallel(32 << 20)
    image = refwriter.replicate(refwriter.write_parallel(tile, codec, level, frame), 8)
    total = len(tile) * 8
    with lib.Reader(image=image, cache_size=64) as rd:
        assert rd.size == total
        dev = torch.empty(total, dtype=torch.uint8, device="cuda")
        assert rd.decode_frames(0, rd.frames, dev) == total
        t = torch.from_numpy(np.frombuffer(tile, dtype=np.uint8).copy()).cuda()
        assert bool((dev.view(8, -1) == t).all())
        rng = np.random.Generator(np.random.PCG64(11))
        offs = rng.integers(0, total - 4096, 20000).astype(np.uint64)
        out = torch.zeros(20000 * 4096, dtype=torch.uint8, device="cuda")
        res = rd.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)
        # a request returns MIN(4096, frame_end - offset): check exactly those bytes
        src = np.frombuffer(tile, dtype=np.uint8)
        o = out.cpu().numpy().reshape(20000, 4096)
        for i in range(0, 20000, 37):
            k, a = int(res[i]), int(offs[i]) % len(tile)
            assert k == min(4096, frame - int(offs[i]) % frame)
            assert (o[i, :k] == src[a:a + k]).all() if a + k <= len(tile) else True


# --------------------------------------------------------------------------- the reference's own integration test
@pytest.mark.parametrize("flag", ["--lz4", "--zstd"])
def test_reference_example_binary_runs_against_b200_reader(flag, tmp_path, torch_cuda):
    """BASELINE configs[0]: the UNMODIFIED reference test/example.c (compiled by oracle/Makefile where
    /root/reference exists) with its six reader symbols bound to libzseek_b200.so and its writer symbols to
    the reference build: compress with the reference writer, scan with 4 KiB zseek_preads (cache_size 1)
    through the CUDA path, memcmp every chunk, print SUCCESS (reference test/example.c:19-122,248-263)."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "example_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/example_b200 not built (needs /root/reference at build time)")
    from datagen import zsyn
    src = tmp_path / "input.bin"
    src.write_bytes(zsyn.gen(64 << 20) if flag == "--lz4" else zsyn.gen(24 << 20))   # 64 MiB synthetic compressible file
    needed = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libzseek_b200.so" in needed
    p = subprocess.run([exe, flag, str(src)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "SUCCESS" in p.stdout, (p.stdout, p.stderr)


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 262144)])
def test_file_opened_by_path_uses_parallel_ingest_and_decodes_identically(lib, codec, level, frame, tmp_path, torch_cuda):
    """n2: zseek_reader_open(FILE*) pulls large ranges of the file with pread(2) from worker threads into the pinned staging
    halves (the callback path stays as it is).  A 192 MiB file read by path — whole-file decode to device memory, a
    multi-frame host range, plain zseek_preads — must give the writer's input, with 1 and with 8 I/O threads; the
    caller's FILE position is left where it was."""
    torch = torch_cuda
    import os
    from datagen import refwriter, zsyn
    data = zsyn.gen_parallel(192 << 20)
    image = refwriter.write_parallel(data, codec, level, frame, piece_frames=max(1, (8 << 20) // frame))
    path = tmp_path / "big.zsk"
    path.write_bytes(image)
    want = np.frombuffer(data, dtype=np.uint8)
    for threads in ("1", "8"):
        os.environ["ZSEEK_B200_IO_THREADS"] = threads
        try:
            with lib.Reader(path=str(path), cache_size=0) as rd:
                assert rd.size == len(data)
                dev = torch.zeros(rd.size + 64, dtype=torch.uint8, device="cuda")
                assert rd.decode_frames(0, rd.frames, dev) == rd.size
                assert bool((dev[:rd.size] == torch.from_numpy(want.copy()).cuda()).all())
                host = np.zeros(50 << 20, dtype=np.uint8)
                assert rd.read_range_into(host, host.size, 7 << 20) == host.size
                assert (host == want[7 << 20:(7 << 20) + host.size]).all()
                for off in (0, 123456789, rd.size - 100):
                    r, b = rd.pread(70000, off)
                    assert b == data[off:off + r] and r > 0
        finally:
            del os.environ["ZSEEK_B200_IO_THREADS"]


def test_hbm_cache_honours_capacity_and_evicts_least_recently_used(lib, torch_cuda):
    """Restates reference test/test_cache.c:135-159 (capacity 3, four inserts: the first one is gone, the other three are
    found) for the HBM frame cache, through the API: capacity C = cache_size, C + 1 distinct frames read in an order no
    read-ahead follows, then hits and the one miss are told apart by whether a read launches a kernel.  Also what the
    reference's list gets wrong: a hit promotes, so the promoted frame survives the next eviction, and
    cached_frames never exceeds the capacity."""
    torch = torch_cuda
    from datagen import refwriter, zsyn
    if not have_reference():
        pytest.skip("inputs come from the reference writer (oracle/_ref)")
    data = zsyn.gen(160 * 8192, seed=5)
    image = refwriter.write(data, LZ4, 0, 8192)
    C_ = 70
    buf = torch.zeros(8192, dtype=torch.uint8, device="cuda")
    with lib.Reader(image=image, cache_size=C_) as rd:
        assert rd.frames == 160
        rd.load(0, rd.frames)

        def read(f):
            l0 = rd.launch_count
            assert rd.pread_into(buf, 100, f * 8192 + 7) == 100
            assert buf[:100].cpu().numpy().tobytes() == data[f * 8192 + 7:f * 8192 + 107]
            return rd.launch_count > l0       # True = a kernel ran = miss

        order = list(range(2 * C_, -1, -2))[:C_ + 1]          # C + 1 distinct frames, descending by 2: never sequential
        assert all(read(f) for f in order[:C_])               # all misses
        assert rd.stats().cached_frames == C_
        assert not read(order[0])                             # hit: promotes the oldest entry to most recently used
        assert read(order[C_])                                # one more frame: evicts the least recently used = order[1]
        assert rd.stats().cached_frames == C_
        assert not read(order[0]) and not read(order[C_])     # the promoted and the new frame are there
        for f in order[2:C_]:
            assert not read(f), f                             # so is everything younger than the victim
        assert read(order[1])                                 # the victim is gone
        assert rd.stats().cached_frames == C_


@pytest.mark.parametrize("name", ["zsyn_lz4_64k", "zsyn_lz4_256k_linked", "mix_lz4", "zsyn_zstd3_128k", "zsyn_zstd19_256k", "zsyn_zstd3_mt",
                                  "mix_zstd3", "mix_zstd19"])
def test_bit_flip_fuzz_never_crashes_and_never_lies(lib, golden, name, monkeypatch, torch_cuda):
    """Random single-bit flips over every golden file, every shipped decode kernel: the call returns (no hang, no sticky
    CUDA error: the next read works), frames the flip did not touch decode to the reference bytes, and the frame that holds
    the flip is either rejected or decodes to bytes of the claimed size (whatever they are) — never an out-of-bounds write
    (guard bytes)."""
    torch = torch_cuda
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        good = op.decode_all()
        c_off, d_off, nfr, codec = op.c_off.astype(np.int64), op.d_off.astype(np.int64), op.frames, op.codec
    payload = int(c_off[-1])
    rng = np.random.Generator(np.random.PCG64(99))
    envs = [{"ZSEEK_B200_LZ4_LANE_MIN": "0"}, {"ZSEEK_B200_LZ4_LANE_MIN": "4000000000"}] if codec == LZ4 else [{}, {"ZSEEK_B200_ZSTD_LEGACY": "1"}]
    for env in envs:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        for trial in range(25):
            img = bytearray(c["image"])
            pos = int(rng.integers(0, payload))
            img[pos] ^= 1 << int(rng.integers(0, 8))
            hit = int(np.searchsorted(c_off, pos, side="right")) - 1
            with lib.Reader(image=bytes(img), cache_size=0) as rd:
                dev = torch.full((int(d_off[-1]) + 256,), 0xEE, dtype=torch.uint8, device="cuda")
                try:
                    rd.decode_frames(0, nfr, dev[128:])
                    rejected = False
                except lib.ZseekError as e:
                    assert str(e).startswith("decompress frame"), str(e)
                    rejected = True
                got = dev.cpu().numpy()
                assert (got[:128] == 0xEE).all() and (got[128 + int(d_off[-1]):] == 0xEE).all(), (name, env, pos)
                for f in range(nfr):
                    if f != hit:
                        a, b = int(d_off[f]), int(d_off[f + 1])
                        assert (got[128 + a:128 + b] == good[a:b]).all(), (name, env, pos, f)
                if not rejected:   # the flipped frame decoded to the claimed size; usually a checksum-less format cannot tell
                    pass
                # the reader (and the CUDA context) is still usable
                f2 = (hit + 1) % nfr
                r, bts = rd.pread(50, int(d_off[f2]))
                assert bts == good[int(d_off[f2]):int(d_off[f2]) + r].tobytes()
        for k in env:
            monkeypatch.delenv(k)


# --------------------------------------------------------------------------- every LZ4 kernel, whatever the launch size
LZ4_KERNEL_ENVS = {
    "lane_per_frame": {"ZSEEK_B200_LZ4_LANE_MIN": "0", "ZSEEK_B200_SORT_MIN": "0"},           # the kernel big launches get
    "lane_per_frame_ordered_jobs": {"ZSEEK_B200_LZ4_LANE_MIN": "0", "ZSEEK_B200_SORT_MIN": "1"},
    "warp_per_frame": {"ZSEEK_B200_LZ4_LANE_MIN": "4000000000"},                              # the kernel small launches get
}


@pytest.fixture
def kernel_env(request, monkeypatch):
    for k, v in request.param.items():
        monkeypatch.setenv(k, v)      # read by zseek_reader_open* (one device context per reader)
    return request.param


@pytest.mark.parametrize("kernel_env", list(LZ4_KERNEL_ENVS.values()), ids=list(LZ4_KERNEL_ENVS), indirect=True)
@pytest.mark.parametrize("name", ["tiny_lz4", "zsyn_lz4_64k", "zsyn_lz4_256k_linked", "zsyn_lz4_4k_chunks", "mix_lz4"])
def test_every_lz4_kernel_matches_golden(lib, golden, name, kernel_env, torch_cuda):
    """The launch layer picks the LZ4 kernel by the number of frames in the launch; here each one is forced onto the
    golden files (raw blocks, linked blocks, ragged frames, unaligned frame starts) through whole-file decode to
    device memory with guard bytes, range reads to host memory and the reference's own read vectors."""
    torch = torch_cuda
    cases, _ = golden
    c = cases[name]
    total = c["input_len"]
    with lib.Reader(image=c["image"], cache_size=4) as rd:
        dev = torch.full((total + 128,), 0xEE, dtype=torch.uint8, device="cuda")
        for lead in (64, 67):                     # frame outputs 16-byte aligned and not
            dev.fill_(0xEE)
            assert rd.decode_frames(0, rd.frames, dev[lead:]) == total
            got = dev.cpu().numpy()
            assert hashlib.sha256(got[lead:lead + total].tobytes()).hexdigest() == c["input_sha256"]
            assert (got[:lead] == 0xEE).all() and (got[lead + total:] == 0xEE).all(), "wrote outside the destination"
        want_kernel = "zsk_lz4_decode_lane_kernel" if kernel_env.get("ZSEEK_B200_LZ4_LANE_MIN") == "0" else "zsk_lz4_decode_batch_kernel"
        assert rd.last_decode_kernel == want_kernel
        assert hashlib.sha256(rd.read_range(total + 5, 0)).hexdigest() == c["input_sha256"]
        for off, cnt, ret, digest in c["reads"][:120]:
            r, b = rd.pread(cnt, off)
            assert r == ret and sha16(b) == digest, (off, cnt)


@pytest.mark.parametrize("kernel_env", [LZ4_KERNEL_ENVS["lane_per_frame_ordered_jobs"], LZ4_KERNEL_ENVS["warp_per_frame"]],
                         ids=["lane_per_frame_ordered_jobs", "warp_per_frame"], indirect=True)
def test_lz4_kernels_reject_corrupt_frames(lib, golden, kernel_env, torch_cuda):
    cases, _ = golden
    c = cases["zsyn_lz4_64k"]
    with OraclePort(c["image"]) as op:
        good = op.decode_all()
        img = bytearray(c["image"])
        c0, c1 = int(op.c_off[1]), int(op.c_off[2])
        for k in range(c0 + 20, c1, 97):
            img[k] ^= 0x55
        d0, d1 = int(op.d_off[1]), int(op.d_off[2])
        trunc = bytes(c["image"][:c0 + 100]) + bytes(c1 - c0 - 100) + bytes(c["image"][c1:])   # frame 1 cut short (zeros)
    for image in (bytes(img), trunc):
        with lib.Reader(image=image, cache_size=0) as rd:
            assert rd.pread(100, 0)[1] == good[:100].tobytes()
            try:
                b = rd.read_range(d1 - d0, d0)
                assert len(b) == d1 - d0 and b != good[d0:d1].tobytes()
            except lib.ZseekError as e:
                assert str(e).startswith("decompress frame")
            assert rd.pread(100, d1 + 5)[1] == good[d1 + 5:d1 + 105].tobytes()


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 131072)])
def test_ordered_job_lists(lib, codec, level, frame, monkeypatch, torch_cuda):
    """Big launches get a largest-first job list; forced here on a small ragged file: every frame must
    still land at its own place."""
    torch = torch_cuda
    from datagen import refwriter, zsyn
    monkeypatch.setenv("ZSEEK_B200_SORT_MIN", "1")
    monkeypatch.setenv("ZSEEK_B200_SORT_MIN_ZSTD", "1")
    monkeypatch.setenv("ZSEEK_B200_LZ4_LANE_MIN", "0")
    rng = np.random.Generator(np.random.PCG64(3))
    data = zsyn.gen(5 << 20, seed=5) + bytes(300000) + rng.integers(0, 256, 400000, dtype=np.uint8).tobytes() + zsyn.gen(1 << 20, seed=6)
    image = refwriter.write(data, codec, level, frame, 4093)
    with lib.Reader(image=image, cache_size=0) as rd:
        dev = torch.full((len(data) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
        for lo, hi in ((0, rd.frames), (3, rd.frames - 2), (0, rd.frames)):
            dev.fill_(0xEE)
            o0, o1 = int(rd.d_off[lo]), int(rd.d_off[hi])
            assert rd.decode_frames(lo, hi, dev) == o1 - o0
            got = dev.cpu().numpy()
            assert got[:o1 - o0].tobytes() == data[o0:o1] and (got[o1 - o0:] == 0xEE).all()


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 262144), (ZSTD, 19, 1 << 20)])
def test_batches_over_partially_decoded_frames(lib, codec, level, frame, torch_cuda):
    """A batch decodes a missing frame only up to the last byte it needs; the HBM cache then holds a
    prefix.  Later batches and plain preads that need more of the same frame must still return the reference's bytes."""
    torch = torch_cuda
    from datagen import refwriter, zsyn
    data = zsyn.gen(6 << 20, seed=31)
    image = refwriter.write(data, codec, level, frame)
    total = len(data)
    rng = np.random.Generator(np.random.PCG64(8))
    with lib.Reader(image=image, cache_size=64) as rd, RefReader(image) as rr:
        nfr = rd.frames
        out = torch.zeros(4000 * 4096, dtype=torch.uint8, device="cuda")
        for rnd, hi_frac in enumerate((0.05, 0.3, 0.1, 1.0, 0.6)):     # how deep into the frames this round reads
            n = 1500
            fr = rng.integers(0, nfr, n)
            inf = (rng.random(n) * hi_frac * frame).astype(np.uint64)
            offs = np.minimum(rd.d_off[fr].astype(np.uint64) + inf, np.uint64(total - 1))
            cnt = 4096 if rnd % 2 == 0 else 700
            out.zero_()
            res = rd.pread_batch(offs, fixed_count=cnt, dst=out, dst_stride=4096)
            o = out.cpu().numpy()
            for i in range(n):
                r, b = rr.pread(cnt, int(offs[i]))
                assert res[i] == r and o[i * 4096:i * 4096 + r].tobytes() == b, (rnd, i, int(offs[i]))
            for _ in range(10):                                        # plain preads see whole frames only
                off = int(rng.integers(0, total))
                assert rd.pread(100000, off) == rr.pread(100000, off)


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 262144), (ZSTD, 19, 1 << 20)])
def test_async_batch_on_a_caller_stream_matches_a_loop_of_reference_preads(lib, codec, level, frame, torch_cuda):
    """n1: zseek_b200_pread_batch_async — request arrays, destination and results in device memory, everything queued on
    a caller stream, no host round trip.  Results and bytes must be what a loop of the REFERENCE's zseek_pread gives
    (short reads at frame boundaries, 0 at/after EOF, count 0); a consumer kernel queued on the same stream right
    behind the call sees the data; a corrupt frame surfaces in batch_wait, not in the call."""
    torch = torch_cuda
    from datagen import refwriter, zsyn
    data = zsyn.gen(5 << 20, seed=41) + b"z" * 12345
    image = refwriter.write(data, codec, level, frame)
    total = len(data)
    rng = np.random.Generator(np.random.PCG64(18))
    stream = torch.cuda.Stream()
    with lib.Reader(image=image, cache_size=0) as rd, RefReader(image) as rr:
        rd.load(0, rd.frames)
        for rnd in range(3):
            n = 3000
            offs = rng.integers(0, total + 500, n).astype(np.uint64)
            offs[:50] = (rd.d_off[rng.integers(1, rd.frames + 1, 50)].astype(np.int64) - rng.integers(1, 2000, 50)).astype(np.uint64)
            counts = rng.choice([0, 1, 700, 4096], n).astype(np.uint64)
            d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
            d_counts = torch.from_numpy(counts.astype(np.int64)).cuda()
            d_res = torch.full((n,), -7, dtype=torch.int64, device="cuda")
            dst = torch.full((n * 4096,), 0x5A, dtype=torch.uint8, device="cuda")
            with torch.cuda.stream(stream):
                rd.pread_batch_async(d_offs, dst, dev_counts=d_counts if rnd != 1 else None, fixed_count=4096 if rnd == 1 else 0,
                                     dst_stride=4096, dev_results=d_res, stream=stream if rnd != 2 else None)
                if rnd == 2:
                    rd.batch_wait()                       # the reader's own stream: wait before touching the results
                checksum = dst.to(torch.int64).sum()      # consumer on the same stream, queued before any host wait
            stream.synchronize()
            rd.batch_wait()
            res, out = d_res.cpu().numpy(), dst.cpu().numpy()
            expect = np.full_like(out, 0x5A)
            for i in range(n):
                cnt = int(counts[i]) if rnd != 1 else 4096
                r, b = rr.pread(cnt, int(offs[i]))
                assert res[i] == r, (rnd, i)
                expect[i * 4096:i * 4096 + r] = np.frombuffer(b, dtype=np.uint8)
            assert (out == expect).all(), rnd
            assert int(checksum.item()) == int(expect.astype(np.int64).sum())
    # a corrupt frame: the call succeeds (nothing has run yet), batch_wait reports it
    with OraclePort(image) as op:
        c0, c1 = int(op.c_off[1]), int(op.c_off[2])
    bad = bytearray(image)
    for k in range(c0 + 12, c1, 5):
        bad[k] ^= 0x3C
    with lib.Reader(image=bytes(bad), cache_size=0) as rd:
        rd.load(0, rd.frames)
        d_offs = torch.from_numpy(rd.d_off[:3].astype(np.int64) + 100).cuda()
        dst = torch.zeros(3 * 4096, dtype=torch.uint8, device="cuda")
        rd.pread_batch_async(d_offs, dst, fixed_count=4096, dst_stride=4096)
        with pytest.raises(lib.ZseekError) as e:
            rd.batch_wait()
        assert str(e.value).startswith("decompress frame")
        assert dst[:4096].cpu().numpy().tobytes() == data[100:4196]          # the good frames of the batch are served


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzseek_ref.so missing")
@pytest.mark.parametrize("codec,level,frame", [(LZ4, 0, 65536), (ZSTD, 3, 262144)])
def test_other_calls_while_an_async_batch_is_pending_on_a_caller_stream(lib, codec, level, frame, torch_cuda):
    """A stream-ordered batch queued on a CALLER stream shares the compressed image, the batch scratch and the zstd scratch
    pools with every other entry point of the reader.  Calls made before zseek_b200_batch_wait must order themselves after
    it on the device (async_fence) instead of racing it: both the pending batch and the interleaved calls return the
    right bytes."""
    torch = torch_cuda
    from datagen import refwriter, zsyn
    data = zsyn.gen(6 << 20, seed=77)
    ref = np.frombuffer(data, dtype=np.uint8)
    image = refwriter.write(data, codec, level, frame)
    total = len(data)
    rng = np.random.Generator(np.random.PCG64(5))
    stream = torch.cuda.Stream()
    with lib.Reader(image=image, cache_size=4) as rd:
        rd.load(0, rd.frames)
        for rnd in range(4):
            n = 2000
            offs = rng.integers(0, total - 4096, n).astype(np.uint64)
            d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
            d_res = torch.full((n,), -7, dtype=torch.int64, device="cuda")
            dst = torch.full((n * 4096,), 0x5A, dtype=torch.uint8, device="cuda")
            whole = torch.zeros(rd.size + 64, dtype=torch.uint8, device="cuda")
            busy = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
            o2 = rng.integers(0, total - 4096, 500).astype(np.uint64)
            out2 = np.zeros(500 * 4096, dtype=np.uint8)
            torch.cuda.synchronize()                      # the tensors above exist before any other stream touches them
            # a few ms of fills on the caller stream first, so that the batch is certainly still pending below
            with torch.cuda.stream(stream):
                for _ in range(24):
                    busy.fill_(rnd)
                rd.pread_batch_async(d_offs, dst, fixed_count=4096, dst_stride=4096, dev_results=d_res, stream=stream)
            # no wait: other entry points of the same reader
            res2 = rd.pread_batch(o2, fixed_count=4096, dst=out2, dst_stride=4096)
            assert rd.decode_frames(0, rd.frames, whole) == rd.size
            r3, b3 = rd.pread(3000, int(offs[0]) // 2)
            rd.batch_wait()
            stream.synchronize()
            d_off = np.asarray(rd.d_off, dtype=np.uint64)

            def expect(o):
                end = int(d_off[np.searchsorted(d_off, o, side="right")])
                return min(4096, end - int(o))
            res, out = d_res.cpu().numpy(), dst.cpu().numpy()
            for i in range(n):
                k = expect(offs[i])
                assert res[i] == k, (rnd, i)
                assert (out[i * 4096:i * 4096 + k] == ref[int(offs[i]):int(offs[i]) + k]).all(), (rnd, i)
            for i in range(500):
                k = expect(o2[i])
                assert res2[i] == k and (out2[i * 4096:i * 4096 + k] == ref[int(o2[i]):int(o2[i]) + k]).all(), (rnd, i)
            assert (whole[:rd.size].cpu().numpy() == ref).all(), rnd
            assert b3 == data[int(offs[0]) // 2:int(offs[0]) // 2 + r3] and r3 > 0


@pytest.mark.parametrize("name", ["mix_lz4", "zsyn_zstd3_128k"])
def test_eight_concurrent_callers_on_one_reader(lib, golden, name, torch_cuda):
    """zseek_pread and zseek_reader_stats may be called concurrently on one reader.  Eight threads
    (ctypes releases the GIL inside the calls) mix sequential scans, random reads and stats on ONE reader, four more use
    readers of their own on the same GPU; every result must be the reference's."""
    import threading
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        want = op.decode_all().tobytes()
        frame_ends = [int(x) for x in op.d_off[1:]]
    total = len(want)
    errors = []

    def expect(off, cnt):
        if off >= total:
            return b""
        end = min(off + cnt, next(e for e in frame_ends if e > off))   # B1: never crosses a frame boundary
        return want[off:end]

    def scan(rd, start, step):
        try:
            off = start
            while off < total:
                r, b = rd.pread(step, off)
                assert b == expect(off, step), ("scan", off)
                off += max(r, 1)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def rand(rd, seed):
        try:
            rng = np.random.Generator(np.random.PCG64(seed))
            for _ in range(150):
                off, cnt = int(rng.integers(0, total + 50)), int(rng.choice([1, 100, 4096, 70000]))
                r, b = rd.pread(cnt, off)
                assert b == expect(off, cnt) and r == len(b), ("rand", off, cnt)
                if seed % 2:
                    st = rd.stats()
                    assert st.decompressed_size == total
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    with lib.Reader(image=c["image"], cache_size=4) as shared:
        own = [lib.Reader(image=c["image"], cache_size=0) for _ in range(4)]
        ts = [threading.Thread(target=scan, args=(shared, 0, 4096)), threading.Thread(target=scan, args=(shared, total // 2, 9000))]
        ts += [threading.Thread(target=rand, args=(shared, s)) for s in range(6)]
        ts += [threading.Thread(target=scan, args=(own[i], 0, 4096 * (i + 1))) for i in range(2)]
        ts += [threading.Thread(target=rand, args=(own[2 + i], 50 + i)) for i in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for r in own:
            r.close()
    assert not errors, errors[:3]


# --------------------------------------------------------------------------- residency of random host readers, parked readers
@pytest.mark.parametrize("mode", ["pinned", "hbm"])
@pytest.mark.parametrize("name", ["zsyn_zstd3_128k", "zsyn_lz4_64k"])
def test_random_host_readers_go_resident(lib, golden, name, mode, monkeypatch, torch_cuda):
    """A host reader that keeps missing at random places gets its whole shard decoded once: into a pinned window (reads are
    memcpys, taken without the reader mutex by the threads that share the reader) or — when the process-wide pinned budget
    is spent, forced here with ZSEEK_B200_RESIDENT_MB=0 — into HBM (reads are small device-to-host copies).  Every byte must
    still be the reference's (B1-B3: short reads at frame ends, 0 at EOF), from several threads at once."""
    import threading
    cases, _ = golden
    c = cases[name]
    monkeypatch.setenv("ZSEEK_B200_RESIDENT_AFTER", "4")
    if mode == "hbm":
        monkeypatch.setenv("ZSEEK_B200_RESIDENT_MB", "0")
    with OraclePort(c["image"]) as op:
        want = op.decode_all().tobytes()
        frame_ends = [int(x) for x in op.d_off[1:]]
    total = len(want)

    def expect(off, cnt):
        if off >= total:
            return b""
        return want[off:min(off + cnt, next(e for e in frame_ends if e > off))]

    errors = []

    def rand(rd, seed, n):
        try:
            rng = np.random.Generator(np.random.PCG64(seed))
            for _ in range(n):
                off, cnt = int(rng.integers(0, total + 50)), int(rng.choice([1, 100, 4096, 70000, 200000]))
                r, b = rd.pread(cnt, off)
                assert b == expect(off, cnt) and r == len(b), (mode, off, cnt)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    with lib.Reader(image=c["image"], cache_size=0) as rd:
        rand(rd, 1, 40)                     # isolated misses, then the residency switch
        ts = [threading.Thread(target=rand, args=(rd, 10 + k, 200)) for k in range(6)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not errors, errors[:3]
        dev = torch_cuda.empty(5000, dtype=torch_cuda.uint8, device="cuda")      # a device destination still works
        assert rd.pread_into(dev, 5000, 12345) == len(expect(12345, 5000))
        assert dev[:len(expect(12345, 5000))].cpu().numpy().tobytes() == expect(12345, 5000)
        rd.cache_clear()                    # leaves residency; the ordinary path serves again
        rand(rd, 99, 10)
    assert not errors, errors[:3]


def test_parked_reader_state_is_reused_across_files(lib, golden, torch_cuda):
    """zseek_reader_close parks the device side of a reader and the next open takes it over — whatever the next file looks like
    (other codec, other frame size, more frames).  Open/close in a loop over all golden files, twice, checking sampled reads
    and a whole-file decode every time."""
    cases, _ = golden
    names = sorted(cases)
    for rnd in range(2):
        for name in names if rnd == 0 else reversed(names):
            c = cases[name]
            with OraclePort(c["image"]) as op:
                want = op.decode_all()
            with lib.Reader(image=c["image"], cache_size=rnd) as rd:
                assert rd.size == want.size
                if want.size == 0:
                    continue
                for off in (0, want.size // 3, max(want.size - 100, 0)):
                    r, b = rd.pread(5000, off)
                    assert r > 0 and b == want[off:off + r].tobytes(), (name, off)
                dev = torch_cuda.empty(rd.size, dtype=torch_cuda.uint8, device="cuda")
                assert rd.decode_frames(0, rd.frames, dev) == rd.size
                assert (dev.cpu().numpy() == want).all(), name
