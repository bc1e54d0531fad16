"""Files the reference writer never emits but the reference reader accepts: built straight from
libzstd / liblz4 by datagen/foreign.py.  CPU part: oracle port == reference reader.  GPU part: CUDA path == both."""
import numpy as np
import pytest

from oracle.pyapi import OraclePort, RefReader, have_reference

CASES = {
    "lz4_256k_blocks": ("lz4", 1 << 20, dict(block_size_id=5)),
    "lz4_1m_blocks_indep": ("lz4", 3 << 20, dict(block_size_id=6, independent=True)),
    "lz4_4m_blocks": ("lz4", 5 << 20, dict(block_size_id=7)),
    "lz4_block_and_content_checksums": ("lz4", 300000, dict(block_checksum=True, content_checksum=True)),
    "lz4_no_content_size_hc": ("lz4", 200000, dict(content_size=False, level=9)),
    "zstd_checksum": ("zstd", 400000, dict(checksum=True)),
    "zstd_no_content_size": ("zstd", 400000, dict(content_size=False)),
    "zstd_small_window_l1": ("zstd", 1 << 20, dict(level=1, window_log=17)),
    "zstd_l12_entry_checksums": ("zstd", 700000, dict(level=12, entry_checksums=True)),
    "zstd_negative_level": ("zstd", 262144, dict(level=-3)),
}


def make(name):
    from datagen import foreign, zsyn
    codec, frame, kw = CASES[name]
    rng = np.random.Generator(np.random.PCG64(len(name)))
    data = zsyn.gen(5 << 20, seed=77) + bytes(70000) + rng.integers(0, 256, 150000, dtype=np.uint8).tobytes() + b"tail"
    return data, foreign.build(data, frame, codec, **dict(kw))


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name", list(CASES))
def test_oracle_port_matches_reference_on_foreign_files(name):
    data, image = make(name)
    with OraclePort(image) as op, RefReader(image) as rr:
        assert op.decode_all().tobytes() == data
        assert rr.pread_full(len(data), 0) == data
        rng = np.random.Generator(np.random.PCG64(1))
        for _ in range(100):
            off, cnt = int(rng.integers(0, len(data) + 10)), int(rng.choice([1, 4096, 1 << 20]))
            assert op.pread(cnt, off) == rr.pread(cnt, off)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_cuda_path_on_foreign_files(lib, name):
    import torch
    data, image = make(name)
    with lib.Reader(image=image, cache_size=2) as rd, OraclePort(image) as op:
        dev = torch.empty(len(data) + 64, dtype=torch.uint8, device="cuda")
        dev.fill_(0xEE)
        assert rd.decode_frames(0, rd.frames, dev) == len(data)
        assert dev[:len(data)].cpu().numpy().tobytes() == data
        assert bool((dev[len(data):] == 0xEE).all())
        rng = np.random.Generator(np.random.PCG64(2))
        for _ in range(60):
            off, cnt = int(rng.integers(0, len(data) + 10)), int(rng.choice([1, 4096, 1 << 20]))
            assert rd.pread(cnt, off) == op.pread(cnt, off)


@pytest.mark.gpu
@pytest.mark.parametrize("codec,kw", [("lz4", dict(block_checksum=True, content_checksum=True)), ("zstd", dict(checksum=True))])
@pytest.mark.parametrize("lane_kernel", [False, True])
def test_cuda_path_rejects_checksum_mismatches(lib, codec, kw, lane_kernel, monkeypatch):
    """The reference (through liblz4 / libzstd) fails a zseek_pread whose frame has a wrong header, block or content
    checksum; so does the CUDA path, frame by frame, while the other frames stay readable."""
    from datagen import foreign, zsyn
    if lane_kernel:
        if codec != "lz4":
            pytest.skip("LZ4 only")
        monkeypatch.setenv("ZSEEK_B200_LZ4_LANE_MIN", "0")
    data = zsyn.gen(400000, seed=13)
    good = foreign.build(data, 100000, codec, **kw)
    with OraclePort(good) as op:
        c_off, d_off = [int(x) for x in op.c_off], [int(x) for x in op.d_off]
    for what, pos in (("content checksum", c_off[2] - 1), ("payload", c_off[1] + (c_off[2] - c_off[1]) // 2)):
        img = bytearray(good)
        img[pos] ^= 0x10                                   # inside frame 1
        with lib.Reader(image=bytes(img), cache_size=2) as rd:
            assert rd.pread(1000, d_off[0] + 5)[1] == data[d_off[0] + 5:d_off[0] + 1005]
            with pytest.raises(lib.ZseekError) as e:
                rd.pread(1000, d_off[1] + 5)
            assert str(e.value).startswith("decompress frame"), (what, str(e.value))
            assert rd.pread(1000, d_off[2] + 5)[1] == data[d_off[2] + 5:d_off[2] + 1005]
            if have_reference():
                with RefReader(bytes(img), cache_size=2) as rr:
                    with pytest.raises(OSError):
                        rr.pread(1000, d_off[1] + 5)


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("codec,kw", [("lz4", dict(block_checksum=True, content_checksum=True)),
                                      ("lz4", dict(content_checksum=True, content_size=False, block_size_id=5)),
                                      ("zstd", dict(checksum=True))])
def test_oracle_port_verifies_checksums_like_the_reference(codec, kw):
    """Same verdict, frame by frame, as the reference's whole-frame (cached) path on files with one flipped bit in a
    frame header, in a frame's payload and in a content checksum (a fresh reference reader per frame)."""
    from datagen import foreign, zsyn
    data = zsyn.gen(350000, seed=17) + bytes(3000)
    good = foreign.build(data, 100000, codec, **kw)
    with OraclePort(good) as op:
        c_off, d_off = [int(x) for x in op.c_off], [int(x) for x in op.d_off]
        nfr = op.frames
    for pos in (None, c_off[0] + 5, c_off[1] + (c_off[2] - c_off[1]) // 2, c_off[3] - 2):
        img = bytearray(good)
        if pos is not None:
            img[pos] ^= 0x20
        for f in range(nfr):
            with RefReader(bytes(img), cache_size=1) as rr, OraclePort(bytes(img)) as op:
                try:
                    want = rr.pread(50, d_off[f])
                except OSError:
                    want = None
                try:
                    got = op.decode_frame(f).tobytes()[:50]
                except OSError:
                    got = None
                assert (got is None) == (want is None), (codec, pos, f)
                if want is not None:
                    assert got == want[1]
