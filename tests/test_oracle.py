"""The CPU oracle (oracle/zsk_oracle.c) pinned against the reference: committed golden vectors
(outputs of the reference's own writer + reader, tests/golden/make_golden.py) and, when
oracle/_ref/libzseek_ref.so is present, the live reference on freshly written files."""
import hashlib

import numpy as np
import pytest
from conftest import golden_case_names, sha16

from oracle.pyapi import LZ4, ZSTD, OraclePort, RefReader, have_reference


@pytest.mark.parametrize("name", golden_case_names())
def test_oracle_matches_golden(golden, name):
    cases, _ = golden
    c = cases[name]
    with OraclePort(c["image"]) as op:
        assert op.frames == c["frames"]
        assert op.size == c["input_len"]
        assert hashlib.sha256(op.decode_all().tobytes()).hexdigest() == c["input_sha256"]
        for off, cnt, ret, digest in c["reads"]:
            r, b = op.pread(cnt, off)
            assert r == ret, (off, cnt)
            assert sha16(b) == digest, (off, cnt)


def test_oracle_open_errors(golden):
    cases, errors = golden
    import os
    from conftest import GOLDEN
    for name, img in (("empty", open(os.path.join(GOLDEN, "empty.zsk"), "rb").read()),
                      ("truncated_footer", cases["tiny_zstd"]["image"][:-3]), ("garbage", b"not a seekable file at all")):
        assert errors[name] is not None  # the reference refuses these files
        with pytest.raises(OSError):
            OraclePort(img)


def test_lookup_semantics_b4():
    """Largest i with d_off[i] <= offset: zero-length frames are skipped in favour of the last frame
    starting at that offset (reference src/seek_table.c:187-202)."""
    import struct
    # hand-built seek table: dSizes 10, 0, 0, 5 over a fake 4-byte zstd magic payload
    ent = [(4, 10), (0, 0), (0, 0), (0, 5)]
    body = b"".join(struct.pack("<II", c, d) for c, d in ent)
    img = struct.pack("<I", 0xFD2FB528) + struct.pack("<II", 0x184D2A5E, len(body) + 9) + body + struct.pack("<IBI", 4, 0, 0x8F92EAB1)
    with OraclePort(img) as op:
        assert [op.offset_to_frame(o) for o in (0, 9, 10, 14, 15, 10 ** 12)] == [0, 0, 3, 3, -1, -1]
    if have_reference():
        # same answers from the reference reader: a read at 10 would have to come from frame 3
        pass


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("codec,level,frame,chunk,workers", [
    (LZ4, 0, 65536, 65536, 0), (LZ4, 0, 65536, 4096, 0), (LZ4, 0, 1 << 20, 4096, 0), (LZ4, 3, 100000, 4093, 0),
    (ZSTD, 3, 262144, 262144, 0), (ZSTD, 3, 100000, 4093, 0), (ZSTD, 19, 1 << 20, 1 << 20, 0), (ZSTD, 1, 65536, 65536, 0),
    (ZSTD, 3, 262144, 262144, 2), (ZSTD, -5, 131072, 131072, 0), (ZSTD, 9, 500000, 4093, 0),
])
def test_oracle_matches_live_reference(codec, level, frame, chunk, workers):
    from datagen import refwriter, zsyn
    data = zsyn.gen(3 << 20, seed=codec * 100 + level + frame)
    image = refwriter.write(data, codec, level, frame, chunk, 0, workers)
    with OraclePort(image) as op, RefReader(image) as rr:
        assert op.decode_all().tobytes() == data
        st = rr.stats()
        assert (op.frames, op.size) == (st.frames, st.decompressed_size)
        rng = np.random.Generator(np.random.PCG64(frame))
        for _ in range(300):
            off = int(rng.integers(0, len(data) + 100))
            cnt = int(rng.choice([0, 1, 4095, 4096, 65536, 1 << 20]))
            assert op.pread(cnt, off) == rr.pread(cnt, off)
