"""Generates the committed golden fixtures (run in the build container, where /root/reference and
hence oracle/_ref/libzseek_ref.so exist):

    python tests/golden/make_golden.py

For every case it writes the file image produced by the REFERENCE WRITER (tests/golden/<name>.zsk)
and records in golden.json what the REFERENCE READER returned for a fixed list of zseek_pread calls
(return value + sha256 of the bytes), the reader stats, and the sha256 of the original input.  The
reference has no golden vectors of its own for this path; these outputs of the
reference itself, run here, are what pins the oracle and the CUDA path on the GPU box, where
/root/reference does not exist.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from datagen import refwriter, zsyn  # noqa: E402
from oracle.pyapi import LZ4, ZSTD, RefReader  # noqa: E402


def mix(scale=1):
    rng = np.random.Generator(np.random.PCG64(7))
    return (bytes(65536 * scale) + rng.integers(0, 256, 65536 * scale, dtype=np.uint8).tobytes()
            + zsyn.gen(196608 * scale) + b"ab" * (20000 * scale)
            + rng.integers(0, 4, 32768 * scale, dtype=np.uint8).tobytes() + b"x" * 77)


def cases():
    z = zsyn.gen(512 << 10)
    m = mix()
    yield "tiny_zstd", b"hello", dict(codec=ZSTD, level=3, min_frame_size=1 << 20, chunk=4096)
    yield "tiny_lz4", b"hello", dict(codec=LZ4, level=0, min_frame_size=1 << 20, chunk=4096)
    yield "zsyn_lz4_64k", z, dict(codec=LZ4, level=0, min_frame_size=65536, chunk=65536)          # direct path
    yield "zsyn_lz4_4k_chunks", z[:300000], dict(codec=LZ4, level=0, min_frame_size=65536, chunk=4096)  # buffered path
    yield "zsyn_lz4_256k_linked", z, dict(codec=LZ4, level=0, min_frame_size=262144, chunk=4096)  # linked blocks
    yield "zsyn_zstd3_128k", z, dict(codec=ZSTD, level=3, min_frame_size=131072, chunk=131072)
    yield "zsyn_zstd3_256k_chunks", z, dict(codec=ZSTD, level=3, min_frame_size=262144, chunk=4096, strategy=1)
    yield "zsyn_zstd19_256k", z, dict(codec=ZSTD, level=19, min_frame_size=262144, chunk=262144)
    yield "zsyn_zstd3_mt", z, dict(codec=ZSTD, level=3, min_frame_size=131072, chunk=131072, nb_workers=2)
    yield "mix_lz4", m, dict(codec=LZ4, level=0, min_frame_size=20000, chunk=4093)
    yield "mix_zstd3", m, dict(codec=ZSTD, level=3, min_frame_size=20000, chunk=4093)
    yield "mix_zstd19", m, dict(codec=ZSTD, level=19, min_frame_size=60000, chunk=4093)


def request_list(d_off, total, seed):
    """Fixed request set: every frame boundary ±{0,1,100}, EOF cases, count 0, seeded random pairs."""
    reqs = []
    for b in d_off:
        for delta in (-100, -1, 0, 1, 100):
            o = int(b) + delta
            if 0 <= o:
                reqs.append((o, 4096))
    reqs += [(total - 1, 10), (total, 10), (total + 10 ** 9, 10), (0, 0), (total // 2, 0), (0, 1), (0, total + 5)]
    rng = np.random.Generator(np.random.PCG64(seed))
    for _ in range(200):
        reqs.append((int(rng.integers(0, total + 50)), int(rng.choice([1, 7, 4095, 4096, 65536, 1 << 20]))))
    return reqs


def main():
    out = {}
    for i, (name, data, p) in enumerate(cases()):
        image = refwriter.write(data, p["codec"], p["level"], p["min_frame_size"], p["chunk"], p.get("strategy", 0),
                                p.get("nb_workers", 0))
        with open(os.path.join(HERE, name + ".zsk"), "wb") as f:
            f.write(image)
        with RefReader(image, cache_size=0) as rr:
            st = rr.stats()
            _, ent = refwriter.split(image)
            d_off = np.concatenate([[0], np.cumsum(ent[:, 1].astype(np.uint64))])
            total = int(st.decompressed_size)
            assert total == len(data)
            assert rr.pread_full(total, 0) == data
            reads = []
            for off, cnt in request_list(d_off, total, 100 + i):
                r, b = rr.pread(cnt, off)
                reads.append([off, cnt, r, hashlib.sha256(b).hexdigest()[:16]])
        out[name] = dict(params=p, input_sha256=hashlib.sha256(data).hexdigest(), input_len=len(data),
                         image_len=len(image), frames=int(st.frames), seek_table_memory=int(st.seek_table_memory),
                         reads=reads)
        print(name, len(data), "->", len(image), "frames", st.frames)
    # files the reference refuses to open (error text is part of the contract)
    empty = refwriter.write(b"", ZSTD, 3, 1 << 20, 4096)
    with open(os.path.join(HERE, "empty.zsk"), "wb") as f:
        f.write(empty)
    bad = {}
    for name, img in (("empty", empty), ("truncated_footer", open(os.path.join(HERE, "tiny_zstd.zsk"), "rb").read()[:-3]),
                      ("garbage", b"not a seekable file at all")):
        try:
            RefReader(img)
            bad[name] = None
        except OSError as e:
            bad[name] = str(e)
    out["_open_errors"] = bad
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(bad)


if __name__ == "__main__":
    main()
