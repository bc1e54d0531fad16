"""zsyn-v1: deterministic synthetic compressible corpus.

Vocabulary of 8,192 pseudo-words (letters drawn with p ∝ 1/rank), four segment kinds drawn with
p = {0.55, 0.30, 0.10, 0.05}: Zipf(1.1) word stream, log records, little-endian int32 random walk,
uniform random bytes.  `gen(n)` is the reference single-stream generator (sha256 of the first
64 MiB = 7154c9d1…bd50a3a under numpy 2.3); `gen_parallel` builds a large tile from independently
seeded chunks in worker processes (same shape, different bytes) so multi-GiB inputs can be produced
on the GPU box within seconds.
"""
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

SEED = 20261018


def gen(nbytes: int, seed: int = SEED) -> bytes:
    rng = np.random.Generator(np.random.PCG64(seed))
    V = 8192
    alpha = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    pa = 1.0 / np.arange(1, 27)
    pa /= pa.sum()
    words = [bytes(rng.choice(alpha, size=l, p=pa)) + b" " for l in rng.integers(2, 13, size=V)]
    pw = 1.0 / np.arange(1, V + 1) ** 1.1
    pw /= pw.sum()
    tmpl = b"2026-10-18T%02d:%02d:%02d.%03dZ host-%03d svc=%s level=%s req=%08x latency_us=%d status=%d msg=\"%s\"\n"
    svcs = [b"gateway", b"auth", b"storage", b"index", b"scheduler"]
    lvls = [b"INFO", b"INFO", b"INFO", b"WARN", b"DEBUG"]
    out = bytearray()
    kinds = rng.choice(4, size=1 << 20, p=[0.55, 0.30, 0.10, 0.05])
    k = 0
    while len(out) < nbytes:
        kind = kinds[k % len(kinds)]
        k += 1
        if kind == 0:  # Zipf word stream, ~48 KiB
            out += b"".join(words[i] for i in rng.choice(V, size=8192, p=pw))
        elif kind == 1:  # log records, ~48 KiB
            r = rng.integers(0, 1 << 31, size=(400, 8))
            for a in r:
                m = b"".join(words[i] for i in rng.choice(64, size=6))
                out += tmpl % (a[0] % 24, a[1] % 60, a[2] % 60, a[3] % 1000, a[4] % 40, svcs[a[5] % 5],
                               lvls[a[6] % 5], a[7], a[0] % 90000, 200 + (a[1] % 7 == 0) * 304, m)
        elif kind == 2:  # int32 random walk, 32 KiB
            out += np.cumsum(rng.integers(-64, 65, size=8192, dtype=np.int32)).astype("<i4").tobytes()
        else:  # incompressible, 16 KiB
            out += rng.integers(0, 256, size=16384, dtype=np.uint8).tobytes()
    return bytes(out[:nbytes])


def _chunk(args):
    n, seed = args
    return gen(n, seed)


def gen_parallel(nbytes: int, chunk: int = 16 << 20, workers: int | None = None, seed: int = SEED) -> bytes:
    """Concatenation of independently seeded `chunk`-byte pieces (seed + piece index)."""
    if nbytes <= chunk:
        return gen(nbytes, seed)
    jobs = []
    off = 0
    i = 0
    while off < nbytes:
        n = min(chunk, nbytes - off)
        jobs.append((n, seed + i))
        off += n
        i += 1
    import os
    workers = workers or min(len(jobs), os.cpu_count() or 1)
    with ProcessPoolExecutor(max_workers=workers) as ex:
        parts = list(ex.map(_chunk, jobs))
    return b"".join(parts)


if __name__ == "__main__":
    open(sys.argv[2], "wb").write(gen(int(sys.argv[1])))
