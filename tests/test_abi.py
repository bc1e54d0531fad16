"""The drop-in boundary: libzseek_b200.so loads and exports every symbol include/*.h declares
(no compute calls — this runs without a GPU), and fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    names = []
    for h in ("zseek.h", "zseek_b200.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"#define[^\n]*", "", text)
        names += re.findall(r"ZSEEK_EXPORT[^;(]*?\b(\w+)\s*\(", text)
    return names


def test_every_declared_symbol_is_exported(lib):
    L = ctypes.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert {"zseek_reader_open_full", "zseek_reader_open", "zseek_reader_close", "zseek_pread", "zseek_read",
            "zseek_reader_stats"} <= set(names)
    assert len(names) >= 16
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ but not exported"


def test_writer_symbols_are_not_provided(lib):
    """The write path stays the reference CPU writer (north_star); a caller needing it links both."""
    L = ctypes.CDLL(lib.LIB_PATH)
    for n in ("zseek_writer_open", "zseek_writer_open_full", "zseek_write", "zseek_writer_close", "zseek_writer_stats"):
        assert not hasattr(L, n)


def test_null_reader_quirks(lib):
    """reference src/decompress.c:809-812 (NULL reader -> 0, not -1), :359-362 (close NULL -> true), :840-848."""
    L = lib.load_library()
    err = ctypes.create_string_buffer(80)
    assert L.zseek_pread(None, None, 10, 0, None, err) == 0
    assert err.value == b"invalid reader"
    assert L.zseek_reader_close(None, None, err) is True
    assert L.zseek_reader_stats(None, None, err) is False and err.value == b"invalid reader"


def test_open_errors_precede_any_device_work(lib, golden):
    """Format errors carry the reference's message text with or without a GPU."""
    cases, errors = golden
    from conftest import GOLDEN
    for name, img in (("empty", open(os.path.join(GOLDEN, "empty.zsk"), "rb").read()),
                      ("truncated_footer", cases["tiny_zstd"]["image"][:-3]), ("garbage", b"not a seekable file at all")):
        with pytest.raises(lib.ZseekError) as e:
            lib.Reader(image=img)
        assert str(e.value) == errors[name]
    with pytest.raises(lib.ZseekError) as e:
        lib.Reader(image=b"ab")
    assert str(e.value) == "unexpected EOF"


def test_no_cpu_fallback_without_gpu(lib, golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    cases, _ = golden
    with pytest.raises(lib.ZseekError) as e:
        lib.Reader(image=cases["tiny_lz4"]["image"])
    assert "context creation failed" in str(e.value)


def test_product_never_touches_the_oracle():
    """Nothing under libzseek_b200/ or include/ may import, link or name the oracle / emulator."""
    bad = []
    for base in ("libzseek_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                    text = open(os.path.join(dp, f)).read()
                    for pat in ("zsk_oracle", "refdrive", "libzseek_ref", "from oracle", "import oracle", "emu_kernels", "hostemu"):
                        if pat in text:
                            bad.append((f, pat))
    assert not bad, bad
    needed = os.popen(f"ldd {os.path.join(ROOT, 'libzseek_b200', 'libzseek_b200.so')} 2>/dev/null").read()
    assert "zstd" not in needed and "lz4" not in needed  # not a recompile of libzstd/liblz4 either


def _open_verdict_b200(lib, image):
    """'ok' when the file was accepted (seek table parsed; without a GPU the open then stops at the device context, which
    comes AFTER every format check), else the error text."""
    try:
        rd = lib.Reader(image=image)
    except lib.ZseekError as e:
        return "ok" if "context creation failed" in str(e) else str(e)
    st = rd.stats()
    rd.close()
    return ("ok", int(st.frames), int(st.decompressed_size))


def test_open_verdicts_match_the_reference_on_mutated_seek_tables(lib, golden):
    """a2/a3/a13: magic sniff and seek-table validation (reference src/decompress.c:261-288, src/seek_table.c:112-176).
    Random byte edits and truncations in the file head and in the seek-table frame (skippable header, entries, footer):
    the product accepts exactly the files the reference accepts and fails with the reference's text otherwise."""
    from oracle.pyapi import RefReader, have_reference
    import numpy as np
    if not have_reference():
        pytest.skip("needs oracle/_ref (the reference build)")
    import torch
    if torch.cuda.is_available():
        # with a device the open goes on to size the HBM slab from the (mutated) entry sizes; an absurd dSize then fails
        # with "buffer creation failed" where the reference opens the file and fails at the first read of that frame
        pytest.skip("verdicts of the format checks are observed without a device")
    cases, _ = golden
    rng = np.random.Generator(np.random.PCG64(2026))
    checked = rejected = 0
    for name in ("tiny_zstd", "tiny_lz4", "mix_lz4", "zsyn_zstd3_128k", "zsyn_lz4_4k_chunks"):
        good = cases[name]["image"]
        n = int.from_bytes(good[-9:-5], "little")
        table = 8 + 8 * n + 9
        for trial in range(120):
            img = bytearray(good)
            kind = trial % 4
            if kind == 0:                                   # one byte somewhere in the seek-table frame
                img[len(img) - 1 - int(rng.integers(0, table))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:                                 # footer / header fields: frame count, descriptor, magics, size
                pos = int(rng.choice([len(img) - 9, len(img) - 8, len(img) - 5, len(img) - 4, len(img) - 1,
                                      len(img) - table, len(img) - table + 4, len(img) - table + 7]))
                img[pos] = int(rng.integers(0, 256))
            elif kind == 2:                                 # the codec magic
                img[int(rng.integers(0, 4))] ^= 1 << int(rng.integers(0, 8))
            else:                                           # truncated or extended tail
                cut = int(rng.integers(1, min(len(img) - 1, table + 20)))
                img = img[:-cut] if trial % 8 == 3 else img + bytes(rng.integers(0, 256, cut, dtype=np.uint8))
            img = bytes(img)
            try:
                rr = RefReader(img)
                st = rr.stats()
                want = ("ok", int(st.frames), int(st.decompressed_size))
                rr.close()
            except OSError as e:
                want = str(e)
                rejected += 1
            got = _open_verdict_b200(lib, img)
            if isinstance(want, tuple) and got == "ok":     # no GPU: accepted is all that can be observed
                got = want
            assert got == want, (name, trial, kind, got, want)
            checked += 1
    assert checked == 600 and rejected > 100
