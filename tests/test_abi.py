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
                    for pat in ("zsk_oracle", "refdrive", "libzseek_ref", "from oracle", "import oracle", "emu_kernels"):
                        if pat in text:
                            bad.append((f, pat))
    assert not bad, bad
    needed = os.popen(f"ldd {os.path.join(ROOT, 'libzseek_b200', 'libzseek_b200.so')} 2>/dev/null").read()
    assert "zstd" not in needed and "lz4" not in needed  # not a recompile of libzstd/liblz4 either
