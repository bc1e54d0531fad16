import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Compile the CPU checkers (oracle port, reference driver, and oracle/_ref when /root/reference exists)."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-s"], check=True)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        meta = json.load(f)
    cases = {}
    for name, m in meta.items():
        if name.startswith("_"):
            continue
        with open(os.path.join(GOLDEN, name + ".zsk"), "rb") as f:
            cases[name] = dict(m, image=f.read())
    return cases, meta["_open_errors"]


def golden_case_names():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return [k for k in json.load(f) if not k.startswith("_")]


@pytest.fixture(scope="session")
def lib():
    """The product library (built for sm_100a in-tree)."""
    import libzseek_b200 as z
    if not os.path.exists(z.LIB_PATH):
        z.build()
    return z


def sha16(b) -> str:
    import hashlib
    return hashlib.sha256(bytes(b)).hexdigest()[:16]
