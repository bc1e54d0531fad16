"""x1 (north_star: "NCCL over NVLink only to gather results to a single device when the caller asks for it"): the
gather paths on real GPUs.  Needs >= 2 GPUs in the box (skipped on the one-GPU boxes of the round-end run); the CPU
version of the same logic runs under gloo in tests/test_host_logic.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 8])
def test_gather_over_nccl(world):
    from oracle.pyapi import have_reference
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    if not have_reference():
        pytest.skip("inputs come from the reference writer (oracle/_ref)")
    port = 29700 + os.getpid() % 200
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_gather_worker.py")],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and f"MGPU_GATHER_OK {world}" in p.stdout, (p.stdout[-2000:], p.stderr[-4000:])
