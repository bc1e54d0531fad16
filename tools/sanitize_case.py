"""Smallest end-to-end case for compute-sanitizer: decode two golden files (LZ4 + zstd) on the GPU and compare."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import libzseek_b200 as z
    gold = os.path.join(ROOT, "tests", "golden")
    meta = json.load(open(os.path.join(gold, "golden.json")))
    for name in sys.argv[1:] or ["zsyn_lz4_256k_linked", "mix_lz4", "zsyn_zstd19_256k", "mix_zstd3"]:
        image = open(os.path.join(gold, name + ".zsk"), "rb").read()
        with z.Reader(image=image, cache_size=4) as rd:
            dev = torch.empty(rd.size, dtype=torch.uint8, device="cuda")
            rd.decode_frames(0, rd.frames, dev)
            ok = hashlib.sha256(dev.cpu().numpy().tobytes()).hexdigest() == meta[name]["input_sha256"]
            offs = np.arange(0, rd.size, 4099, dtype=np.uint64)
            out = torch.zeros(len(offs) * 512, dtype=torch.uint8, device="cuda")
            rd.pread_batch(offs, fixed_count=512, dst=out, dst_stride=512)
            print(name, "ok" if ok else "MISMATCH", flush=True)
            assert ok


if __name__ == "__main__":
    main()
