"""What can HBM deliver when every access is a random 32 / 64 / 128-byte read?  The lane-per-frame LZ4 kernel fetches ~330 M
match sources per 4 GiB from 65,536 different 64 KiB windows — none of them in L2 — so this, not the streaming copy bandwidth,
is the ceiling of its far loads.  Probe: torch.index_select of random rows of a 4 GiB table (a library gather kernel; probe only).

    python tools/dram_random_probe.py
"""
import json
import torch


def main():
    dev = "cuda"
    out = {}
    table_bytes = 4 << 30
    for row in (32, 64, 128):
        n_rows = table_bytes // row
        tab = torch.empty(n_rows, row // 4, dtype=torch.float32, device=dev).normal_()
        m = 64 << 20
        idx = torch.randint(0, n_rows, (m,), device=dev)
        dst = torch.empty(m, row // 4, dtype=torch.float32, device=dev)
        torch.index_select(tab, 0, idx, out=dst)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            torch.index_select(tab, 0, idx, out=dst)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[f"row_{row}B"] = {"reads_per_s": round(m / ms * 1e3), "read_GBps": round(m * row / ms / 1e6, 1), "ms": round(ms, 3)}
        del tab, dst, idx
    print(json.dumps(out))


if __name__ == "__main__":
    main()
