"""What bounds host-buffer reads when all GPUs of the box copy at once?  One process per GPU (torch.distributed.run): every
rank times bare pinned 1 GiB copies (H2D, D2H, both directions at once) ALONE (the other ranks idle at a barrier) and then
ALL TOGETHER.  Rank 0 prints one JSON object: per-rank GB/s alone, per-rank and aggregate GB/s together, the NUMA node /
CPU affinity of every GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe8.py
"""
import json
import os
import time


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << 30
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def run(kind, reps=4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return reps * n * (2 if kind == "both" else 1) / dt / 1e9

    res = {}
    for kind in ("h2d", "d2h", "both"):
        run(kind, 1)
        alone = 0.0
        for r in range(world):        # one rank at a time
            barrier()
            if r == rank:
                alone = run(kind)
            barrier()
        barrier()
        together = run(kind)          # everybody at once
        barrier()
        res[kind] = (alone, together)
    vals = torch.tensor([x for k in ("h2d", "d2h", "both") for x in res[k]], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, vals)
    else:
        allv = [vals]
    if rank == 0:
        out = {"gpus": world, "bytes_per_copy": n, "host_cores": os.cpu_count()}
        try:
            import pynvml
            pynvml.nvmlInit()
            out["gpu_cpu_affinity_words"] = [[int(w) for w in pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(i), (os.cpu_count() + 63) // 64)] for i in range(world)]
        except Exception as e:  # noqa: BLE001
            out["gpu_cpu_affinity_words"] = str(e)
        try:
            out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
        except OSError:
            out["numa_nodes"] = None
        for i, kind in enumerate(("h2d", "d2h", "both")):
            alone = [round(float(v[2 * i]), 1) for v in allv]
            tog = [round(float(v[2 * i + 1]), 1) for v in allv]
            out[kind] = {"alone_GBps_per_rank": alone, "together_GBps_per_rank": tog, "together_aggregate_GBps": round(sum(tog), 1),
                         "sum_of_alone_GBps": round(sum(alone), 1)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
