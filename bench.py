#!/usr/bin/env python
"""bench.py — headline benchmark of the seekable-format READ path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size-gib G] [--tile-mib T]

Workload at every N (config.workload): BASELINE.json configs[1] — LZ4 level 0 seekable file of zsyn-v1
data, 64 KiB frames, G GiB per GPU (default 4; frame-range sharded, per-GPU work fixed => "weak"),
whole-file sequential decode.  A "step" = one pass of the hot path over that file.

  value    whole-job decompressed GB/s with the compressed image already resident in HBM and the
           output going to HBM; timed with CUDA events on the launching stream, max over ranks.
  e2e      same metric through the C-ABI with HOST buffers: compressed image in pinned host memory,
           zseek_b200_read_range into a pinned host buffer, H2D + decode + D2H inside the timed region.
  roofline the LZ4 decode kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json):
           algorithmic bytes = C + D per launch.
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, one reader per thread.
  extra    configs[2] (zstd level 3, 256 KiB frames) decode GB/s and the batched random 4 KiB pread
           rate (configs[3] shape, scaled to the zstd file that is resident), each with its own
           roofline fraction and CPU figure.

--impl reference times the reference's own CPU implementation of the same workload on all host
cores (rank 0 only) and prints the same line with "impl": "reference".

Inputs are produced by the reference CPU writer (north_star) with the tile-and-replicate construction
of SURVEY.md §8d: a T-MiB zsyn-v1 tile is written once, its compressed frames are replicated.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LZ4, ZSTD = 1, 0
GB = 1e9


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def profile_traffic(kernel):
    """Per-launch DRAM traffic from the committed ncu capture, if one exists for this kernel."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


# ----------------------------------------------------------------------------- inputs
def build_inputs(args, rank, world):
    """Returns dict name -> (image ndarray, total decompressed bytes, tile bytes)."""
    from datagen import refwriter, zsyn
    tile_bytes = args.tile_mib << 20
    total = int(args.size_gib * (1 << 30))
    reps = max(1, total // tile_bytes)
    cache = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir(),
                         f"zsk_bench_{os.environ.get('MASTER_PORT', 'solo')}_{args.tile_mib}")
    paths = {k: f"{cache}_{k}.zsk" for k in ("lz4", "zstd3")}
    done = cache + ".done"
    if rank == 0 and not os.path.exists(done):
        t0 = time.time()
        tile = zsyn.gen_parallel(tile_bytes)
        log(f"[bench] zsyn-v1 tile {args.tile_mib} MiB generated in {time.time() - t0:.1f}s")
        t0 = time.time()
        one = {"lz4": refwriter.write_parallel(tile, LZ4, 0, 65536, piece_frames=1024),
               "zstd3": refwriter.write_parallel(tile, ZSTD, 3, 262144, strategy=0, piece_frames=256)}
        log(f"[bench] reference writer: lz4 ratio {tile_bytes / len(one['lz4']):.3f}, zstd3 ratio "
            f"{tile_bytes / len(one['zstd3']):.3f} in {time.time() - t0:.1f}s")
        for k, img in one.items():
            with open(paths[k], "wb") as f:
                f.write(img)
        open(done, "w").write("ok")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    out = {}
    for k in ("lz4", "zstd3"):
        one = open(paths[k], "rb").read()
        img = refwriter.replicate(one, reps)
        out[k] = (np.frombuffer(img, dtype=np.uint8), reps * tile_bytes)
    return out, cache


def cleanup_inputs(cache, rank):
    if rank == 0:
        for suffix in ("_lz4.zsk", "_zstd3.zsk", ".done"):
            try:
                os.remove(cache + suffix)
            except OSError:
                pass


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up; only samples whose timestamp falls inside the
    timed region (mark_start .. mark_end) are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name).read().strip().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), float(c[3]), c[4:8]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        scope = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to the busiest samples of the run
            inside = sorted(rows, key=lambda r: -r[3])[:max(1, len(rows) // 4)]
            scope = "highest-power quarter of the run (timed region shorter than one sample)"
        sm = sorted(r[1] for r in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[4][j].startswith("Active") for r in inside)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": inside[0][2], "power_w_max": max(r[3] for r in inside),
                "samples": len(inside), "scope": scope, "reasons": reasons}


# ----------------------------------------------------------------------------- CPU (reference) legs
def cpu_scan(image, total, threads, repeats=3):
    from oracle.pyapi import RefDrive
    best = None
    for _ in range(repeats):
        t, nbytes = RefDrive.scan(image, total, threads, req=1 << 20, cache_size=0, pin=True)
        assert nbytes == total
        best = t if best is None else min(best, t)
    return best


def cpu_random(image, offsets, count, threads):
    from oracle.pyapi import RefDrive
    t, ops = RefDrive.random(image, offsets, count, threads, cache_size=0, pin=True)
    return t, ops


def gen_offsets(n, total, count, seed=1):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, total - count, n, dtype=np.uint64)


# ----------------------------------------------------------------------------- GPU arm
def run_b200(args, rank, world):
    import torch
    import torch.distributed as dist
    import libzseek_b200 as z

    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ["ZSEEK_B200_DEVICE"] = str(local)
    inputs, cache = build_inputs(args, rank, world)
    peak, peak_src = measured_peak()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    results = {}
    launches_total = 0
    clocks = None
    for name in ("lz4", "zstd3"):
        image, total = inputs[name]
        pinned = torch.from_numpy(image.copy()).pin_memory()       # compressed file image in pinned host memory
        rd = z.Reader(image=pinned, cache_size=0)
        C = int(rd.c_off[-1])
        dev_out = torch.empty(total + 64, dtype=torch.uint8, device="cuda")
        sampler = ClockSampler(local) if (name == "lz4" and rank == 0) else None
        rd.load(0, rd.frames)                                       # compressed image resident in HBM
        for _ in range(args.warmup):
            rd.decode_frames(0, rd.frames, dev_out)
        barrier()
        l0 = rd.launch_count
        if sampler:
            sampler.mark_start()
        rd.timer_start()
        wall0 = time.perf_counter()
        kernel_ms = 0.0
        for _ in range(args.steps):
            rd.decode_frames(0, rd.frames, dev_out)                 # inputs (C+D = 6 GB) >> L2, no flush needed
            kernel_ms += rd.last_decode_ms
        kernel_name = rd.last_decode_kernel
        dev_ms = rd.timer_stop()
        barrier()
        wall = time.perf_counter() - wall0
        if sampler:
            sampler.mark_end()
            clocks = sampler.stop()
        launches = rd.launch_count - l0
        dev_ms = max_over_ranks(dev_ms)
        # spot-check the bytes that were just timed against the source tile (full parity lives in tests/)
        chk = dev_out[:1 << 20].cpu().numpy()
        results[name] = dict(total=total, C=C, dev_ms=dev_ms, kernel_ms=max_over_ranks(kernel_ms), wall=wall, launches=launches,
                             frames=rd.frames, first_mib=chk, kernel=kernel_name)
        launches_total += launches if name == "lz4" else 0
        # ---- e2e: host buffers, H2D + decode + D2H inside the timed region (through zseek_b200_read_range)
        if name == "lz4":
            rd2 = z.Reader(image=pinned, cache_size=0)
            host_out = torch.empty(total, dtype=torch.uint8).pin_memory()
            for _ in range(2):
                rd2.unload()
                rd2.read_range_into(host_out, total, 0)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                rd2.unload()
                got = rd2.read_range_into(host_out, total, 0)
                assert got == total
            barrier()
            e2e_t = max_over_ranks(time.perf_counter() - t0)
            results[name]["e2e_s"] = e2e_t
            results[name]["host_first_mib"] = host_out[:1 << 20].numpy().copy()
            rd2.close()
            del host_out
        # ---- batched random 4 KiB preads over the zstd file (config 4 shape)
        if name == "zstd3":
            n_req = args.random_ops
            offs = gen_offsets(n_req, total, 4096)
            out = torch.empty(n_req * 4096, dtype=torch.uint8, device="cuda")
            rdc = z.Reader(image=pinned, cache_size=rd.frames)       # decoded-frame cache can hold the file
            rdc.load(0, rdc.frames)
            rdc.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)  # untimed: one-time allocation of the batch buffers
            rdc.cache_clear()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = rdc.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)   # cold: decodes every touched frame
            torch.cuda.synchronize()
            cold = time.perf_counter() - t0
            sample = out[:4096 * 64].cpu().numpy().copy()
            warm = []
            for _ in range(3):
                t0 = time.perf_counter()
                rdc.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)
                torch.cuda.synchronize()
                warm.append(time.perf_counter() - t0)
            lat = []
            for b in range(100):
                rdc.cache_clear()
                t0 = time.perf_counter()
                rdc.pread_batch(gen_offsets(10000, total, 4096, seed=100 + b), fixed_count=4096, dst=out, dst_stride=4096)
                torch.cuda.synchronize()
                lat.append(time.perf_counter() - t0)
            lat.sort()
            results["random"] = dict(n=n_req, cold_s=cold, warm_s=min(warm), p50_ms=lat[len(lat) // 2] * 1e3,
                                     p99_ms=lat[-1] * 1e3, short_reads=int((res < 4096).sum()), offs=offs,
                                     sample=sample, res=res[:64].copy())
            rdc.close()
            del out
        rd.close()
        del dev_out, pinned
        torch.cuda.empty_cache()

    # ---- CPU baseline (reference build) on rank 0, bounded sample
    cpu = {}
    if rank == 0:
        from oracle.pyapi import RefReader
        threads = os.cpu_count() or 1
        for name in ("lz4", "zstd3"):
            image, total = inputs[name]
            sample = min(total, (1 << 30) if name == "lz4" else (1 << 30))
            # sample = the first `sample` decompressed bytes of the same file
            if world == 1:   # the CPU legs are timed at N = 1 only (other ranks would share the host cores)
                t = cpu_scan(image, sample, threads)
                t1 = cpu_scan(image, min(sample, 256 << 20), 1, repeats=1)
                cpu[name] = dict(gbps=round(sample / t / GB, 3), gbps_1t=round(min(sample, 256 << 20) / t1 / GB, 3), threads=threads, sample=sample)
            else:
                cpu[name] = dict(gbps=None, gbps_1t=None, threads=0, sample=sample)
            # correctness spot-check of what was timed on the GPU, against the reference reader
            with RefReader(image) as rr:
                want = np.frombuffer(rr.pread_full(1 << 20, 0), dtype=np.uint8)
            assert (results[name]["first_mib"] == want).all(), f"{name}: GPU output differs from the reference"
            if "host_first_mib" in results[name]:
                assert (results[name]["host_first_mib"] == want).all(), f"{name}: e2e output differs from the reference"
        image, total = inputs["zstd3"]
        r = results["random"]
        n_cpu = min(r["n"], 20000)
        if world == 1:
            t, ops = cpu_random(image, r["offs"][:n_cpu], 4096, threads)
            cpu["random"] = dict(ops=round(ops / t), threads=threads, n=n_cpu)
        else:
            cpu["random"] = dict(ops=None, threads=0, n=n_cpu)
        with RefReader(image) as rr:
            for i in range(64):
                k, b = rr.pread(4096, int(r["offs"][i]))
                assert k == r["res"][i] and r["sample"][i * 4096:i * 4096 + k].tobytes() == b, "random batch differs from reference"

    if rank == 0:
        lz, zs, rn = results["lz4"], results["zstd3"], results["random"]
        units = lz["total"] * world * args.steps
        value = units / (lz["dev_ms"] / 1e3) / GB
        alg = (lz["C"] + lz["total"]) * args.steps
        achieved = alg / (lz["kernel_ms"] / 1e3) / GB
        zs_value = zs["total"] * world * args.steps / (zs["dev_ms"] / 1e3) / GB
        zs_ach = (zs["C"] + zs["total"]) * args.steps / (zs["kernel_ms"] / 1e3) / GB
        line = {
            "metric": "decompressed_GBps", "value": round(value, 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(lz["dev_ms"] / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: LZ4 level 0 seekable file, zsyn-v1, 64 KiB frames, "
                                   f"{args.size_gib:g} GiB per GPU, full sequential decode",
                       "frames_per_gpu": lz["frames"], "compressed_bytes_per_gpu": lz["C"], "decompressed_bytes_per_gpu": lz["total"],
                       "tile_mib": args.tile_mib, "sharding": f"frame-range x{world}, no collective",
                       "l2": "inputs (C+D per step) larger than L2; no flush needed", "timer": "CUDA events on the launching stream"},
            "clocks": clocks,
            "e2e": {"value": round(lz["total"] * world * args.steps / lz["e2e_s"] / GB, 2), "unit": "GB/s",
                    "h2d_bytes_per_step": lz["C"], "d2h_bytes_per_step": lz["total"],
                    "api": "zseek_b200_read_range, pinned host image -> pinned host buffer"},
            "gpu_launches": launches_total,
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": profile_traffic(lz["kernel"]),
                         "kernel": lz["kernel"], "algorithmic_bytes_per_launch": lz["C"] + lz["total"],
                         "peak_source": peak_src},
            "cpu_baseline": {"value": cpu["lz4"]["gbps"], "unit": "GB/s", "cores": cpu["lz4"]["threads"], "kind": "reference",
                             "sample": (f"first {cpu['lz4']['sample'] >> 20} MiB of the same file, one reader per thread, 1 MiB zseek_pread "
                                        f"requests, cache_size 0, best of 3; 1 thread: {cpu['lz4']['gbps_1t']} GB/s") if world == 1 else
                                       "timed at N = 1 only; see the --impl reference line of this N"},
            "extra": {
                "zstd3_256k": {"workload": "BASELINE configs[2]: zstd level 3, 256 KiB frames", "value": round(zs_value, 2), "unit": "GB/s",
                               "ms_per_step": round(zs["dev_ms"] / args.steps, 3), "frames": zs["frames"], "compressed_bytes": zs["C"],
                               "roofline": {"bound": "hbm", "achieved": round(zs_ach, 1), "peak": peak, "frac": round(zs_ach / peak, 4),
                                            "kernel": zs["kernel"], "traffic": profile_traffic(zs["kernel"])},
                               "cpu_baseline": {"value": cpu["zstd3"]["gbps"], "unit": "GB/s", "cores": cpu["zstd3"]["threads"],
                                                "kind": "reference", "one_thread": cpu["zstd3"]["gbps_1t"]}},
                "random_4k": {"workload": f"BASELINE configs[3] shape: {rn['n']} x 4 KiB zseek_pread requests, uniform byte offsets, over the "
                                          f"{zs['total'] >> 30} GiB zstd-3 file (rank 0)",
                              "ops_per_s_cold": round(rn["n"] / rn["cold_s"]), "ops_per_s_warm_cache": round(rn["n"] / rn["warm_s"]),
                              "batch_10k_cold_p50_ms": round(rn["p50_ms"], 3), "batch_10k_cold_p99_ms": round(rn["p99_ms"], 3),
                              "short_reads_at_frame_boundaries": rn["short_reads"],
                              "cpu_baseline": {"value": cpu["random"]["ops"], "unit": "ops/s", "cores": cpu["random"]["threads"],
                                               "kind": "reference", "sample": f"first {cpu['random']['n']} requests"}},
                "wall_s_lz4_timed_region": round(lz["wall"], 4), "kernel_ms_lz4_sum": round(lz["kernel_ms"], 3),
            },
        }
        emit(line)
    barrier()
    cleanup_inputs(cache, rank)


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    inputs, cache = build_inputs(args, 0, 1)
    image, total = inputs["lz4"]
    threads = os.cpu_count() or 1
    sample = min(total, 1 << 30)
    from oracle.pyapi import RefDrive
    for _ in range(args.warmup):
        RefDrive.scan(image, sample, threads, req=1 << 20, cache_size=0, pin=True)
    t0 = time.perf_counter()
    inner = 0.0
    for _ in range(args.steps):
        t, nbytes = RefDrive.scan(image, sample, threads, req=1 << 20, cache_size=0, pin=True)
        assert nbytes == sample
        inner += t
    wall = time.perf_counter() - t0
    value = sample * args.steps / inner / GB
    desc = (f"first {sample >> 20} MiB of the same {args.size_gib:g} GiB file per step, {threads} pinned threads, one reference reader "
            f"per thread over a RAM image (memcpy pread), 1 MiB zseek_pread requests, cache_size 0")
    line = {"impl": "reference", "metric": "decompressed_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(inner / args.steps * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: LZ4 level 0 seekable file, zsyn-v1, 64 KiB frames, {args.size_gib:g} GiB per GPU, "
                                   f"full sequential decode", "tile_mib": args.tile_mib},
            "cpu_baseline": {"value": round(value, 3), "unit": "GB/s", "cores": threads, "kind": "reference", "sample": desc},
            "e2e": {"value": round(value, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "extra": {"wall_s": round(wall, 3)}}
    emit(line)
    cleanup_inputs(cache, 0)


_REAL_STDOUT = None


def claim_stdout():
    """Libraries (NCCL prints its version banner) must not write to the stdout the driver parses: fd 1 is pointed
    at stderr for the whole run and the single JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def bind_to_gpu_numa_node(local):
    """N > 1: run this rank (and first-touch its pinned buffers) on the cores next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            log(f"[bench] local rank {local}: bound to {len(cpus)} cores near GPU {local}")
    except Exception as e:  # best effort
        log(f"[bench] local rank {local}: no NUMA binding ({e})")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size-gib", type=float, default=4.0)
    ap.add_argument("--tile-mib", type=int, default=512)
    ap.add_argument("--random-ops", type=int, default=1000000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        bind_to_gpu_numa_node(int(os.environ.get("LOCAL_RANK", 0)))
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    run_b200(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
