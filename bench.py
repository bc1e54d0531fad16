#!/usr/bin/env python
"""bench.py — headline benchmark of the seekable-format READ path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size-gib G] [--tile-mib T]

Workload at every N (config.workload): BASELINE.json configs[1] — LZ4 level 0 seekable file of zsyn-v1
data, 64 KiB frames, G GiB per GPU (default 4; frame-range sharded, per-GPU work fixed => "weak"),
whole-file sequential decode.  A "step" = one pass of the hot path over that file.

The corpus is UNIQUE data by default (tile = file size: no frame occurs twice); --tile-mib T builds the
file from a T-MiB tile instead (SURVEY.md §8d allows tiles >= 1 GiB), and at N = 1 the line carries an
A/B of the headline on a 512 MiB tile x 8 (`extra.replica_ab`) so the effect of replicas on the
size-ordered job list of the lane-per-frame LZ4 kernel is on record.

  value    whole-job decompressed GB/s with the compressed image already resident in HBM and the
           output going to HBM; timed with CUDA events on the launching stream, max over ranks.
  e2e      same metric through the C-ABI with HOST buffers: compressed image in pinned host memory,
           zseek_b200_read_range into a pinned host buffer, H2D + decode + D2H inside the timed region.
  roofline the LZ4 decode kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json):
           algorithmic bytes = C + D per launch.
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, one reader per thread;
           warm-up pass, then median (value) and best of 5.
  verified_bytes  every byte the timed loops produced is compared with the writer's input after the
           loop (device compare for HBM outputs, upload + device compare for the e2e host buffer);
           all 1 M random results are checked (lengths against B1 arithmetic, bytes on the device).
  extra    configs[2] (zstd-3, 256 KiB frames; value, e2e, roofline), configs[3] at its stated shape
           (1 M x 4 KiB over a 16 GiB zstd-3 file of 65,536 frames, rank 0), configs[4] (1 MiB frames,
           LZ4 + zstd-19, 8 GiB per GPU, every rank), the replica A/B, and the same-run bare pinned-copy
           ceiling the e2e figures are measured against.

--impl reference times the reference's own CPU implementation of the same workload on all host
cores (rank 0 only; each step decodes the first 1 GiB of the same file) and prints the same line with
"impl": "reference".

Inputs are produced by the reference CPU writer (north_star) once per box and kept under /dev/shm for
the other arm / the other N of the same round.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LZ4, ZSTD = 1, 0
GB = 1e9
MIB = 1 << 20


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
    """Per-launch DRAM traffic of this kernel at the bench shape from the committed ncu capture (profiles/traffic.json)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


def workload_config(args):
    """config of the JSON line — identical in both arms."""
    return {"workload": f"BASELINE configs[1]: LZ4 level 0 seekable file, zsyn-v1, 64 KiB frames, "
                        f"{args.size_gib:g} GiB per GPU, full sequential decode",
            "tile_mib": args.tile_mib, "unique_data": args.tile_mib * MIB >= int(args.size_gib * (1 << 30))}


# ----------------------------------------------------------------------------- inputs
class Corpus:
    """Files of one round under /dev/shm (or the temp dir): the raw tile and the reference writer's images of it."""
    C5_LZ4_TILE_MIB = 1024     # configs[4]: 1 MiB frames; LZ4 over the first 1 GiB of the tile, zstd-19 over the first 256 MiB
    C5_ZSTD_TILE_MIB = 256

    def __init__(self, args):
        self.args = args
        self.tile = args.tile_mib * MIB
        self.total = int(args.size_gib * (1 << 30))
        assert self.total % self.tile == 0, "--size-gib must be a multiple of --tile-mib"
        self.reps = self.total // self.tile
        need = int(self.tile * 2.2) + (2 << 30)
        base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need else tempfile.gettempdir()
        self.dir = os.path.join(base, f"zsk_bench_r2_{args.tile_mib}")
        self.c5_lz4_tile = min(self.tile, self.C5_LZ4_TILE_MIB * MIB)
        self.c5_zstd_tile = min(self.tile, self.C5_ZSTD_TILE_MIB * MIB)

    def path(self, name):
        return os.path.join(self.dir, name)

    def build(self, need_all=True):
        """rank 0 only; idempotent."""
        from datagen import refwriter, zsyn
        if os.path.exists(self.path("done_all" if need_all else "done_lz4")):
            return
        os.makedirs(self.dir, exist_ok=True)

        def put(name, data):
            with open(self.path(name + ".tmp"), "wb") as f:
                f.write(data)
            os.replace(self.path(name + ".tmp"), self.path(name))

        t0 = time.time()
        if os.path.exists(self.path("raw.bin")):
            tile = open(self.path("raw.bin"), "rb").read()
        else:
            tile = zsyn.gen_parallel(self.tile)
            put("raw.bin", tile)
        log(f"[bench] zsyn-v1 tile {self.args.tile_mib} MiB ready in {time.time() - t0:.1f}s")
        if not os.path.exists(self.path("lz4.zsk")):
            t0 = time.time()
            img = refwriter.write_parallel(tile, LZ4, 0, 65536, piece_frames=1024)
            put("lz4.zsk", img)
            log(f"[bench] reference writer lz4 64 KiB frames: ratio {self.tile / len(img):.3f} in {time.time() - t0:.1f}s")
        open(self.path("done_lz4"), "w").write("ok")
        if not need_all:
            return
        jobs = [("zstd3.zsk", ZSTD, 3, 262144, self.tile, 256), ("lz4_1m.zsk", LZ4, 0, MIB, self.c5_lz4_tile, 16),
                ("zstd19_1m.zsk", ZSTD, 19, MIB, self.c5_zstd_tile, 1)]
        for name, codec, level, frame, nbytes, piece in jobs:
            if os.path.exists(self.path(name)):
                continue
            t0 = time.time()
            img = refwriter.write_parallel(tile[:nbytes], codec, level, frame, strategy=0, piece_frames=piece)
            put(name, img)
            log(f"[bench] reference writer {name}: ratio {nbytes / len(img):.3f} in {time.time() - t0:.1f}s")
        open(self.path("done_all"), "w").write("ok")

    def image(self, name, reps):
        from datagen import refwriter
        one = open(self.path(name), "rb").read()
        return np.frombuffer(one if reps == 1 else refwriter.replicate(one, reps), dtype=np.uint8)

    def raw(self):
        return np.memmap(self.path("raw.bin"), dtype=np.uint8, mode="r")


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
def cpu_scan_stats(image, sample, threads, runs=5):
    """Warm-up pass, then `runs` timed passes of the reference over the first `sample` decompressed bytes:
    -> (median GB/s, best GB/s)."""
    from oracle.pyapi import RefDrive
    RefDrive.scan(image, sample, threads, req=1 << 20, cache_size=0, pin=True)
    ts = []
    for _ in range(runs):
        t, nbytes = RefDrive.scan(image, sample, threads, req=1 << 20, cache_size=0, pin=True)
        assert nbytes == sample
        ts.append(t)
    ts.sort()
    return sample / ts[len(ts) // 2] / GB, sample / ts[0] / GB


def cpu_leg(corpus, fname):
    """The reference on all host cores (and on one) over the first GiB of one corpus file: warm-up, median and best of 5."""
    image = corpus.image(fname, 1)
    threads = os.cpu_count() or 1
    sample = min(corpus.tile, 1 << 30)
    med, best = cpu_scan_stats(image, sample, threads)
    med1, _ = cpu_scan_stats(image, min(sample, 256 * MIB), 1, runs=3)
    return dict(gbps=round(med, 3), best=round(best, 3), gbps_1t=round(med1, 3), threads=threads, sample=sample)


def cpu_leg_subprocess(args, fname):
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", fname, "--size-gib", str(args.size_gib), "--tile-mib", str(args.tile_mib)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if p.returncode != 0:
        raise RuntimeError(p.stderr[-300:])
    return json.loads(p.stdout.strip().splitlines()[-1])


def step_stats(ms):
    """best / median / worst of the per-step device times of one leg (rank 0)."""
    s = sorted(ms)
    return {"best": round(s[0], 3), "median": round(s[len(s) // 2], 3), "worst": round(s[-1], 3), "n": len(s)}


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
    corpus = Corpus(args)
    if rank == 0:
        corpus.build()
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

    def min_over_ranks(x):
        return -max_over_ranks(-x)

    barrier()
    tile = corpus.tile
    raw_dev = torch.from_numpy(np.ascontiguousarray(corpus.raw())).cuda()       # the writer's input, for verification

    def verify_device(dev, nbytes, period):
        """bytes of dev[:nbytes] equal to the (period-tiled) writer input; compares everything."""
        ok = 0
        step = 256 * MIB
        for o in range(0, nbytes, step):
            n = min(step, nbytes - o)
            p = o % period
            assert p + n <= period, "compare chunks must not straddle the tile (tiles are multiples of 256 MiB, or the file is one tile)"
            if bool((dev[o:o + n] == raw_dev[p:p + n]).all()):
                ok += n
        return ok

    def pcie_ceiling():
        """bare pinned copies of 1 GiB, same run: (H2D GB/s, D2H GB/s)"""
        n = 1 << 30
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        d = torch.empty(n, dtype=torch.uint8, device="cuda")
        out = []
        for dst, src in ((d, h), (h, d)):
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            out.append(3 * n / (time.perf_counter() - t0) / GB)
        return out

    def decode_leg(image, total, period, steps, warmup, sampler=None):
        """HBM -> HBM whole-file decode: CUDA-event time of `steps` passes (max over ranks), everything verified."""
        pinned = torch.from_numpy(np.ascontiguousarray(image)).pin_memory()
        rd = z.Reader(image=pinned, cache_size=0)
        C = int(rd.c_off[-1])
        dev_out = torch.zeros(total + 64, dtype=torch.uint8, device="cuda")
        rd.load(0, rd.frames)
        for _ in range(warmup):
            rd.decode_frames(0, rd.frames, dev_out)
        dev_out.zero_()                                                       # the verified bytes are those of the timed passes
        barrier()
        l0 = rd.launch_count
        if sampler:
            sampler.mark_start()
        rd.timer_start()
        kernel_ms, step_ms = 0.0, []
        for _ in range(steps):
            rd.decode_frames(0, rd.frames, dev_out)                           # inputs (C + D per step) >> L2: no flush needed
            kernel_ms += rd.last_decode_ms
            step_ms.append(float(rd.last_decode_ms))                          # CUDA events around this step's launches
        dev_ms = rd.timer_stop()
        barrier()
        if sampler:
            sampler.mark_end()
        res = dict(total=total, C=C, frames=rd.frames, dev_ms=max_over_ranks(dev_ms), kernel_ms=max_over_ranks(kernel_ms),
                   launches=rd.launch_count - l0, kernel=rd.last_decode_kernel, step_ms=step_ms,
                   verified=int(min_over_ranks(verify_device(dev_out, total, period))))
        return res, rd, pinned, dev_out

    def e2e_leg(pinned, total, period, steps, scratch_dev):
        """host image -> host buffer through zseek_b200_read_range; H2D + decode + D2H inside the timed region."""
        rd2 = z.Reader(image=pinned, cache_size=0)
        host_out = torch.zeros(total, dtype=torch.uint8).pin_memory()
        for _ in range(2):
            rd2.unload()
            rd2.read_range_into(host_out, total, 0)
        host_out.zero_()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rd2.unload()
            got = rd2.read_range_into(host_out, total, 0)
            assert got == total
        barrier()
        t = max_over_ranks(time.perf_counter() - t0)
        scratch_dev[:total].copy_(host_out, non_blocking=False)               # verify what arrived in the host buffer
        ok = int(min_over_ranks(verify_device(scratch_dev, total, period)))
        rd2.close()
        del host_out
        return t, ok

    results, clocks = {}, None
    # ---------------- configs[1] (headline) and configs[2]
    for name, fname in (("lz4", "lz4.zsk"), ("zstd3", "zstd3.zsk")):
        image = corpus.image(fname, corpus.reps)
        sampler = ClockSampler(local) if (name == "lz4" and rank == 0) else None
        res, rd, pinned, dev_out = decode_leg(image, corpus.total, tile, args.steps, args.warmup, sampler)
        if sampler:
            clocks = sampler.stop()
        rd.close()
        res["e2e_s"], res["e2e_verified"] = e2e_leg(pinned, corpus.total, tile, args.steps, dev_out)
        results[name] = res
        log(f"[bench] {name}: {res['total'] * args.steps / res['dev_ms'] / 1e6:.1f} GB/s per GPU, e2e "
            f"{res['total'] * args.steps / res['e2e_s'] / GB:.1f} GB/s, verified {res['verified']}/{res['total']} + {res['e2e_verified']}")
        if name == "lz4" and world == 1 and corpus.tile > 512 * MIB:
            # A/B: the same kernel on a 512 MiB tile x (size / 512 MiB) — replicas of a frame sit next to each other in the
            # size-ordered job list and run in lock-step in one warp of the lane-per-frame kernel
            from datagen import refwriter
            payload, ent = refwriter.split(image.tobytes())
            nfr = (512 * MIB) // 65536
            small = payload[:int(ent[:nfr, 0].sum())] + refwriter.seek_table(ent[:nfr])
            ab_img = np.frombuffer(refwriter.replicate(small, corpus.total // (512 * MIB)), dtype=np.uint8)
            del payload
            del dev_out, pinned
            ab, rd_ab, p_ab, o_ab = decode_leg(ab_img, corpus.total, 512 * MIB, args.steps, 2)
            rd_ab.close()
            results["replica_ab"] = ab
            del p_ab, o_ab, ab_img
        else:
            del dev_out, pinned
        del image
        torch.cuda.empty_cache()
    pcie = pcie_ceiling() if rank == 0 else None

    # ---------------- configs[4]: 1 MiB frames, LZ4 + zstd-19, 8 GiB per GPU, every rank its own shard
    c5_total = int(args.c5_gib * (1 << 30))
    for name, fname, period in (("lz4_1m", "lz4_1m.zsk", corpus.c5_lz4_tile), ("zstd19_1m", "zstd19_1m.zsk", corpus.c5_zstd_tile)):
        if c5_total <= 0:
            break
        image = corpus.image(fname, max(1, c5_total // period))
        res, rd, pinned, dev_out = decode_leg(image, (c5_total // period) * period if c5_total >= period else period, period, 3, 2)
        rd.close()
        results[name] = res
        log(f"[bench] configs[4] {name}: {res['total'] * 3 / res['dev_ms'] / 1e6:.1f} GB/s per GPU, verified {res['verified']}/{res['total']}")
        del dev_out, pinned, image
        torch.cuda.empty_cache()

    # ---------------- x1: gather of the decoded shards to rank 0 over NVLink (N > 1): the 4 GiB LZ4 file sharded by frame range
    if world > 1:
        from libzseek_b200.sharding import decode_and_gather, gather_to, shard_range
        image = corpus.image("lz4.zsk", corpus.reps)
        pinned = torch.from_numpy(np.ascontiguousarray(image)).pin_memory()
        rd = z.Reader(image=pinned, cache_size=0)
        rd.set_shard(rank, world)
        lo, hi = shard_range(rd.frames, rank, world)
        d_off = rd.d_off
        nbytes = int(d_off[hi] - d_off[lo])
        rd.load(lo, hi)
        whole = torch.zeros(corpus.total + 64, dtype=torch.uint8, device="cuda") if rank == 0 else None
        local_buf = torch.zeros(nbytes + 64, dtype=torch.uint8, device="cuda")
        chunk = 128 * MIB

        def decode_chunk(f0, f1, view):
            rd.decode_frames(f0, f1, view)

        def timed(fn, reps=3):
            fn()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            e1.synchronize()
            barrier()
            return max_over_ranks(e0.elapsed_time(e1) / reps)

        from libzseek_b200.sharding import chunk_plan
        plan = chunk_plan(d_off, lo, hi, chunk)
        base = int(d_off[lo])

        def decode_only():
            for f0, f1 in plan:
                decode_chunk(f0, f1, local_buf[int(d_off[f0]) - base:int(d_off[f1]) - base])

        def gather_only():
            gather_to(local_buf[:nbytes], d_off, dst_rank=0, out=whole)

        def pipeline():
            decode_and_gather(decode_chunk, d_off, out=whole, local=local_buf, dst_rank=0, chunk_bytes=chunk)

        t_dec, t_gat, t_pipe = timed(decode_only), timed(gather_only), timed(pipeline)
        if rank == 0:
            whole.zero_()
        pipeline()
        torch.cuda.synchronize()
        barrier()
        ok = verify_device(whole, corpus.total, tile) if rank == 0 else 0
        moved = corpus.total - (int(d_off[shard_range(rd.frames, 0, world)[1]]) - int(d_off[shard_range(rd.frames, 0, world)[0]]))
        results["gather"] = dict(decode_ms=t_dec, gather_ms=t_gat, pipeline_ms=t_pipe, moved=moved, verified=ok, total=corpus.total)
        rd.close()
        del whole, local_buf, pinned, image
        torch.cuda.empty_cache()

    # ---------------- configs[3]: 1 M x 4 KiB random preads over a 16 GiB zstd-3 file (rank 0)
    if rank == 0 and args.random_ops > 0:
        rtotal = int(args.random_gib * (1 << 30))
        rreps = max(1, rtotal // tile)
        rtotal = rreps * tile
        image = corpus.image("zstd3.zsk", rreps)
        pinned = torch.from_numpy(np.ascontiguousarray(image)).pin_memory()
        n_req = args.random_ops
        offs = gen_offsets(n_req, rtotal, 4096)
        out = torch.zeros(n_req * 4096, dtype=torch.uint8, device="cuda")
        rdc = z.Reader(image=pinned, cache_size=1 << 30)                         # the decoded-frame cache can hold the file
        d_off = np.asarray(rdc.d_off, dtype=np.uint64)
        rdc.load(0, rdc.frames)
        rdc.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)         # untimed, full size: batch buffers and scratch pools exist afterwards (cache-cold, not allocator-cold)
        rdc.cache_clear()
        out.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = rdc.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)   # cold: decodes every touched frame
        torch.cuda.synchronize()
        cold = time.perf_counter() - t0
        # every request: length by B1 arithmetic, bytes against the writer's input (device gather)
        ends = d_off[np.searchsorted(d_off, offs, side="right")]
        want = np.minimum(np.uint64(4096), ends - offs).astype(np.int64)
        len_ok = int((res == want).sum())
        ar = torch.arange(4096, device="cuda", dtype=torch.int64)
        bytes_ok, reqs_ok = 0, 0
        for o in range(0, n_req, 8192):
            k = min(8192, n_req - o)
            so = torch.from_numpy(offs[o:o + k].astype(np.int64)).cuda()
            ln = torch.from_numpy(want[o:o + k]).cuda()
            exp = raw_dev[(so[:, None] + ar[None, :]) % tile]
            got = out[o * 4096:(o + k) * 4096].view(k, 4096)
            good = ((exp == got) | (ar[None, :] >= ln[:, None])).all(dim=1)
            reqs_ok += int(good.sum())
            bytes_ok += int(ln[good].sum())
        warm = []
        for _ in range(3):
            t0 = time.perf_counter()
            rdc.pread_batch(offs, fixed_count=4096, dst=out, dst_stride=4096)
            torch.cuda.synchronize()
            warm.append(time.perf_counter() - t0)
        lat = []
        for b in range(100):
            rdc.cache_clear()
            o10 = gen_offsets(10000, rtotal, 4096, seed=100 + b)
            t0 = time.perf_counter()
            rdc.pread_batch(o10, fixed_count=4096, dst=out, dst_stride=4096)
            torch.cuda.synchronize()
            lat.append(time.perf_counter() - t0)
        lat.sort()
        # the same cold 10,000-request batches through the stream-ordered API: request arrays resident in HBM, no host round trip
        d_batches = [torch.from_numpy(gen_offsets(10000, rtotal, 4096, seed=100 + b).astype(np.int64)).cuda() for b in range(100)]
        d_res = torch.zeros(10000, dtype=torch.int64, device="cuda")
        rdc.pread_batch_async(d_batches[0], out, fixed_count=4096, dst_stride=4096, dev_results=d_res)
        rdc.batch_wait()
        alat = []
        for b in range(100):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rdc.pread_batch_async(d_batches[b], out, fixed_count=4096, dst_stride=4096, dev_results=d_res)
            rdc.batch_wait()
            alat.append(time.perf_counter() - t0)
        alat.sort()
        # check the last async batch completely against the writer's input
        o10 = d_batches[99]
        ends10 = torch.from_numpy(d_off[np.searchsorted(d_off, o10.cpu().numpy().astype(np.uint64), side="right")].astype(np.int64)).cuda()
        want10 = torch.minimum(torch.full_like(o10, 4096), ends10 - o10)
        exp = raw_dev[(o10[:, None] + ar[None, :]) % tile]
        got = out[:10000 * 4096].view(10000, 4096)
        async_ok = int((((exp == got) | (ar[None, :] >= want10[:, None])).all(dim=1) & (d_res == want10)).sum())
        del d_batches
        results["random"] = dict(n=n_req, total=rtotal, frames=rdc.frames, cold_s=cold, warm_s=min(warm), p50_ms=lat[len(lat) // 2] * 1e3,
                                 p99_ms=lat[98] * 1e3, async_p50_ms=alat[len(alat) // 2] * 1e3, async_p99_ms=alat[98] * 1e3, async_ok=async_ok, short_reads=int((res < 4096).sum()), lengths_ok=len_ok, requests_ok=reqs_ok,
                                 verified_bytes=bytes_ok, expected_bytes=int(want.sum()), C=int(rdc.c_off[-1]))
        rdc.close()
        del out, pinned
        torch.cuda.empty_cache()
        log(f"[bench] random: cold {n_req / cold / 1e6:.2f} M ops/s, warm {n_req / min(warm) / 1e6:.1f} M ops/s, p50 {results['random']['p50_ms']:.2f} ms, "
            f"{reqs_ok}/{n_req} requests verified")
        # CPU figure for the same request list (bounded sample)
        if world == 1:
            from oracle.pyapi import RefDrive
            threads = os.cpu_count() or 1
            n_cpu = min(n_req, 20000)
            t, ops = RefDrive.random(image, offs[:n_cpu], 4096, threads, cache_size=0, pin=True)
            results["random"]["cpu"] = dict(ops=round(ops / t), threads=threads, n=n_cpu)
        del image

    # ---------------- CPU baseline (the reference build) on rank 0, N = 1 only (other ranks would share the host cores)
    cpu = {}
    if rank == 0 and world == 1:
        # The multi-thread figure of these VMs is bimodal between invocations seconds apart (LZ4, 16 threads, same box:
        # 37 / 45 / 53 GB/s; within one invocation median ~ best), whichever process runs it (profiles/README.md,
        # r02_final*).  So the leg is timed twice — in a process of its own (the harness of --impl reference) and in
        # this process — 5 passes each after a warm-up, and `value` is the BETTER median: the CPU gets its best showing.
        for name, fname in (("lz4", "lz4.zsk"), ("zstd3", "zstd3.zsk")):
            legs = []
            try:
                legs.append(dict(cpu_leg_subprocess(args, fname), process="own process"))
            except Exception as e:  # noqa: BLE001
                log(f"[bench] CPU leg {name} in a subprocess failed ({e})")
            legs.append(dict(cpu_leg(corpus, fname), process="bench process"))
            cpu[name] = dict(max(legs, key=lambda c: c["gbps"]))
            cpu[name]["all"] = [{"timed_in": c["process"], "median": c["gbps"], "best": c["best"]} for c in legs]

    dropin = None
    if rank == 0 and world == 1 and args.dropin_mib > 0:
        # the six-symbol drop-in path under the caller's own threads, both libraries, same C harness (tools/dropin_bench.py)
        try:
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dropin_bench.py"), str(args.dropin_mib)], capture_output=True, text=True, timeout=900)
            dropin = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-300:]}
        except Exception as e:  # noqa: BLE001
            dropin = {"error": str(e)[:300]}
    if rank == 0:
        lz, zs = results["lz4"], results["zstd3"]

        def gbps(r, steps=args.steps, key="dev_ms"):
            return r["total"] * world * steps / (r[key] / 1e3) / GB

        def roof(r, steps=args.steps):
            ach = (r["C"] + r["total"]) * steps / (r["kernel_ms"] / 1e3) / GB
            return {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                    "frac_of_nominal_8000": round(ach / 8000.0, 4),
                    "traffic": profile_traffic(r["kernel"]), "kernel": r["kernel"], "algorithmic_bytes_per_launch": r["C"] + r["total"]}

        config = workload_config(args)
        config.update({"frames_per_gpu": lz["frames"], "compressed_bytes_per_gpu": lz["C"], "decompressed_bytes_per_gpu": lz["total"],
                       "sharding": f"frame-range x{world}, no collective",
                       "l2": "inputs (C+D per step) larger than L2; no flush needed", "timer": "CUDA events on the launching stream"})
        extra = {
            "zstd3_256k": {"workload": "BASELINE configs[2]: zstd level 3, 256 KiB frames, same size", "value": round(gbps(zs), 2), "unit": "GB/s",
                           "ms_per_step": round(zs["dev_ms"] / args.steps, 3), "frames": zs["frames"], "compressed_bytes": zs["C"],
                           "verified_bytes": zs["verified"], "roofline": roof(zs),
                           "e2e": {"value": round(zs["total"] * world * args.steps / zs["e2e_s"] / GB, 2), "unit": "GB/s",
                                   "verified_bytes": zs["e2e_verified"], "h2d_bytes_per_step": zs["C"], "d2h_bytes_per_step": zs["total"]},
                           "cpu_baseline": ({"value": cpu["zstd3"]["gbps"], "best": cpu["zstd3"]["best"], "unit": "GB/s", "cores": cpu["zstd3"]["threads"],
                                             "kind": "reference", "one_thread": cpu["zstd3"]["gbps_1t"], "timed_in": cpu["zstd3"]["process"],
                                             "all_timings": cpu["zstd3"]["all"]} if cpu else None)},
            "kernel_ms_lz4_sum": round(lz["kernel_ms"], 3),
            # SURVEY §8d "best and median of >= 5": per-step kernel times of this rank (events around each step's launches)
            "step_ms_lz4": step_stats(lz["step_ms"]), "step_ms_zstd3": step_stats(zs["step_ms"]),
        }
        if dropin is not None:
            extra["dropin_plain_zseek_pread"] = dropin
        if pcie:
            extra["pinned_copy_ceiling"] = {"h2d_GBps": round(pcie[0], 1), "d2h_GBps": round(pcie[1], 1),
                                            "e2e_lz4_fraction_of_d2h": round(lz["total"] * args.steps / lz["e2e_s"] / GB / pcie[1], 3),
                                            "note": "bare 1 GiB pinned copies on rank 0, same run; a host-buffer read cannot beat the D2H figure"}
        if "replica_ab" in results:
            ab = results["replica_ab"]
            extra["replica_ab"] = {"tile_mib": 512, "value": round(gbps(ab), 2), "unit": "GB/s", "ms_per_step": round(ab["dev_ms"] / args.steps, 3),
                                   "verified_bytes": ab["verified"], "note": "same kernel and size, file = 512 MiB tile replicated; the headline uses unique data"}
        for k in ("lz4_1m", "zstd19_1m"):
            if k in results:
                r = results[k]
                extra.setdefault("config5_1mib_frames", {})[k] = {
                    "workload": f"BASELINE configs[4]: {k}, 1 MiB frames, {r['total'] / (1 << 30):g} GiB per GPU x {world} GPUs, frame-range shards",
                    "value": round(gbps(r, 3), 2), "per_gpu": round(gbps(r, 3) / world, 2), "unit": "GB/s", "ms_per_step": round(r["dev_ms"] / 3, 3),
                    "frames_per_gpu": r["frames"], "verified_bytes": r["verified"], "decompressed_bytes_per_gpu": r["total"], "roofline": roof(r, 3)}
        if "gather" in results:
            g = results["gather"]
            extra["gather_to_rank0"] = {
                "workload": f"the {g['total'] >> 30} GiB LZ4 file sharded by frame range over {world} GPUs, decoded shards collected on rank 0 (NCCL send/recv over NVLink)",
                "gather_GBps_into_rank0": round(g["moved"] / (g["gather_ms"] / 1e3) / GB, 1), "gather_ms": round(g["gather_ms"], 3),
                "decode_ms": round(g["decode_ms"], 3), "decode_then_send_pipeline_ms": round(g["pipeline_ms"], 3),
                "fraction_of_gather_hidden_behind_decode": round(max(0.0, min(1.0, (g["decode_ms"] + g["gather_ms"] - g["pipeline_ms"]) / g["gather_ms"])), 3),
                "verified_bytes": g["verified"], "chunk_mib": 128, "timer": "CUDA events on the current stream, max over ranks"}
        if "random" in results:
            rn = results["random"]
            extra["random_4k"] = {
                "workload": f"BASELINE configs[3]: {rn['n']} x 4 KiB zseek_pread requests, uniform byte offsets, over a {rn['total'] >> 30} GiB zstd-3 "
                            f"file of {rn['frames']} frames (rank 0)",
                "ops_per_s_cold": round(rn["n"] / rn["cold_s"]), "ops_per_s_warm_cache": round(rn["n"] / rn["warm_s"]),
                "batch_10k_cold_p50_ms": round(rn["async_p50_ms"], 3), "batch_10k_cold_p99_ms": round(rn["async_p99_ms"], 3),
                "batch_10k_api": "zseek_b200_pread_batch_async, device-resident requests, call + zseek_b200_batch_wait timed on the host; "
                                 f"last batch verified: {rn['async_ok']}/10000 requests",
                "batch_10k_cold_p50_ms_host_arrays_lru_cache": round(rn["p50_ms"], 3), "batch_10k_cold_p99_ms_host_arrays_lru_cache": round(rn["p99_ms"], 3),
                "short_reads_at_frame_boundaries": rn["short_reads"], "lengths_ok": rn["lengths_ok"], "requests_verified": rn["requests_ok"],
                "verified_bytes": rn["verified_bytes"], "expected_bytes": rn["expected_bytes"],
                "roofline_bound_ops_per_s": round(rn["n"] / ((rn["C"] + rn["total"] + 2 * rn["expected_bytes"]) / (peak * GB))),
                "cpu_baseline": ({"value": rn["cpu"]["ops"], "unit": "ops/s", "cores": rn["cpu"]["threads"], "kind": "reference",
                                  "sample": f"first {rn['cpu']['n']} requests"} if "cpu" in rn else None)}
        line = {
            "metric": "decompressed_GBps", "value": round(gbps(lz), 2), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(lz["dev_ms"] / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config, "clocks": clocks,
            "verified_bytes": lz["verified"],
            "e2e": {"value": round(lz["total"] * world * args.steps / lz["e2e_s"] / GB, 2), "unit": "GB/s",
                    "h2d_bytes_per_step": lz["C"], "d2h_bytes_per_step": lz["total"], "verified_bytes": lz["e2e_verified"],
                    "api": "zseek_b200_read_range, pinned host image -> pinned host buffer"},
            "gpu_launches": lz["launches"],
            "roofline": dict(roof(lz), peak_source=peak_src),
            "cpu_baseline": ({"value": cpu["lz4"]["gbps"], "best": cpu["lz4"]["best"], "unit": "GB/s", "cores": cpu["lz4"]["threads"], "kind": "reference",
                              "sample": (f"first {cpu['lz4']['sample'] >> 20} MiB of the same file, one reader per thread, 1 MiB zseek_pread requests, "
                                         f"cache_size 0, warm-up + 5 passes, timed twice (own process / bench process): value = the better median "
                                         f"({cpu['lz4']['process']}), best alongside; 1 thread: {cpu['lz4']['gbps_1t']} GB/s"),
                              "all_timings": cpu["lz4"]["all"]}
                             if cpu else {"value": None, "unit": "GB/s", "cores": 0, "kind": "reference",
                                          "sample": "timed at N = 1 only; see the --impl reference line of this N"}),
            "extra": extra,
        }
        emit(line)
    # no collective after this point: the rank-0-only legs above (random reads, drop-in harness) take minutes, and a rank that
    # waits for them inside an NCCL barrier gains nothing; every rank leaves on its own (main)
    torch.cuda.synchronize()


# ----------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    corpus = Corpus(args)
    corpus.build(need_all=False)
    image = corpus.image("lz4.zsk", 1)
    threads = os.cpu_count() or 1
    sample = min(corpus.tile, 1 << 30)
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
    desc = (f"each step = the first {sample >> 20} MiB of the same {args.size_gib:g} GiB file (a rate), {threads} pinned threads, one reference reader "
            f"per thread over a RAM image (memcpy pread), 1 MiB zseek_pread requests, cache_size 0")
    line = {"impl": "reference", "metric": "decompressed_GBps", "value": round(value, 3), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(inner / args.steps * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": round(value, 3), "unit": "GB/s", "cores": threads, "kind": "reference", "sample": desc},
            "e2e": {"value": round(value, 3), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "extra": {"wall_s": round(wall, 3), "sample": desc}}
    emit(line)


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
    ap.add_argument("--tile-mib", type=int, default=0, help="0 = unique data (tile = file size)")
    ap.add_argument("--random-ops", type=int, default=1000000)
    ap.add_argument("--random-gib", type=float, default=16.0, help="size of the zstd-3 file of the random-read leg (configs[3]: 16)")
    ap.add_argument("--c5-gib", type=float, default=8.0, help="per-GPU size of the configs[4] legs (0 = skip)")
    ap.add_argument("--dropin-mib", type=int, default=1024, help="file size of the plain-zseek_pread drop-in leg at N = 1 (0 = skip)")
    ap.add_argument("--cpu-leg", default="", help="internal: time the reference on the host cores over this corpus file and print one JSON object")
    args = ap.parse_args()
    if args.tile_mib <= 0:
        args.tile_mib = int(args.size_gib * 1024)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    claim_stdout()
    if args.cpu_leg:
        emit(cpu_leg(Corpus(args), args.cpu_leg))
        return
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
        # Leave without tearing the process group down collectively: ranks finish at different times (see run_b200), and a
        # 2-GPU run of this round sat in the final barrier / destroy_process_group until the box's time limit although the
        # result line had been printed.  The communicator dies with the process.
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
