"""D2H bandwidth of pinned copies: one stream vs two concurrent streams, whole buffer vs 512 MiB pieces."""
import time
import torch

n = 4 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = [torch.cuda.Stream() for _ in range(4)]


def run(name, fn):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    print(f"{name:40s} {n / best / 1e9:6.2f} GB/s ({best * 1e3:.1f} ms)", flush=True)


def one():
    with torch.cuda.stream(s[0]):
        h.copy_(d, non_blocking=True)


def pieces(k, nstreams):
    def f():
        step = n // k
        for i in range(k):
            with torch.cuda.stream(s[i % nstreams]):
                h[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
    return f


run("one copy, one stream", one)
run("8 pieces, one stream", pieces(8, 1))
run("8 pieces, two streams", pieces(8, 2))
run("8 pieces, four streams", pieces(8, 4))
run("64 pieces, two streams", pieces(64, 2))
