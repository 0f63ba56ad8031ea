"""H2D bandwidth from pinned memory with 1, 2, 4 concurrent copy streams.  GPU box: python tools/h2d_probe.py"""
import time, torch
n = 338 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    cuts = [n * i // k for i in range(k + 1)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for s, lo, hi in zip(streams, cuts[:-1], cuts[1:]):
            with torch.cuda.stream(s):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{k} stream(s): {dt*1e3:.2f} ms = {n*4/dt/1e9:.1f} GB/s")
# chunked on one stream (as the solver does): 40 chunks
s = torch.cuda.Stream()
for chunks in (8, 40):
    cuts = [n * i // chunks for i in range(chunks + 1)]
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for lo, hi in zip(cuts[:-1], cuts[1:]):
                d[lo:hi].copy_(h[lo:hi], non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"1 stream, {chunks} chunks: {dt*1e3:.2f} ms = {n*4/dt/1e9:.1f} GB/s")
