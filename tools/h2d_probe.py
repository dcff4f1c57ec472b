"""H2D rate of pinned host memory: one stream vs the same bytes split over 2 / 4 streams (is 47 GB/s the link or the DMA setup?)."""
import torch, time
n = 363_000_000 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory(); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
def run(ns, reps=10):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    per = n // ns
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * per:(i + 1) * per].copy_(h[i * per:(i + 1) * per], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{ns} stream(s): {n * 8 / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms for {n * 8 / 1e6:.0f} MB)")
for ns in (1, 2, 4, 1):
    run(ns)
# device-to-host
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): h.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"D2H: {n * 8 / dt / 1e9:.1f} GB/s")
