"""Host-to-device rate of a 4 GiB pinned buffer as one copy and as 2 / 4 concurrent copies on separate streams
(is the 55 GB/s of the e2e leg the link or one DMA engine?)."""
import time, torch
n = 1 << 29                                  # 4 GiB of f64
h = torch.empty(n, dtype=torch.float64, pin_memory=True); h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
for parts in (1, 2, 4, 1):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = n // parts
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * step:(i + 1) * step].copy_(h[i * step:(i + 1) * step], non_blocking=True)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"{parts} concurrent copies: {best * 1e3:7.2f} ms  {n * 8 / best / 1e9:6.1f} GB/s", flush=True)
