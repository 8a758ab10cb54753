"""device-to-host bandwidth of one 0.9 GB copy against the same bytes split over several streams (pinned destination)"""
import time, torch
n = 225767424
src = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
dst = torch.empty(n, dtype=torch.float32).pin_memory()
for parts in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (n + parts - 1) // parts
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            a, b = i * step, min(n, (i + 1) * step)
            with torch.cuda.stream(s):
                dst[a:b].copy_(src[a:b], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{parts} stream(s): {best*1e3:.2f} ms  {n*4/best/1e9:.1f} GB/s")
