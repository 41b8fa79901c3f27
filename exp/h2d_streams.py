"""pinned host -> device copy rate of ~294 MB as one copy, three copies on one stream, and split over 2 / 4 streams"""
import torch, time
n = 294 * (1 << 20)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(parts, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    chunk = n // parts
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(6):
        t0 = time.perf_counter()
        for i in range(parts):
            with torch.cuda.stream(ss[i % streams]):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{parts:3d} parts on {streams} stream(s): {best * 1e3:6.3f} ms  {n / best / 1e9:6.2f} GB/s")
for parts, streams in ((1, 1), (3, 1), (2, 2), (4, 2), (4, 4), (8, 4), (16, 2)):
    run(parts, streams)
