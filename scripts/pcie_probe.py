"""Raw D2H/H2D bandwidth of this box with the bench's transfer sizes (58 MB row buffers, 2 MB frames)."""
import time, torch
dev = torch.device("cuda", 0)
N = 1920 * 1080
src = [torch.empty((N, 7), dtype=torch.float32, device=dev) for _ in range(4)]
dst = [torch.empty((N, 7), dtype=torch.float32).pin_memory() for _ in range(8)]
fr_h = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(8)]
fr_d = [torch.empty(N, dtype=torch.uint8, device=dev) for _ in range(8)]
s = [torch.cuda.Stream() for _ in range(3)]
def run(nstreams, with_h2d, reps=16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(reps):
        with torch.cuda.stream(s[i % nstreams]):
            dst[i % 8].copy_(src[i % 4], non_blocking=True)
        if with_h2d:
            with torch.cuda.stream(s[2]):
                fr_d[i % 8].copy_(fr_h[i % 8], non_blocking=True)
                fr_d[(i + 1) % 8].copy_(fr_h[(i + 1) % 8], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    return reps * N * 28 / dt / 1e9
for ns in (1, 2):
    for h in (False, True):
        run(ns, h)
        print(f"D2H streams={ns} concurrent_h2d={h}: {run(ns, h):.1f} GB/s")
