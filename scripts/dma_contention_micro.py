"""Does a device-to-pinned-host DMA slow down an HBM-bound kernel running next to it?  torch only."""
import torch, time
dev = torch.device("cuda", 0)
a = torch.empty(64 << 20, dtype=torch.float32, device=dev)      # 256 MB
b = torch.empty_like(a)
src = torch.empty(1920 * 1080 * 7, dtype=torch.float32, device=dev)
dst = torch.empty(1920 * 1080 * 7, dtype=torch.float32).pin_memory()
hsrc = torch.empty(1920 * 1080 * 7, dtype=torch.float32).pin_memory()
ddst = torch.empty_like(src)
cs = torch.cuda.Stream()
def timed_copy(bg):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if bg == "d2h":
        with torch.cuda.stream(cs):
            for _ in range(4): dst.copy_(src, non_blocking=True)
    elif bg == "h2d":
        with torch.cuda.stream(cs):
            for _ in range(4): ddst.copy_(hsrc, non_blocking=True)
    time.sleep(0.0005)
    e0.record()
    for _ in range(10): b.copy_(a)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
for bg in (None, "d2h", "h2d", None, "d2h", "h2d"):
    timed_copy(bg)
    ms = timed_copy(bg)
    print(f"background={bg}: 256 MB device copy {ms:.3f} ms  ({2 * 256 / 1024 / ms * 1000:.0f} GB/s)")
