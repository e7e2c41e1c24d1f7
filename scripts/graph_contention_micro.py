"""Launch latency under PCIe load: 40 small kernels per 'pair' as stream launches vs. one CUDA graph, with and
without a background D2H DMA saturating the link.  torch only."""
import torch, time
dev = torch.device("cuda", 0)
x = torch.zeros(1 << 20, device=dev)
src = torch.empty(1920 * 1080 * 7, dtype=torch.float32, device=dev)
dst = torch.empty(1920 * 1080 * 7, dtype=torch.float32).pin_memory()
cs, ws = torch.cuda.Stream(), torch.cuda.Stream()
def body():
    for _ in range(40): x.add_(1.0)
with torch.cuda.stream(ws):
    body(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=ws):
        body()
def run(use_graph, bg, reps=20):
    torch.cuda.synchronize()
    if bg:
        with torch.cuda.stream(cs):
            for _ in range(8): dst.copy_(src, non_blocking=True)
    time.sleep(0.0005)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ws):
        e0.record()
        for _ in range(reps):
            if use_graph: g.replay()
            else: body()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000 / 40
for ug in (False, True):
    for bg in (False, True):
        run(ug, bg)
        print(f"graph={ug} background_d2h={bg}: {run(ug, bg):.2f} us per kernel")
