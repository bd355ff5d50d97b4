import time, torch
dev = torch.device("cuda:0")
src = torch.randn(1, 8, 1152, 1600).pin_memory()
dst = torch.empty_like(src, device=dev)
side = torch.cuda.Stream()
torch.cuda.synchronize()
for label, ctx in (("current stream", None), ("side stream", side)):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if ctx is None:
            dst.copy_(src, non_blocking=True)
        else:
            with torch.cuda.stream(ctx):
                dst.copy_(src, non_blocking=True)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{label}: host {1e3*(t1-t0):.3f} ms, total {1e3*(t2-t0):.3f} ms, pinned={src.is_pinned()}")
x = src.to(dev, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); x = src.to(dev, non_blocking=True); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f".to(): host {1e3*(t1-t0):.3f} ms")
