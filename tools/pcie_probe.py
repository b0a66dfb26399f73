"""Host<->device copy bandwidth with pinned buffers (what bounds the un-hidden part of nrx_forward_host)."""
import torch, time
for mb in (8, 20, 64):
    n = mb * (1 << 20)
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
        for _ in range(3): dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): dst.copy_(src, non_blocking=True)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} {mb} MB: {10 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9:.1f} GB/s")
# two copy streams side by side (is one copy engine the limit?)
n = 8 << 20
hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(2)]
ds = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(2)]
ss = [torch.cuda.Stream() for _ in range(2)]
for k in (1, 2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        for i in range(k):
            with torch.cuda.stream(ss[i]): ds[i].copy_(hs[i], non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"H2D 8 MB x {k} stream(s): {20 * k * n / dt / 1e9:.1f} GB/s aggregate")
# one cold copy (first touch of a fresh pinned buffer) vs warm
h = torch.empty(21 << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(21 << 20, dtype=torch.uint8, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print(f"H2D 21 MB single copy #{rep}: {(time.perf_counter() - t0) * 1e3:.3f} ms")
