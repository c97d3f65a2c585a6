"""Probe: pinned-memory PCIe copy rates of this box (bounds the end-to-end number of bench.py)."""
import torch
dev = torch.device("cuda:0")
n = 106_561_536 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory()
d = torch.empty(n, dtype=torch.float64, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: h.copy_(d, non_blocking=True)); print(f"D2H {n*8/1e6:.0f} MB: {ms:.3f} ms  {n*8/ms/1e6:.1f} GB/s")
ms = timeit(lambda: d.copy_(h, non_blocking=True)); print(f"H2D {n*8/1e6:.0f} MB: {ms:.3f} ms  {n*8/ms/1e6:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.float64).pin_memory(); d2 = torch.empty(n, dtype=torch.float64, device=dev)
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
ms = timeit(both); print(f"D2H + H2D concurrently: {ms:.3f} ms  {n*8/ms/1e6:.1f} GB/s each way")
