"""Probe: pinned-memory D2H / H2D rates when every rank of a torchrun job copies at the same time (what bounds the
multi-GPU `e2e` of bench.py).   torchrun --nproc-per-node N tools/pcie_probe_n.py [--bind]"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
bound = False
if "--bind" in sys.argv:
    from quattro_transformer_ilqr_b200.sharding import bind_host_to_gpu
    bound = bind_host_to_gpu(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = torch.device(f"cuda:{local}")
n = 106_561_536 // 8
h = torch.empty(n, dtype=torch.float64).pin_memory(); d = torch.empty(n, dtype=torch.float64, device=dev)
h2 = torch.empty(n // 4, dtype=torch.float64).pin_memory(); d2 = torch.empty(n // 4, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(reps, up):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        if up:
            with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for up in (False, True):
    run(3, up)
    ms = run(20, up)
    t = torch.tensor([ms], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{world} rank(s){' bound' if bound else ''}: D2H 106.6 MB{' + H2D 26.6 MB' if up else ''} per rank: {t.item():.3f} ms -> {n*8/t.item()/1e6:.1f} GB/s per rank, {world*n*8/t.item()/1e6:.1f} GB/s total", flush=True)
if world > 1: dist.destroy_process_group()
