#!/usr/bin/env python3
"""Benchmark of the batched iLQR hot path (BASELINE.json metric: trajectory-iterations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload quadrotor_classic] [--impl reference]

Workload (default, BASELINE.json configs[2]): quadrotor 12-state / 4-control classic iLQR, horizon 50,
batch 16 384 trajectories PER GPU (weak scaling; trajectories are independent MPC instances, so
ranks share nothing and there is no collective on the path — one NCCL all_gather of the final
costs closes the timed region).  A "step" is ONE iLQR iteration over the whole batch in
fixed-iteration mode (every trajectory forced active): linearise -> Riccati -> line search; the
problem is reset (nominal rollout) every 10 steps so the steps cover the first ten iterations of a
solve, like SURVEY.md 8(d) prescribes.  The expansion buffer alone is 1.36 GB per pass, far larger
than the 126 MB L2, so no explicit flush is needed between iterations.

Printed JSON (one line, rank 0): value = device-timed throughput with inputs resident in HBM;
e2e = the same metric with host (pinned) inputs copied in and results copied out every step;
roofline = dominant kernel vs MEASURED_PEAKS.json; cpu_baseline = the scalar oracle port (the
reference's algorithm and cost profile, oracle/ilqr_port.py) on this box's host cores.
`--impl reference` times that CPU path alone (the reference is pure Python + NumPy and cannot
travel to the GPU box; the port is pinned bit-for-bit to it in tests/test_oracle.py).
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (system, horizon, per-GPU batch, transformer?)
    "quadrotor_classic": ("quadrotor", 50, 16384, False),
    "cartpole_classic": ("cartpole", 30, 16384, False),
    "quadrotor_tf": ("quadrotor", 50, 8192, True),
    "cartpole_tf": ("cartpole", 30, 4096, True),
}
RESET_EVERY = 10
# algorithmic bytes per trajectory per launch, SURVEY.md section 8(d) (quadrotor / cart-pole)
ALGO_BYTES = {
    "quadrotor": {"linearize": 87072, "riccati": 94224, "linesearch": 16924, "rollout": 3300},
    "cartpole": {"linearize": 6216, "riccati": 6200, "linesearch": 1860, "rollout": 636},
}
# x,u are kept in fp64 on the device: their share of the figures above doubles
FP64_EXTRA = {"quadrotor": {"linearize": 3248, "riccati": 0, "linesearch": 3248 + 3276, "rollout": 3300},
              "cartpole": {"linearize": 616, "riccati": 0, "linesearch": 616 + 620, "rollout": 636}}
# Block path (quadrotor, Euler, diagonal weights; include/quattro_b200.h QX_LAYOUT_BLK): the expansion is 64
# floats per (t, b) instead of 416 and the terminal 16 floats, so the algorithmic traffic of the backward pass is
#   linearise: read x,u fp64 6 496 B + write 50 x 192 + 64 B            = 16 160 B per trajectory
#   Riccati  : read 50 x 192 + 64 B      + write gains 50 x 208 B       = 20 064 B per trajectory
BLK_BYTES = {"linearize": 6496 + 50 * 192 + 64, "riccati": 50 * 192 + 64 + 50 * 208}


def workload_config(workload, world, B=None):
    """The `config` object of the JSON line: the workload only (both arms print the same object)."""
    system, N, B_default, use_tf = WORKLOADS[workload]
    B = B or B_default
    n, m = (12, 4) if system == "quadrotor" else (4, 1)
    return {"workload": f"{workload}: {system} n={n} m={m} horizon={N}, {B} trajectories per GPU, "
                        f"fixed-iteration mode, problem reset every {RESET_EVERY} steps"
                        + (", gains of steps 0..N-P-1 predicted by the transformer (random-init fp32 weights)" if use_tf else ""),
            "global_batch": world * B, "parallelism": f"dp{world} (independent trajectory shards)",
            "l2": "no flush needed: the working set of an iteration (x, u ping-pong 13 KB + expansion 13 KB + gains 10 KB "
                  "per quadrotor trajectory = 0.6 GB at 16 384; transformer activations 0.4 MB per trajectory) is larger than "
                  "the 126 MB L2"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# dense-equivalent flops of one transformer forward per trajectory (causal mask not exploited, all L
# rows in every layer, as PyTorch does)
TF_FLOPS = {"quadrotor": 135635456.0, "cartpole": 53757952.0}


def tensor_peak():
    """Tensor peak = the measured dense bf16 / fp16 figure (burst: the forward is timed alone, between events): the dense
    layers and the attention products all run as kind::f16 products of fp16 (hi, lo) halves.  Every product is evaluated
    as three MMAs; algorithmic flops are counted once."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["bf16_tflops"]), "measured bf16 burst (MEASURED_PEAKS.json): the kind::f16 rate every product of the forward runs at"
    return 1650.0, "fallback bf16 figure (B200_PROFILING.md)"


def ncu_reference(kernel, blk, system):
    """DRAM traffic and pipe utilisation of the dominant kernel from the committed `ncu --set full` summary of the
    same workload (profiles/): counters cannot be read live, the launch duration is (CUDA events, above)."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_full_classic_summary.txt")
    sym = {("riccati", True): "riccati_quad_thr_kernel", ("linearize", True): "linearize_quad_blk_kernel",
           ("ls_phaseA", True): "ls_qd_phaseA_kernel", ("ls_phaseB", True): "ls_qd_phaseB_kernel",
           ("ls_phaseW", True): "ls_qd_phaseW_kernel"}.get((kernel, blk))
    if system != "quadrotor" or sym is None or not os.path.exists(path):
        return None
    vals, inside = {}, False
    for line in open(path):
        if line.startswith("-----"):
            inside = False
        elif line.startswith(sym) or (" " + sym) in line[:80] or ("::" + sym) in line[:80]:
            inside = True
        elif inside and line.startswith("  "):
            parts = line.split()
            if len(parts) >= 2:
                try:
                    vals[parts[0]] = (float(parts[1]), parts[2] if len(parts) > 2 else "")
                except ValueError:
                    pass
    if "dram__bytes_read.sum" not in vals:
        return None
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}

    def byt(k):
        v, u = vals[k]
        return v * scale.get(u, 1.0)
    return {"source": "profiles/r2_ncu_full_classic_summary.txt", "captured": "iteration 6 of a forced solve (13 220 trajectories queued)",
            "traffic": byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum"),
            "fma_pipe_pct": vals.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", (None,))[0],
            "fp64_pipe_pct": vals.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", (None,))[0],
            "duration_us": vals.get("gpu__time_duration.sum", (None,))[0],
            "issue_active_pct": vals.get("smsp__issue_active.avg.pct_of_peak_sustained_active", (None,))[0],
            "dram_pct": vals.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", (None,))[0]}


# ------------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    system, x0, iters = args
    from oracle.ilqr_port import ILQRPort
    from oracle.models import make_model
    model = make_model(system)
    N = {"cartpole": 30, "quadrotor": 50}[system]
    p = ILQRPort(model, x0, [np.zeros(model.m) for _ in range(N)], N, max_iter=iters, tol=0.0)
    p.optimize(model.x_ref)
    return p.total_iter + 1


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference(system, steps, warmup, iters_per_step=1):
    """Each step: every host core runs `iters_per_step` iLQR iterations of one trajectory."""
    from quattro_transformer_ilqr_b200.sampling import sample_x0
    cores = cpu_cores()
    x0 = sample_x0(system, cores, seed=0)
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(_cpu_worker, [(system, x, iters_per_step) for x in x0])
        t0 = time.perf_counter()
        done = 0
        for _ in range(steps):
            done += sum(pool.map(_cpu_worker, [(system, x, iters_per_step) for x in x0]))
        dt = time.perf_counter() - t0
    return done / dt, cores, dt / steps, f"{cores} trajectories x {iters_per_step} iteration(s) per step, one process per core"


# ------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """Samples SM clocks and throttle reasons from a thread while the timed region runs (the recipe's clocks
    line).  NVML (nvidia_ml_py) when importable - a sample costs microseconds, so even a 100 ms region gets
    dozens - else the nvidia-smi query of the recipe."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.index, self._stop, self.thread, self.max_mhz = [], index, threading.Event(), None, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical GPUs: honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                phys = int(vis.split(",")[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8))),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))]
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                mhz = int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(reasons_fn(self.handle))
                self.rows.append([str(mhz), str(self.max_mhz)] + ["Active" if mask & b else "Not Active" for b in bits])
            except Exception:
                return
            self._stop.wait(0.002)

    def _poll_smi(self):
        cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)]
        while not self._stop.is_set():
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=5).stdout.strip()
                self.rows.append([c.strip() for c in out.split(",")])
            except (OSError, subprocess.SubprocessError):
                return
            self._stop.wait(0.02)

    def start(self):
        self.thread = threading.Thread(target=self._poll_nvml if self.nvml else self._poll_smi, daemon=True)
        self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join(timeout=10)
        rows = [r for r in self.rows if len(r) == 6 and r[0].isdigit()]
        if not rows:
            return None
        sm = sorted(int(r[0]) for r in rows)
        reasons = [n for i, n in enumerate(self.NAMES) if any(r[2 + i] == "Active" for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]), "reasons": reasons, "samples": len(rows),
                "source": "nvml" if self.nvml else "nvidia-smi"}


# ------------------------------------------------------------------------------- GPU arm
def gpu_bench(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    from quattro_transformer_ilqr_b200.sampling import sample_x0
    from quattro_transformer_ilqr_b200 import BatchedILQR, DeviceModel, TransformerILQR, TransformerPredictor

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from quattro_transformer_ilqr_b200.sharding import bind_host_to_gpu
    numa_bound = bind_host_to_gpu(local) if world > 1 else False     # before any pinned allocation (first touch)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()

    system, N, B, use_tf = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    model = DeviceModel(system)
    n, m = model.n, model.m
    tf = None
    if use_tf:
        tf = TransformerILQR(1, 1).load(os.path.join(ROOT, "tests", "golden", "checkpoints", system))
        torch.manual_seed(0)                                  # random-init fp32 weights of the repo architecture
        tf.model = TransformerPredictor(tf.state_dim, tf.control_dim, tf.d_model, tf.nhead, tf.num_decoder_layers,
                                        tf.dim_feedforward, tf.dropout, tf.max_seq_len, tf.target_len, tf.prompt_len).eval()
        tf.invalidate()
        shift = np.zeros(n)
        if system == "quadrotor":
            shift[2] = 0.5 - 0.5                              # state_offset - x_ref
        tf.set_shift(shift, dev)
    eng = BatchedILQR(model, B, N, device=dev, tf=tf)
    # each rank owns a contiguous shard of the global batch (seeded per shard)
    x0_h = torch.as_tensor(sample_x0(system, B, seed=rank)).pin_memory()
    u0_h = torch.zeros((B, N, m), dtype=torch.float64).pin_memory()
    x0_d, u0_d = x0_h.to(dev), u0_h.to(dev)
    tol = 1e-3 if system == "quadrotor" else 1e-1

    def run_steps(k, state):
        done = 0
        while done < k:
            if state["i"] % RESET_EVERY == 0:
                eng.set_problem(x0_d, u0_d)
            # two iterations per CUDA-graph replay where the step count and the reset cadence allow it
            if args.graph and done + 2 <= k and state["i"] % RESET_EVERY <= RESET_EVERY - 2:
                eng.iterate_pair(tol, force=True)
                state["i"] += 2
                done += 2
            else:
                eng.iterate(tol, force=True)
                state["i"] += 1
                done += 1

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    state = {"i": 0}
    run_steps(args.warmup, state)
    if args.graph:
        # capture the iteration-pair graphs for both buffer parities now, whatever W was: a capture costs milliseconds
        # and must not land inside the timed region
        for parity in (0, 1):
            eng.set_problem(x0_d, u0_d)
            if eng.cur != parity:
                eng.iterate(tol, force=True)
            for _ in range(3):
                eng.iterate_pair(tol, force=True)
    sampler = ClockSampler(local)
    sync_all()
    if rank == 0:
        sampler.start()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    state["i"] = 0
    ev0.record()
    run_steps(args.steps, state)
    if world > 1:                                            # the path's only collective: final gather of costs
        all_cost = torch.empty(world * B, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(all_cost, eng.cost)
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches - launches0
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)
    if args.skip_extras:        # profiling runs (ncu): the timed loop only
        if rank == 0:
            args.emit(json.dumps({"value": value, "ms_per_step": ms / args.steps, "gpu_launches": launches, "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the host-facing call: pinned inputs in, results out, every step -------
    # StreamedILQR (quattro_transformer_ilqr_b200/pipeline.py) cuts the batch into slices with their own engine and
    # stream, so uploads, iterations and downloads of different slices overlap (transformer mode: two slices sharing the
    # wrapper's weights, each engine with its own activations workspace).
    u_out = torch.empty((B, N, m), dtype=torch.float64).pin_memory()
    x_out = torch.empty((B, N + 1, n), dtype=torch.float64).pin_memory()
    c_out = torch.empty(B, dtype=torch.float64).pin_memory()
    chunks = args.e2e_chunks if tf is None else min(args.e2e_chunks, 2)     # transformer mode: two slices (each owns a workspace)
    if chunks > 1:
        from quattro_transformer_ilqr_b200.pipeline import StreamedILQR
        pipe = StreamedILQR(model, B, N, chunks=chunks, device=dev, tf=tf)

        def e2e_step():
            # join=False: consecutive batches pipeline (batch k's downloads overlap batch k+1's uploads and kernels);
            # e2e_join() below makes the timing event wait for every slice stream
            pipe.run(x0_h, u0_h, u_out, x_out, c_out, iters=1, tol=tol, force=True, join=False)
        e2e_join = pipe.join
    else:
        def e2e_join():
            pass

        def e2e_step():
            eng.set_problem(x0_h, u0_h)                           # H2D + nominal rollout
            eng.iterate(tol, force=True)
            u_out.copy_(eng.u_batch(), non_blocking=True)         # D2H
            x_out.copy_(eng.x_batch(), non_blocking=True)
            c_out.copy_(eng.cost, non_blocking=True)

    for _ in range(3):
        e2e_step()
    e2e_join()
    sync_all()
    e2e_steps = max(3, min(args.steps, 20))
    ev0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_join()
    ev1.record()
    sync_all()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / (float(t.item()) * 1e-3)
    h2d = x0_h.numel() * 8 + u0_h.numel() * 8
    d2h = (u_out.numel() + x_out.numel() + c_out.numel()) * 8

    # ---- the same path at the granularity of the reference's API call: optimize() runs a window of iterations per call
    # (quattro_ilqr_tf.py:426-479), so the problem goes in once and the solution comes out once per RESET_EVERY
    # iterations - the window `value` is measured on.  Reported beside `e2e`, which moves everything EVERY iteration.
    e2e_solve = None
    if chunks > 1:
        def solve_call():
            pipe.run(x0_h, u0_h, u_out, x_out, c_out, iters=RESET_EVERY, tol=tol, force=True, join=False)
        for _ in range(2):
            solve_call()
        e2e_join()
        sync_all()
        calls = max(2, min(args.steps, 20) // 2)
        ev0.record()
        for _ in range(calls):
            solve_call()
        e2e_join()
        ev1.record()
        sync_all()
        t2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_solve = {"value": world * B * RESET_EVERY * calls / (float(t2.item()) * 1e-3), "unit": "trajectory-iterations/s",
                     "iterations_per_call": RESET_EVERY, "calls": calls, "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": d2h,
                     "note": "one optimize()-like call per window: pinned x0, u_init in, u, x, cost out, RESET_EVERY forced "
                             "iterations in between; consecutive calls pipeline over the slice streams"}

    # ---- per-kernel durations (CUDA events on the launching stream), rank 0 only --------------------
    # One window of RESET_EVERY forced iterations from the reset - the iterations the timed region consists of - after
    # an untimed one.  The staged quadrotor line search is three launches inside ONE C-ABI call; the library records
    # caller-owned events between them (qx_linesearch_stage_events), so each launch is timed on its own.
    kernels = {}
    if rank == 0:
        import ctypes
        names = ["linearize", "riccati"] + (["transformer"] if use_tf else []) + ["linesearch"]
        acc = {k: [] for k in names}
        staged = system == "quadrotor" and eng.ls_workspace is not None and bool(eng.params.diag_cost)
        stage_acc = {"ls_phaseA": [], "ls_phaseB": [], "ls_phaseW": []}
        queued = []
        if staged:
            sev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            for e in sev:
                e.record()                           # creates the cudaEvent_t handles
            handles = (ctypes.c_void_p * 4)(*[e.cuda_event for e in sev])
            qlib_h = eng.lib
            assert qlib_h.qx_linesearch_stage_events(handles, 4) == 0
        start = 0 if tf is None else N - tf.prompt_len
        try:
            for window in range(2):
                eng.set_problem(x0_d, u0_d)
                for it in range(RESET_EVERY):
                    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
                    evs[0].record()
                    eng.linearize(start); evs[1].record()
                    eng.riccati(start); evs[2].record()
                    j = 2
                    if use_tf:
                        tf.forward_device(eng); j += 1; evs[j].record()
                    eng.linesearch(tol, force=True); evs[j + 1].record()
                    torch.cuda.synchronize(dev)
                    if window == 1:
                        for q, k in enumerate(names):
                            acc[k].append(evs[q].elapsed_time(evs[q + 1]))
                        if staged:
                            for q, k in enumerate(stage_acc):
                                stage_acc[k].append(sev[q].elapsed_time(sev[q + 1]))
                            queued.append(int(eng.ls_workspace[:4].view(torch.int32)[0].item()))
        finally:                                     # never leave the hook pointing at events that are about to die
            if staged:
                assert eng.lib.qx_linesearch_stage_events(None, 0) == 0
        hbm, how = peaks()
        from quattro_transformer_ilqr_b200 import _lib as qlib
        blk = eng.exp_layout == qlib.QX_LAYOUT_BLK
        for k in names:
            ms_k = float(np.mean(acc[k]))
            entry = {"ms": round(ms_k, 4)}
            if k in ALGO_BYTES[system]:
                frac_T = (N - (0 if tf is None else N - tf.prompt_len)) / N if k in ("linearize", "riccati") and use_tf else 1.0
                by = (ALGO_BYTES[system][k] + FP64_EXTRA[system][k]) * frac_T
                if blk and k in BLK_BYTES:
                    by = BLK_BYTES[k] * frac_T
                gbs = by * B / (ms_k * 1e-3) / 1e9
                entry.update({"algo_bytes_per_traj": by, "GBps": round(gbs, 1), "frac_hbm": round(gbs / hbm, 4)})
            kernels[k] = entry
        if staged:
            # algorithmic bytes per launch: A rolls every trajectory (read x, u, K, k 16 896 B, write x^, u^, cost
            # 6 552 B); B reads the nominal and the gains of a QUEUED trajectory once (its five probes share the lines)
            # and writes five costs; W rolls the queued ones again.  Queue lengths are read back per iteration.
            ls_in, ls_out = 16896.0, 6552.0
            per = {"ls_phaseA": [(ls_in + ls_out) * B] * len(queued),
                   "ls_phaseB": [(ls_in + 40.0) * q for q in queued],
                   "ls_phaseW": [(ls_in + ls_out) * q for q in queued]}
            for k, t in stage_acc.items():
                ms_k = float(np.mean(t))
                by = float(np.mean(per[k]))
                gbs = float(np.sum(per[k])) / (float(np.sum(t)) * 1e-3) / 1e9
                kernels[k] = {"ms": round(ms_k, 4), "ms_max": round(float(np.max(t)), 4), "algo_bytes_per_launch": by,
                              "GBps": round(gbs, 1), "frac_hbm": round(gbs / hbm, 4)}
            kernels["ls_phaseB"]["queued_per_iteration"] = queued

    # ---- run-to-convergence mode: one full solve with the reference's stopping rule, rank 0 ----------
    converge = None
    if rank == 0:
        eng.set_problem(x0_d, u0_d)
        eng.solve(max_iter=100, tol=tol, check_every=4, compact=args.compact)   # untimed: allocates the cached engines, if any
        eng.set_problem(x0_d, u0_d)
        torch.cuda.synchronize(dev)
        ev0.record()
        batch_iters = eng.solve(max_iter=100, tol=tol, check_every=4, compact=args.compact)
        ev1.record()
        torch.cuda.synchronize(dev)
        ms_c = ev0.elapsed_time(ev1)
        done = int(eng.iters.sum().item())
        converge = {"value": done / (ms_c * 1e-3), "unit": "trajectory-iterations/s", "ms": round(ms_c, 3),
                    "trajectory_iterations": done, "batch_iterations": batch_iters,
                    "finished": int((eng.status != 0).sum().item()),
                    "note": "one solve to convergence (tol of the reference driver), active set checked every 4 iterations"
                            + (", compacted into smaller engines once at most half of it is left" if args.compact else "")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # dominant KERNEL = the single launch with the longest mean duration over the window (linearise, Riccati and - where
    # the library can time them separately - the three stages of the line search; otherwise the line search, which is
    # then a sequence of launches, is reported in `kernels` only)
    stage_names = [k for k in ("ls_phaseA", "ls_phaseB", "ls_phaseW") if k in kernels]
    single = [k for k in kernels if "GBps" in kernels[k] and k != "linesearch"] or [k for k in kernels if "GBps" in kernels[k]]
    dom = max(single, key=lambda k: kernels[k]["ms"])
    step_ms = sum(kernels[k]["ms"] for k in kernels if k not in stage_names)
    for k in kernels:
        kernels[k]["share_of_step"] = round(kernels[k]["ms"] / step_ms, 3)
    hbm, how = peaks()
    ncu = ncu_reference(dom, blk, system) if B == WORKLOADS[args.workload][2] else None
    if use_tf:
        # transformer mode: the forward (dense layers + attention on tcgen05) is 95 % of the step; it is a sequence of
        # 16 launches, reported as one tensor-bound unit
        tpeak, thow = tensor_peak()
        tf_ms = kernels["transformer"]["ms"]
        achieved = TF_FLOPS[system] * B / (tf_ms * 1e-3) / 1e12
        kernels["transformer"].update({"algo_flops_per_traj": TF_FLOPS[system], "TFLOPs": round(achieved, 1),
                                       "frac_tensor": round(achieved / tpeak, 4)})
        roofline = {"kernel": "transformer forward (three-term fp16 splits on tcgen05: dense layers, fused FFN, attention on fp16-pair q, k, v)", "bound": "tensor",
                    "achieved": round(achieved, 1), "peak": tpeak, "unit": "TFLOP/s", "frac": round(achieved / tpeak, 4),
                    "traffic": None, "peak_source": thow, "algo_flops_per_launch": TF_FLOPS[system] * B,
                    "note": "fp32-accurate products cost three tensor-core MMAs each: the tensor pipe does 3x the algorithmic "
                            "flops counted here; per-kernel tensor-pipe activity is in profiles/r2_ncu_full_transformer_p16_summary.txt"}
    else:
        per_launch = kernels[dom].get("algo_bytes_per_launch", kernels[dom].get("algo_bytes_per_traj", 0.0) * B)
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["GBps"], "peak": hbm, "unit": "GB/s",
                    "frac": kernels[dom]["frac_hbm"], "traffic": ncu["traffic"] if ncu else None, "peak_source": how,
                    "algo_bytes_per_launch": per_launch, "ms": kernels[dom]["ms"],
                    "share_of_step": kernels[dom]["share_of_step"]}
    if ncu and not use_tf:
        roofline["ncu"] = ncu
    if not use_tf:
        notes = {
            "riccati": "one thread per trajectory, 50 dependent steps: at 16 384 trajectories there is one warp per scheduler "
                       "and the kernel runs at its fp32 latency floor (SURVEY 8d asks for both figures: ncu.fma_pipe_pct, "
                       "ncu.issue_active_pct); from 32 768 trajectories on it is HBM-bound",
            "ls_phaseB": "costs-only probes of the five smaller alphas, one thread per (queued trajectory, alpha): the launch does "
                         "not run in iterations 0-2 of the window (nothing is rejected there) and is the longest launch from "
                         "iteration 3 on (ms_max).  It moves little (the five probes of a trajectory share its nominal and gain "
                         "lines in L2) and is bound by the fp64 pipe, not by HBM: ncu.fp64_pipe_pct is the figure that says how "
                         "close it is to its own ceiling; frac is the algorithmic bytes against the copy peak, as the bench "
                         "contract defines it",
            "ls_phaseA": "alpha = 1 rollout of every trajectory, one thread each: one warp per scheduler at 16 384 trajectories "
                         "(fp64 latency floor); HBM-bound from 32 768 trajectories on",
        }
        if dom in notes:
            roofline["note"] = notes[dom]
        roofline["all_launches"] = {k: {"ms": kernels[k]["ms"], "frac_hbm": kernels[k]["frac_hbm"],
                                        "share_of_step": kernels[k]["share_of_step"]}
                                    for k in single}
    if world == 1:      # the CPU leg runs at N = 1 only (at N > 1 the other ranks' host threads share the cores)
        cpu_val, cores, _, sample = cpu_reference(system, steps=2, warmup=1, iters_per_step=2)
        cpu_baseline = {"value": cpu_val, "unit": "trajectory-iterations/s", "cores": cores, "kind": "port", "sample": sample}
    else:
        cpu_baseline = None
    out = {
        "metric": "batched iLQR trajectory-iterations/sec (quadrotor, horizon 50)" if system == "quadrotor"
        else "batched iLQR trajectory-iterations/sec (cart-pole, horizon 30)",
        "value": value, "unit": "trajectory-iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (linearise/Riccati/transformer) + f64 (rollouts, costs)", "data": "synthetic",
        "config": workload_config(args.workload, world, B),
        "implementation": {"backward_pass": "block records (48 floats per step)" if eng.exp_layout == 2 else "dense records",
                           "launch": "two iterations per CUDA-graph replay" if args.graph else "one launch per kernel"},
        "e2e": {"value": e2e_value, "unit": "trajectory-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "pipeline": f"{chunks} slices on {chunks} streams, consecutive batches overlap" if chunks > 1 else "single stream",
                "host_numa_bound": numa_bound},
        "e2e_solve": e2e_solve,
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernels": kernels, "converge": converge,
        "cpu_baseline": cpu_baseline,
    }
    args.emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def reference_arm(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    system, N, B, _ = WORKLOADS[args.workload]
    t0 = time.perf_counter()
    val, cores, step_s, sample = cpu_reference(system, args.steps, args.warmup)
    n, m = (12, 4) if system == "quadrotor" else (4, 1)
    args.emit(json.dumps({
        "impl": "reference",
        "metric": "batched iLQR trajectory-iterations/sec (quadrotor, horizon 50)" if system == "quadrotor"
        else "batched iLQR trajectory-iterations/sec (cart-pole, horizon 30)",
        "value": val, "unit": "trajectory-iterations/s", "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same config object as the GPU arm's line (the driver compares them); what this arm runs of it - the scalar
        # finite-difference port, a bounded sample per step - is said in `implementation` and `cpu_baseline.sample`
        "config": workload_config(args.workload, int(os.environ.get("WORLD_SIZE", args.gpus)), args.batch or None),
        "implementation": {"cpu_path": "scalar finite-difference iLQR (oracle/ilqr_port.py, pinned bit for bit to the "
                                       "reference's quattro_ilqr_tf.py), one process per host core"},
        "cpu_baseline": {"value": val, "unit": "trajectory-iterations/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "trajectory-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }))


def _json_only_stdout():
    """Route everything libraries print on fd 1 (NCCL's version banner, ...) to stderr; return a writer for
    the one JSON line the driver reads from stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(line):
        real.write(line + "\n")
        real.flush()
    return emit


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="quadrotor_classic", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="slices (streams) of the end-to-end leg")
    ap.add_argument("--compact", action="store_true", help="run-to-convergence leg with active-set compaction")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every kernel of the timed loop individually")
    ap.add_argument("--skip-extras", action="store_true", help="timed loop only (for ncu): no e2e / per-kernel / CPU legs")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)        # timing rule: at least three untimed warm-up steps (reported as run)
    args.emit = _json_only_stdout()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_bench(args)


if __name__ == "__main__":
    main()
