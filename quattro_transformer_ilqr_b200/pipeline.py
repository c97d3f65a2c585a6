"""Host-resident batches: chunked, stream-pipelined front end of ``BatchedILQR``.

The reference's callers hand ``iLQR_TF`` host arrays and get host arrays back
(quattro_ilqr_tf/quattro_ilqr_tf.py:66-86, :477-479).  For a large batch of MPC instances whose states live on
the host, the PCIe transfers (x0 and u_init in; u_seq, x_seq and the cost out: 8.2 KB per quadrotor trajectory,
6.5 KB of it outbound) cost more than the iteration itself, so the batch is cut into ``chunks`` contiguous
slices, each with its own engine and CUDA stream: while slice k iterates, slice k+1 is being uploaded and slice
k-1 downloaded (the two copy engines and the SMs run concurrently).  Trajectories are independent, so the
slicing does not change any result (same property as the multi-GPU sharding, tests/test_gpu_solver.py).

    pipe = StreamedILQR(DeviceModel("quadrotor"), B, N, chunks=4)
    pipe.run(x0_pinned, u_pinned, u_out_pinned, x_out_pinned, cost_out_pinned, iters=1, tol=1e-3, force=True)

All host tensors should be pinned (``torch.Tensor.pin_memory``) - pageable memory makes the copies synchronous.
"""
from __future__ import annotations

import torch

from .engine import BatchedILQR
from .sharding import shard_bounds


class StreamedILQR:
    def __init__(self, model, B, N, chunks=4, device=None, tf=None, **engine_kw):
        if chunks < 1 or chunks > B:
            raise ValueError("need 1 <= chunks <= B")
        self.B, self.N, self.chunks = int(B), int(N), int(chunks)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.bounds = [shard_bounds(self.B, c, self.chunks) for c in range(self.chunks)]
        self.engines = [BatchedILQR(model, hi - lo, N, device=self.device, tf=tf, **engine_kw) for lo, hi in self.bounds]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.chunks)]
        self.n, self.m = self.engines[0].n, self.engines[0].m
        if tf is not None:
            # transformer mode: the slices share the wrapper's device weights (uploaded and split here, once, before any
            # slice stream can ask for them) and its state shift; each engine owns its activations workspace
            tf._device_weights(self.device)
            self.set_shift(tf._shift.cpu().numpy() if tf._shift is not None else None)
            torch.cuda.current_stream(self.device).synchronize()

    def set_shift(self, shift=None):
        """State shift of the transformer input (x_seq_err = x_seq + shift, quattro_ilqr_tf.py:504/:532), one private
        copy per slice engine, created on the current stream: call it before ``run`` (which makes every slice stream
        wait for the current one), never concurrently with it."""
        import numpy as np
        v = np.zeros(self.n) if shift is None else np.asarray(shift, dtype=np.float64).reshape(-1)
        for eng in self.engines:
            eng._tf_shift = torch.as_tensor(v, dtype=torch.float64).to(self.device)

    @property
    def launches(self):
        return sum(e.launches for e in self.engines)

    def run(self, x0, u_init, u_out, x_out, cost_out, iters=None, max_iter=100, tol=1e-3, force=False, join=True):
        """Upload, solve and download every slice on its own stream.

        ``iters`` forced iterations (``force=True``: benchmark mode, every trajectory stays active) or, with
        ``iters=None``, a solve to convergence with the reference's stopping rule.  x0 (B, n), u_init (B, N, m)
        in; u_out (B, N, m), x_out (B, N+1, n), cost_out (B,) out - host tensors, float64.  Returns after the
        work has been ENQUEUED; call ``synchronize()`` (or record an event on ``streams``) before reading.

        ``join=True`` makes the current stream wait for every slice, so later work on it sees the results.  A caller
        that feeds one batch after the other (a serving loop) passes ``join=False`` and calls ``join()`` once at the
        end: the slice streams are in order by themselves, and without the per-call join the downloads of batch k
        overlap the uploads and iterations of batch k+1 (the D2H engine never idles between batches)."""
        cur = torch.cuda.current_stream(self.device)
        for (lo, hi), eng, st in zip(self.bounds, self.engines, self.streams):
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                eng.set_problem(x0[lo:hi], u_init[lo:hi])                       # H2D + nominal rollout
                if iters is None:
                    eng.solve(max_iter=max_iter, tol=tol)
                else:
                    for _ in range(iters):
                        eng.iterate(tol, force=force)
                u_out[lo:hi].copy_(eng.u_batch(), non_blocking=True)           # D2H
                x_out[lo:hi].copy_(eng.x_batch(), non_blocking=True)
                cost_out[lo:hi].copy_(eng.cost, non_blocking=True)
        if join:
            self.join()

    def join(self):
        """The current stream waits for everything enqueued on the slice streams."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def synchronize(self):
        for st in self.streams:
            st.synchronize()
