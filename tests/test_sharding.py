"""CPU tests of the multi-GPU plumbing: world_size-2/3 gloo processes shard a batch, "solve" their
shard with the oracle, and gather; the result must equal the unsharded solve row for row."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quattro_transformer_ilqr_b200.sharding import gather_rows, shard, shard_bounds


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 16, 65536, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ilqr_batched as ob
    from oracle.models import make_model, sample_x0
    model = make_model("cartpole")
    x0 = sample_x0("cartpole", total, seed=5)
    mine = shard(x0, rank, world)
    res = ob.optimize(model, mine, np.zeros((len(mine), 30, 1)), tol=1e-1)
    cost = gather_rows(torch.as_tensor(res["cost"]), total)
    iters = gather_rows(torch.as_tensor(res["iters"]), total)
    u0 = gather_rows(torch.as_tensor(res["u"][:, 0]), total)
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), cost=cost.numpy(), iters=iters.numpy(), u0=u0.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 64), (3, 50)])
def test_sharded_solve_equals_unsharded(tmp_path, world, total):
    from oracle import ilqr_batched as ob
    from oracle.models import make_model, sample_x0
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npz")
    ref = ob.optimize(make_model("cartpole"), sample_x0("cartpole", total, seed=5), np.zeros((total, 30, 1)), tol=1e-1)
    assert np.array_equal(got["cost"], ref["cost"])          # no cross-trajectory arithmetic: bit-identical
    assert np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["u0"], ref["u"][:, 0])
