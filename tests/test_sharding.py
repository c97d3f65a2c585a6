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


def test_batched_logs_split_into_reference_entries(tmp_path):
    """datasets.split_batched_logs: one reference-schema entry per (trajectory, active iteration); the pickle
    stream round-trips and feeds _create_dataset's expectations (x_seq (T, n), k_seq (T, m), K_seq (T, m, n))."""
    import numpy as np
    from quattro_transformer_ilqr_b200 import datasets
    B, N, n, m = 3, 5, 4, 1
    rng = np.random.default_rng(0)
    logs = []
    for it in range(2):
        found = np.array([True, it == 0, False])
        logs.append({"iteration": it, "x_seq": rng.normal(size=(B, N + 1, n)), "u_seq": rng.normal(size=(B, N, m)),
                     "current_cost": rng.uniform(1, 2, B), "k_seq": rng.normal(size=(B, N, m)),
                     "K_seq": rng.normal(size=(B, N, m, n)), "alpha": np.where(found, 0.5, np.nan),
                     "new_x_seq": rng.normal(size=(B, N + 1, n)), "new_u_seq": rng.normal(size=(B, N, m)),
                     "new_cost": np.where(found, 1.0, np.nan), "found_update": found,
                     "active": np.array([True, True, it == 0])})
    entries = datasets.split_batched_logs(logs)
    assert len(entries) == 5                                     # trajectory 2 is inactive in iteration 1
    e = entries[0]
    assert e["x_seq"].shape == (N + 1, n) and isinstance(e["k_seq"], list) and e["K_seq"][0].shape == (m, n)
    assert isinstance(e["current_cost"], float) and e["alpha"] == 0.5 and e["found_update"] is True
    last = entries[-1]                                           # trajectory 2, iteration 0: no update found
    assert last["alpha"] is None and last["new_x_seq"] is None and last["new_u_seq"] is None and last["new_cost"] is None
    path = tmp_path / "logs.pkl"
    datasets.dump_entries(entries[:2], path)
    datasets.dump_entries(entries[2:], path, append=True)
    back = datasets.load_entries(path)
    assert len(back) == 5 and np.array_equal(back[3]["x_seq"], entries[3]["x_seq"])
    df = datasets.entries_to_dataframe(back)
    assert {"x_seq", "k_seq", "K_seq", "iteration"} <= set(df.columns)
    assert np.array(df.iloc[0]["K_seq"], dtype=np.float32).shape == (N, m, n)
