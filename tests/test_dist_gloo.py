"""world_size-2 gloo tests (CPU) of the population-sharding host logic: record packing, the per-generation
all-gather, and elite-state migration."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pgmorl_b200 import dist as pd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_tasks, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        M, I = 3, 4
        mine = pd.shard_tasks(n_tasks, world, rank)
        rng = lambda t: np.random.RandomState(100 + t)
        local = pd.pack_records(mine, [t + 50 for t in mine], [rng(t).rand(M) for t in mine],
                                [rng(t).rand(I, M) for t in mine])
        table = pd.all_gather_records(local, n_tasks)
        recs = pd.unpack_records(table, M)
        assert [r[0] for r in recs] == list(range(n_tasks)) and [r[1] for r in recs] == [t + 50 for t in range(n_tasks)]
        for t, _, w, o in recs:
            assert np.array_equal(w, rng(t).rand(M)) and o.shape == (I, M)
        np.save(os.path.join(out_dir, f"table{rank}.npy"), table)
        # migration: elite of new task i currently lives on rank owners[i]
        n_par = 7
        owners = [(3 * i + 1) % world for i in range(n_tasks)]
        plan = pd.plan_migration(owners, world)
        got = {}
        state = lambda task: torch.arange(3 * n_par + 2, dtype=torch.float64) + 1000.0 * task + 1e-13   # needs float64 to survive
        pd.migrate_states(plan, state, lambda task, t: got.__setitem__(task, t.clone()), 3 * n_par + 2)
        expect = [task for task, src, dst in plan if dst == rank]
        assert sorted(got) == sorted(expect)
        for task, t in got.items():
            assert torch.equal(t.cpu(), state(task))
    finally:
        dist.destroy_process_group()


def test_shard_and_plan_logic():
    assert pd.shard_tasks(7, 2, 0) == [0, 2, 4, 6] and pd.shard_tasks(7, 2, 1) == [1, 3, 5]
    assert sorted(sum((pd.shard_tasks(64, 8, r) for r in range(8)), [])) == list(range(64))
    assert pd.plan_migration([0, 0, 1, 1], 2) == [(1, 0, 1), (2, 1, 0)]
    local = pd.pack_records([2, 0], [5, 6], [np.array([.2, .8])] * 2, [np.arange(6.).reshape(3, 2)] * 2)
    table = pd.all_gather_records(local, 2)            # world size 1: just sorted by task id
    assert table[:, 0].tolist() == [0.0, 2.0]
    assert pd.unpack_records(table, 2)[1][3].shape == (3, 2)


def test_all_gather_and_migration_world2(tmp_path):
    n_tasks = 7                                            # uneven shards: 4 + 3
    mp.spawn(_worker, args=(2, _free_port(), n_tasks, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "table0.npy"), np.load(tmp_path / "table1.npy")
    assert a.shape == (n_tasks, 2 + 3 + 4 * 3) and np.array_equal(a, b)    # identical metadata on every rank
