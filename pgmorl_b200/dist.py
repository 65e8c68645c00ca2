"""Population sharding across GPUs (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box,
gloo in the CPU tests).

Tasks are independent during MOPG (no cross-task term in morl/mopg.py or algo/ppo.py), so rank r of W owns the
tasks {i : i % W == r}, keeps their parameters / Adam state / rollout buffers in its own HBM and runs K1-K3 with
no communication for a whole generation. ONE exchange per generation (the reference's results_queue traffic,
morl/morl.py:93-118): an all-gather of a packed float64 record per task -- scalarisation weight, parent
opt-graph node, and the objective vector of every iteration -- after which every rank holds identical
OptGraph / Population / EP metadata and runs the deterministic float64 selection (K4 + K5) redundantly, so
all ranks agree on the next (elite, weight) tasks without another collective. Policy state moves only when a
selected elite is owned by a different rank than the task it will train as: point-to-point send/recv of
(params, exp_avg, exp_avg_sq, step, lr).
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def owner_of(task_id, world_size):
    return task_id % world_size


def shard_tasks(n_tasks, world_size, rank):
    """Task ids owned by `rank` (round robin keeps shards within one task of each other)."""
    return [i for i in range(n_tasks) if owner_of(i, world_size) == rank]


def pack_records(task_ids, parent_nodes, weights, objs_per_iter):
    """-> float64 [n_local, 2 + M + I*M]: task id, parent node, weight[M], objs[I, M] row-major."""
    rows = []
    for t, p, w, o in zip(task_ids, parent_nodes, weights, objs_per_iter):
        rows.append(np.concatenate([[float(t), float(p)], np.asarray(w, dtype=np.float64).reshape(-1),
                                    np.asarray(o, dtype=np.float64).reshape(-1)]))
    return np.array(rows, dtype=np.float64).reshape(len(rows), -1)


def unpack_records(table, M):
    """Inverse of pack_records, rows sorted by task id: list of (task_id, parent, weight[M], objs[I,M])."""
    out = []
    for row in table[np.argsort(table[:, 0], kind="stable")]:
        out.append((int(row[0]), int(row[1]), row[2:2 + M].copy(), row[2 + M:].reshape(-1, M).copy()))
    return out


def all_gather_records(local, n_tasks, device=None):
    """All-gather the per-task records of every rank; returns the [n_tasks, R] table in task-id order, identical
    (bit for bit) on every rank. `local` is a float64 numpy array [n_local, R]; shards may be uneven."""
    rank, W = world()
    local = np.asarray(local, dtype=np.float64)
    R = local.shape[1]
    if W == 1:
        return local[np.argsort(local[:, 0], kind="stable")]
    per = (n_tasks + W - 1) // W
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.full((per, R), -1.0, dtype=torch.float64, device=dev)       # task id -1 marks padding
    send[:len(local)] = torch.from_numpy(local).to(dev)
    recv = torch.empty(W * per, R, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    table = recv.cpu().numpy()
    table = table[table[:, 0] >= 0]
    assert len(table) == n_tasks, (len(table), n_tasks)
    return table[np.argsort(table[:, 0], kind="stable")]


def plan_migration(elite_owner_ranks, world_size):
    """New task i is trained by rank i % W; its elite's state lives on elite_owner_ranks[i].
    Returns [(new_task, src_rank, dst_rank)] for the states that must move."""
    return [(i, src, owner_of(i, world_size)) for i, src in enumerate(elite_owner_ranks)
            if src != owner_of(i, world_size)]


def migrate_states(plan, get_state, put_state, n_par, device=None):
    """Execute a migration plan. get_state(new_task) -> float32 tensor [3*n_par + 2] (params, exp_avg, exp_avg_sq,
    step, lr) on the source rank; put_state(new_task, tensor) on the destination rank."""
    rank, W = world()
    if W == 1:
        return
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    ops, bufs = [], []
    for task, src, dst in plan:
        if rank == src:
            t = get_state(task).to(dev).contiguous()
            ops.append(dist.P2POp(dist.isend, t, dst, tag=task)); bufs.append((None, t))
        elif rank == dst:
            t = torch.empty(3 * n_par + 2, dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, t, src, tag=task)); bufs.append((task, t))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for task, t in bufs:
        if task is not None:
            put_state(task, t)
