"""Population sharding across GPUs (one process per GPU, torch.distributed; NCCL over NVLink on the B200 box,
gloo in the CPU tests).

Tasks are independent during MOPG (no cross-task term in morl/mopg.py or algo/ppo.py), so rank r of W owns the
tasks {i : i % W == r}, keeps their parameters / Adam state / rollout buffers in its own HBM and runs K1-K3 with
no communication for a whole generation. ONE exchange per generation (the reference's results_queue traffic,
morl/morl.py:93-118): an all-gather of a packed float64 record per task -- scalarisation weight, parent
opt-graph node, and the objective vector of every iteration -- after which every rank holds identical
OptGraph / Population / EP metadata and runs the deterministic float64 selection (K4 + K5) redundantly, so
all ranks agree on the next (elite, weight) tasks without another collective. Policy state moves only when a
selected elite is owned by a different rank than the task it will train as: point-to-point send/recv of ONE float64
vector per elite (params, exp_avg, exp_avg_sq, step, lr and the running observation / return / objective moments,
`pack_sample_state`). The generation loop that uses these helpers is `pgmorl_b200.morl.run` under torchrun.
"""
import numpy as np
import torch
import torch.distributed as dist

from ._nvtx import rng as _nvtx


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
    with _nvtx("dist.all_gather_records"):
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


RMS_KEYS = ('ob_rms', 'ret_rms', 'obj_rms')


def _rms_cap(dims):
    """Largest element count of each running-moment vector (ob_rms [O]; ret_rms scalar; obj_rms scalar until its first
    update, then [M] -- running_mean_std.py:4-9, a2c/envs.py:197-211)."""
    return {'ob_rms': dims.obs, 'ret_rms': 1, 'obj_rms': dims.obj}


def sample_state_len(dims):
    """Length of pack_sample_state() for the network shape `dims` (same for every sample of a run)."""
    return 3 * dims.n_par + 2 + sum(3 + 2 * k for k in _rms_cap(dims).values())


def pack_sample_state(sample):
    """Everything a Sample needs to keep training on another rank, as ONE fixed-length float64 vector (lossless: the
    float32 parameters / Adam moments are exactly representable, step and learning rate travel as float64, and the
    float64 running moments of the observation / return / objective normalisation -- `env_params`, morl/sample.py:12,
    mopg.py:70-75 -- go bit for bit):
        [params n | exp_avg n | exp_avg_sq n | step | lr | per rms key: present, ndim, count, mean (padded), var (padded)]"""
    ac, opt = sample.actor_critic, sample.agent.optimizer
    parts = [ac.flat.detach().to(torch.float64).cpu(), opt.exp_avg.detach().to(torch.float64).cpu(),
             opt.exp_avg_sq.detach().to(torch.float64).cpu(),
             torch.tensor([float(opt.step_count), float(opt.param_groups[0]["lr"])], dtype=torch.float64)]
    for key, cap in _rms_cap(ac.dims).items():
        rms = sample.env_params[key]
        row = np.zeros(3 + 2 * cap)
        if rms is not None:
            mean, var = np.asarray(rms.mean, dtype=np.float64), np.asarray(rms.var, dtype=np.float64)
            assert mean.size <= cap and var.shape == mean.shape, (key, mean.shape, cap)
            row[0], row[1], row[2] = 1.0, float(mean.ndim), float(getattr(rms, 'count', 0.0))
            row[3:3 + mean.size] = mean.reshape(-1)
            row[3 + cap:3 + cap + var.size] = var.reshape(-1)
        parts.append(torch.from_numpy(row))
    out = torch.cat(parts).contiguous()
    assert out.numel() == sample_state_len(ac.dims)
    return out


def unpack_sample_state(payload, template, objs=None, optgraph_id=None):
    """Inverse of pack_sample_state: a full Sample built from a deep copy of `template` (any local Sample of the same
    run: it provides the object structure -- Policy, PPO agent, running-moment classes) with every number replaced."""
    from .sample import Sample
    payload = payload.detach().to(torch.float64).cpu()
    s = Sample.copy_from(template)
    dims = s.actor_critic.dims
    n, dev = dims.n_par, s.actor_critic.flat.device
    s.actor_critic.flat = payload[:n].to(torch.float32).to(dev).contiguous()
    opt = s.agent.optimizer
    opt.exp_avg = payload[n:2 * n].to(torch.float32).to(dev).contiguous()
    opt.exp_avg_sq = payload[2 * n:3 * n].to(torch.float32).to(dev).contiguous()
    opt.step_count = int(payload[3 * n].item())
    opt.param_groups[0]["lr"] = float(payload[3 * n + 1].item())
    off = 3 * n + 2
    for key, cap in _rms_cap(dims).items():
        row = payload[off:off + 3 + 2 * cap].numpy()
        off += 3 + 2 * cap
        rms = s.env_params[key]
        assert (row[0] != 0.0) == (rms is not None), f"env_params[{key}] presence differs between ranks"
        if rms is None:
            continue
        k = cap if row[1] else 1
        shape = (k,) if row[1] else ()
        if hasattr(rms, 'count'):
            rms.count = float(row[2])
        rms.mean = row[3:3 + k].copy().reshape(shape)
        rms.var = row[3 + cap:3 + cap + k].copy().reshape(shape)
    s.objs, s.optgraph_id = objs, optgraph_id
    return s


def migrate_states(plan, get_state, put_state, n_state, device=None, dtype=torch.float64):
    """Execute a migration plan. get_state(new_task) -> tensor [n_state] (pack_sample_state) on the source rank;
    put_state(new_task, tensor) on the destination rank. One batched isend/irecv round (NCCL: point-to-point over
    NVLink, grouped)."""
    rank, W = world()
    if W == 1:
        return
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    ops, bufs = [], []
    for task, src, dst in plan:
        if rank == src:
            t = get_state(task).to(device=dev, dtype=dtype).contiguous()
            assert t.numel() == n_state, (t.numel(), n_state)
            ops.append(dist.P2POp(dist.isend, t, dst, tag=task)); bufs.append((None, t))
        elif rank == dst:
            t = torch.empty(n_state, dtype=dtype, device=dev)
            ops.append(dist.P2POp(dist.irecv, t, src, tag=task)); bufs.append((task, t))
    if ops:
        with _nvtx("dist.migrate_states"):
            for req in dist.batch_isend_irecv(ops):
                req.wait()
    for task, t in bufs:
        if task is not None:
            put_state(task, t)


def warm_up_p2p(device=None, n_elems=1, rounds=2):
    """Establish the point-to-point connection of every rank pair once (NCCL opens a pair's channels lazily at first use,
    ~20-50 ms each: without this the first generations of an 8-GPU run spend up to 1.5 s in `migrate_states` while new
    pairs keep appearing). `rounds` batched exchanges of `n_elems` float64 with every peer -- pass the size of one migrated
    state (`sample_state_len`) so that the channels a message of that size uses are all open."""
    rank, W = world()
    if W == 1:
        return
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.full((max(int(n_elems), 1),), float(rank), dtype=torch.float64, device=dev)
    recv = [torch.empty_like(send) for _ in range(W)]
    for _ in range(rounds):
        ops = []
        for peer in range(W):
            if peer != rank:
                ops.append(dist.P2POp(dist.isend, send, peer))
                ops.append(dist.P2POp(dist.irecv, recv[peer], peer))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    assert all(int(recv[p][0].item()) == p for p in range(W) if p != rank)


def all_gather_rows(local, n_total, device=None):
    """All-gather a float64 table whose row i lives on rank i % W (position i // W there): `local` = this rank's rows in
    that order; returns the [n_total, R] table in global row order, bit-identical on every rank."""
    rank, W = world()
    local = np.asarray(local, dtype=np.float64)
    if W == 1:
        return local
    R = local.shape[1]
    per = (n_total + W - 1) // W
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.zeros(per, R, dtype=torch.float64, device=dev)
    send[:len(local)] = torch.from_numpy(local).to(dev)
    recv = torch.empty(W * per, R, dtype=torch.float64, device=dev)
    with _nvtx("dist.all_gather_rows"):
        dist.all_gather_into_tensor(recv, send)
    table = recv.cpu().numpy().reshape(W, per, R)
    return np.stack([table[i % W, i // W] for i in range(n_total)]) if n_total else np.zeros((0, R))


def all_reduce_rows(rows, device=None):
    """Sum a float64 table over the ranks (each rank fills the rows it owns, zeros elsewhere)."""
    rank, W = world()
    rows = np.asarray(rows, dtype=np.float64)
    if W == 1:
        return rows
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.from_numpy(rows.copy()).to(dev)
    dist.all_reduce(t)
    return t.cpu().numpy()
