"""The PG-MORL generation loop on a persistent device population (counterpart of morl/morl.py:28-239).

Same sequence and bookkeeping as the reference's `run(args)` -- warm-up tasks, then per generation: MOPG on every
(policy, weight) task, opt-graph / Pareto-archive / population update, task selection, result files -- but the
process-per-task spawn / Queue / pickle round trip (morl.py:79-99) is replaced by ONE batched call,
`mopg.mopg_population_update`, that advances all tasks together on the GPU. The text files written per generation
and the final archive use the reference's formats (morl.py:181-239) so its plotting scripts read them unchanged.
"""
import os
import pickle
import time
from copy import deepcopy

import numpy as np
import torch

from . import dist as pdist
from .ep import EP
from .mopg import mopg_population_update
from .opt_graph import OptGraph
from .population_2d import Population as Population2d
from .population_3d import Population as Population3d
from .sample import Sample, Task
from .scalarization_methods import WeightedSumScalarization
from .utils import generate_weights_batch_dfs, print_info
from .warm_up import initialize_warm_up_batch


def _grid_scalarizations(args, template):
    """One scalarisation per weight of the evenly spaced simplex grid (used by the moead / ra baselines)."""
    weights_batch, out = [], []
    generate_weights_batch_dfs(0, args.obj_num, args.min_weight, args.max_weight, args.delta_weight, [], weights_batch)
    for weights in weights_batch:
        s = deepcopy(template)
        s.update_weights(weights)
        out.append(s)
    return out


def select_tasks(args, population, ep, opt_graph, template, iteration, rl_num_updates, total_num_updates, last_offspring_batch):
    """-> (elite_batch, scalarization_batch, predicted_offspring_objs or None) for args.selection_method
    (morl.py:128-169)."""
    method = args.selection_method
    if method == 'prediction-guided':
        return population.prediction_guided_selection(args, iteration, ep, opt_graph, template)
    if method == 'random':
        return population.random_selection(args, template) + (None,)
    if method == 'moead':        # for each grid weight, the population member with the best scalarised value
        scals, elites = _grid_scalarizations(args, template), []
        for s in scals:
            values = [float(s.evaluate(torch.as_tensor(np.asarray(sample.objs, dtype=np.float64)))) for sample in population.sample_batch]
            best, best_value = None, -np.inf
            for sample, value in zip(population.sample_batch, values):      # strict '>' : first maximum wins
                if value > best_value:
                    best, best_value = sample, value
            elites.append(best)
        return elites, scals, None
    if method == 'ra':           # every task keeps its own last offspring and its grid weight
        return last_offspring_batch, _grid_scalarizations(args, template), None
    if method == 'pfa':          # weights slide from one grid point towards the next as training progresses
        if args.obj_num > 2:
            raise NotImplementedError
        ratio = np.clip((iteration + rl_num_updates + args.update_iter - args.warmup_iter)
                        / (total_num_updates - args.warmup_iter), 0.0, 1.0)
        scals = []
        for w0 in np.arange(args.min_weight, args.max_weight + 0.5 * args.delta_weight, args.delta_weight):
            w = np.clip(w0 + ratio * args.delta_weight, args.min_weight, args.max_weight)
            s = deepcopy(template)
            s.update_weights(np.array([abs(w), abs(1.0 - w)]))
            scals.append(s)
        return last_offspring_batch, scals, None
    raise NotImplementedError(method)


def _rows(fp, rows, n):
    fmt = '{:5f}' + (n - 1) * ',{:5f}' + '\n'
    for r in rows:
        fp.write(fmt.format(*r))


def save_generation(args, iteration, ep, population, opt_graph, elite_batch, scalarization_batch,
                    predicted_offspring_objs, all_offspring_batch):
    """The per-generation text dumps of morl.py:181-218, same file names and number formats."""
    M = args.obj_num
    base = os.path.join(args.save_dir, str(iteration))
    for sub in ('ep', 'population', 'elites'):
        os.makedirs(os.path.join(base, sub), exist_ok=True)
    with open(os.path.join(base, 'ep', 'objs.txt'), 'w') as fp:
        _rows(fp, ep.obj_batch, M)
    with open(os.path.join(base, 'population', 'objs.txt'), 'w') as fp:
        _rows(fp, [s.objs for s in population.sample_batch], M)
    with open(os.path.join(base, 'population', 'optgraph.txt'), 'w') as fp:
        fp.write('{}\n'.format(len(opt_graph.objs)))
        fmt = '{:5f}' + (M - 1) * ',{:5f}' + ';{:5f}' + (M - 1) * ',{:5f}' + ';{}\n'
        for w, o, prev in zip(opt_graph.weights, opt_graph.objs, opt_graph.prev):
            fp.write(fmt.format(*w, *o, prev))
        fp.write('{}\n'.format(len(population.sample_batch)))
        for s in population.sample_batch:
            fp.write('{}\n'.format(s.optgraph_id))
    with open(os.path.join(base, 'elites', 'elites.txt'), 'w') as fp:
        _rows(fp, [e.objs for e in elite_batch], M)
    with open(os.path.join(base, 'elites', 'weights.txt'), 'w') as fp:
        _rows(fp, [s.weights for s in scalarization_batch], M)
    if args.selection_method == 'prediction-guided':
        with open(os.path.join(base, 'elites', 'predictions.txt'), 'w') as fp:
            _rows(fp, predicted_offspring_objs, M)
    with open(os.path.join(base, 'elites', 'offsprings.txt'), 'w') as fp:
        _rows(fp, [s.objs for task in all_offspring_batch for s in task], M)


def save_final(args, ep):
    """Final archive: policies, env params and objectives (morl.py:222-239). Sharded runs: every rank writes the
    EP_policy_i / EP_env_params_i files of the archive members whose state it owns (the ranks of one node share the
    file system); rank 0 writes the two text files, for which the objective moments of all members are summed over
    the ranks first."""
    rank, W = pdist.world()
    final = os.path.join(args.save_dir, 'final')
    os.makedirs(final, exist_ok=True)
    for i, sample in enumerate(ep.sample_batch):
        if sample.is_stub:
            continue
        torch.save(sample.actor_critic.state_dict(), os.path.join(final, 'EP_policy_{}.pt'.format(i)))
        with open(os.path.join(final, 'EP_env_params_{}.pkl'.format(i)), 'wb') as fp:
            pickle.dump(sample.env_params, fp)
    obj_rms = None
    if args.obj_rms:
        M = args.obj_num
        rows = np.zeros((len(ep.sample_batch), 2 * M))
        for i, sample in enumerate(ep.sample_batch):
            if not sample.is_stub:
                rows[i, :M] = np.asarray(sample.env_params['obj_rms'].mean, dtype=np.float64) * np.ones(M)
                rows[i, M:] = np.asarray(sample.env_params['obj_rms'].var, dtype=np.float64) * np.ones(M)
        obj_rms = pdist.all_reduce_rows(rows) if W > 1 else None
    if rank != 0:
        return
    with open(os.path.join(final, 'objs.txt'), 'w') as fp:
        _rows(fp, ep.obj_batch, args.obj_num)
    if args.obj_rms:
        with open(os.path.join(final, 'env_params.txt'), 'w') as fp:
            for i, sample in enumerate(ep.sample_batch):
                if W > 1:
                    mean, var = obj_rms[i, :args.obj_num], obj_rms[i, args.obj_num:]
                else:
                    mean, var = sample.env_params['obj_rms'].mean, sample.env_params['obj_rms'].var
                fp.write('obj_rms: mean: {} var: {}\n'.format(mean, var))


def run(args, device="cuda", cluster=0, save=True):
    """One PG-MORL run; returns (ep, population, opt_graph, timings) where timings lists, per generation,
    the seconds spent in the MOPG stage, the exchange and in task selection.

    Under torchrun (torch.distributed initialised, W ranks, one GPU each) the population is SHARDED (SURVEY 8(e),
    morl/morl.py:84-143 is the part that changes): task i of a generation trains on rank i % W, which holds the elite's
    policy / Adam state / running moments in its HBM; after MOPG the ranks all-gather ONE packed float64 record per task
    (task id, parent node, weight, objective vector of every iteration), every rank rebuilds the same offspring list --
    full Samples for its own tasks, metadata stubs for the others -- and runs the opt-graph / archive / population
    update and the deterministic FP64 selection redundantly, so all ranks pick the same (elite, weight) pairs without a
    second collective. An elite whose state lives on another rank than its next task moves point-to-point as one
    lossless float64 vector (`dist.pack_sample_state`). With W = 1 none of this triggers and the loop is the reference's."""
    rank, W = pdist.world()
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    device = torch.device(device)
    template = WeightedSumScalarization(num_objs=args.obj_num, weights=np.ones(args.obj_num) / args.obj_num)
    total_num_updates = int(args.num_env_steps) // args.num_steps // args.num_processes
    start_time = time.time()
    ep = EP()
    if args.obj_num == 2:
        population = Population2d(args)
    elif args.obj_num > 2:
        population = Population3d(args)
    else:
        raise NotImplementedError
    opt_graph = OptGraph()

    # every rank builds the whole warm-up batch (same RNG streams as the single-process run), then keeps the state of
    # the tasks it owns
    elite_batch, scalarization_batch = initialize_warm_up_batch(args, device)
    rl_num_updates = args.warmup_iter
    for sample, scalarization in zip(elite_batch, scalarization_batch):
        sample.optgraph_id = opt_graph.insert(deepcopy(scalarization.weights), deepcopy(sample.objs), -1)
    state_template = None
    if W > 1:
        state_template = Sample.copy_from(elite_batch[0])
        # open every rank pair's channels once, up front, at the size of one migrated state
        pdist.warm_up_p2p(device if device.type == "cuda" else None,
                          n_elems=pdist.sample_state_len(state_template.actor_critic.dims))
        for i, sample in enumerate(elite_batch):
            owner = pdist.owner_of(i, W)
            elite_batch[i] = sample if owner == rank else Sample.stub(sample.objs, sample.optgraph_id)
            elite_batch[i].owner = owner
    say = print_info if rank == 0 else (lambda *a, **k: None)

    episode, iteration, timings = 0, 0, []
    while iteration < total_num_updates:
        say('\n------------------------------- Warm-up Stage -------------------------------' if episode == 0 else
            '\n-------------------- Evolutionary Stage: Generation {:3} --------------------'.format(episode))
        episode += 1
        n_tasks = len(elite_batch)
        t0 = time.time()
        n_migrated = 0
        if W > 1:
            elite_batch, n_migrated = _migrate_elites(elite_batch, state_template, W, rank)
        t_migrate = time.time() - t0
        mine = pdist.shard_tasks(n_tasks, W, rank)
        # the deep copies of Task(...) are made for the tasks this rank trains (all of them when W = 1)
        task_batch = {i: Task(elite_batch[i], scalarization_batch[i]) for i in mine}
        t0 = time.time()
        produced = mopg_population_update(args, [task_batch[i] for i in mine], device, iteration, rl_num_updates,
                                          start_time, cluster=cluster) if mine else []
        t_mopg = time.time() - t0
        t0 = time.time()
        local_offspring = {i: [Sample.copy_from(s) for s in offsprings] for i, offsprings in zip(mine, produced)}
        if W > 1:
            all_offspring_batch = _exchange_offspring(args, elite_batch, scalarization_batch, local_offspring, mine, W, rank)
        else:
            all_offspring_batch = [local_offspring[i] for i in range(n_tasks)]
        t_exchange = time.time() - t0

        # every intermediate policy feeds the archive; every update_iter-th one becomes an opt-graph node and a
        # population candidate (morl.py:101-118)
        all_sample_batch, offspring_batch, last_offspring_batch = [], [], [None] * n_tasks
        for task_id, offsprings in enumerate(all_offspring_batch):
            prev_node_id = elite_batch[task_id].optgraph_id
            opt_weights = deepcopy(scalarization_batch[task_id].weights).detach().numpy()
            for i, sample in enumerate(offsprings):
                all_sample_batch.append(sample)
                if (i + 1) % args.update_iter == 0:
                    prev_node_id = opt_graph.insert(opt_weights, deepcopy(sample.objs), prev_node_id)
                    sample.optgraph_id = prev_node_id
                    offspring_batch.append(sample)
            last_offspring_batch[task_id] = offsprings[-1]

        # (the reference updates the archive first; the two updates are independent, and with the population done first
        # the model fits of the selection can already run on the device while the archive is filtered)
        population.update(offspring_batch)
        if args.selection_method == 'prediction-guided':
            population.prefetch_fits(args, opt_graph)
        ep.update(all_sample_batch)

        t0 = time.time()
        elite_batch, scalarization_batch, predicted = select_tasks(args, population, ep, opt_graph, template, iteration,
                                                                   rl_num_updates, total_num_updates, last_offspring_batch)
        timings.append({"mopg_s": t_mopg, "selection_s": time.time() - t0, "exchange_s": t_exchange, "migrate_s": t_migrate,
                        "migrated": n_migrated, "elite_nodes": [e.optgraph_id for e in elite_batch],
                        "weights": [np.asarray(s.weights, dtype=np.float64).copy() for s in scalarization_batch]})
        say('Selected Tasks:')
        for elite, scalarization in zip(elite_batch, scalarization_batch):
            say('objs = {}, weight = {}'.format(elite.objs, scalarization.weights))

        iteration = min(iteration + rl_num_updates, total_num_updates)
        rl_num_updates = args.update_iter
        if save and rank == 0:
            save_generation(args, iteration, ep, population, opt_graph, elite_batch, scalarization_batch, predicted,
                            all_offspring_batch)
    if save:
        save_final(args, ep)
    return ep, population, opt_graph, timings


def _migrate_elites(elite_batch, state_template, W, rank):
    """Move the state of every elite whose owner is not the rank of its next task (task i trains on rank i % W).
    Returns the elite list in which the entries this rank trains are full Samples, and the number of states moved."""
    dims = state_template.actor_critic.dims
    plan = pdist.plan_migration([e.owner for e in elite_batch], W)
    out = list(elite_batch)

    def put(task, payload):
        stub = elite_batch[task]
        out[task] = pdist.unpack_sample_state(payload, state_template, objs=deepcopy(stub.objs), optgraph_id=stub.optgraph_id)
        out[task].owner = rank

    pdist.migrate_states(plan, lambda task: pdist.pack_sample_state(elite_batch[task]), put, pdist.sample_state_len(dims))
    for i in pdist.shard_tasks(len(out), W, rank):
        assert not out[i].is_stub, f"rank {rank}: elite of task {i} has no state after migration"
    return out, len(plan)


def _exchange_offspring(args, elite_batch, scalarization_batch, local_offspring, mine, W, rank):
    """The results_queue traffic of morl.py:93-99 as one all-gather: packed float64 record per task -> offspring list
    of EVERY task on every rank (full Samples for local tasks, stubs with the gathered objectives for the others)."""
    M, n_tasks = args.obj_num, len(elite_batch)
    weights = lambda i: np.asarray(scalarization_batch[i].weights, dtype=np.float64)
    local = pdist.pack_records(mine, [elite_batch[i].optgraph_id for i in mine], [weights(i) for i in mine],
                               [np.stack([np.asarray(s.objs, dtype=np.float64) for s in local_offspring[i]]) for i in mine])
    n_iter = {len(local_offspring[i]) for i in mine}
    assert len(n_iter) <= 1
    if not mine:                                    # fewer tasks than ranks: contribute an empty, correctly shaped shard
        raise NotImplementedError("sharded run needs at least one task per rank")
    table = pdist.all_gather_records(local, n_tasks)
    out = []
    for task_id, parent, w, objs in pdist.unpack_records(table, M):
        owner = pdist.owner_of(task_id, W)
        assert parent == elite_batch[task_id].optgraph_id and np.array_equal(w, weights(task_id)), \
            f"rank {rank}: replicated metadata of task {task_id} diverged"
        if owner == rank:
            samples = local_offspring[task_id]
            for s, o in zip(samples, objs):
                assert np.array_equal(np.asarray(s.objs, dtype=np.float64), o)
        else:
            samples = [Sample.stub(o.copy()) for o in objs]
        for s in samples:
            s.owner = owner
        out.append(samples)
    return out
