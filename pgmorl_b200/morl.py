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
    """Final archive: policies, env params and objectives (morl.py:222-239)."""
    final = os.path.join(args.save_dir, 'final')
    os.makedirs(final, exist_ok=True)
    for i, sample in enumerate(ep.sample_batch):
        torch.save(sample.actor_critic.state_dict(), os.path.join(final, 'EP_policy_{}.pt'.format(i)))
        with open(os.path.join(final, 'EP_env_params_{}.pkl'.format(i)), 'wb') as fp:
            pickle.dump(sample.env_params, fp)
    with open(os.path.join(final, 'objs.txt'), 'w') as fp:
        _rows(fp, ep.obj_batch, args.obj_num)
    if args.obj_rms:
        with open(os.path.join(final, 'env_params.txt'), 'w') as fp:
            for sample in ep.sample_batch:
                fp.write('obj_rms: mean: {} var: {}\n'.format(sample.env_params['obj_rms'].mean,
                                                              sample.env_params['obj_rms'].var))


def run(args, device="cuda", cluster=0, save=True):
    """One PG-MORL run; returns (ep, population, opt_graph, timings) where timings lists, per generation,
    the seconds spent in the MOPG stage and in task selection."""
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

    elite_batch, scalarization_batch = initialize_warm_up_batch(args, device)
    rl_num_updates = args.warmup_iter
    for sample, scalarization in zip(elite_batch, scalarization_batch):
        sample.optgraph_id = opt_graph.insert(deepcopy(scalarization.weights), deepcopy(sample.objs), -1)

    episode, iteration, timings = 0, 0, []
    while iteration < total_num_updates:
        print_info('\n------------------------------- Warm-up Stage -------------------------------' if episode == 0 else
                   '\n-------------------- Evolutionary Stage: Generation {:3} --------------------'.format(episode))
        episode += 1
        task_batch = [Task(elite, scalarization) for elite, scalarization in zip(elite_batch, scalarization_batch)]
        t0 = time.time()
        produced = mopg_population_update(args, task_batch, device, iteration, rl_num_updates, start_time, cluster=cluster)
        t_mopg = time.time() - t0
        all_offspring_batch = [[Sample.copy_from(s) for s in offsprings] for offsprings in produced]

        # every intermediate policy feeds the archive; every update_iter-th one becomes an opt-graph node and a
        # population candidate (morl.py:101-118)
        all_sample_batch, offspring_batch, last_offspring_batch = [], [], [None] * len(task_batch)
        for task_id, offsprings in enumerate(all_offspring_batch):
            prev_node_id = task_batch[task_id].sample.optgraph_id
            opt_weights = deepcopy(task_batch[task_id].scalarization.weights).detach().numpy()
            for i, sample in enumerate(offsprings):
                all_sample_batch.append(sample)
                if (i + 1) % args.update_iter == 0:
                    prev_node_id = opt_graph.insert(opt_weights, deepcopy(sample.objs), prev_node_id)
                    sample.optgraph_id = prev_node_id
                    offspring_batch.append(sample)
            last_offspring_batch[task_id] = offsprings[-1]

        ep.update(all_sample_batch)
        population.update(offspring_batch)

        t0 = time.time()
        elite_batch, scalarization_batch, predicted = select_tasks(args, population, ep, opt_graph, template, iteration,
                                                                   rl_num_updates, total_num_updates, last_offspring_batch)
        timings.append({"mopg_s": t_mopg, "selection_s": time.time() - t0})
        print_info('Selected Tasks:')
        for elite, scalarization in zip(elite_batch, scalarization_batch):
            print_info('objs = {}, weight = {}'.format(elite.objs, scalarization.weights))

        iteration = min(iteration + rl_num_updates, total_num_updates)
        rl_num_updates = args.update_iter
        if save:
            save_generation(args, iteration, ep, population, opt_graph, elite_batch, scalarization_batch, predicted,
                            all_offspring_batch)
    if save:
        save_final(args, ep)
    return ep, population, opt_graph, timings
