"""Performance-buffer population and prediction-guided task selection for 2 objectives
(mirror of morl/population_2d.py:123-319, same public API).

Host: buffer bookkeeping and candidate-weight generation (tiny, sequential). Device: every model fit
of the call in one K4 launch, and the whole greedy scoring loop (exact sort-based hypervolume,
sparsity, Pareto filtering, arg-max) in K5 -- float64, bit-exact with the reference's arithmetic."""
from copy import deepcopy

import numpy as np

from . import kernels as K
from .prediction import Candidates, predict_candidates, prefetch
from .utils import get_ep_indices, norm2, rownorm


class Population:
    def __init__(self, args):
        self.sample_batch = []
        self.pbuffer_num = args.pbuffer_num
        self.pbuffer_size = args.pbuffer_size
        self.dtheta = np.pi / 2.0 / self.pbuffer_num
        self.z_min = np.zeros(args.obj_num)
        self.pbuffers = None
        self.pbuffer_dist = None
        self.last_fits = None          # record of the most recent K4 launch (diagnostics / tests)

    # ------------------------------------------------------------------ performance buffers
    def insert_pbuffer(self, index, objs):
        """Angular buffer = floor(arccos(f1/|f|) / dtheta); keep the pbuffer_size farthest points per buffer,
        strict '<' so earlier samples win ties (population_2d.py:136-165)."""
        f = objs - self.z_min
        if np.min(f) < 1e-7:
            return False
        dist = np.linalg.norm(f)
        buffer_id = int(np.arccos(np.clip(f[1] / dist, -1.0, 1.0)) // self.dtheta)
        if buffer_id < 0 or buffer_id >= self.pbuffer_num:
            return False
        return self._insert_sorted(index, dist, buffer_id)

    def _insert_sorted(self, index, dist, buffer_id):
        ids, dists = self.pbuffers[buffer_id], self.pbuffer_dist[buffer_id]
        pos = next((i for i, dcur in enumerate(dists) if dcur < dist), None)
        if pos is not None:
            ids.insert(pos, index); dists.insert(pos, dist)
            del ids[self.pbuffer_size:], dists[self.pbuffer_size:]
            return True
        if len(ids) < self.pbuffer_size:
            ids.append(index); dists.append(dist)
            return True
        return False

    def update(self, sample_batch):
        """population = performance-buffer selection over (population + offspring) (population_2d.py:169-183)."""
        everyone = self.sample_batch + sample_batch
        self.pbuffers = [[] for _ in range(self.pbuffer_num)]
        self.pbuffer_dist = [[] for _ in range(self.pbuffer_num)]
        if everyone:
            # distance and angular buffer of every sample at once (the expressions of insert_pbuffer, element by element,
            # with the bits of the scalar calls: utils.rownorm); the insertions themselves stay sequential, in order
            F = np.array([np.asarray(s.objs, dtype=np.float64) for s in everyone]).reshape(len(everyone), -1) - self.z_min
            with np.errstate(all="ignore"):
                dist = rownorm(F)
                bid = np.arccos(np.clip(F[:, 1] / dist, -1.0, 1.0)) // self.dtheta
            ok = (F.min(axis=1) >= 1e-7) & (bid >= 0) & (bid < self.pbuffer_num)
            for i in np.nonzero(ok)[0].tolist():
                self._insert_sorted(i, float(dist[i]), int(bid[i]))
        self.sample_batch = [everyone[i] for buf in self.pbuffers for i in buf]

    # ------------------------------------------------------------------ metrics (device)
    def compute_hypervolume(self, objs_batch):
        return K.front_metrics(np.array(objs_batch, dtype=np.float64))[0]

    def compute_sparsity(self, objs_batch):
        return K.front_metrics(np.array(objs_batch, dtype=np.float64))[1]

    def _score(self, candidates, mask, virtual_ep_objs_batch):
        mask = np.asarray(mask, dtype=bool)
        hv, sp = np.zeros(len(candidates)), np.zeros(len(candidates))
        if mask.any():
            pred = np.array([candidates[i]['prediction'] for i in np.nonzero(mask)[0]], dtype=np.float64)
            _, h, s, _ = K.select_greedy(np.array(virtual_ep_objs_batch, dtype=np.float64).reshape(-1, 2), pred, 0.0, 1)
            hv[mask], sp[mask] = h[0], s[0]
        return hv, sp

    def _score_fork(self, cand_pred, mask, virtual_ep_objs_batch):
        """Scorer of the reference's fork copy (WorkingMorl/morl/population_2d.py:207-226): update_ep with its 1e-5
        tolerances, InnerHyperVolume (round(hv, 4)) and the M-D sparsity. The dimension sweep over 2-objective points is
        the per-slice area routine of the 3-objective kernel, so the points are lifted to z = 1 (x * 1.0 and + 0.0 are
        exact) and scored by the 3-D kernel: one CTA per candidate, one launch per round."""
        mask = np.asarray(mask, dtype=bool)
        hv, sp = np.zeros(len(cand_pred)), np.zeros(len(cand_pred))
        if mask.any():
            lift = lambda a: np.concatenate([np.asarray(a, dtype=np.float64).reshape(-1, 2), np.ones((len(a), 1))], axis=1)
            _, h, s, _ = K.select_greedy(lift(virtual_ep_objs_batch), lift(cand_pred[mask]), 0.0, 1)
            hv[mask], sp[mask] = h[0], s[0]
        return hv, sp

    def evaluate_hv(self, candidates, mask, virtual_ep_objs_batch):
        return self._score(candidates, mask, virtual_ep_objs_batch)[0].tolist()

    def evaluate_sparsity(self, candidates, mask, virtual_ep_objs_batch):
        return self._score(candidates, mask, virtual_ep_objs_batch)[1].tolist()

    # ------------------------------------------------------------------ selection
    def _test_weights(self, opt_graph, sample, num_weights):
        """num_weights unit vectors on an arc of +-45 degrees around the sample's last weight; drop those
        outside the first quadrant or within 1e-3 of an existing successor's weight (population_2d.py:238-254)."""
        center = opt_graph.weights[sample.optgraph_id]
        angle_center = np.arctan2(center[1], center[0])
        lo, hi = angle_center - np.pi / 4., angle_center + np.pi / 4.
        succ_w = []
        for s in opt_graph.succ[sample.optgraph_id]:
            w = deepcopy(opt_graph.weights[s])
            succ_w.append(w / np.linalg.norm(w))
        out = []
        for i in range(num_weights):
            angle = lo + (hi - lo) / (num_weights - 1) * i
            weight = np.array([np.cos(angle), np.sin(angle)])
            if weight[0] >= -1e-7 and weight[1] >= -1e-7:
                if not any(norm2(w - weight) < 1e-3 for w in succ_w):
                    out.append(weight)
        return out

    def _test_weights_batch(self, view, node_ids, num_weights):
        """`_test_weights` for many members at once, same arithmetic element by element (the row-wise helpers of
        utils.py give the bits of the scalar calls): -> (tests [n, num_weights, 2] with the kept weights first, in arc
        order, padding = 1; counts [n])."""
        n = len(node_ids)
        center = view.weights[node_ids]
        angle_center = np.arctan2(center[:, 1], center[:, 0])
        lo, hi = angle_center - np.pi / 4., angle_center + np.pi / 4.
        angle = lo[:, None] + ((hi - lo) / (num_weights - 1))[:, None] * np.arange(num_weights)[None, :]
        weight = np.stack([np.cos(angle), np.sin(angle)], axis=2)
        keep = (weight[:, :, 0] >= -1e-7) & (weight[:, :, 1] >= -1e-7)
        member, succ = view.successors_of(node_ids)
        if len(member):
            w = view.weights[succ]
            succ_w = w / rownorm(w)[:, None]
            used = rownorm(succ_w[:, None, :] - weight[member]) < 1e-3            # [pairs, num_weights]
            hit = np.zeros((n, num_weights), dtype=np.int64)
            np.add.at(hit, member, used)
            keep &= hit == 0
        counts = keep.sum(axis=1)
        order = np.argsort(~keep, axis=1, kind="stable")                          # kept weights first, arc order preserved
        tests = np.take_along_axis(weight, order[:, :, None], axis=1)
        tests[np.arange(num_weights)[None, :] >= counts[:, None]] = 1.0
        return tests, counts

    def prefetch_fits(self, args, opt_graph):
        """Optional: launch the model fits of the current population now (after `update`), so that they run under whatever
        the caller does before `prediction_guided_selection` (archive update, logging); results are identical."""
        self._pending = prefetch(opt_graph, self.sample_batch, args.obj_num, bool(getattr(args, 'fork_scoring', False)))

    def prediction_guided_selection(self, args, iteration, ep, opt_graph, scalarization_template):
        """Returns (elite_batch, scalarization_batch, predicted_offspring_objs) (population_2d.py:229-304)."""
        N = args.num_tasks
        # ---- prediction: candidates = (sample, weight) pairs with their predicted objectives
        # the fits need only the opt-graph: launch all of them (K4) first and enumerate the test weights while they run
        fork = bool(getattr(args, 'fork_scoring', False))     # the WorkingMorl/ copy's variant of this routine
        prep = {}

        def while_fitting():            # host work that does not need the fits, done while they run
            prep['virtual_ep'] = np.array([np.asarray(s.objs, dtype=np.float64) for s in ep.sample_batch]).reshape(-1, args.obj_num)
            prep['scalarizations'] = [deepcopy(scalarization_template) for _ in range(N)]
        pending, self._pending = getattr(self, '_pending', None), None
        tests, counts, pred, self.last_fits = predict_candidates(
            opt_graph, self.sample_batch, lambda view, ids: self._test_weights_batch(view, ids, args.num_weight_candidates),
            args.obj_num, cap_threshold=fork, max_tests=args.num_weight_candidates, zero_if_degenerate=fork,
            pending=pending, while_fitting=while_fitting)
        candidates = Candidates(self.sample_batch, tests, counts, pred)
        # ---- optimisation: greedy knapsack on the device
        virtual_ep = prep['virtual_ep']
        elite_batch, scalarization_batch, predicted_offspring_objs = [], [], []
        if len(candidates) == 0:
            print('Too few candidates')
            return elite_batch, scalarization_batch, predicted_offspring_objs
        cand_pred = np.ascontiguousarray(candidates.prediction, dtype=np.float64)
        if fork:
            best_ids, self.last_hv, self.last_sparsity = self._greedy_fork(virtual_ep, cand_pred, args.sparsity, N)
        else:
            best_ids, self.last_hv, self.last_sparsity, _ = K.select_greedy(virtual_ep, cand_pred, args.sparsity, N)
        for best_id in best_ids:
            if best_id == -1:
                print('Too few candidates')
                break
            c = candidates[int(best_id)]
            elite_batch.append(c['sample'])
            scalarization = prep['scalarizations'][len(scalarization_batch)]
            scalarization.update_weights(c['weight'] / np.sum(c['weight']))
            scalarization_batch.append(scalarization)
            predicted_offspring_objs.append(deepcopy(c['prediction']))
        self.last_candidates = candidates
        return elite_batch, scalarization_batch, predicted_offspring_objs

    def _greedy_fork(self, virtual_ep, cand_pred, alpha, N):
        """Greedy loop of the fork copy (WorkingMorl/morl/population_2d.py:266-306): scores from `_score_fork`, arg-max of
        hv - alpha * sparsity with strict `>`, virtual front rebuilt with get_ep_indices after every pick."""
        mask = np.ones(len(cand_pred), dtype=bool)
        vep = [np.asarray(p, dtype=np.float64) for p in virtual_ep]
        best_ids, hvs, sps = [], [], []
        for _ in range(N):
            hv, sp = self._score_fork(cand_pred, mask, vep)
            hvs.append(hv); sps.append(sp)
            best_id, max_metrics = -1, -np.inf
            for i in np.nonzero(mask)[0]:
                if hv[i] - alpha * sp[i] > max_metrics:
                    max_metrics, best_id = hv[i] - alpha * sp[i], int(i)
            best_ids.append(best_id)
            if best_id == -1:
                break
            mask[best_id] = False
            batch = np.array(vep + [cand_pred[best_id]])
            vep = [batch[i] for i in get_ep_indices(batch)]
        return best_ids, np.array(hvs), np.array(sps)

    def random_selection(self, args, scalarization_template):
        """population_2d.py:309-319 (numpy global RNG, like the reference)."""
        elite_batch, scalarization_batch = [], []
        for _ in range(args.num_tasks):
            elite_batch.append(self.sample_batch[np.random.choice(len(self.sample_batch))])
            weights = np.random.uniform(args.min_weight, args.max_weight, args.obj_num)
            scalarization = deepcopy(scalarization_template)
            scalarization.update_weights(weights / np.sum(weights))
            scalarization_batch.append(scalarization)
        return elite_batch, scalarization_batch
