"""Performance-buffer population and prediction-guided task selection for >= 3 objectives
(mirror of morl/population_3d.py:115-345, same public API).

The reference scores every candidate in its own OS process (population_3d.py:216-237); here one CTA per
candidate computes update_ep + the exact slice-based hypervolume + sparsity, and the greedy loop never
leaves the device (K5). All model fits of the call run in one K4 launch."""
from copy import deepcopy

import numpy as np

from . import kernels as K
from .prediction import Candidates, predict_candidates, prefetch
from .utils import generate_weights_batch_dfs, norm2, rowdot, rownorm


class Population:
    def __init__(self, args):
        self.sample_batch = []
        self.pbuffer_size = args.pbuffer_size
        self.obj_num = args.obj_num
        self.z_min = np.zeros(args.obj_num)
        self.pbuffer_vec = []
        generate_weights_batch_dfs(0, args.obj_num, 0.0, 1.0, 1.0 / (args.pbuffer_num - 1), [], self.pbuffer_vec)
        for i in range(len(self.pbuffer_vec)):
            self.pbuffer_vec[i] = self.pbuffer_vec[i] / np.linalg.norm(self.pbuffer_vec[i])
        self.pbuffer_num = len(self.pbuffer_vec)
        self.pbuffers = [[] for _ in range(self.pbuffer_num)]
        self.pbuffer_dist = [[] for _ in range(self.pbuffer_num)]
        self.last_fits = None

    # ------------------------------------------------------------------ performance buffers
    def find_buffer_id(self, f):
        """First direction with the largest dot product (population_3d.py:129-135)."""
        best, buffer_id = -np.inf, -1
        for i in range(self.pbuffer_num):
            dot = np.dot(self.pbuffer_vec[i], f)
            if dot > best:
                best, buffer_id = dot, i
        return buffer_id

    def insert_pbuffer(self, index, objs, enforce):
        """population_3d.py:137-172."""
        f = objs - self.z_min
        if np.min(f) < 1e-7:
            return False
        return self._insert_sorted(index, np.linalg.norm(f), self.find_buffer_id(f), enforce)

    def _insert_sorted(self, index, dist, buffer_id, enforce):
        ids, dists = self.pbuffers[buffer_id], self.pbuffer_dist[buffer_id]
        pos = next((i for i, dcur in enumerate(dists) if dcur < dist), None)
        if enforce:
            if pos is None:
                ids.append(index); dists.append(dist)
            else:
                ids.insert(pos, index); dists.insert(pos, dist)
            return True
        if pos is not None:
            ids.insert(pos, index); dists.insert(pos, dist)
            del ids[self.pbuffer_size:], dists[self.pbuffer_size:]
            return True
        if len(ids) < self.pbuffer_size:
            ids.append(index); dists.append(dist)
            return True
        return False

    def update(self, sample_batch):
        everyone = self.sample_batch + sample_batch
        self.pbuffers = [[] for _ in range(self.pbuffer_num)]
        self.pbuffer_dist = [[] for _ in range(self.pbuffer_num)]
        if everyone:
            # distance and buffer (first direction with the largest dot product) of every sample at once -- the reference's
            # loop is n_samples x pbuffer_num scalar np.dot calls (100 ms per generation at full size); utils.rowdot /
            # rownorm give the bits of those calls, np.argmax the first maximum like the strict '>' scan. The insertions
            # themselves stay sequential, in order.
            F = np.array([np.asarray(s.objs, dtype=np.float64) for s in everyone]).reshape(len(everyone), -1) - self.z_min
            dist = rownorm(F)
            bid = np.argmax(rowdot(np.array(self.pbuffer_vec)[None, :, :], F[:, None, :]), axis=1)
            for i in np.nonzero(F.min(axis=1) >= 1e-7)[0].tolist():
                self._insert_sorted(i, float(dist[i]), int(bid[i]), False)
        self.sample_batch = [everyone[i] for buf in self.pbuffers for i in buf]

    # ------------------------------------------------------------------ metrics (device)
    def evaluate_hypervolume_sparsity(self, candidates, mask, virtual_ep_objs_batch):
        mask = np.asarray(mask, dtype=bool)
        hv, sp = np.zeros(len(candidates)), np.zeros(len(candidates))
        if mask.any():
            pred = np.array([candidates[i]['prediction'] for i in np.nonzero(mask)[0]], dtype=np.float64)
            ep = np.array(virtual_ep_objs_batch, dtype=np.float64).reshape(-1, self.obj_num)
            _, h, s, _ = K.select_greedy(ep, pred, 0.0, 1)
            hv[mask], sp[mask] = h[0], s[0]
        return hv.tolist(), sp.tolist()

    def evaluate_hypervolume_sparsity_parallel(self, args, candidates, mask, virtual_ep_objs_batch):
        """Same call as the reference's process-per-candidate scorer; candidates are CTAs here."""
        return self.evaluate_hypervolume_sparsity(candidates, mask, virtual_ep_objs_batch)

    def evaluate_hv(self, candidates, mask, virtual_ep_objs_batch):
        return self.evaluate_hypervolume_sparsity(candidates, mask, virtual_ep_objs_batch)[0]

    def evaluate_sparsity(self, candidates, mask, virtual_ep_objs_batch):
        return self.evaluate_hypervolume_sparsity(candidates, mask, virtual_ep_objs_batch)[1]

    # ------------------------------------------------------------------ selection
    def _test_weights(self, args, opt_graph, sample, grid, grid_arr=None, grid_norm=None):
        """Centre weight + randomly ordered simplex-grid weights within 45 degrees of it, up to
        num_weight_candidates, skipping weights an existing successor already used (population_3d.py:248-284).
        Consumes numpy's global RNG exactly like the reference (np.random.shuffle).
        `grid_arr` / `grid_norm` (the grid as one array and its row norms) only save work: a weight whose largest
        coordinate difference from a vector exceeds 1e-3 cannot be within 1e-3 of it in L2 norm, so the exact test
        (same expression as the reference) runs only for the others; decisions are unchanged."""
        num_weights = args.num_weight_candidates
        center = opt_graph.weights[sample.optgraph_id]
        center = center / np.sum(center)
        succ_w = []
        for s in opt_graph.succ[sample.optgraph_id]:
            w = deepcopy(opt_graph.weights[s])
            succ_w.append(w / np.sum(w))
        if grid_arr is None:
            grid_arr = np.array(grid, dtype=np.float64)
            grid_norm = [norm2(w) for w in grid]
        close_to_center = (np.abs(grid_arr - center).max(axis=1) <= 1.0e-3).tolist()
        close_to_succ = [(np.abs(grid_arr - w).max(axis=1) <= 1.0e-3).tolist() for w in succ_w]
        used_exact = lambda weight: any(norm2(w - weight) < 1e-3 for w in succ_w)
        out = []
        if not used_exact(center):
            out.append(center)
        order = np.arange(len(grid))                  # same dtype and length as the reference's list-built array
        np.random.shuffle(order)
        norm_center = norm2(center)
        quarter = np.pi / 4.0
        for i in order.tolist():
            if len(out) >= num_weights:
                break
            weight = grid[i]
            if close_to_center[i] and norm2(weight - center) < 1e-3:
                continue
            cosine = np.dot(center, weight) / norm_center / grid_norm[i]
            angle = np.arccos(min(max(cosine, -1.0), 1.0))
            if angle < quarter:
                if any(c[i] and norm2(w - weight) < 1e-3 for c, w in zip(close_to_succ, succ_w)):
                    continue
                out.append(weight)
        return out

    def _test_weights_batch(self, args, view, node_ids, grid_arr, grid_norm):
        """`_test_weights` for many members: the geometry (angle to the centre weight, weights already used by a
        successor, the grid point that coincides with the centre) as whole-population array expressions with the bits
        of the scalar calls (utils.rowdot / rownorm), then per member only the reference's np.random.shuffle -- it
        consumes the global RNG, so it is called once per member in order -- and the pick of the first eligible grid
        weights in shuffled order. -> (tests [n, num_weight_candidates + 1, M], counts [n]); padding = 1."""
        num_weights = args.num_weight_candidates
        n, G, M = len(node_ids), len(grid_arr), grid_arr.shape[1]
        c = view.weights[node_ids]
        center = c / c.sum(axis=1, keepdims=True)
        step = args.delta_weight / 2.0
        # a grid point within 1e-3 (L2) of a vector is the lattice point nearest to it (spacing >> 2e-3): look that one up
        # and run the reference's exact test on it alone
        if getattr(self, '_lattice_of', None) is not grid_arr:
            self._lattice = {tuple(key): i for i, key in enumerate(np.rint(grid_arr / step).astype(np.int64).tolist())}
            self._lattice_of = grid_arr
        lattice = self._lattice
        assert step > 4e-3 and len(lattice) == G

        def coincident(vectors):
            """index of the grid point with norm2(grid - vector) < 1e-3, or -1"""
            keys = np.rint(vectors / step).astype(np.int64).tolist()
            idx = np.array([lattice.get(tuple(k), -1) for k in keys], dtype=np.int64)
            near = idx >= 0
            if near.any():
                exact = rownorm(grid_arr[idx[near]] - vectors[near]) < 1e-3
                idx[np.nonzero(near)[0][~exact]] = -1
            return idx

        excluded = np.zeros((n, G), dtype=bool)
        at_center = coincident(center)
        excluded[np.nonzero(at_center >= 0)[0], at_center[at_center >= 0]] = True
        has_center = np.ones(n, dtype=bool)
        member, succ = view.successors_of(node_ids)
        if len(member):
            w = view.weights[succ]
            succ_w = w / w.sum(axis=1, keepdims=True)
            used_center = rownorm(succ_w - center[member]) < 1e-3
            has_center[member[used_center]] = False
            at_succ = coincident(succ_w)
            excluded[member[at_succ >= 0], at_succ[at_succ >= 0]] = True
        cosine = rowdot(center[:, None, :], grid_arr[None, :, :]) / rownorm(center)[:, None] / grid_norm[None, :]
        angle = np.arccos(np.minimum(np.maximum(cosine, -1.0), 1.0))
        eligible = (angle < np.pi / 4.0) & ~excluded
        # one np.random.shuffle per member, in member order: the reference's RNG stream (np.random.permutation(G) is
        # shuffle(arange(G)): same dtype and length as the reference's list-built array)
        orders = np.empty((n, G), dtype=np.int64)
        for b in range(n):
            orders[b] = np.random.permutation(G)
        # ... then, for all members at once, the first eligible grid weights in shuffled order up to the cap
        elig = np.take_along_axis(eligible, orders, axis=1)
        nth = np.cumsum(elig, axis=1)                                # 1-based rank among the eligible ones
        k0 = has_center.astype(np.int64)
        take = elig & (nth <= (num_weights - k0)[:, None])
        tests = np.ones((n, num_weights + 1, M))
        tests[has_center, 0] = center[has_center]
        rows, cols = np.nonzero(take)                                # row-major: shuffled order within each member
        tests[rows, k0[rows] + nth[rows, cols] - 1] = grid_arr[orders[rows, cols]]
        counts = k0 + take.sum(axis=1)
        return tests, counts

    def prefetch_fits(self, args, opt_graph):
        """Optional: launch the model fits of the current population now (after `update`), so that they run under whatever
        the caller does before `prediction_guided_selection` (archive update, logging); results are identical."""
        self._pending = prefetch(opt_graph, self.sample_batch, args.obj_num, True)

    def prediction_guided_selection(self, args, iteration, ep, opt_graph, scalarization_template):
        """Returns (elite_batch, scalarization_batch, predicted_offspring_objs) (population_3d.py:239-333)."""
        N = args.num_tasks
        # the reference rebuilds this simplex grid for every sample (population_3d.py:262-263); it is a pure function of the
        # arguments (no RNG), so it is enumerated once
        key = (args.obj_num, args.delta_weight)
        if getattr(self, '_grid_key', None) != key:
            grid = []
            generate_weights_batch_dfs(0, args.obj_num, 0.0, 1.0, args.delta_weight / 2.0, [], grid)
            self._grid_arr = np.array(grid, dtype=np.float64)
            self._grid_norm = rownorm(self._grid_arr)
            self._grid_key = key
        grid_arr, grid_norm = self._grid_arr, self._grid_norm
        # the fits need only the opt-graph: launch all of them (K4) first and enumerate the test weights while they run
        prep = {}

        def while_fitting():            # host work that does not need the fits, done while they run
            prep['virtual_ep'] = np.array([np.asarray(s.objs, dtype=np.float64) for s in ep.sample_batch]).reshape(-1, args.obj_num)
            prep['scalarizations'] = [deepcopy(scalarization_template) for _ in range(N)]
        pending, self._pending = getattr(self, '_pending', None), None
        tests, counts, pred, self.last_fits = predict_candidates(
            opt_graph, self.sample_batch, lambda view, ids: self._test_weights_batch(args, view, ids, grid_arr, grid_norm),
            args.obj_num, cap_threshold=True, max_tests=args.num_weight_candidates + 1, tests_in_lockstep=True,
            pending=pending, while_fitting=while_fitting)
        candidates = Candidates(self.sample_batch, tests, counts, pred)
        virtual_ep = prep['virtual_ep']
        elite_batch, scalarization_batch, predicted_offspring_objs = [], [], []
        if len(candidates) == 0:
            print('Too few candidates')
            return elite_batch, scalarization_batch, predicted_offspring_objs
        cand_pred = np.ascontiguousarray(candidates.prediction, dtype=np.float64)
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

    def random_selection(self, args, scalarization_template):
        elite_batch, scalarization_batch = [], []
        for _ in range(args.num_tasks):
            elite_batch.append(self.sample_batch[np.random.choice(len(self.sample_batch))])
            weights = np.random.uniform(args.min_weight, args.max_weight, args.obj_num)
            scalarization = deepcopy(scalarization_template)
            scalarization.update_weights(weights / np.sum(weights))
            scalarization_batch.append(scalarization)
        return elite_batch, scalarization_batch
