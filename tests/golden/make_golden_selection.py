"""Generate tests/golden/selection_*.npz by running the UNMODIFIED reference selection code
(morl/population_2d.py, population_3d.py, ep.py, utils.py, hypervolume.py, opt_graph.py) on
synthetic optimisation histories (pgmorl_b200.synth_envs.run_selection_history).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_selection.py

Per generation the file stores the inputs of `prediction_guided_selection` (opt-graph, population,
EP) and what the reference did with them: candidate list, every scipy least_squares fit
(inputs, theta, status, nfev), per-round hv / sparsity arrays, chosen candidate ids, elites,
weights and predicted objectives; plus known-answer values of the helper functions.

One deliberate substitution, stated here: for 3 objectives the reference scores candidates with one
OS process per candidate (population_3d.py:216-237); the golden run calls the reference's own serial
scorer `evaluate_hypervolume_sparsity` (population_3d.py:206-214) instead -- same arithmetic.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import ref_import  # noqa: E402

ref_import.install()

import population_2d as ref2d  # noqa: E402
import population_3d as ref3d  # noqa: E402
import utils as ref_utils  # noqa: E402
from ep import EP as RefEP  # noqa: E402
from hypervolume import InnerHyperVolume  # noqa: E402
from opt_graph import OptGraph as RefOptGraph  # noqa: E402
from scalarization_methods import WeightedSumScalarization as RefScal  # noqa: E402
from scipy.optimize import least_squares as scipy_lsq  # noqa: E402

sys.path.insert(0, os.path.join(HERE, ".."))
import synth_envs  # noqa: E402


def record_history(M, generations, seed, **argkw):
    mod = ref2d if M == 2 else ref3d
    args = synth_envs.SelectionArgs(M, **argkw)
    fits, rounds, cands = [], [], []

    def lsq(fun, x0, **kw):
        res = scipy_lsq(fun, x0, **kw)
        x, y = kw["args"]
        w = fun(np.array([0.0, 1.0, 0.0, 1.0]), x, np.zeros_like(y))   # A=0, c=1 -> residual = w
        fits.append(dict(x=np.array(x), y=np.array(y), w=np.array(w), ub=np.array(kw["bounds"][1]),
                         theta=res.x.copy(), status=res.status, nfev=res.nfev, cost=res.cost))
        return res

    mod.least_squares = lsq

    class Pop(mod.Population):
        def evaluate_hv(self, candidates, mask, vep):
            hv = super().evaluate_hv(candidates, mask, vep)
            rounds.append(dict(hv=np.array(hv), mask=np.array(mask), vep=np.array(vep)))
            cands.clear(); cands.extend(candidates)
            return hv

        def evaluate_sparsity(self, candidates, mask, vep):
            sp = super().evaluate_sparsity(candidates, mask, vep)
            rounds[-1]["sparsity"] = np.array(sp)
            return sp

        def evaluate_hypervolume_sparsity_parallel(self, args, candidates, mask, vep):
            hv, sp = self.evaluate_hypervolume_sparsity(candidates, mask, vep)
            rounds.append(dict(hv=np.array(hv), sparsity=np.array(sp), mask=np.array(mask), vep=np.array(vep)))
            cands.clear(); cands.extend(candidates)
            return hv, sp

    blob = {"meta": np.array([M, generations, seed]),
            "args": np.array([args.num_tasks, args.num_weight_candidates, args.pbuffer_num, args.pbuffer_size],
                             dtype=np.int64),
            "args_f": np.array([args.sparsity, args.delta_weight])}

    def on_gen(g, st):
        gr, pop, ep = st["graph"], st["population"], st["ep"]
        pre = f"g{g}_"
        blob[pre + "graph_w"] = np.array([np.asarray(w, dtype=np.float64) for w in gr.weights])
        blob[pre + "graph_objs"] = np.array(gr.objs)
        blob[pre + "graph_prev"] = np.array(gr.prev)
        blob[pre + "pop_ids"] = np.array([s.optgraph_id for s in pop.sample_batch])
        blob[pre + "pop_objs"] = np.array([s.objs for s in pop.sample_batch])
        blob[pre + "ep_objs"] = np.array(ep.obj_batch)
        blob[pre + "elite_ids"] = np.array([s.optgraph_id for s in st["elites"]])
        blob[pre + "elite_w"] = np.array([sc.weights.numpy() for sc in st["scalarizations"]])
        blob[pre + "predicted"] = np.array(st["predicted"])
        blob[pre + "cand_pred"] = np.array([c["prediction"] for c in cands])
        blob[pre + "cand_weight"] = np.array([c["weight"] for c in cands])
        blob[pre + "cand_node"] = np.array([c["sample"].optgraph_id for c in cands])
        blob[pre + "n_fits"] = np.array(len(fits))
        for i, f in enumerate(fits):
            for k, v in f.items():
                blob[f"{pre}fit{i}_{k}"] = np.asarray(v)
        blob[pre + "n_rounds"] = np.array(len(rounds))
        for i, r in enumerate(rounds):
            for k, v in r.items():
                blob[f"{pre}round{i}_{k}"] = np.asarray(v)
        fits.clear(); rounds.clear(); cands.clear()

    classes = dict(EP=RefEP, Population=Pop, OptGraph=RefOptGraph, Scalarization=RefScal)
    synth_envs.run_selection_history(classes, args, generations, seed, on_generation=on_gen)
    return blob


def helper_kats():
    """Known-answer tests of the helper functions, computed by the reference."""
    rng = np.random.RandomState(0)
    out = {}
    pop2 = ref2d.Population(synth_envs.SelectionArgs(2))
    for i in range(40):
        n = int(rng.randint(1, 60))
        M = 2 if i % 2 == 0 else 3
        pts = rng.uniform(-0.5 if i % 5 == 0 else 0.0, 10.0, (n, M))
        if i % 3 == 0 and n > 4:                       # duplicates and ties on single coordinates
            pts[1] = pts[0]; pts[2, 0] = pts[3, 0]; pts[n - 1, -1] = pts[0, -1]
        if i % 7 == 0:
            pts = np.round(pts, 1)
        out[f"kat{i}_pts"] = pts
        out[f"kat{i}_ep_idx"] = np.array(ref_utils.get_ep_indices(pts), dtype=np.int64)
        if M == 2:
            out[f"kat{i}_hv"] = np.array(pop2.compute_hypervolume(pts))
            out[f"kat{i}_sp"] = np.array(pop2.compute_sparsity(pts))
        else:
            ep = pts[np.array(ref_utils.get_ep_indices(pts), dtype=int)] if n else pts
            if len(ep):
                out[f"kat{i}_hv"] = np.array(ref_utils.compute_hypervolume(ep.tolist()))
                out[f"kat{i}_hv_all"] = np.array(InnerHyperVolume(np.zeros(3)).compute(pts[(pts >= 0).all(1)].tolist()))
                out[f"kat{i}_sp"] = np.array(ref_utils.compute_sparsity(ep.tolist()))
                new = rng.uniform(0, 10, 3)
                out[f"kat{i}_new"] = new
                out[f"kat{i}_upd"] = np.array(ref_utils.update_ep([p for p in ep], new))
    out["n_kat"] = np.array(40)
    for name, (M, d) in {"w2": (2, 0.2), "w3a": (3, 0.25), "w3b": (3, 0.125), "w3c": (3, 1.0 / 19)}.items():
        wb = []
        ref_utils.generate_weights_batch_dfs(0, M, 0.0, 1.0, d, [], wb)
        out["weights_" + name] = np.array(wb)
    return out


if __name__ == "__main__":
    for name, M, gens, seed in [("selection_2d.npz", 2, 7, 3), ("selection_3d.npz", 3, 5, 5)]:
        blob = record_history(M, gens, seed)
        path = os.path.join(HERE, name)
        np.savez_compressed(path, **blob)
        print(name, os.path.getsize(path) // 1024, "KiB", "fits/gen", [int(blob[f"g{g}_n_fits"]) for g in range(gens)])
    if "--kats" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "selection_kats.npz"), **helper_kats())
        print("selection_kats.npz")
