"""world_size-2 gloo test (CPU) of the sharded selection front-end (pgmorl_b200/prediction.py::predict_candidates):
member i's fit inputs, fits and predictions are computed on rank i % W and one all-gather of the padded candidate table
makes every rank hold the same candidates as a single process. The three device entry points the path calls are replaced
by the oracle (scalar neighbourhood scan, scipy fit) -- the kernels themselves are GPU-tested; this checks the host logic
around them, on both the 2-objective (tests enumerated per rank) and the 3-objective (RNG in lockstep) variants."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _install_oracle_stubs(setattr_=setattr):
    """kernels.fit_inputs_launch / fit_hyperbolic_launch_packed / fit_hyperbolic_collect on the CPU, from the oracle.
    `setattr_`: pytest's monkeypatch.setattr in the test process (undone afterwards), plain setattr in the workers."""
    import contextlib
    import torch
    from oracle import selection_oracle as so
    from pgmorl_b200 import kernels as K, prediction as P

    class _Graph:           # the oracle's scan wants the list form; rebuild it from the flat arrays the product passes
        pass

    def fit_inputs_launch(objs, parent, edge_w, edge_dy, node_ids, cap_threshold, force_global_scratch=False):
        view = _Graph()
        view.objs, view.parent, view.edge_w, view.edge_dy = objs, np.asarray(parent), edge_w, edge_dy
        M = objs.shape[1]
        scans = [so.fit_inputs(view, int(k), M, cap_threshold, with_steps=True) for k in node_ids]
        kmax = max([len(e) for _, _, e in scans] + [1])
        source = np.zeros((len(scans), kmax), dtype=np.int32)
        for b, (_, _, e) in enumerate(scans):
            source[b, :len(e)] = view.parent[e]
        return dict(n=len(scans), M=M, Kmax=kmax, klen=np.array([len(e) for _, _, e in scans], dtype=np.int32),
                    steps=np.array([s for _, s, _ in scans], dtype=np.int32), source=source, scans=scans)

    def fit_hyperbolic_launch_packed(front, coef):
        theta = []
        for b, (out, _, e) in enumerate(front["scans"]):
            for x, y, w, ub in out:
                assert np.array_equal(w, coef[b, :len(e)])          # the batched Gaussian weights = the scalar ones
                theta.append(so.fit_scipy(x, y, w, ub).x)
        return dict(theta=np.array(theta).reshape(-1, 4))

    def fit_hyperbolic_collect(handle):
        t = handle["theta"]
        return t, np.ones(len(t), dtype=np.int64), np.ones(len(t), dtype=np.int64), np.zeros(len(t))

    setattr_(K, "fit_inputs_launch", fit_inputs_launch)
    setattr_(K, "fit_hyperbolic_launch_packed", fit_hyperbolic_launch_packed)
    setattr_(K, "fit_hyperbolic_collect", fit_hyperbolic_collect)
    setattr_(P, "fit_stream", lambda: None)
    setattr_(torch.cuda, "stream", lambda s: contextlib.nullcontext())


def _candidates(name, M):
    import torch
    from tests.helpers import rebuild_selection_state
    from pgmorl_b200.prediction import predict_candidates
    from pgmorl_b200.utils import generate_weights_batch_dfs, rownorm
    torch.set_default_dtype(torch.float64)
    z = np.load(os.path.join(GOLDEN, name))
    g = int(z["meta"][1]) - 1
    args, graph, pop, ep = rebuild_selection_state(z, g, M)
    np.random.seed(1000 + g)
    if M == 2:
        make = lambda view, ids: pop._test_weights_batch(view, ids, args.num_weight_candidates)
        return predict_candidates(graph, pop.sample_batch, make, M, cap_threshold=False, max_tests=args.num_weight_candidates)[:3]
    grid = []
    generate_weights_batch_dfs(0, M, 0.0, 1.0, args.delta_weight / 2.0, [], grid)
    grid_arr = np.array(grid, dtype=np.float64)
    make = lambda view, ids: pop._test_weights_batch(args, view, ids, grid_arr, rownorm(grid_arr))
    return predict_candidates(graph, pop.sample_batch, make, M, cap_threshold=True, max_tests=args.num_weight_candidates + 1,
                              tests_in_lockstep=True)[:3]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _install_oracle_stubs()
        for name, M in (("selection_2d.npz", 2), ("selection_3d.npz", 3)):
            tests, counts, pred = _candidates(name, M)
            np.savez(os.path.join(out_dir, f"cand{M}_{rank}.npz"), tests=tests, counts=counts, pred=pred)
    finally:
        dist.destroy_process_group()


def test_sharded_candidates_equal_single_process(tmp_path, monkeypatch):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    import torch
    _install_oracle_stubs(monkeypatch.setattr)
    monkeypatch.setattr(torch, "set_default_dtype", lambda d: None)   # (the workers switch to float64 like the reference; not needed here)
    for name, M in (("selection_2d.npz", 2), ("selection_3d.npz", 3)):
        tests, counts, pred = _candidates(name, M)                    # single process (no process group)
        valid = np.arange(tests.shape[1])[None, :] < counts[:, None]
        z = np.load(os.path.join(GOLDEN, name))
        g = int(z["meta"][1]) - 1
        assert np.array_equal(tests[valid], z[f"g{g}_cand_weight"])   # ... which are the reference's candidates
        for rank in range(2):
            r = np.load(os.path.join(str(tmp_path), f"cand{M}_{rank}.npz"))
            assert np.array_equal(r["counts"], counts)
            assert np.array_equal(r["tests"][valid], tests[valid]) and np.array_equal(r["pred"][valid], pred[valid])
