"""One rank of the sharded PG-MORL run used by tests/test_gpu_dist_run.py (launched with torch.distributed.run, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/dist_run_worker.py SAVE_DIR [selection_method]

Runs pgmorl_b200.morl.run on the replay environments of the driver-loop golden (tests/golden/run_2d*), one GPU per rank,
and leaves next to the run's files a pickle per rank with the selected (elite node, weight) pairs of every generation and
the packed metadata every rank ended with."""
import os
import pickle
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from pgmorl_b200 import mopg, morl
    from pgmorl_b200.layout import NetDims
    import synth_envs as synthetic
    save_dir = sys.argv[1]
    method = sys.argv[2] if len(sys.argv) > 2 else "prediction-guided"
    rank, local = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d = NetDims(17, 6, 2)
    args = synthetic.run_args_2d(save_dir)
    args.selection_method = method
    mopg.set_env_hooks(
        make_vec_envs=lambda **kw: synthetic.SeededReplayVecEnv(d, args.num_steps, args.num_processes, [1.3, 0.7], base_seed=500),
        gym_make=lambda name: synthetic.ToyEvalEnv(d))
    ep, population, graph, timings = morl.run(args, device=dev)
    with open(os.path.join(save_dir, f"rank{rank}.pkl"), "wb") as fp:
        pickle.dump({"timings": timings, "ep_objs": np.asarray(ep.obj_batch),
                     "pop_objs": np.array([s.objs for s in population.sample_batch]),
                     "pop_nodes": [s.optgraph_id for s in population.sample_batch],
                     "graph_objs": np.array(graph.objs), "graph_prev": list(graph.prev),
                     "owned_ep": [i for i, s in enumerate(ep.sample_batch) if not s.is_stub]}, fp)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
