#!/usr/bin/env python
"""Benchmark of the PG-MORL hot path (BASELINE.json metric: MOPG env-steps/s; configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|...]

One "step" = one MOPG iteration of the population shard on one batch of synthetic
trajectories: rollout inference (K1) + vector GAE / advantage (K2) + PPO update of
E epochs x B minibatches with Adam (K3). Workload at N=1 = BASELINE.json configs[1]:
HalfCheetah shape, 6 tasks x 4 envs x 2048 steps, 10 epochs x 32 minibatches. With N > 1
(torchrun, one rank per GPU) every rank owns its own 6 tasks (weak scaling) and the ranks
all-gather the per-task objective/loss records once per generation (every 20th step).

Prints ONE JSON line (rank 0). `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same through the public host-buffer API (pinned H2D of every input + D2H of the
losses inside the timed region). `--impl reference` times the reference's CPU path (oracle
port with the reference's op sequence, process per task, one thread each) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (env shape, P per GPU, T, N, E, B, gamma)
    "c2": ("halfcheetah", 6, 2048, 4, 10, 32, 0.995),       # BASELINE.json configs[1]
    "walker64": ("walker2d", 64, 2048, 4, 10, 32, 0.995),    # population sweep points (configs[4])
    "walker256": ("walker2d", 256, 2048, 4, 10, 32, 0.995),
    "walker1024": ("walker2d", 1024, 2048, 4, 10, 32, 0.995),
    "hopper3": ("hopper3", 15, 2048, 4, 10, 32, 0.995),
    "humanoid8": ("humanoid", 8, 2048, 8, 10, 32, 0.99),     # BASELINE.json configs[3] shape: 64 tasks over 8 GPUs
    "humanoid16": ("humanoid", 16, 2048, 8, 10, 32, 0.99),   # ... over 4 GPUs
    "humanoid32": ("humanoid", 32, 2048, 8, 10, 32, 0.99),   # ... over 2 GPUs
}
GEN_ITERS = 20           # update_iter of the reference's launch scripts (scripts/walker2d-v2.py:40)


def flops_per_env_step(d, E):
    """Algorithmic FLOPs: K1 forward once; K3 fwd + weight-grad + input-grad x E."""
    O, A, M, H = d.obs, d.act, d.obj, d.hidden
    actor, critic = O * H + H * H + H * A, O * H + H * H + H * M
    fwd = 2 * (actor + critic)
    per_epoch = fwd + fwd + 2 * (H * H + H * A) + 2 * (H * H + H * M)
    return fwd, per_epoch * E


def bytes_per_env_step(d, E):
    """Algorithmic HBM bytes: K1 128 B, K2 36 B, K3 per epoch one 29-float record + 4 B index (Walker dims)."""
    O, A, M = d.obs, d.act, d.obj
    k1 = 4 * (O + A + A + 1 + M)
    k2 = 4 * (M + M + 2 + M + 1)
    k3 = E * (4 * (O + A + 1 + 2 * M + 1) + 4)
    return k1, k2, k3


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_inputs(d, P, T, N, E, seed):
    from pgmorl_b200 import synthetic
    traj = synthetic.make_trajectories(P, T, N, d, seed=seed)
    eps, perm = synthetic.host_rng_streams(0, T, N, d.act, E)
    w = synthetic.simplex_weights(d.obj, 0.2 if d.obj == 2 else 0.25)
    weights = np.stack([w[p % len(w)] for p in range(P)])
    obj_var = np.tile(np.array([1.3, 0.7, 0.9][:d.obj]), (P, 1))
    flats = [synthetic.init_policy_flat(d, seed=1000 + p).numpy() for p in range(P)]
    return traj, eps, perm, weights, obj_var, flats


def cpu_reference_leg(d, P, T, N, E, B, gamma, steps, warmup):
    """Reference CPU path on the host cores (oracle port, reference op sequence, process per task,
    one torch thread each -- morl/morl.py:34,84-88). Returns (env_steps_per_s, ms_per_step, cores, sample)."""
    from oracle.mopg_torch_port import timed_population_iteration
    cores = os.cpu_count() or 1
    procs = min(P, cores)
    traj, eps, perm, weights, obj_var, flats = synthetic_inputs(d, P, T, N, E, seed=1)
    trajs = [{k: v[p].numpy() for k, v in traj.items()} for p in range(P)]
    dims = (d.obs, d.act, d.obj)
    kw = dict(gamma=gamma, lam=0.95, ppo_epoch=E, num_mini_batch=B)
    for _ in range(warmup):
        timed_population_iteration(flats[:procs], dims, trajs[:procs], 0, 3e-4, weights, obj_var, procs, **kw)
    walls = []
    for i in range(steps):
        wall, per_task, _ = timed_population_iteration(flats, dims, trajs, i, 3e-4, weights, obj_var, procs, **kw)
        walls.append(wall)
    ms = 1e3 * float(np.mean(walls))
    return P * T * N / (ms / 1e3), ms, procs


def selection_leg(cpu=True):
    """Secondary BASELINE metric: prediction-guided selection ms/generation on the recorded synthetic histories
    (tests/golden/selection_{2d,3d}.npz, last generation): product path = host candidate generation + K4 fits + K5
    greedy loop; CPU = oracle port (scipy least_squares + python scoring; 3-D scoring timed on ONE of the 15 rounds
    and scaled, stated in `cpu_sample`)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.helpers import rebuild_selection_state
    from pgmorl_b200.scalarization_methods import WeightedSumScalarization
    out = {}
    torch.set_default_dtype(torch.float64)
    try:
        for name, M in (("selection_2d.npz", 2), ("selection_3d.npz", 3)):
            z = np.load(os.path.join(ROOT, "tests", "golden", name))
            g = int(z["meta"][1]) - 1
            times = []
            for rep in range(4):
                args_s, graph, pop, ep = rebuild_selection_state(z, g, M)
                np.random.seed(1000 + g)
                template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                elites, scals, pred = pop.prediction_guided_selection(args_s, g, ep, graph, template)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            key = f"{M}d"
            out[key] = {"ms_per_generation": 1e3 * min(times[1:]), "n_pop": len(pop.sample_batch),
                        "candidates": len(pop.last_candidates), "fits": len(pop.last_fits["x"]),
                        "archive": int(len(ep.obj_batch)), "tasks": args_s.num_tasks}
            if cpu:
                from oracle import selection_oracle as so
                f = pop.last_fits
                t0 = time.perf_counter()
                for x, y, w, ub in zip(f["x"], f["y"], f["w"], f["ub"]):
                    so.fit_scipy(x, y, w, ub)
                t_fit = time.perf_counter() - t0
                cand = np.array([c["prediction"] for c in pop.last_candidates])
                t0 = time.perf_counter()
                if M == 2:
                    so.greedy_select_2d(z[f"g{g}_round0_vep"], cand, args_s.sparsity, args_s.num_tasks)
                    t_sel = time.perf_counter() - t0
                else:
                    so.greedy_select_3d(z[f"g{g}_round0_vep"], cand, args_s.sparsity, 1)
                    t_sel = (time.perf_counter() - t0) * args_s.num_tasks
                out[key]["cpu_ms_per_generation"] = 1e3 * (t_fit + t_sel)
                out[key]["cpu_sample"] = ("scipy least_squares on every fit + python scoring, 1 core"
                                          + ("" if M == 2 else "; scoring timed on 1 of 15 greedy rounds and scaled"))
        # full performance buffers (n_pop 200 / 1 400 candidates / archive 300 and
        # n_pop 420 / 2 940 candidates / archive 500), built directly by synthetic.make_selection_state
        from pgmorl_b200.synthetic import make_selection_state
        for M, n_pop, n_ep in ((2, 200, 300), (3, 420, 500)):
            times = []
            for rep in range(3):
                args_s, graph, pop, ep = make_selection_state(M, n_pop, n_ep, seed=3)
                np.random.seed(7)
                template = WeightedSumScalarization(num_objs=M, weights=np.ones(M) / M)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                pop.prediction_guided_selection(args_s, 0, ep, graph, template)
                torch.cuda.synchronize()
                times.append(time.perf_counter() - t0)
            key = f"{M}d_full"
            out[key] = {"ms_per_generation": 1e3 * min(times[1:]), "n_pop": n_pop, "candidates": len(pop.last_candidates),
                        "fits": len(pop.last_fits["x"]), "archive": n_ep, "tasks": args_s.num_tasks}
            if cpu:
                from oracle import selection_oracle as so
                f = pop.last_fits
                sub = range(0, len(f["x"]), 8)                       # every 8th fit, scaled
                t0 = time.perf_counter()
                for i in sub:
                    so.fit_scipy(f["x"][i], f["y"][i], f["w"][i], f["ub"][i])
                t_fit = (time.perf_counter() - t0) * len(f["x"]) / len(sub)
                cand = np.array([c["prediction"] for c in pop.last_candidates])
                csub = cand[:: max(1, len(cand) // 40)]              # ~40 candidates of one greedy round, scaled
                t0 = time.perf_counter()
                (so.greedy_select_2d if M == 2 else so.greedy_select_3d)(ep.obj_batch, csub, args_s.sparsity, 1)
                t_sel = (time.perf_counter() - t0) * len(cand) / len(csub) * args_s.num_tasks
                out[key]["cpu_ms_per_generation"] = 1e3 * (t_fit + t_sel)
                out[key]["cpu_sample"] = ("scipy least_squares on every 8th fit and python scoring of ~40 candidates of one "
                                          "greedy round, both scaled to the full counts, 1 core")
    finally:
        torch.set_default_dtype(torch.float32)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--cluster", type=int, default=int(os.environ.get("PGM_PPO_CLUSTER", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selection", action="store_true")
    ap.add_argument("--full-sample", action="store_true",
                    help="reference arm: time the full-size iteration instead of the T/4 bounded sample")
    args = ap.parse_args()
    if args.impl == "reference":
        # CPU-only arm: hide the GPUs before torch is imported (fork-based worker pools cannot follow the CUDA
        # autograd threads torch starts when a device is visible)
        os.environ["CUDA_VISIBLE_DEVICES"] = ""

    from pgmorl_b200.layout import ENV_SHAPES
    env, P, T, N, E, B, gamma = CONFIGS[args.config]
    d = ENV_SHAPES[env]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever native libraries write to file descriptor 1 (NCCL prints its version
    # banner / debug output there) is sent to stderr, and the JSON line goes to a private duplicate of the real stdout
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    S = T * N
    metric = "MOPG env-steps/sec (rollout infer + GAE + PPO update)"
    workload = {"workload": f"{env}-shape population MOPG update: {P} tasks/GPU x {N} envs x {T} steps, "
                            f"{E} PPO epochs x {B} minibatches, obs {d.obs} act {d.act} obj {d.obj}, 64-64 tanh MLP",
                "tasks_per_gpu": P, "envs": N, "steps": T, "ppo_epochs": E, "minibatches": B,
                "l2": "flushed between timed iterations (256 MiB write)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # bounded sample: same per-row work (minibatch of 256 rows, 10 epochs) on T/4 steps
        Ts, Bs = (T, B) if args.full_sample else (max(T // 4, 1), max(B // 4, 1))
        v, ms, cores = cpu_reference_leg(d, P, Ts, N, E, Bs, gamma, args.steps, min(args.warmup, 1))
        sample = (f"each step = one MOPG iteration of {P} tasks on T={Ts} steps x {N} envs with {Bs} minibatches "
                  f"of {Ts * N // Bs} rows x {E} epochs (same minibatch size as the full workload), "
                  f"process per task, 1 torch thread each")
        line = {"impl": "reference", "metric": metric, "value": v, "unit": "env-steps/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload,
                "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
        return 0

    # ------------------------------------------------------------------ our arm (B200)
    import torch
    import torch.distributed as dist
    from pgmorl_b200.population_state import PopulationMOPG

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pop = PopulationMOPG(d, P, T, N, ppo_epoch=E, num_mini_batch=B, gamma=gamma, device=dev, cluster=args.cluster)
    traj, eps, perm, weights, obj_var, flats = synthetic_inputs(d, P, T, N, E, seed=1 + rank)
    for p in range(P):
        pop.load_task(p, flats[p], weights=weights[p], obj_var=obj_var[p])
    pop.set_lr(3e-4)
    host = dict(obs=traj["obs"], rewards=traj["rewards"], masks=traj["masks"], bad_masks=traj["bad_masks"],
                eps=eps.to(torch.float32), perm=perm.to(torch.int32))
    pop.upload(**host)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gather_buf = torch.empty(world * P, 3, device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def exchange(i):
        # once per generation the ranks all-gather the per-task records
        if world > 1 and (i + 1) % GEN_ITERS == 0:
            dist.all_gather_into_tensor(gather_buf, pop.losses)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def run(steps, e2e):
        """-> (device ms per step, list of per-stage ms or None)"""
        marks = []
        if e2e:
            pop.upload_staged_async()                  # inputs of the first iteration
        for i in range(steps):
            flush.fill_(i & 0xFF)                      # evict L2 between timed iterations (not timed)
            s, e = ev(), ev()
            if e2e:
                # every timed iteration: wait for ITS inputs' H2D copy, start the next iteration's copy on the copy stream
                # (it runs under this iteration's kernels), compute, read the losses back to the host
                s.record()
                pop.swap_inputs()
                pop.upload_staged_async()
                pop.step()
                pop._h_losses.copy_(pop.losses, non_blocking=True)
                exchange(i)
                e.record()
                e.synchronize()                        # the caller reads the losses on the host
            else:
                s.record()
                pop.step()
                exchange(i)
                e.record()
            marks.append((s, e))
        torch.cuda.synchronize()
        return sum(s.elapsed_time(e) for s, e in marks) / steps

    # stage breakdown (K1 / K2 / K3) with events around each launch, a few steps, outside the timed runs
    def stage_times(steps=10):
        from pgmorl_b200 import kernels as K
        acc = np.zeros(3)
        for i in range(steps):
            flush.fill_(i)
            e0, e1, e2, e3 = ev(), ev(), ev(), ev()
            e0.record()
            K.policy_forward(pop.params, pop.obs, d, eps=pop.eps, rows_a=S, out=(pop.value, pop.action, pop.logp))
            e1.record()
            K.gae_adv(pop.rewards, pop.value.view(P, T + 1, N, d.obj), pop.masks, pop.bad_masks, gamma, 0.95,
                      weights=pop.weights, obj_var=pop.obj_var, out=(pop.returns, pop.adv))
            e2.record()
            K.ppo_update(pop.params, pop.adam_m, pop.adam_v, pop.adam_step, pop.lr, pop.obs, pop.action, pop.logp,
                         pop.value, pop.returns.view(P, S, d.obj), pop.adv.view(P, S), pop.perm, B, d,
                         hyper=pop.hyper, workspace=pop.workspace, cluster=args.cluster, losses=pop.losses)
            e3.record()
            torch.cuda.synchronize()
            acc += (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3))
        return acc / steps

    run(args.warmup, False)
    run(min(args.warmup, 3), True)
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    t0 = time.perf_counter()
    ms = run(args.steps, False)
    barrier()
    ms_e2e = run(args.steps, True)
    barrier()
    t1 = time.perf_counter()
    clk = clocks.stop(t0, t1)
    stages = stage_times()
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    finite = bool(torch.isfinite(pop.params).all() and torch.isfinite(pop.losses).all())

    if rank == 0:
        env_steps = world * P * S
        fwd, upd = flops_per_env_step(d, E)
        k1b, k2b, k3b = bytes_per_env_step(d, E)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        ffma_peak = 2 * 128 * 148 * sm_max * 1e6 / 1e12          # FP32 FFMA TFLOP/s at the max SM clock
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        k3_ms = float(stages[2])
        k3_tflops = P * S * upd / (k3_ms * 1e-3) / 1e12
        # which K3 kernel ran: the tensor-core path (tcgen05 on FP16 operand pairs) is chosen explicitly (cluster 32) or by
        # the library from 8 tasks on for the shapes it is built for; otherwise the FP32 FFMA cluster kernels
        wide = (d.obs, d.act, d.obj) == (376, 17, 2)          # Humanoid: the streamed tensor-core kernel (csrc/k3_tcw.cuh)
        tc = (args.cluster in (32, 64) or (wide and args.cluster in (0, 128)) or
              (args.cluster == 0 and P >= 8 and (d.obs, d.act, d.obj) in ((17, 6, 2), (11, 3, 3))))
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` captures (profiles/): only the
        # two configurations that were captured, null otherwise
        traffic = {("c2", False): 7.5e6, ("walker64", True): 81.4e6}.get((args.config, tc))
        common = {"unit": "TFLOP/s", "achieved": k3_tflops, "traffic": traffic,
                  "traffic_source": ("profiles/README_r01.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
                                     if traffic else None),
                  "algorithmic_bytes_per_launch": P * S * k3b / E, "algorithmic_flops_per_launch": P * S * upd,
                  "launch_ms": k3_ms, "hbm_achieved_gbs": P * S * k3b / (k3_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                  "hbm_peak_source": "measured" if peaks else "fallback",
                  "fp32_ffma_peak_tflops": ffma_peak, "frac_of_fp32_ffma_peak": k3_tflops / ffma_peak}
        if tc:
            tpeak = peaks.get("bf16_tflops_sustained", 1400.0)    # FP16 and BF16 UMMA run at the same rate
            roofline = dict(common, kernel="k3_tcw_kernel" if wide else "k3_tc_kernel", bound="tensor", peak=tpeak, frac=k3_tflops / tpeak,
                            peak_source=("measured" if peaks else "fallback") + " dense 16-bit tensor peak (sustained)",
                            note="algorithmic FLOPs; the kernel executes 3 MMAs per product (FP16 operand pairs, FP32-level "
                                 "accuracy) on 128xNx16 tiles with N <= 64, which are shared-memory-operand bound (113 B/clk "
                                 "measured, profiles/tc_mma_bench.py), and its tanh epilogues are XU-pipe bound: see "
                                 "profiles/README_r01.md")
        else:
            roofline = dict(common, kernel="k3_ppo_fast_kernel", bound="fp32-ffma", peak=ffma_peak, frac=k3_tflops / ffma_peak,
                            peak_source=f"2*128 lanes*148 SMs*{sm_max:.0f} MHz (no measured FP32 peak in MEASURED_PEAKS.json)",
                            note="small populations are latency/occupancy bound: 6 chains of 320 dependent Adam steps")
        line = {
            "metric": metric, "value": env_steps / (ms * 1e-3), "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload,
            "clocks": clk,
            "e2e": {"value": env_steps / (ms_e2e * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": pop.h2d_bytes, "d2h_bytes_per_step": pop.d2h_bytes},
            "gpu_launches": pop.GPU_LAUNCHES_PER_STEP * args.steps * 2,
            "roofline": roofline,
            "stages_ms": {"k1_forward": float(stages[0]), "k2_gae_adv": float(stages[1]), "k3_pack_ppo": k3_ms},
            "stage_hbm_gbs": {"k1_forward": P * (S + N) * k1b / (stages[0] * 1e-3) / 1e9,
                              "k2_gae_adv": P * S * k2b / (stages[1] * 1e-3) / 1e9},
            "ppo_cluster": args.cluster, "k3_path": "tensor-core" if tc else "fp32-ffma", "finite": finite,
        }
        if world == 1 and not args.no_selection:
            try:
                line["selection"] = selection_leg(cpu=not args.no_cpu_baseline)
            except Exception as ex:
                line["selection"] = {"error": repr(ex)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            # fresh CPU-only process: fork-based worker pools cannot follow CUDA/autograd use in this one
            env_cpu = dict(os.environ, CUDA_VISIBLE_DEVICES="")
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                                    "--warmup", "0", "--config", args.config, "--full-sample"],
                                   capture_output=True, text=True, env=env_cpu, timeout=900)
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
                line["cpu_baseline"]["ms_per_step"] = ref["ms_per_step"]
            except Exception as ex:   # keep the GPU numbers even if the CPU leg fails
                line["cpu_baseline"] = {"error": repr(ex)[:200]}
        json_out.write(json.dumps(line) + "\n")
        json_out.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
