import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from tests.test_gpu_kernels import _ppo_inputs, DIMS, dev, rel_err
from oracle import mopg_oracle as orc
from pgmorl_b200 import kernels as K
for cluster in (32, 64):
    for name, P, T, N, mb in [("walker", 2, 64, 4, 256), ("walker", 1, 30, 4, 100), ("hopper3", 2, 48, 2, 64)]:
        d = DIMS[name]
        cur, pk, perm = _ppo_inputs(d, P, T, N, seed=11)
        S = T * N
        idx = np.random.RandomState(0).permutation(S)[:mb]
        hyper = K.PpoHyper(entropy_coef=0.01)
        g, losses = K.ppo_grad(dev(cur), dev(pk["obs"]), dev(pk["action"]), dev(pk["logp"]), dev(pk["value"]),
                               dev(pk["returns"]), dev(pk["adv"]), dev(idx, torch.int32), d, hyper=hyper, cluster=cluster)
        torch.cuda.synchronize()
        for p in range(P):
            net = orc.Net(cur[p].copy(), d.obs, d.act, d.obj)
            gref, lref = orc.ppo_grad(net, pk["obs"][p][:S][idx].astype(np.float64), pk["action"][p][idx],
                                      pk["logp"][p][idx], pk["value"][p][:S][idx], pk["returns"][p][idx],
                                      pk["adv"][p][idx], ecoef=0.01)
            gg = g[p].cpu().numpy()
            from pgmorl_b200.layout import param_layout
            lay, _ = param_layout(d)
            per = {k: float(np.abs(gg[o:o+int(np.prod(sh))] - gref[o:o+int(np.prod(sh))]).max() / (np.abs(gref[o:o+int(np.prod(sh))]).max() + 1e-30)) for k, (o, sh) in lay.items()}
            print(cluster, name, mb, p, "grad rel", rel_err(gg, gref), "loss rel", rel_err(losses[p].cpu().numpy(), np.array(lref)))
            if cluster >= 32: print("   per tensor:", {k: f"{v:.1e}" for k, v in per.items()})
    for name, P, T, N, B in [("walker", 3, 64, 4, 4), ("hopper3", 2, 48, 2, 3), ("walker", 2, 160, 4, 2)]:
        d = DIMS[name]
        cur, pk, perm = _ppo_inputs(d, P, T, N, seed=13)
        S = T * N
        rng = np.random.RandomState(2)
        m0 = rng.randn(P, d.n_par) * 1e-3
        v0 = rng.rand(P, d.n_par) * 1e-5
        step0 = np.array([0, 7, 640][:P], dtype=np.int32)
        lr = np.array([3e-4, 2.5e-4, 1e-4][:P])
        gp, gm, gv = dev(cur), dev(m0), dev(v0)
        gstep = dev(step0, torch.int32)
        losses = K.ppo_update(gp, gm, gv, gstep, dev(lr, torch.float64), dev(pk["obs"]), dev(pk["action"]),
                              dev(pk["logp"]), dev(pk["value"]), dev(pk["returns"]), dev(pk["adv"]),
                              dev(perm[None], torch.int32), B, d, cluster=cluster)
        torch.cuda.synchronize()
        for p in range(P):
            flat = dev(cur[p]).cpu().numpy().astype(np.float64)
            m = dev(m0[p]).cpu().numpy().astype(np.float64)
            v = dev(v0[p]).cpu().numpy().astype(np.float64)
            obs3 = pk["obs"][p].reshape(T + 1, N, d.obs).astype(np.float64)
            step, lref = orc.ppo_update(flat, m, v, int(step0[p]), lr[p], (d.obs, d.act, d.obj), obs3,
                                        pk["action"][p].reshape(T, N, -1), pk["logp"][p].reshape(T, N),
                                        pk["value"][p].reshape(T + 1, N, -1), pk["returns"][p].reshape(T, N, -1),
                                        pk["adv"][p].reshape(T, N), perm, B)
            print(cluster, name, B, p, "step", int(gstep[p]), step, "param", rel_err(gp[p].cpu().numpy(), flat), "m", rel_err(gm[p].cpu().numpy(), m),
                  "v", rel_err(gv[p].cpu().numpy(), v), "loss", rel_err(losses[p].cpu().numpy(), np.array(lref)))
