"""Pin which recorded model fits K4 does NOT reproduce to rtol 1e-6, next to scipy's own sensitivity on the same fits
(protocol (ii)). Needs a B200 (runs K4) and scipy:

    gpurun -- python tests/golden/make_k4_allowlist.py gpurun_out/k4_disagreeing_fits.json
    cp gpurun_out/k4_disagreeing_fits.json tests/golden/

For every recorded history (selection_2d / selection_3d / selection_2d_fork .npz, outputs of the unmodified reference):
  disagree   flat indices (generation-major, fit order of the file) of the fits whose K4 theta differs from the recorded
             scipy theta by more than rtol 1e-6 / atol 1e-9 -- the explicit allow-list tests/test_gpu_selection.py checks
  sensitive  indices of the fits where scipy ITSELF moves by more than that when x0 = ones(4) is perturbed by a relative
             1e-13 (the chaotic fits: the solver's path, not its input, decides which local solution is reached)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import selection_oracle as so  # noqa: E402


def scipy_fit(x, y, w, ub, x0):
    from scipy.optimize import least_squares
    return least_squares(lambda p, xx, yy: so.residual(p, xx, yy, w), x0, loss="soft_l1", f_scale=20.0, args=(x, y),
                         jac=lambda p, xx, yy: so.jacobian(p, xx, yy, w), bounds=(so.LB, ub)).x


def main(out_path):
    from pgmorl_b200 import kernels as K
    out = {}
    for name in ("selection_2d.npz", "selection_3d.npz", "selection_2d_fork.npz"):
        z = np.load(os.path.join(ROOT, "tests", "golden", name))
        xs, ys, ws, ubs, ref = [], [], [], [], []
        for g in range(int(z["meta"][1])):
            for i in range(int(z[f"g{g}_n_fits"])):
                pre = f"g{g}_fit{i}_"
                xs.append(z[pre + "x"]); ys.append(z[pre + "y"]); ws.append(z[pre + "w"]); ubs.append(z[pre + "ub"])
                ref.append(z[pre + "theta"])
        theta, status, nfev, cost = K.fit_hyperbolic(xs, ys, ws, ubs)
        ref = np.array(ref)
        close = np.isclose(theta, ref, rtol=1e-6, atol=1e-9).all(axis=1)
        sens = []
        for i, (x, y, w, ub) in enumerate(zip(xs, ys, ws, ubs)):
            a = scipy_fit(x, y, w, ub, np.ones(4))
            b = scipy_fit(x, y, w, ub, np.ones(4) * (1.0 + 1e-13))
            if not np.isclose(a, b, rtol=1e-6, atol=1e-9).all():
                sens.append(i)
        dis = np.nonzero(~close)[0].tolist()
        out[name] = {"n_fits": len(ref), "disagree": dis, "sensitive": sens,
                     "disagree_and_sensitive": sorted(set(dis) & set(sens))}
        print(name, len(ref), "fits; K4 disagrees on", len(dis), "; scipy self-sensitive on", len(sens), "; both", len(set(dis) & set(sens)))
    with open(out_path, "w") as fp:
        json.dump(out, fp, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
