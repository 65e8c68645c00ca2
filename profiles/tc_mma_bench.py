"""How tcgen05.mma (kind::f16, shared-memory operands) paces small dependent MMAs on a B200.
    gpurun -- python profiles/tc_mma_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgmorl_b200 import _lib  # noqa: E402


def run(M, N, a_mn, b_mn, nmma, nacc, a_step=0, b_step=0):
    out = torch.zeros(8, dtype=torch.float32, device="cuda")
    _lib.check_diag(_lib.diag_lib().pgm_tc_mma_bench(_lib.ptr(out), M, N, a_mn, b_mn, nmma, nacc, a_step, b_step, None))
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    return o[4], o[5]


print("shape            majors  nmma nacc  total  issue  per-MMA")
for (M, N, am, bm) in [(128, 64, 0, 0), (128, 16, 0, 0), (128, 64, 0, 1), (64, 64, 1, 1), (64, 8, 1, 1), (128, 32, 1, 1), (128, 128, 0, 0), (128, 256, 0, 0)]:
    for nacc in (1, 3):
        if max(N, 128) * (nacc - 1) + N > 512:
            continue
        t, iss = run(M, N, am, bm, 48, nacc, a_step=32 if not am else 2048, b_step=32 if not bm else 2048)
        byt = (M + N) * 32
        print(f"{M:3d}x{N:3d}x16   {'MN' if am else 'K '} {'MN' if bm else 'K '}   48   {nacc}   {t:6.0f} {iss:6.0f}  {t / 48:6.1f} cycles/MMA  {byt * 48 / t:6.1f} B/clk")
