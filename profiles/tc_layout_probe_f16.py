"""kind::f16 operand layouts (SWIZZLE_128B, 16-bit): one physical [row][64 halfwords] image read as K-major and as
MN-major.   gpurun -- python profiles/tc_layout_probe_f16.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgmorl_b200 import _lib  # noqa: E402


def run(M, N, a_mn, b_mn, fillA, fillB, RA, RB, lboA, sboA, lboB, sboB, ltA=0, ltB=0, offA=0, offB=0):
    out = torch.zeros(128 * N, dtype=torch.float32, device="cuda")
    _lib.check_diag(_lib.diag_lib().pgm_tc_layout_probe(_lib.ptr(out), M, N, a_mn, b_mn, fillA, fillB, RA, RB,
                                              lboA, sboA, lboB, sboB, 0, ltA, ltB, 0, 1, offA, offB, None))
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(128, N)


def hw(r, f, blk):
    """halfword index of (row r, feature f) in a [rows][64 halfwords] SW128 image; blk = bytes per 64-feature block"""
    byte = (f // 64) * blk + r * 128 + ((((f % 64) // 8) ^ (r % 8)) * 16) + (f % 8) * 2
    return (byte // 2) % 2048


np.set_printoptions(linewidth=220, suppress=True)
# identity B: K-major no-swizzle fp16, N=16, R=16: lbo = 16 rows*16B = 256, sbo = 128
IDB = dict(fillB=1, RB=16, lboB=256, sboB=128)
# (1) A MN-major SW128 (lt=2): M = features (128 = 2 blocks of 64, block stride 4096 B here), K = rows
D = run(128, 16, 1, 0, 0, 1, 128, 16, 4096, 1024, 256, 128, ltA=2)
exp = np.array([[hw(k, m, 4096) for k in range(16)] for m in range(128)], dtype=np.float32)
print("A MN-major SW128 f16 (lbo=4096 block, sbo=1024): match", np.array_equal(exp, D[:, :16]), int((exp != D[:, :16]).sum()))
print("  k=0 m=0..15", D[:16, 0], "\n  m=0 k=0..15", D[0, :16], "\n  m=64 k=0:", D[64, 0], " k=9 m=0..9:", D[:10, 9])
D = run(128, 16, 1, 0, 0, 1, 128, 16, 1024, 4096, 256, 128, ltA=2)
print("  swapped lbo/sbo: k=0 m=0,64:", D[0, 0], D[64, 0], " m=0 k=8:", D[0, 8])
# (2) A K-major SW128 (lt=2): M = rows, K = features; sbo = 1024 (8 rows x 128 B); k window at +32*j bytes
for off in (0, 32, 96):
    D = run(128, 16, 0, 0, 0, 1, 128, 16, 16, 1024, 256, 128, ltA=2, offA=off)
    exp = np.array([[hw(m, off // 2 + k, 0) for k in range(16)] for m in range(128)], dtype=np.float32)
    print(f"A K-major SW128 f16 off={off}: match", np.array_equal(exp, D[:, :16]), int((exp != D[:, :16]).sum()), D[:2, :4], D[9, :4])
# (3) B MN-major SW128, N = 64 / 32 (+64 B) / 8 (+48 B): A = K-major identity (no swizzle, R = 128: lbo 2048, sbo 128)
for (N, off) in ((64, 0), (32, 64), (8, 48), (8, 112)):
    D = run(128, N, 0, 1, 1, 0, 128, 64, 2048, 128, 4096, 1024, ltB=2, offB=off)
    exp = np.array([[hw(k, off // 2 + n, 4096) for n in range(N)] for k in range(16)], dtype=np.float32)
    print(f"B MN-major SW128 f16 N={N} off={off}: match", np.array_equal(exp, D[:16, :]), int((exp != D[:16, :]).sum()))
# (4) B K-major SW128 (rows = N, 64 features): N = 64
D = run(128, 64, 0, 0, 1, 0, 128, 64, 2048, 128, 16, 1024, ltB=2)
exp = np.array([[hw(n, k, 0) for n in range(64)] for k in range(16)], dtype=np.float32)
print("B K-major SW128 f16 N=64: match", np.array_equal(exp, D[:16, :]), int((exp != D[:16, :]).sum()))
# (5) M = 64, A MN-major SW128
D = run(64, 16, 1, 0, 0, 1, 64, 16, 4096, 1024, 256, 128, ltA=2)
lanes = [32 * (j // 16) + j % 16 for j in range(64)]
exp = np.array([[hw(k, m, 4096) for k in range(16)] for m in range(64)], dtype=np.float32)
print("M=64 A MN-major SW128 f16: match", np.array_equal(exp, D[lanes, :16]))
# (6) K-major A windows that start at 16-byte (not 32-byte) offsets, and one that runs past the 128-byte row
for off in (48, 80, 112):
    D = run(128, 16, 0, 0, 0, 1, 128, 16, 16, 1024, 256, 128, ltA=2, offA=off)
    exp = np.array([[hw(m, off // 2 + k, 0) if off // 2 + k < 64 else -1 for k in range(16)] for m in range(128)], dtype=np.float32)
    ok = (exp == D[:, :16]) | (exp < 0)
    print(f"A K-major SW128 f16 off={off}: in-row part matches:", bool(ok.all()), " row0:", D[0, :16])
# (7) B K-major SW128 N=16 whose second 8-row group lives elsewhere (SBO = 2048): rows 8..15 = words at +2048 B
D = run(128, 16, 0, 0, 1, 0, 128, 16, 2048, 128, 16, 2048, ltB=2)
print("B K-major N=16 sbo=2048: k=0 n=0..15:", D[0, :16], " expect n>=8 ->", [hw(n - 8, 0, 0) + 1024 for n in range(8, 16)])
# (8) B MN-major SW128 (N = 64 features, K = 16 rows) whose second 8-K-row group lives at +2048 B
D = run(128, 64, 0, 1, 1, 0, 128, 64, 2048, 128, 16384, 2048, ltB=2)
print("B MN-major sbo=2048: n=0 k=0..15:", D[:16, 0], " expect k>=8 ->", [(hw(k - 8, 0, 0) + 1024) % 2048 for k in range(8, 16)])
