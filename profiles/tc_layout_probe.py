"""Discover how tcgen05.mma (kind::tf32) addresses shared-memory operands: run single MMAs whose operand words hold
their own word index and print the (m, k) -> word map the hardware used.
    gpurun -- python profiles/tc_layout_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgmorl_b200 import _lib  # noqa: E402


def run(M, N, a_mn, b_mn, fillA, fillB, RA, RB, lboA, sboA, lboB, sboB, lane_off=0, ltA=0, ltB=0, a_tmem=0):
    out = torch.zeros(128 * N, dtype=torch.float32, device="cuda")
    _lib.check_diag(_lib.diag_lib().pgm_tc_layout_probe(_lib.ptr(out), M, N, a_mn, b_mn, fillA, fillB, RA, RB,
                                              lboA, sboA, lboB, sboB, lane_off, ltA, ltB, a_tmem, None))
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(128, N)


def model(r, f, lbo, sbo):
    """hypothesis for SWIZZLE_128B_BASE32B MN-major: word index of (k-row r, mn f)"""
    byte = (f // 32) * lbo + (r // 4) * sbo + (r % 4) * 128 + ((((f % 32) // 8) ^ (r % 4)) * 32) + (f % 8) * 4
    return (byte // 4) % 2048


np.set_printoptions(linewidth=220, suppress=True)
# (1) A MN-major SW128_32B (layout type 1), index-filled; B = K-major identity (R=16, N=16): D[m][n<8] = A_logical[m][k=n]
for (lbo, sbo) in [(4096, 512), (512, 4096)]:
    D = run(128, 16, 1, 0, 0, 1, 128, 16, lbo, sbo, 256, 128, ltA=1)
    print(f"A MN-major lt=1 lbo={lbo} sbo={sbo}")
    print("  k=0, m=0..39:", D[:40, 0])
    print("  m=0, k=0..7 :", D[0, :8])
    print("  k=1, m=0..39:", D[:40, 1])
    print("  k=5, m=0..39:", D[:40, 5])
    print("  m=32,64,96 k=0:", D[32, 0], D[64, 0], D[96, 0])
    exp = np.array([[model(k, m, lbo, sbo) for k in range(8)] for m in range(128)], dtype=np.float32)
    print("  matches hypothesis:", np.array_equal(exp, D[:, :8]), " mismatches:", int((exp != D[:, :8]).sum()))
# (2) B MN-major lt=1 index-filled (N=64); A = K-major identity (R=128): D[m<8][n] = B_logical[n][k=m]
for (lbo, sbo) in [(4096, 512)]:
    D = run(128, 64, 0, 1, 1, 0, 128, 64, 2048, 128, lbo, sbo, ltB=1)
    exp = np.array([[model(k, n, lbo, sbo) for n in range(64)] for k in range(8)], dtype=np.float32)
    print(f"B MN-major lt=1 lbo={lbo} sbo={sbo}: matches hypothesis:", np.array_equal(exp, D[:8, :]), int((exp != D[:8, :]).sum()))
    print("  k=0, n=0..39:", D[0, :40])
# (3) M=64 A MN-major lt=1
D = run(64, 16, 1, 0, 0, 1, 64, 16, 4096, 512, 256, 128, ltA=1)
lanes = [32 * (j // 16) + j % 16 for j in range(64)]
exp = np.array([[model(k, m, 4096, 512) for k in range(8)] for m in range(64)], dtype=np.float32)
print("M=64 A MN-major lt=1: matches:", np.array_equal(exp, D[lanes, :8]))
# (4) A from TMEM (lane m, col k = m*8+k), B = K-major identity N=16
D = run(128, 16, 0, 0, 0, 1, 128, 16, 0, 0, 256, 128, a_tmem=1)
exp = np.array([[m * 8 + k for k in range(8)] for m in range(128)], dtype=np.float32)
print("A from TMEM: matches:", np.array_equal(exp, D[:, :8]), D[:3, :8])
# (5) SBO = 0 aliasing for a K-major B with N=16 (rows 8..15 alias rows 0..7): A = K-major identity; B index-filled
D = run(128, 16, 0, 0, 1, 0, 128, 16, 2048, 128, 128, 0)
print("B K-major N=16, lbo=128 sbo=0: k=0 n=0..15:", D[0, :16], " k=4:", D[4, :16])
# (6) no-swizzle MN-major really unsupported? (expect zeros)
D = run(128, 16, 1, 0, 0, 1, 128, 16, 128, 256, 256, 128)
print("A MN-major lt=0: all zero:", bool((D[:, :8] == 0).all()))
for lt in (2, 4, 6):
    D = run(128, 16, 1, 0, 0, 1, 128, 16, 4096, 1024, 256, 128, ltA=lt)
    print(f"A MN-major lt={lt}: k=0 m=0..11:", D[:12, 0], "k=1:", D[:4, 1], "m=32:", D[32, 0])
