"""tcgen05 building blocks (csrc/tc.cuh) checked on the device: exact integer GEMMs through every operand view the
K3 tensor-core kernel uses, and the precision of the 3-way TF32 split against FP64."""
import pytest
import torch

from pgmorl_b200 import _lib

pytestmark = pytest.mark.gpu


def test_tc_selftest():
    out = torch.zeros(16, dtype=torch.float32, device="cuda")
    _lib.check_diag(_lib.diag_lib().pgm_tc_selftest(_lib.ptr(out), 16, None))
    torch.cuda.synchronize()
    o = out.cpu().tolist()
    print("tc selftest:", o)
    assert o[0] == 0 and o[1] == 0 and o[2] == 0 and o[3] == 0 and o[4] == 0, o   # exact integer GEMMs
    assert o[12] == 0, o                                                            # TMEM store -> load round trip
    assert o[5] < 2e-3 and o[6] < 2e-6, o     # 1 x TF32 ~ 2^-11; 3 x TF32 split within ~10x of FP32 FMA (o[7])
