#!/usr/bin/env python3
"""Fixed workload for ncu: the fused VP-SDE update (step_kernel<SDE>, Philox noise in registers) at an HBM-bound size:
n = 32768 samples = 1.5 GiB of algorithmic traffic per launch (read x, read eps, write x)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
from toycrystals_b200 import _cabi  # noqa: E402
from toycrystals_b200.models import sde_score_model as shim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
torch.manual_seed(1)
m = shim.CondUNetTiny(4, 4, 96, 128, 8, 8, precision="bf16").cuda().eval()
h = m.engine_handle(shim.VPSDE(0.1, 30.0))
x = torch.randn((n, 1, 64, 64), device="cuda")
e = torch.randn_like(x)
L = _cabi.lib()
for r in range(6):
    _cabi.check(L.tcs_sde_update(h, x.data_ptr(), e.data_ptr(), None, n, 0.5, 0.499, 1234, 0, r, None))
torch.cuda.synchronize()
print("ok", float(x.mean()))
