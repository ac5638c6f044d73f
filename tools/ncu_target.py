#!/usr/bin/env python3
"""Small, fixed workload for ncu: n samples, a few reverse-SDE steps, plain launches (no graph)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
from toycrystals_b200.models import sde_score_model as shim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
precision = sys.argv[4] if len(sys.argv) > 4 else "bf16"
engine = sys.argv[5] if len(sys.argv) > 5 else "auto"
torch.manual_seed(1)
m = shim.CondUNetTiny(4, 4, 96, 128, 8, 8, precision=precision, engine=engine, use_graph=False, chunk=chunk).cuda().eval()
sde = shim.VPSDE(0.1, 30.0)
yc, yk = shim.condition_grid(m, n, 3.141592653589793 / 3, "cuda")
for _ in range(2):
    x = shim.sample_reverse_sde_euler_maruyama(m, sde, yc, yk, (n, 1, 64, 64), n_steps=steps, guidance_scale=1.5,
                                               t_end=0.005, seed=1)
torch.cuda.synchronize()
print("ok", float(x.mean()), m.launch_count())
