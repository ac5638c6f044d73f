#!/usr/bin/env python3
"""Tiny end-to-end run of both sampling paths for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
from toycrystals_b200.models import diffusion_prior as pshim  # noqa: E402
from toycrystals_b200.models import sde_score_model as shim  # noqa: E402
from toycrystals_b200.models import vae as vshim  # noqa: E402

torch.manual_seed(1)
m = shim.CondUNetTiny(4, 4, 96, 128, 8, 8, precision="bf16", use_graph=False).cuda().eval()
sde = shim.VPSDE(0.1, 30.0)
n = 3
yc, yk = shim.condition_grid(m, n, 3.141592653589793 / 3, "cuda")
x = shim.sample_reverse_sde_euler_maruyama(m, sde, yc, yk, (n, 1, 64, 64), n_steps=1, guidance_scale=1.5, t_end=0.005, seed=1)
x2 = shim.sample_probability_flow_ode(m, sde, yc, yk, (n, 1, 64, 64), n_steps=1, guidance_scale=0.0, t_end=0.005, seed=1)
torch.manual_seed(0)
prior = pshim.DiffusionPriorFiLM(32, 4, 4, 64, 1024, 8, 64, precision="bf16", use_graph=False).cuda().eval()
torch.manual_seed(2)
vae = vshim.CondVAE(z_dim=32, n_types=4, y_cont_dim=4).cuda().eval()
sched = pshim.DiffusionSchedule.linear(1000, 1e-4, 0.05, torch.device("cuda"))
img = pshim.sample_images(vae, prior, sched, yc, yk, torch.zeros(32, device="cuda"), torch.ones(32, device="cuda"), 2, seed=1)
torch.cuda.synchronize()
print("ok", float(x.mean()), float(x2.mean()), float(img.mean()))
