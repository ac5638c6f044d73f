#!/usr/bin/env python3
"""Small, fixed workload for ncu on the latent-prior path: n samples, a few DDIM steps + decode, plain launches (no graph)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
from toycrystals_b200.models import diffusion_prior as pshim  # noqa: E402
from toycrystals_b200.models import vae as vshim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
prior = pshim.DiffusionPriorFiLM(32, 4, 4, 64, 1024, 8, 64, precision="bf16", use_graph=False).cuda().eval()
torch.manual_seed(2)
vae = vshim.CondVAE(z_dim=32, n_types=4, y_cont_dim=4).cuda().eval()
sched = pshim.DiffusionSchedule.linear(1000, 1e-4, 0.05, torch.device("cuda"))
y_cat = (torch.arange(n, device="cuda") % 4).to(torch.int64)
y_cont = torch.zeros((n, 4), device="cuda")
y_cont[:, 1] = torch.linspace(0.0, 3.141592653589793 / 3, n, device="cuda")
zm, zs = torch.zeros(32, device="cuda"), torch.ones(32, device="cuda")
for _ in range(2):
    x = pshim.sample_images(vae, prior, sched, y_cat, y_cont, zm, zs, steps, seed=1)
torch.cuda.synchronize()
print("ok", float(x.mean()), prior.launch_count(), vae.launch_count())
