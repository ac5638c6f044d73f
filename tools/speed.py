#!/usr/bin/env python3
"""Quick throughput probe: ms per CFG evaluation and conv TFLOP/s for a few (chunk, env) settings."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
from toycrystals_b200.models import sde_score_model as shim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [64]
sde = shim.VPSDE(0.1, 30.0)
for chunk in chunks:
    torch.manual_seed(1)
    m = shim.CondUNetTiny(4, 4, 96, 128, 8, 8, precision="bf16", chunk=chunk).cuda().eval()
    yc, yk = shim.condition_grid(m, n, 3.141592653589793 / 3, "cuda")
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        shim.sample_reverse_sde_euler_maruyama(m, sde, yc, yk, (n, 1, 64, 64), n_steps=steps, guidance_scale=1.5,
                                               t_end=0.005, seed=1)
        torch.cuda.synchronize(); dt = time.time() - t0
        if rep:
            best = min(best, dt)
    per = best / (steps + 1)
    print(f"MSUB={os.environ.get('TCS_MSUB','2')} chunk={chunk} n={n}: {per*1e3:.2f} ms/eval, "
          f"{2*n*7.092e-3/per:.0f} TFLOP/s conv, ~{n/(per*301):.1f} samples/s @300 steps", flush=True)
    del m
