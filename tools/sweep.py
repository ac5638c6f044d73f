#!/usr/bin/env python3
"""BASELINE configs[4]: batch-size sweep of ONE score-net step (CFG network evaluation on 2n images + fused
SDE update) for bf16/tcgen05 and fp32/FFMA.  Prints one JSON object."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
from toycrystals_b200.models import sde_score_model as shim  # noqa: E402

sde = shim.VPSDE(0.1, 30.0)
out = {"unit": "ms per reverse-SDE step (CFG evaluation of n samples + update)", "rows": []}
for precision, sizes in (("bf16", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]), ("fp32", [64, 128, 256, 512, 1024, 2048])):
    torch.manual_seed(1)
    m = shim.CondUNetTiny(4, 4, 96, 128, 8, 8, precision=precision).cuda().eval()
    for n in sizes:
        yc, yk = shim.condition_grid(m, n, 3.141592653589793 / 3, "cuda")
        steps = 4 if precision == "bf16" else 2
        ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            shim.sample_reverse_sde_euler_maruyama(m, sde, yc, yk, (n, 1, 64, 64), n_steps=steps, guidance_scale=1.5,
                                                   t_end=0.005, seed=1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / (steps + 1))
        ms = min(ts[1:])
        out["rows"].append({"precision": precision, "n": n, "ms_per_step": round(ms, 3),
                            "samples_per_s_at_300_steps": round(n / (ms * 1e-3 * 301), 2),
                            "conv_tflops": round(2 * n * 7.092e-3 / (ms * 1e-3), 1)})
        print(out["rows"][-1], file=sys.stderr, flush=True)
    del m
print(json.dumps(out))
