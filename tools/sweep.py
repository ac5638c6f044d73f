#!/usr/bin/env python3
"""BASELINE configs[4]: batch-size sweep 64..16384 of ONE score-net step (CFG network evaluation on 2n images + fused
SDE update): bf16 / tcgen05, fp32 / tcgen05 (bf16x3 split), fp32 / FFMA, beside the reference's own sampler as eager
PyTorch on the same B200 (cuDNN TF32 = PyTorch's default, and IEEE fp32) driven from oracle/_ref.  Prints one JSON object.

    python tools/sweep.py [--quick]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"))
sys.path.insert(0, ROOT)
from toycrystals_b200.models import sde_score_model as shim  # noqa: E402

quick = "--quick" in sys.argv
sde = shim.VPSDE(0.1, 30.0)
out = {"unit": "ms per reverse-SDE step (CFG evaluation of n samples + update)", "rows": []}
ALL = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
modes = (("bf16", "tcgen05", ALL), ("fp32", "tcgen05", ALL if not quick else ALL[:6]), ("fp32", "simt", ALL[:6] if not quick else ALL[:3]))
for precision, engine, sizes in modes:
    torch.manual_seed(1)
    m = shim.CondUNetTiny(4, 4, 96, 128, 8, 8, precision=precision, engine=engine).cuda().eval()
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
        out["rows"].append({"impl": "libtcs", "precision": precision, "engine": engine, "n": n, "ms_per_step": round(ms, 3),
                            "samples_per_s_at_300_steps": round(n / (ms * 1e-3 * 301), 2),
                            "conv_tflops": round(2 * n * 7.092e-3 / (ms * 1e-3), 1)})
        print(out["rows"][-1], file=sys.stderr, flush=True)
    del m
    torch.cuda.empty_cache()

# the reference's own module, eager PyTorch on the GPU (the sizes that fit: its forward materialises the 16 constant
# maps and both skip concats in fp32, ~30 MB per sample)
import bench  # noqa: E402
for tf32 in (True, False):
    for n in ([64, 256, 1024, 4096] if not quick else [64, 1024]):
        try:
            r = bench.cuda_eager_reference_samples_per_sec(n=n, evals=4 if n >= 4096 else 6, tf32=tf32)
            out["rows"].append({"impl": "reference (eager PyTorch, oracle/_ref)", "precision": "tf32 convs" if tf32 else "ieee fp32",
                                "n": n, "ms_per_step": round(n / r["value"] / 301 * 1e3, 3),
                                "samples_per_s_at_300_steps": round(r["value"], 3)})
        except Exception as e:  # noqa: BLE001
            out["rows"].append({"impl": "reference", "n": n, "tf32": tf32, "error": str(e)[:200]})
            torch.cuda.empty_cache()
        print(out["rows"][-1], file=sys.stderr, flush=True)
print(json.dumps(out))
