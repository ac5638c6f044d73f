#!/usr/bin/env python3
"""Diagnostic run for the GPU box: per-conv and per-layer error tables for every engine.
Prints, never asserts — meant to make one gpurun call as informative as possible."""
import math
import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("vae-diffusion-toy-crystals_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import parity_utils as pu  # noqa: E402
import toycrystals_oracle as orc  # noqa: E402
from test_gpu_parity import CONV_CASES, MODES  # noqa: E402


SEL = [int(a) for a in sys.argv[2:]] if len(sys.argv) > 2 else [0, 1, 2]
MODES = [m for i, m in enumerate(MODES) if i in SEL]


def conv_table():
    for precision, engine in MODES:
        for case in CONV_CASES:
            c0, c1, cout, k, s, res = case
            g = torch.Generator().manual_seed(1)
            B = 3
            in0 = torch.randn((B, res * s, res * s, c0), generator=g)
            in1 = torch.randn((B, res * s, res * s, c1), generator=g) if c1 else None
            w = torch.randn((cout, c0 + c1, k, k), generator=g) / math.sqrt((c0 + c1) * k * k)
            b = torch.randn((cout,), generator=g)
            ref = pu.conv_reference(in0, in1, w, b, k, s, round_bf16=(precision == "bf16"))
            row = []
            for epi in (0, 1, 2):
                try:
                    out, _ = pu.debug_conv(engine, precision, in0.cuda(), None if in1 is None else in1.cuda(), w.cuda(),
                                           b.cuda(), k, s, epi)
                    row.append("%.2e" % pu.rel_l2(out, ref))
                except Exception as e:  # noqa: BLE001
                    row.append("ERR:" + str(e)[:70])
            print(f"conv {precision:4s}/{engine:7s} {str(case):32s} raw/padded/plain rel-L2: {row}", flush=True)


def layer_table():
    g = pu.golden("score_fwd.pt")
    sd = orc.default_init_state_dict(0)
    x, t = g["x"] * 37.0, torch.full((3,), 0.37)
    taps = {}
    with torch.no_grad():
        orc.score_net(sd, pu.CFG, x.double(), t.double(), g["y_cat"], g["y_cont"].double(), taps)
    for precision, engine in MODES:
        try:
            m = pu.model(precision, engine)
            errs = []
            for name, C, res in pu.LAYERS:
                try:
                    got = pu.debug_layer(m, name, x.cuda(), t.cuda(), g["y_cat"].cuda(), g["y_cont"].cuda(), C, res)
                except NotImplementedError:
                    continue
                errs.append((name, pu.rel_l2(got, taps[name].permute(0, 2, 3, 1))))
            print(f"layers {precision}/{engine}: " + ", ".join(f"{n}={e:.1e}" for n, e in errs), flush=True)
            for case in g["cases"]:
                xx = (g["x"] * case["scale"]).cuda()
                tt = torch.full((3,), case["t"]).cuda()
                from toycrystals_b200.models.sde_score_model import predict_eps_cfg
                e = predict_eps_cfg(m, xx, tt, g["y_cat"].cuda(), g["y_cont"].cuda(), 1.5)
                print(f"  eps_cfg15 t={case['t']}: rel-L2 {pu.rel_l2(e, case['eps_cfg15']):.3e}", flush=True)
        except Exception:  # noqa: BLE001
            traceback.print_exc()


def quick_speed():
    from toycrystals_b200.models import sde_score_model as shim
    sde = shim.VPSDE(0.1, 30.0)
    for precision, engine, n in (("bf16", "tcgen05", 256), ("bf16", "simt", 32), ("fp32", "simt", 32)):
        try:
            m = pu.model(precision, engine, seed=1)
            y_cat, y_cont = orc.condition_grid(n, 4, 4)
            yc, yk = y_cat.cuda(), y_cont.cuda()
            for steps in (2, 6):
                torch.cuda.synchronize(); t0 = time.time()
                shim.sample_reverse_sde_euler_maruyama(m, sde, yc, yk, (n, 1, 64, 64), n_steps=steps, guidance_scale=1.5,
                                                       t_end=0.005, seed=1)
                torch.cuda.synchronize(); dt = time.time() - t0
                print(f"speed {precision}/{engine} n={n} steps={steps}: {dt*1e3:.1f} ms "
                      f"-> {dt/(steps+1)*1e3:.2f} ms per CFG evaluation of {n} samples "
                      f"({2*n*7.092e-3*(steps+1)/dt:.1f} TFLOP/s conv)", flush=True)
        except Exception:  # noqa: BLE001
            traceback.print_exc()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    which = sys.argv[1:2] or ["conv", "layers", "speed"]
    if "conv" in which:
        conv_table()
    if "layers" in which:
        layer_table()
    if "speed" in which:
        quick_speed()
