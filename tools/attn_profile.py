#!/usr/bin/env python3
"""clock64 phase profile of the fused attention block (attn_tc.cu): CTA 0, second image of its loop."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vae-diffusion-toy-crystals_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity_utils as pu  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
g = torch.Generator().manual_seed(5)
x = torch.randn((B, 16, 16, 192), generator=g)
args = (x, 1.0 + 0.2 * torch.randn(192, generator=g), 0.1 * torch.randn(192, generator=g),
        torch.randn((576, 192), generator=g) * (2.0 / math.sqrt(192)), 0.1 * torch.randn(576, generator=g),
        torch.randn((192, 192), generator=g) / math.sqrt(192), 0.1 * torch.randn(192, generator=g))
out, d = pu.debug_attn_block(*args, want_dbg=True)
CTRL = ["start", "phase0", "qkv0 done, W1 req"] + [f"h{h}:{n}" for h in range(4) for n in
                                                  ("keys here", "S issued", "V pushed, qkv+1 issued", "PV issued", "head done")] + \
       ["proj issued", "image done"]
WORK = ["start", "phase0", "head 0 converted"] + [f"h{h}:{n}" for h in range(4) for n in
                                                 ("S ready", "softmax", "q,k next", "O ready", "head done")] + \
       ["Y ready", "rows staged", "written"]
for role, names in ((0, CTRL), (1, WORK)):
    t = d["prof"][role].tolist()
    print("control lane" if role == 0 else "worker lane (warp 1)")
    for i, n in enumerate(names):
        if t[i] == 0:
            break
        print(f"  {n:24s} +{t[i] - t[i - 1] if i else 0:7d}   @{t[i] - t[0]:8d}")
