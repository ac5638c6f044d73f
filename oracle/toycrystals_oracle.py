"""CPU oracle for the ToyCrystals VP-SDE sampling hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it.  The product path (``toycrystals_b200``) never does and
fails loudly when its CUDA library is missing.

It is a from-scratch, state-dict-driven, dtype-generic restatement (plain
``torch.nn.functional`` calls, no nn.Module graph) of the algorithm in the
reference file ``src/toycrystals/models/sde_score_model.py``:

  * ``time_features``      <- timestep_embedding            (:17-32)
  * ``condition_vector``   <- ConditionEmbedding.forward    (:69-82)  incl. the
                              theta-aliasing quirk  y[:,2] = cos(sin(theta))
  * ``score_net``          <- CondUNetTiny.forward          (:227-266)
  * ``Schedule``           <- VPSDE                         (:273-298)
  * ``eps_cfg``            <- predict_eps_cfg               (:402-423)
  * ``sample``             <- sample_probability_flow_ode   (:452-504) and
                              sample_reverse_sde_euler_maruyama (:507-569)
  * ``condition_grid``     <- save_sde_samples              (:317-321)
  * ``default_init_state_dict`` reproduces ``CondUNetTiny(...)`` default init
    under ``torch.manual_seed(seed)`` by creating torch layers in the
    constructor's order (:180-225).

Parity pin: the reference ships no tests / golden vectors ("parity unpinned" by
the reference itself).  The oracle is therefore pinned against the reference
*module itself*, imported read-only in the build container by
``oracle/gen_golden.py``; in fp32 on CPU the two agree BIT-EXACTLY (asserted by
that script) and the resulting vectors are committed under ``tests/golden/``.

Because every op is dtype-generic, the same code run in float64 gives a
higher-precision "truth" that the reference cannot provide (its
timestep_embedding forces fp32); tests use it to rank fp32 implementations.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]

DEFAULT_CFG = dict(n_types=4, y_cont_dim=4, base_ch=96, emb_dim=128, cond_ch=8, time_ch=8)
N_HEADS = 4
GN_EPS = 1e-5


# --------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------
def _layer_plan(cfg: dict) -> List[tuple]:
    """(state-dict prefix, kind, ctor args) in the reference constructor's order."""
    e, b = cfg["emb_dim"], cfg["base_ch"]
    cin = 1 + cfg["cond_ch"] + cfg["time_ch"]
    plan: List[tuple] = [
        ("cond_emb.cat_emb", "emb", (cfg["n_types"] + 1, e)),
        ("cond_emb.cont_mlp.0", "lin", (cfg["y_cont_dim"], e)),
        ("cond_emb.cont_mlp.2", "lin", (e, e)),
        ("cond_emb.out.1", "lin", (2 * e, e)),
        ("time_mlp.0", "lin", (e, e)),
        ("time_mlp.2", "lin", (e, e)),
        ("to_cond_map", "lin", (e, cfg["cond_ch"])),
        ("to_time_map", "lin", (e, cfg["time_ch"])),
    ]

    def block(name, i, o):
        return [
            (f"{name}.net.0", "conv", (i, o, 3)),
            (f"{name}.net.1", "gn", (o,)),
            (f"{name}.net.3", "conv", (o, o, 3)),
            (f"{name}.net.4", "gn", (o,)),
        ]

    plan += block("down1", cin, b)
    plan += [("ds1", "conv", (b, b, 4))]
    plan += block("down2", b, 2 * b)
    plan += [("ds2", "conv", (2 * b, 2 * b, 4))]
    plan += block("mid", 2 * b, 2 * b)
    plan += [("attn.norm", "gn", (2 * b,)), ("attn.qkv", "conv", (2 * b, 6 * b, 1)),
             ("attn.proj", "conv", (2 * b, 2 * b, 1))]
    plan += [("us2_conv", "conv", (2 * b, 2 * b, 3))]
    plan += block("up2", 4 * b, b)
    plan += [("us1_conv", "conv", (b, b, 3))]
    plan += block("up1", 2 * b, b)
    plan += [("out", "conv", (b, 1, 3))]
    return plan


def default_init_state_dict(seed: int, cfg: Optional[dict] = None) -> StateDict:
    """State dict equal to ``torch.manual_seed(seed); CondUNetTiny(**cfg).state_dict()``."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    torch.manual_seed(seed)
    sd: StateDict = {}
    for name, kind, a in _layer_plan(cfg):
        if kind == "emb":
            m = torch.nn.Embedding(*a)
        elif kind == "lin":
            m = torch.nn.Linear(*a)
        elif kind == "conv":
            m = torch.nn.Conv2d(a[0], a[1], kernel_size=a[2])
        else:
            m = torch.nn.GroupNorm(8, a[0])
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def make_checkpoint(cfg: Optional[dict] = None, seed_model: int = 0, seed_ema: int = 1,
                    beta_min: float = 0.1, beta_max: float = 30.0) -> dict:
    """Payload in the layout the training script writes (train_sde_score_model.py:45-54,179-192)."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    config = dict(img_ch=1, **cfg, beta_min=beta_min, beta_max=beta_max, t_power=1.0, p_uncond=0.1)
    return {
        "epoch_next": 1,
        "model": default_init_state_dict(seed_model, cfg),
        "opt": {},
        "loss_hist": [],
        "config": config,
        "ema": default_init_state_dict(seed_ema, cfg),
    }


# --------------------------------------------------------------------------
# network
# --------------------------------------------------------------------------
def _w(sd: StateDict, key: str, like: torch.Tensor) -> torch.Tensor:
    return sd[key].to(device=like.device, dtype=like.dtype)


def _linear(sd, name, x):
    return F.linear(x, _w(sd, name + ".weight", x), _w(sd, name + ".bias", x))


def _conv_circ(sd, name, x, stride=1, pad=1):
    if pad:
        x = F.pad(x, (pad, pad, pad, pad), mode="circular")
    return F.conv2d(x, _w(sd, name + ".weight", x), _w(sd, name + ".bias", x), stride=stride)


def _gn(sd, name, x):
    return F.group_norm(x, 8, _w(sd, name + ".weight", x), _w(sd, name + ".bias", x), eps=GN_EPS)


def _conv_block(sd, name, x, taps: Optional[dict]):
    for conv, norm in ((".net.0", ".net.1"), (".net.3", ".net.4")):
        x = _conv_circ(sd, name + conv, x)
        if taps is not None:
            taps[name + conv + ".raw"] = x
        x = F.silu(_gn(sd, name + norm, x))
        if taps is not None:
            taps[name + conv + ".act"] = x
    return x


def time_features(t: torch.Tensor, dim: int, dtype=torch.float32) -> torch.Tensor:
    half = dim // 2
    idx = torch.arange(half, device=t.device, dtype=dtype)
    freqs = torch.exp(-math.log(10_000.0) * idx / max(half - 1, 1))
    ang = (2.0 * math.pi) * t.to(dtype).unsqueeze(1) * freqs.unsqueeze(0)
    out = torch.cat([torch.cos(ang), torch.sin(ang)], dim=1)
    if dim % 2 == 1:
        out = F.pad(out, (0, 1))
    return out


def condition_vector(sd: StateDict, cfg: dict, y_cat: torch.Tensor, y_cont: torch.Tensor,
                     dtype=torch.float32) -> torch.Tensor:
    idx = y_cat.clamp(min=0, max=cfg["n_types"]).to(torch.long)
    y = y_cont.to(dtype).clone()
    s = torch.sin(y[:, 1])
    y[:, 1] = s
    y[:, 2] = torch.cos(s)  # the reference reads the already-overwritten column (:76-78)
    e_cat = F.embedding(idx, _w(sd, "cond_emb.cat_emb.weight", y))
    e_cont = _linear(sd, "cond_emb.cont_mlp.2", F.silu(_linear(sd, "cond_emb.cont_mlp.0", y)))
    return _linear(sd, "cond_emb.out.1", F.silu(torch.cat([e_cat, e_cont], dim=1)))


def self_attention(sd: StateDict, x: torch.Tensor) -> torch.Tensor:
    B, C, H, W = x.shape
    d = C // N_HEADS
    qkv = _conv_circ(sd, "attn.qkv", _gn(sd, "attn.norm", x), pad=0)
    q, k, v = (z.reshape(B, N_HEADS, d, H * W).transpose(2, 3) for z in torch.chunk(qkv, 3, dim=1))
    y = F.scaled_dot_product_attention(q, k, v)
    y = y.transpose(2, 3).contiguous().reshape(B, C, H, W)
    return x + _conv_circ(sd, "attn.proj", y, pad=0)


def _up2x(x):
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def score_net(sd: StateDict, cfg: dict, x_t: torch.Tensor, t: torch.Tensor, y_cat: torch.Tensor,
              y_cont: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """eps_hat = CondUNetTiny(x_t, t, y_cat, y_cont); optional per-layer ``taps`` for debugging."""
    dt = x_t.dtype
    B, _, H, W = x_t.shape
    e = cfg["emb_dim"]
    t_emb = _linear(sd, "time_mlp.2", F.silu(_linear(sd, "time_mlp.0", time_features(t, e, dt))))
    c_emb = condition_vector(sd, cfg, y_cat, y_cont, dt)
    t_map = _linear(sd, "to_time_map", t_emb)[:, :, None, None].expand(-1, -1, H, W)
    c_map = _linear(sd, "to_cond_map", c_emb)[:, :, None, None].expand(-1, -1, H, W)
    x = torch.cat([x_t, t_map, c_map], dim=1)

    h1 = _conv_block(sd, "down1", x, taps)
    h = _conv_circ(sd, "ds1", h1, stride=2)
    if taps is not None:
        taps["ds1"] = h
    h2 = _conv_block(sd, "down2", h, taps)
    h = _conv_circ(sd, "ds2", h2, stride=2)
    if taps is not None:
        taps["ds2"] = h
    h = _conv_block(sd, "mid", h, taps)
    h = self_attention(sd, h)
    if taps is not None:
        taps["attn"] = h
    h = _conv_circ(sd, "us2_conv", _up2x(h))
    if taps is not None:
        taps["us2_conv"] = h
    h = _conv_block(sd, "up2", torch.cat([h, h2], dim=1), taps)
    h = _conv_circ(sd, "us1_conv", _up2x(h))
    if taps is not None:
        taps["us1_conv"] = h
    h = _conv_block(sd, "up1", torch.cat([h, h1], dim=1), taps)
    return _conv_circ(sd, "out", h)


# --------------------------------------------------------------------------
# schedule + samplers
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Schedule:
    beta_min: float = 0.1
    beta_max: float = 20.0

    def beta(self, t):
        return self.beta_min + t * (self.beta_max - self.beta_min)

    def int_beta(self, t):
        return self.beta_min * t + 0.5 * (self.beta_max - self.beta_min) * (t ** 2)

    def alpha(self, t):
        return torch.exp(-0.5 * self.int_beta(t))

    def sigma(self, t):
        a = self.alpha(t)
        return torch.sqrt(torch.clamp(1.0 - a * a, min=1e-8))


def time_grid(n_steps: int, t_end: float, device=None, dtype=torch.float32) -> torch.Tensor:
    u = torch.linspace(0.0, 1.0, n_steps + 1, device=device, dtype=dtype)
    return t_end + (1.0 - t_end) * (1.0 - u) ** 2


def condition_grid(n: int, n_types: int, y_cont_dim: int, theta_max: float = math.pi / 3.0,
                   device=None):
    y_cat = torch.tensor([i % n_types for i in range(n)], device=device, dtype=torch.int64)
    y_cont = torch.zeros((n, y_cont_dim), device=device)
    y_cont[:, 1] = torch.linspace(0.0, theta_max, steps=n, device=device)
    return y_cat, y_cont


def eps_cfg(sd, cfg, x_t, t, y_cat, y_cont, guidance: float) -> torch.Tensor:
    if guidance <= 0.0:
        return score_net(sd, cfg, x_t, t, y_cat, y_cont)
    e_u = score_net(sd, cfg, x_t, t, torch.full_like(y_cat, cfg["n_types"]), torch.zeros_like(y_cont))
    e_c = score_net(sd, cfg, x_t, t, y_cat, y_cont)
    return e_u + guidance * (e_c - e_u)


@dataclass
class SampleTrace:
    image: torch.Tensor            # [n,1,H,W] in [0,1]
    x0_hat: torch.Tensor           # pre-clamp projection
    eps: List[torch.Tensor]        # one entry per network evaluation (CFG-combined)
    x_in: List[torch.Tensor]       # the x_t each evaluation saw (for teacher forcing)
    t_in: List[float]
    ts: torch.Tensor


@torch.no_grad()
def sample(sd: StateDict, cfg: dict, sched: Schedule, y_cat, y_cont, x_init: torch.Tensor,
           sampler: str, n_steps: int, guidance: float, t_end: float,
           noise: Optional[Sequence[torch.Tensor]] = None, keep_trace: bool = True) -> SampleTrace:
    """Both reference samplers with the random draws injected.

    ``x_init`` replaces the reference's ``torch.randn(img_shape)`` and ``noise[i]`` its i-th
    ``torch.randn_like(x)`` ("sde" only; one per step, the last step included)."""
    if not (0.0 < float(t_end) < 1.0):
        raise ValueError(f"t_end must be in (0,1), got {t_end}")
    if sampler not in ("ode", "sde"):
        raise ValueError(f"Unknown sampler='{sampler}'. Use 'ode' or 'sde'.")
    x = x_init.clone()
    B = x.shape[0]
    ts = time_grid(n_steps, float(t_end), device=x.device, dtype=x.dtype)
    tr = SampleTrace(image=x, x0_hat=x, eps=[], x_in=[], t_in=[], ts=ts)

    def evaluate(xx, t):
        e = eps_cfg(sd, cfg, xx, t, y_cat, y_cont, guidance)
        if keep_trace:
            tr.eps.append(e.clone()); tr.x_in.append(xx.clone()); tr.t_in.append(float(t[0]))
        return e

    def pf_drift(xx, t):
        beta = sched.beta(t).view(B, 1, 1, 1)
        sig = sched.sigma(t).view(B, 1, 1, 1)
        score = -evaluate(xx, t) / sig
        return -0.5 * beta * xx - 0.5 * beta * score

    for i in range(n_steps):
        t, t_next = ts[i].expand(B), ts[i + 1].expand(B)
        dt = (t_next - t).view(B, 1, 1, 1)
        if sampler == "ode":
            d0 = pf_drift(x, t)
            x_pred = x + d0 * dt
            d1 = pf_drift(x_pred, t_next)
            x = x + 0.5 * (d0 + d1) * dt
        else:
            beta = sched.beta(t).view(B, 1, 1, 1)
            sig = sched.sigma(t).view(B, 1, 1, 1)
            g = torch.sqrt(beta)
            score = -evaluate(x, t) / sig
            drift = (-0.5 * beta * x) - (beta * score)
            x = x + drift * dt + g * torch.sqrt(torch.abs(dt)) * noise[i]

    t_fin = ts[-1].expand(B)
    a = sched.alpha(t_fin).view(B, 1, 1, 1)
    s = sched.sigma(t_fin).view(B, 1, 1, 1)
    x0_hat = (x - s * evaluate(x, t_fin)) / torch.clamp(a, min=1e-6)
    tr.x0_hat = x0_hat
    tr.image = ((x0_hat + 1.0) * 0.5).clamp(0.0, 1.0)
    return tr


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||2 / ||b||2 in float64 (the parity metric of SURVEY 8c)."""
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300))
