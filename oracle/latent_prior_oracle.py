"""CPU oracle for the "next" row of SURVEY 8(f): latent diffusion prior (FiLM residual MLP, DDIM eta=0) followed
by the conditional VAE decoder — BASELINE configs[3].

TEST INFRASTRUCTURE ONLY (same rule as toycrystals_oracle.py: imported by tests/, smoke and bench CPU legs only).

State-dict-driven, dtype-generic restatement in torch.nn.functional of
  * ``prior_time_features`` <- timestep_embedding             src/toycrystals/models/diffusion_prior.py:11-25
                               (int64 timesteps, [sin, cos] order, no 2*pi — unlike the score net's embedding)
  * ``film_prior``          <- DiffusionPriorFiLM.forward      :108-127   (FiLMResBlock :39-54)
  * ``DdpmSchedule``        <- DiffusionSchedule.linear        :177-188
  * ``ddim_sample``         <- DiffusionSchedule.ddim_sample   :200-252
  * ``vae_decode``          <- CondVAE.decode (eval mode)      src/toycrystals/models/vae.py:62-70
  * ``sample_images``       <- save_diffusion_samples          scripts/train_diffusion_prior.py:61-95
  * default inits in constructor order (DiffusionPriorFiLM :62-106, CondVAE :9-43).
Pinned bit-exactly (fp32, CPU) against the reference module by oracle/gen_golden_prior.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

StateDict = Dict[str, torch.Tensor]
PRIOR_CFG = dict(z_dim=32, n_types=4, y_cont_dim=4, t_emb_dim=64, width=1024, n_blocks=8, y_cat_emb_dim=64)
VAE_CFG = dict(z_dim=32, n_types=4, y_cont_dim=4)
LN_EPS = 1e-5


def _collect(plan, seed) -> StateDict:
    torch.manual_seed(seed)
    sd: StateDict = {}
    for name, make in plan:
        for k, v in make().state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def prior_default_init(seed: int, cfg: Optional[dict] = None) -> StateDict:
    c = dict(PRIOR_CFG, **(cfg or {}))
    w, e = c["width"], c["y_cat_emb_dim"]
    L, nn = torch.nn.Linear, torch.nn
    plan = [("y_cat_emb", lambda: nn.Embedding(c["n_types"], e)),
            ("y_cont_mlp.0", lambda: L(c["y_cont_dim"], e)), ("y_cont_mlp.2", lambda: L(e, e)),
            ("y_fuse.0", lambda: L(2 * e, w)), ("y_fuse.2", lambda: L(w, w)),
            ("t_mlp.0", lambda: L(c["t_emb_dim"], w)), ("t_mlp.2", lambda: L(w, w)),
            ("in_proj", lambda: L(c["z_dim"], w))]
    for i in range(c["n_blocks"]):
        plan += [(f"blocks.{i}.norm", lambda: nn.LayerNorm(w)), (f"blocks.{i}.fc1", lambda: L(w, 4 * w)),
                 (f"blocks.{i}.fc2", lambda: L(4 * w, w)), (f"blocks.{i}.cond", lambda: L(2 * w, 2 * w))]
    plan += [("out_norm", lambda: nn.LayerNorm(w)), ("out_proj", lambda: L(w, c["z_dim"]))]
    return _collect(plan, seed)


def vae_default_init(seed: int, cfg: Optional[dict] = None) -> StateDict:
    c = dict(VAE_CFG, **(cfg or {}))
    nn = torch.nn
    yd = c["n_types"] + c["y_cont_dim"]
    chans = [1, 32, 64, 128, 256]
    plan = [(f"enc.{2 * i}", (lambda i=i: nn.Conv2d(chans[i], chans[i + 1], 4, 2, 1))) for i in range(4)]
    plan += [("enc_fc", lambda: nn.Linear(256 * 16 + yd, 256)), ("mu", lambda: nn.Linear(256, c["z_dim"])),
             ("logvar", lambda: nn.Linear(256, c["z_dim"])), ("dec_fc", lambda: nn.Linear(c["z_dim"] + yd, 256 * 16))]
    dch = [256, 128, 64, 32, 1]
    plan += [(f"dec.{2 * i}", (lambda i=i: nn.ConvTranspose2d(dch[i], dch[i + 1], 4, 2, 1))) for i in range(4)]
    return _collect(plan, seed)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"].to(x), sd[name + ".bias"].to(x))


def prior_time_features(t: torch.Tensor, dim: int, dtype=torch.float32) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(torch.linspace(0, math.log(10_000), steps=half, device=t.device, dtype=dtype) * (-1.0))
    args = t.to(dtype)[:, None] * freqs[None, :]
    emb = torch.cat([torch.sin(args), torch.cos(args)], dim=1)
    if dim % 2 == 1:
        emb = torch.cat([emb, torch.zeros((emb.shape[0], 1), device=t.device, dtype=dtype)], dim=1)
    return emb


def film_prior(sd: StateDict, cfg: dict, z_t, t, y_cat, y_cont) -> torch.Tensor:
    dt = z_t.dtype
    w = cfg["width"]
    te = prior_time_features(t, cfg["t_emb_dim"], dt)
    t_feat = _lin(sd, "t_mlp.2", F.silu(_lin(sd, "t_mlp.0", te)))
    yc = F.embedding(y_cat, sd["y_cat_emb.weight"].to(dt))
    yk = _lin(sd, "y_cont_mlp.2", F.silu(_lin(sd, "y_cont_mlp.0", y_cont.to(dt))))
    y_feat = _lin(sd, "y_fuse.2", F.silu(_lin(sd, "y_fuse.0", torch.cat([yc, yk], dim=-1))))
    cond = torch.cat([t_feat, y_feat], dim=-1)
    h = _lin(sd, "in_proj", z_t)
    for i in range(cfg["n_blocks"]):
        p = f"blocks.{i}."
        u = F.layer_norm(h, (w,), sd[p + "norm.weight"].to(dt), sd[p + "norm.bias"].to(dt), LN_EPS)
        gamma, beta = _lin(sd, p + "cond", cond).chunk(2, dim=-1)
        u = u * (1.0 + gamma) + beta
        h = h + _lin(sd, p + "fc2", F.silu(_lin(sd, p + "fc1", u)))
    h = F.layer_norm(h, (w,), sd["out_norm.weight"].to(dt), sd["out_norm.bias"].to(dt), LN_EPS)
    return _lin(sd, "out_proj", h)


@dataclass
class DdpmSchedule:
    betas: torch.Tensor
    alpha_bars: torch.Tensor

    @staticmethod
    def linear(T: int, beta_start: float, beta_end: float, dtype=torch.float32) -> "DdpmSchedule":
        betas = torch.linspace(beta_start, beta_end, steps=T, dtype=dtype)
        return DdpmSchedule(betas, torch.cumprod(1.0 - betas, dim=0))

    def timesteps(self, n_steps: int) -> torch.Tensor:
        T = int(self.betas.shape[0])
        ts = torch.round(torch.linspace(T - 1, 0, steps=n_steps)).to(torch.int64)
        return torch.unique_consecutive(ts)


@torch.no_grad()
def ddim_sample(sd, cfg, sched: DdpmSchedule, y_cat, y_cont, z_init, n_steps: int = 50, eps_trace: Optional[List] = None):
    """eta = 0 DDIM with the initial latent injected (the reference draws torch.randn((B, z_dim)))."""
    z = z_init.clone()
    B = z.shape[0]
    ts = sched.timesteps(n_steps)
    n = int(ts.numel())
    for i in range(n):
        t = ts[i].repeat(B)
        eps = film_prior(sd, cfg, z, t, y_cat, y_cont)
        if eps_trace is not None:
            eps_trace.append(eps.clone())
        abar = sched.alpha_bars.to(z)[t].unsqueeze(1)
        z0 = (z - torch.sqrt(1.0 - abar) * eps) / (torch.sqrt(abar) + 1e-8)
        if i == n - 1:
            return z0
        abar_p = sched.alpha_bars.to(z)[ts[i + 1].repeat(B)].unsqueeze(1)
        z = torch.sqrt(abar_p) * z0 + torch.sqrt(1.0 - abar_p) * eps
    return z


def vae_decode(sd: StateDict, cfg: dict, z, y_cat, y_cont) -> torch.Tensor:
    dt = z.dtype
    y = torch.cat([F.one_hot(y_cat, num_classes=cfg["n_types"]).to(dt), y_cont.to(dt)], dim=1)
    h = _lin(sd, "dec_fc", torch.cat([z, y], dim=1)).view(-1, 256, 4, 4)
    for i in range(4):
        h = F.conv_transpose2d(h, sd[f"dec.{2 * i}.weight"].to(dt), sd[f"dec.{2 * i}.bias"].to(dt), stride=2, padding=1)
        h = torch.sigmoid(h) if i == 3 else F.relu(h)
    return h


@torch.no_grad()
def sample_images(psd, pcfg, vsd, vcfg, sched, y_cat, y_cont, z_init, z_mean, z_std, n_steps: int = 50):
    z_norm = ddim_sample(psd, pcfg, sched, y_cat, y_cont, z_init, n_steps)
    z = z_norm * z_std + z_mean
    return z_norm, vae_decode(vsd, vcfg, z, y_cat, y_cont)
