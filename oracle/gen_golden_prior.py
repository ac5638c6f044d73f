#!/usr/bin/env python3
"""Golden vectors for SURVEY 8(f)-1 (latent prior DDIM + CondVAE decode) from the UNMODIFIED reference; asserts the
oracle (oracle/latent_prior_oracle.py) is bit-exact to it in fp32 on CPU.  Writes tests/golden/prior.pt."""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, HERE)
from toycrystals.models.diffusion_prior import DiffusionPriorFiLM, DiffusionSchedule  # noqa: E402
from toycrystals.models.vae import CondVAE  # noqa: E402
import latent_prior_oracle as po  # noqa: E402
import toycrystals_oracle as orc  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)
prior = DiffusionPriorFiLM(**po.PRIOR_CFG).eval()
psd = po.prior_default_init(0)
assert list(prior.state_dict().keys()) == list(psd.keys())
assert all(torch.equal(v, psd[k]) for k, v in prior.state_dict().items())
torch.manual_seed(2)
vae = CondVAE(z_dim=32, n_types=4, y_cont_dim=4).eval()
vsd = po.vae_default_init(2)
assert list(vae.state_dict().keys()) == list(vsd.keys())
assert all(torch.equal(v, vsd[k]) for k, v in vae.state_dict().items())
print("default inits: oracle == reference (prior %d params, vae %d params)"
      % (sum(v.numel() for v in psd.values()), sum(v.numel() for v in vsd.values())))

n = 8
gen = torch.Generator().manual_seed(4321)
y_cat, y_cont = orc.condition_grid(n, 4, 4)
z_init = torch.randn((n, 32), generator=gen)
z_mean, z_std = torch.randn((32,), generator=gen) * 0.3, torch.rand((32,), generator=gen) + 0.5
sched_ref = DiffusionSchedule.linear(T=1000, beta_start=1e-4, beta_end=0.05, device=torch.device("cpu"))
sched = po.DdpmSchedule.linear(1000, 1e-4, 0.05)
assert torch.equal(sched.alpha_bars, sched_ref.alpha_bars)

# single forwards
fw = []
for tval in (999, 500, 0):
    t = torch.full((n,), tval, dtype=torch.int64)
    with torch.no_grad():
        e_ref = prior(z_init * 1.7, t, y_cat, y_cont)
        e_orc = po.film_prior(psd, po.PRIOR_CFG, z_init * 1.7, t, y_cat, y_cont)
    assert torch.equal(e_ref, e_orc)
    fw.append(dict(t=tval, eps=e_ref))
# DDIM with the initial draw injected
real = torch.randn
torch.randn = lambda *a, **k: z_init.clone()
try:
    z_ref = sched_ref.ddim_sample(prior, y_cat=y_cat, y_cont=y_cont, n_steps=50, eta=0.0)
finally:
    torch.randn = real
trace = []
z_orc = po.ddim_sample(psd, po.PRIOR_CFG, sched, y_cat, y_cont, z_init, 50, trace)
assert torch.equal(z_ref, z_orc), "ddim: oracle != reference"
with torch.no_grad():
    x_ref = vae.decode(z_ref * z_std + z_mean, y_cat, y_cont)
_, x_orc = po.sample_images(psd, po.PRIOR_CFG, vsd, po.VAE_CFG, sched, y_cat, y_cont, z_init, z_mean, z_std, 50)
assert torch.equal(x_ref, x_orc), "decode: oracle != reference"
# the random-init DDIM latents are ~1e6 (1/sqrt(abar_T) amplification), which saturates the decoder's sigmoid: pin the
# decoder separately on moderate latents
with torch.no_grad():
    x_dec_ref = vae.decode(z_init, y_cat, y_cont)
x_dec_orc = po.vae_decode(vsd, po.VAE_CFG, z_init, y_cat, y_cont)
assert torch.equal(x_dec_ref, x_dec_orc), "decode(z_init): oracle != reference"
out = dict(x_dec=x_dec_ref, n=n, y_cat=y_cat, y_cont=y_cont, z_init=z_init, z_mean=z_mean, z_std=z_std, forwards=fw, z0=z_ref, x=x_ref,
           eps_trace=torch.stack(trace), timesteps=sched.timesteps(50), T=1000, beta_start=1e-4, beta_end=0.05,
           seed_prior=0, seed_vae=2)
path = os.path.join(os.path.dirname(HERE), "tests", "golden", "prior.pt")
torch.save(out, path)
print("ddim steps", len(trace), "| z0 absmax", float(z_ref.abs().max()), "| x range", float(x_ref.min()), float(x_ref.max()),
      "|", os.path.getsize(path), "bytes")
