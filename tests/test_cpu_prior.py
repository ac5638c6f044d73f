"""CPU tests of the latent-prior row (BASELINE configs[3]): the oracle against the golden vectors generated
from the unmodified reference (oracle/gen_golden_prior.py), libtcs' host-side DDIM schedule against torch, and the
Python shim's reference surface.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import latent_prior_oracle as po
import toycrystals_oracle as orc
from toycrystals_b200 import _cabi
from toycrystals_b200.models import diffusion_prior as pshim
from toycrystals_b200.models import vae as vshim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prior.pt")
torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN, weights_only=False)


def test_oracle_reproduces_the_reference_golden_vectors(gold):
    psd, vsd = po.prior_default_init(gold["seed_prior"]), po.vae_default_init(gold["seed_vae"])
    n = gold["n"]
    for fw in gold["forwards"]:   # single evaluations of DiffusionPriorFiLM.forward
        t = torch.full((n,), fw["t"], dtype=torch.int64)
        eps = po.film_prior(psd, po.PRIOR_CFG, gold["z_init"] * 1.7, t, gold["y_cat"], gold["y_cont"])
        assert torch.equal(eps, fw["eps"])
    sched = po.DdpmSchedule.linear(gold["T"], gold["beta_start"], gold["beta_end"])
    assert torch.equal(sched.timesteps(50), gold["timesteps"])
    # first 3 DDIM evaluations bit-exact (the full 50-step run is asserted by the generator script itself)
    z, ts = gold["z_init"].clone(), gold["timesteps"]
    for i in range(3):
        t = ts[i].repeat(n)
        eps = po.film_prior(psd, po.PRIOR_CFG, z, t, gold["y_cat"], gold["y_cont"])
        assert torch.equal(eps, gold["eps_trace"][i])
        abar, abar_p = sched.alpha_bars[t].unsqueeze(1), sched.alpha_bars[ts[i + 1].repeat(n)].unsqueeze(1)
        z0 = (z - torch.sqrt(1.0 - abar) * eps) / (torch.sqrt(abar) + 1e-8)
        z = torch.sqrt(abar_p) * z0 + torch.sqrt(1.0 - abar_p) * eps
    x = po.vae_decode(vsd, po.VAE_CFG, gold["z_init"], gold["y_cat"], gold["y_cont"])
    assert torch.equal(x, gold["x_dec"])
    x = po.vae_decode(vsd, po.VAE_CFG, gold["z0"] * gold["z_std"] + gold["z_mean"], gold["y_cat"], gold["y_cont"])
    assert torch.equal(x, gold["x"])


def test_host_ddim_schedule_matches_torch():
    L = _cabi.lib()
    for T, b0, b1 in ((1000, 1e-4, 0.05), (200, 1e-4, 1.0), (50, 1e-3, 0.02)):
        buf = (C.c_float * T)()
        assert L.tcs_prior_schedule_host(T, b0, b1, buf) == 0
        mine = np.frombuffer(buf, dtype=np.float32)
        want = po.DdpmSchedule.linear(T, b0, b1).alpha_bars.numpy()
        # linspace: torch's vectorised CPU kernel vs the scalar formula (1 ulp of a beta, see test_cpu_oracle); the
        # cumulative product then agrees to a few ulp
        np.testing.assert_allclose(mine, want, rtol=3e-6, atol=1e-37)
        assert mine[0] == want[0]
    for T, n_steps in ((1000, 50), (1000, 1000), (1000, 2000), (200, 50), (200, 7), (10, 50), (1000, 1), (1000, 2)):
        buf, cnt = (C.c_int64 * max(n_steps, 1))(), C.c_int32()
        assert L.tcs_prior_timesteps_host(T, n_steps, buf, C.byref(cnt)) == 0
        ts = torch.unique_consecutive(torch.round(torch.linspace(T - 1, 0, steps=n_steps)).to(torch.int64))
        assert list(buf[:cnt.value]) == ts.tolist(), (T, n_steps)


def test_shims_keep_the_reference_state_dict_and_default_init():
    torch.manual_seed(0)
    m = pshim.DiffusionPriorFiLM(**po.PRIOR_CFG)
    sd = po.prior_default_init(0)
    assert list(m.state_dict().keys()) == list(sd.keys())
    assert all(torch.equal(v, sd[k]) for k, v in m.state_dict().items())
    torch.manual_seed(2)
    v = vshim.CondVAE(z_dim=32, n_types=4, y_cont_dim=4)
    vsd = po.vae_default_init(2)
    assert list(v.state_dict().keys()) == list(vsd.keys())
    assert all(torch.equal(t, vsd[k]) for k, t in v.state_dict().items())
    # reference constructor defaults
    d = pshim.DiffusionPriorFiLM(32, 4, 4)
    assert d.blocks[0].fc1.weight.shape == (1024, 256) and len(d.blocks) == 6
    assert vshim.CondVAE().z_dim == 16


def test_shims_have_no_cpu_path_and_keep_reference_errors():
    m = pshim.DiffusionPriorFiLM(32, 4, 4, width=256)
    y_cat, y_cont = orc.condition_grid(2, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 32), torch.zeros(2, dtype=torch.long), y_cat, y_cont)
    s = pshim.DiffusionSchedule.linear(T=1000, beta_start=1e-4, beta_end=0.05, device=torch.device("cpu"))
    ref = po.DdpmSchedule.linear(1000, 1e-4, 0.05)
    assert torch.equal(s.alpha_bars, ref.alpha_bars) and torch.equal(s.betas, ref.betas)
    with pytest.raises(NotImplementedError, match="eta != 0"):
        s.ddim_sample(m, y_cat, y_cont, eta=0.5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s.ddim_sample(m, y_cat, y_cont)
    v = vshim.CondVAE(z_dim=32).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        v.decode(torch.zeros(2, 32), y_cat, y_cont)
    with pytest.raises(NotImplementedError):
        v.encode(torch.zeros(2, 1, 64, 64), y_cat, y_cont)
    with pytest.raises(NotImplementedError, match="condition dropout"):
        vshim.CondVAE(z_dim=32).train().decode(torch.zeros(2, 32), y_cat, y_cont)


def test_prior_create_fails_loudly_without_gpu_or_on_bad_config():
    L = _cabi.lib()
    cfg = _cabi.TcsPriorConfig()
    L.tcs_prior_default_config(C.byref(cfg))
    assert (cfg.z_dim, cfg.width, cfg.n_blocks, cfg.T, cfg.beta_end) == (32, 1024, 8, 1000, 0.05)
    h = C.c_void_p()
    cfg.width = 300
    assert L.tcs_prior_create(C.byref(h), C.byref(cfg)) == _cabi.ERR_UNSUPPORTED
    assert b"width" in L.tcs_last_error()
    cfg.width = 1024
    cfg.z_dim = 48
    assert L.tcs_prior_create(C.byref(h), C.byref(cfg)) == _cabi.ERR_UNSUPPORTED
    if not torch.cuda.is_available():
        cfg.z_dim = 32
        assert L.tcs_prior_create(C.byref(h), C.byref(cfg)) == _cabi.ERR_CUDA
        assert b"no CPU fallback" in L.tcs_last_error()
        vcfg = _cabi.TcsVaeConfig(32, 4, 4, 0)
        assert L.tcs_vae_create(C.byref(h), C.byref(vcfg)) == _cabi.ERR_CUDA
