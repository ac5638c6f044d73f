"""GPU parity tests of the latent-prior row (BASELINE configs[3]): libtcs (through the Python shim, i.e.
through the C ABI) against the oracle and the golden vectors generated from the unmodified reference.

Tolerances (same rule as the score network): per-evaluation eps rel-L2 <= 1e-4 in fp32 mode (FFMA GEMMs) and <= 2e-2 in
bf16 mode (tcgen05 GEMMs), measured teacher-forced at the z_t the CUDA path saw.  With random-init weights the DDIM
latents grow to ~1e6 (the 1/sqrt(abar_T) amplification of the schedule), so final latents are compared by rel-L2.
The decoder is fp32: <= 2e-5 absolute on sigmoid outputs."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import latent_prior_oracle as po
import philox_ref
import toycrystals_oracle as orc
from toycrystals_b200 import _cabi
from toycrystals_b200.models import diffusion_prior as pshim
from toycrystals_b200.models import vae as vshim

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prior.pt")
TOL = {"fp32": 1e-4, "bf16": 2e-2}
_cache = {}


def gold():
    if "gold" not in _cache:
        _cache["gold"] = torch.load(GOLDEN, weights_only=False)
    return _cache["gold"]


def prior(precision, cfg=None, seed=0, use_graph=True):
    key = ("prior", precision, tuple(sorted((cfg or {}).items())), seed, use_graph)
    if key not in _cache:
        c = dict(po.PRIOR_CFG, **(cfg or {}))
        m = pshim.DiffusionPriorFiLM(**c, precision=precision, use_graph=use_graph)
        m.load_state_dict(po.prior_default_init(seed, cfg))
        _cache[key] = (m.to("cuda").eval(), po.prior_default_init(seed, cfg), c)
    return _cache[key]


def vae(precision="fp32"):
    key = ("vae", precision)
    if key not in _cache:
        v = vshim.CondVAE(z_dim=32, n_types=4, y_cont_dim=4, precision=precision)
        v.load_state_dict(po.vae_default_init(2))
        _cache[key] = v.to("cuda").eval()
    return _cache[key]


def sched_gpu():
    return pshim.DiffusionSchedule.linear(T=1000, beta_start=1e-4, beta_end=0.05, device=torch.device("cuda"))


def debug_linear(engine, A, W, bias, silu=False, acc_into=None, bf16_out=False):
    M, K = A.shape
    N = W.shape[0]
    out = acc_into.clone() if acc_into is not None else torch.full((M, N), float("nan"), device="cuda")
    _cabi.check(_cabi.lib().tcs_debug_linear(engine, M, N, K, A.contiguous().data_ptr(), W.contiguous().data_ptr(),
                                             None if bias is None else bias.data_ptr(), out.data_ptr(), int(silu),
                                             int(acc_into is not None), int(bf16_out), None))
    return out


# ---- the dense layer alone -----------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 192), (1, 256, 1024), (4096, 1024, 1024), (257, 4096, 1024),
                                   (640, 1024, 4096)])
def test_linear_tcgen05_matches_fp64_on_bf16_rounded_operands(M, N, K):
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn((M, K), generator=g).cuda()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).cuda()
    b = torch.randn((N,), generator=g).cuda()
    want = A.bfloat16().double() @ W.bfloat16().double().T + b.double()
    got = debug_linear(_cabi.ENGINE_TCGEN05, A, W, b)
    assert orc.rel_l2(got, want) < 1e-5   # fp32 accumulation over K terms
    assert float((got.double() - want).abs().max()) < 1e-4 * float(want.abs().max())
    # SiLU + bf16 store (fc1), residual accumulate (fc2), no bias
    got = debug_linear(_cabi.ENGINE_TCGEN05, A, W, b, silu=True, bf16_out=True)
    assert orc.rel_l2(got, torch.nn.functional.silu(want)) < 4e-3
    base = torch.randn((M, N), generator=g).cuda()
    got = debug_linear(_cabi.ENGINE_TCGEN05, A, W, None, acc_into=base)
    assert orc.rel_l2(got, base.double() + want - b.double()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(50, 1024, 64), (300, 96, 128), (64, 16384, 1024), (1, 32, 16), (513, 64, 2048)])
def test_linear_ffma_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g).cuda()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).cuda()
    b = torch.randn((N,), generator=g).cuda()
    want = A.double() @ W.double().T + b.double()
    assert orc.rel_l2(debug_linear(_cabi.ENGINE_SIMT, A, W, b), want) < 1e-6
    assert orc.rel_l2(debug_linear(_cabi.ENGINE_SIMT, A, W, b, silu=True), torch.nn.functional.silu(want)) < 1e-6
    base = torch.randn((M, N), generator=g).cuda()
    assert orc.rel_l2(debug_linear(_cabi.ENGINE_SIMT, A, W, b, acc_into=base), base.double() + want) < 1e-6
    assert orc.rel_l2(debug_linear(_cabi.ENGINE_SIMT, A, W, b, bf16_out=True), want) < 4e-3


# ---- DiffusionPriorFiLM.forward ---------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_matches_the_reference_golden_vectors(precision):
    G = gold()
    m, _, _ = prior(precision)
    for fw in G["forwards"]:
        t = torch.full((G["n"],), fw["t"], dtype=torch.int64)
        eps = m((G["z_init"] * 1.7).cuda(), t.cuda(), G["y_cat"].cuda(), G["y_cont"].cuda())
        assert eps.shape == (G["n"], 32)
        assert orc.rel_l2(eps, fw["eps"]) < TOL[precision], (precision, fw["t"], orc.rel_l2(eps, fw["eps"]))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_with_per_sample_timesteps_and_ragged_batch(precision):
    m, sd, cfg = prior(precision)
    n = 261   # not a multiple of the 128-row tile
    g = torch.Generator().manual_seed(5)
    z = torch.randn((n, 32), generator=g) * 3.0
    t = torch.randint(0, 1000, (n,), generator=g)
    y_cat, y_cont = orc.condition_grid(n, 4, 4)
    want = po.film_prior(sd, cfg, z, t, y_cat, y_cont)
    got = m(z.cuda(), t.cuda(), y_cat.cuda(), y_cont.cuda())
    assert orc.rel_l2(got, want) < TOL[precision]
    # rows are independent: a sub-batch gives bit-identical rows
    sub = m(z[40:90].cuda(), t[40:90].cuda(), y_cat[40:90].cuda(), y_cont[40:90].cuda())
    assert torch.equal(sub, got[40:90])


def test_forward_other_widths():
    for width, blocks in ((256, 2), (512, 3)):
        cfg = dict(width=width, n_blocks=blocks)
        for precision in ("fp32", "bf16"):
            m, sd, c = prior(precision, cfg, seed=3)
            n = 130
            g = torch.Generator().manual_seed(width)
            z, t = torch.randn((n, 32), generator=g), torch.randint(0, 1000, (n,), generator=g)
            y_cat, y_cont = orc.condition_grid(n, 4, 4)
            want = po.film_prior(sd, c, z, t, y_cat, y_cont)
            got = m(z.cuda(), t.cuda(), y_cat.cuda(), y_cont.cuda())
            assert orc.rel_l2(got, want) < TOL[precision], (width, precision)


# ---- DDIM ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,use_graph", [("fp32", True), ("bf16", True), ("bf16", False)])
def test_ddim_against_the_reference_golden_run(precision, use_graph):
    G = gold()
    m, sd, cfg = prior(precision, use_graph=use_graph)
    s = sched_gpu()
    z0, tr = s.ddim_sample(m, G["y_cat"].cuda(), G["y_cont"].cuda(), n_steps=50, z_init=G["z_init"].cuda(), return_trace=True)
    assert torch.equal(tr.timesteps, G["timesteps"])
    S = tr.eps.shape[0]
    assert S == G["eps_trace"].shape[0] == 50
    assert torch.equal(tr.z_in[0].cpu(), G["z_init"])
    # (1) every evaluation, teacher-forced: the oracle evaluates the z_t the CUDA path saw
    worst = 0.0
    for k in range(S):
        t = G["timesteps"][k].repeat(G["n"])
        want = po.film_prior(sd, cfg, tr.z_in[k].cpu(), t, G["y_cat"], G["y_cont"])
        worst = max(worst, orc.rel_l2(tr.eps[k], want))
    assert worst < TOL[precision], (precision, worst)
    # (2) the update is the reference's fp32 expression, bit for bit, given (z_t, eps) and the library's alpha_bars
    # (torch.linspace's CUDA formula; its CPU kernel differs by 1 ulp per SIMD width, see tests/test_cpu_prior.py)
    buf = (C.c_float * 1000)()
    assert _cabi.lib().tcs_prior_schedule_host(1000, 1e-4, 0.05, buf) == 0
    ab = torch.from_numpy(np.frombuffer(buf, dtype=np.float32).copy())
    for k in range(S):
        z, eps = tr.z_in[k].cpu(), tr.eps[k].cpu()
        a_t = ab[G["timesteps"][k]]
        z0p = (z - torch.sqrt(1.0 - a_t) * eps) / (torch.sqrt(a_t) + 1e-8)
        if k == S - 1:
            assert torch.equal(z0p, z0.cpu())
        else:
            a_p = ab[G["timesteps"][k + 1]]
            assert torch.equal(torch.sqrt(a_p) * z0p + torch.sqrt(1.0 - a_p) * eps, tr.z_in[k + 1].cpu()), k
    # (3) free-running result against the reference's own run
    err = orc.rel_l2(z0, G["z0"])
    assert err < (2e-3 if precision == "fp32" else 1e-1), (precision, err)
    if precision == "fp32":
        assert orc.rel_l2(tr.eps, G["eps_trace"]) < 1e-3


def test_ddim_is_deterministic_shard_invariant_and_philox_keyed():
    m, _, _ = prior("bf16")
    s = sched_gpu()
    n = 300
    y_cat, y_cont = (t.cuda() for t in orc.condition_grid(n, 4, 4))
    a, tra = s.ddim_sample(m, y_cat, y_cont, n_steps=12, seed=77, return_trace=True)
    b = s.ddim_sample(m, y_cat, y_cont, n_steps=12, seed=77)
    assert torch.equal(a, b)
    lo = s.ddim_sample(m, y_cat[:100], y_cont[:100], n_steps=12, seed=77, global_index_offset=0)
    hi = s.ddim_sample(m, y_cat[100:], y_cont[100:], n_steps=12, seed=77, global_index_offset=100)
    assert torch.equal(torch.cat([lo, hi]), a)
    c = s.ddim_sample(m, y_cat, y_cont, n_steps=12, seed=78)
    assert not torch.equal(a, c)
    # initial draw = the Philox stream of the score sampler: word 0, groups 0..7 of sample (offset + i)
    for i in (0, 1, 299):
        want = philox_ref.normal_image(77, i, 0).reshape(-1)[:32]
        np.testing.assert_allclose(tra.z_in[0, i].cpu().numpy(), want, atol=2e-5, rtol=0)
    assert tra.timesteps.tolist() == torch.unique_consecutive(torch.round(torch.linspace(999, 0, 12)).long()).tolist()


# ---- CondVAE.decode -----------------------------------------------------------------------------------------------
def test_decode_matches_the_reference_golden_vectors():
    G = gold()
    v = vae()
    x = v.decode(G["z_init"].cuda(), G["y_cat"].cuda(), G["y_cont"].cuda())
    assert x.shape == (G["n"], 1, 64, 64)
    assert float((x.cpu() - G["x_dec"]).abs().max()) < 2e-5
    # the sampling call: normalised latents + un-standardisation fused into the decoder's first kernel
    x = v.decode(G["z0"].cuda(), G["y_cat"].cuda(), G["y_cont"].cuda(), z_mean=G["z_mean"].cuda(), z_std=G["z_std"].cuda())
    bad = ((x.cpu() - G["x"]).abs() > 1e-3).float().mean()
    assert float(bad) < 1e-3   # saturated sigmoid of ~1e6 latents: pixels are 0/1 up to sign flips of tiny sums
    # per-layer check against torch on a ragged batch
    vsd = po.vae_default_init(2)
    n = 37
    g = torch.Generator().manual_seed(9)
    z = torch.randn((n, 32), generator=g) * 2.0
    y_cat, y_cont = orc.condition_grid(n, 4, 4)
    want = po.vae_decode(vsd, po.VAE_CFG, z.double(), y_cat, y_cont.double())
    got = v.decode(z.cuda(), y_cat.cuda(), y_cont.cuda())
    assert float((got.cpu().double() - want).abs().max()) < 2e-5


def test_decode_on_the_tensor_cores():
    """bf16 mode: the three wide ConvTranspose2d stages run as tcgen05 parity-class GEMMs (bf16 activations)."""
    G = gold()
    v = vae("bf16")
    x = v.decode(G["z_init"].cuda(), G["y_cat"].cuda(), G["y_cont"].cuda())
    d = (x.cpu() - G["x_dec"]).abs()
    assert float(d.max()) < 1e-2 and float(d.mean()) < 1e-3, (float(d.max()), float(d.mean()))
    # ragged batches around the tile sizes (8 images per 128-row tile in the first stage, 2 in the second)
    vsd = po.vae_default_init(2)
    g = torch.Generator().manual_seed(11)
    for n in (1, 7, 9, 130):
        z = torch.randn((n, 32), generator=g) * 2.0
        y_cat, y_cont = orc.condition_grid(n, 4, 4)
        want = po.vae_decode(vsd, po.VAE_CFG, z, y_cat, y_cont)
        got = v.decode(z.cuda(), y_cat.cuda(), y_cont.cuda()).cpu()
        d = (got - want).abs()
        assert float(d.max()) < 1e-2 and float(d.mean()) < 1e-3, (n, float(d.max()), float(d.mean()))
    # each image is independent of its batch
    a = v.decode(z[:50].cuda(), y_cat[:50].cuda(), y_cont[:50].cuda())
    assert torch.equal(a.cpu(), got[:50])


def test_sample_images_end_to_end():
    G = gold()
    m, _, _ = prior("fp32")
    x = pshim.sample_images(vae(), m, sched_gpu(), G["y_cat"].cuda(), G["y_cont"].cuda(), G["z_mean"].cuda(), G["z_std"].cuda(),
                            50, z_init=G["z_init"].cuda())
    assert x.shape == (G["n"], 1, 64, 64) and float(x.min()) >= 0.0 and float(x.max()) <= 1.0
    assert float(((x.cpu() - G["x"]).abs() > 1e-3).float().mean()) < 2e-2
    assert m.launch_count() > 0 and vae().launch_count() > 0


def test_ddim_jobs_larger_than_one_chunk():
    """tcs_prior_ddim_sample runs jobs above 16384 samples in chunks: rows on both sides of the boundary must equal the
    same rows sampled on their own (conditions, Philox offsets, output and trace offsets per chunk)."""
    m, _, _ = prior("bf16")
    s = sched_gpu()
    n = 16384 + 70
    y_cat, y_cont = (t.cuda() for t in orc.condition_grid(n, 4, 4))
    z0, tr = s.ddim_sample(m, y_cat, y_cont, n_steps=3, seed=5, return_trace=True)
    lo, hi = 16384 - 40, 16384 + 70
    sub, trs = s.ddim_sample(m, y_cat[lo:hi], y_cont[lo:hi], n_steps=3, seed=5, global_index_offset=lo, return_trace=True)
    assert torch.equal(sub, z0[lo:hi])
    assert torch.equal(trs.eps, tr.eps[:, lo:hi]) and torch.equal(trs.z_in, tr.z_in[:, lo:hi])
    assert bool(torch.isfinite(z0).all())
    # forward (per-sample timesteps) across the same boundary
    g = torch.Generator().manual_seed(3)
    z = torch.randn((n, 32), generator=g).cuda()
    t = torch.randint(0, 1000, (n,), generator=g).cuda()
    e = m(z, t, y_cat, y_cont)
    e_sub = m(z[lo:hi], t[lo:hi], y_cat[lo:hi], y_cont[lo:hi])
    assert torch.equal(e_sub, e[lo:hi])


# ---- round 2: BASELINE configs[3] at full size against the oracle ---------------------------------------------------
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_full_size_forward_and_decode_against_the_oracle(precision):
    """n = 4096 (configs[3]): eps at three timesteps and the decode against the oracle evaluated by PyTorch on the GPU in
    IEEE fp32 (TF32 off) - the golden vectors hold a tiny n only."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")
    m, sd, cfg = prior(precision)
    sdc = {k: v.cuda() for k, v in sd.items()}
    n = 4096
    g = torch.Generator().manual_seed(17)
    z = (torch.randn((n, 32), generator=g) * 2.0).cuda()
    y_cat, y_cont = (t.cuda() for t in orc.condition_grid(n, 4, 4))
    for tv in (999, 500, 0):
        t = torch.full((n,), tv, dtype=torch.int64, device="cuda")
        with torch.no_grad():
            want = po.film_prior(sdc, cfg, z, t, y_cat, y_cont)
        got = m(z, t, y_cat, y_cont)
        err = orc.rel_l2(got, want)
        rows = ((got - want).norm(dim=1) / want.norm(dim=1)).max().item()
        print(f"prior {precision} n={n} t={tv}: eps rel-L2 {err:.3e}, worst row {rows:.3e}")
        assert err < TOL[precision], (tv, err)
        assert rows < 3 * TOL[precision], (tv, rows)
    v = vae(precision)
    vsd = {k: t.cuda() for k, t in po.vae_default_init(2).items()}
    zn = torch.randn((n, 32), generator=g).cuda()
    with torch.no_grad():
        want = po.vae_decode(vsd, po.VAE_CFG, zn, y_cat, y_cont)
    got = v.decode(zn, y_cat, y_cont)
    tol = 2e-5 if precision == "fp32" else 1e-2
    assert float((got - want).abs().max()) < tol, float((got - want).abs().max())
    with pytest.raises(ValueError, match="both z_mean and z_std"):
        v.decode(zn, y_cat, y_cont, z_std=torch.ones(32, device="cuda"))
