"""Drop-in for the decoder of the reference ``CondVAE`` (``src/toycrystals/models/vae.py:9-70``): the last stage of the
latent-diffusion-prior sampling path (BASELINE configs[3]).

``CondVAE`` keeps the reference constructor, construction order (default initialisation under a torch seed is
identical) and ``state_dict()`` keys/shapes, so reference checkpoints load verbatim.  ``decode`` (eval mode) runs in
libtcs (dec_fc + four ConvTranspose2d(4,2,1) stages as parity-class GEMMs); the encoder / training members are out of
scope of this package and raise.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.nn as nn

from .. import _cabi
from .sde_score_model import _as_f32, _ptr, _stream_ptr


_PRECISIONS = {"fp32": _cabi.FP32, "bf16": _cabi.BF16}


class CondVAE(nn.Module):
    """``precision`` (keyword-only addition): "bf16" = the three wide ConvTranspose2d stages on the tensor cores (tcgen05,
    bf16 activations/weights, fp32 accumulation; default), "fp32" = FFMA kernels (parity mode).  TCS_PRECISION overrides."""

    def __init__(self, z_dim: int = 16, n_types: int = 4, y_cont_dim: int = 4, cond_drop: float = 0.1, *,
                 precision: Optional[str] = None) -> None:
        super().__init__()
        self.precision = (precision or os.environ.get("TCS_PRECISION", "bf16")).lower()
        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}, got {self.precision!r}")
        self.z_dim = z_dim
        self.n_types = n_types
        self.y_cont_dim = y_cont_dim
        self.y_dim = n_types + y_cont_dim
        self.cond_drop = float(cond_drop)
        self.enc = nn.Sequential(
            nn.Conv2d(1, 32, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(32, 64, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(64, 128, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(128, 256, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
        )
        self.enc_fc = nn.Linear(256 * 4 * 4 + self.y_dim, 256)
        self.mu = nn.Linear(256, z_dim)
        self.logvar = nn.Linear(256, z_dim)
        self.dec_fc = nn.Linear(z_dim + self.y_dim, 256 * 4 * 4)
        self.dec = nn.Sequential(
            nn.ConvTranspose2d(256, 128, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(128, 64, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(64, 32, kernel_size=4, stride=2, padding=1), nn.ReLU(inplace=True),
            nn.ConvTranspose2d(32, 1, kernel_size=4, stride=2, padding=1), nn.Sigmoid(),
        )
        self._handle: Optional[C.c_void_p] = None
        self._handle_key = None

    def _release(self):
        if self._handle is not None:
            _cabi.lib().tcs_vae_destroy(self._handle)
            self._handle = None
            self._handle_key = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def engine_handle(self) -> C.c_void_p:
        ps = [p for n, p in self.named_parameters() if n.startswith("dec")]
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("toycrystals_b200 runs on a CUDA (B200, sm_100a) device only; move the model with "
                               ".to('cuda') — there is no CPU fallback (use the reference package for --device cpu)")
        key = (dev, self.precision, tuple((p.data_ptr(), p._version) for p in ps))
        if self._handle is not None and key == self._handle_key:
            return self._handle
        self._release()
        L = _cabi.lib()
        cfg = _cabi.TcsVaeConfig(int(self.z_dim), int(self.n_types), int(self.y_cont_dim),
                                 dev.index if dev.index is not None else torch.cuda.current_device(),
                                 _PRECISIONS[self.precision])
        h = C.c_void_p()
        _cabi.check(L.tcs_vae_create(C.byref(h), C.byref(cfg)))
        try:
            torch.cuda.synchronize(dev)
            for name, t in self.state_dict().items():
                w = t.detach().to(torch.float32).contiguous()
                shape = (C.c_int64 * w.dim())(*w.shape)
                _cabi.check(L.tcs_vae_set_weight(h, name.encode(), w.data_ptr(), shape, w.dim()))
            _cabi.check(L.tcs_vae_finalize_weights(h))
        except Exception:
            L.tcs_vae_destroy(h)
            raise
        self._handle, self._handle_key = h, key
        return h

    def launch_count(self) -> int:
        return 0 if self._handle is None else int(_cabi.lib().tcs_vae_launch_count(self._handle))

    def decode(self, z: torch.Tensor, y_cat: torch.Tensor, y_cont: torch.Tensor, *, z_mean: Optional[torch.Tensor] = None,
               z_std: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x = dec(dec_fc([z, onehot(y_cat), y_cont])) -> [B,1,64,64].  With ``z_mean``/``z_std`` (keyword-only
        addition) ``z`` is a standardised latent and ``z * z_std + z_mean`` is fused into the first kernel."""
        if self.training and self.cond_drop > 0.0:
            raise NotImplementedError("condition dropout is a training-time feature: call .eval() (training is out of "
                                      "scope of toycrystals_b200)")
        if (z_mean is None) != (z_std is None):
            raise ValueError("decode: give both z_mean and z_std or neither")
        h = self.engine_handle()
        dev = self.dec_fc.weight.device
        B = int(z.shape[0])
        zz = _as_f32(z, dev)
        yc = y_cat.to(device=dev, dtype=torch.int64).contiguous()
        yk = _as_f32(y_cont, dev)
        zm = None if z_mean is None else _as_f32(z_mean, dev).reshape(-1)
        zs = None if z_std is None else _as_f32(z_std, dev).reshape(-1)
        out = torch.empty((B, 1, 64, 64), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().tcs_vae_decode(h, zz.data_ptr(), yc.data_ptr(), yk.data_ptr(), B, _ptr(zm), _ptr(zs),
                                                   out.data_ptr(), _stream_ptr(dev)))
        return out

    def encode(self, x, y_cat, y_cont):
        raise NotImplementedError("CondVAE.encode is on the training path, which toycrystals_b200 does not cover")

    def reparameterise(self, mu, logvar):
        raise NotImplementedError("CondVAE.reparameterise is on the training path, which toycrystals_b200 does not cover")

    def forward(self, x, y_cat, y_cont):
        raise NotImplementedError("CondVAE.forward is on the training path, which toycrystals_b200 does not cover")
