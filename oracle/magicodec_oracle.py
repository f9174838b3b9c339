"""ORACLE — test infrastructure only.  ARCHITECTURE UNPINNED BY THE REFERENCE; attention / rotary / RMSNorm semantics
PINNED to the reference's named kernel dependency, flash-attention 2.8.3 (tests/test_gpu_flash_attn_pin.py, run on B200:
this module's SDPA + band mask == flash_attn_func(causal=True, window_size=(32, 0)) up to flash-attn's bf16 output
rounding, apply_rope == flash_attn.layers.rotary.apply_rotary_emb(interleaved=False) to 5e-7, _RMSNorm ==
flash_attn.ops.triton.layer_norm.rms_norm_fn to 1e-6; magicodec_build.sh:4-16 builds exactly those kernels).

fp32 PyTorch restatement of the MagiCodec network behind
/root/reference/realtime_codec_agent/audio_tokenizer.py.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this file; the product package never does.

Why "unpinned": the arithmetic of the path lives in three third-party
dependencies that are absent from /root/reference and from this image:
  * Ereboas/MagiCodec            git clone, unpinned HEAD  (magicodec_build.sh:2)
  * codec-bpe[magicodec]         pip, unpinned              (requirements.txt:2)
  * Dao-AILab/flash-attention    commit 92dd570 + csrc/{rotary,layer_norm,fused_dense_lib}
                                                           (magicodec_build.sh:4-16)
and the reference ships no tests, fixtures or golden vectors.
What IS pinned is the seam the reference wrapper drives, and this module
implements exactly that duck type (audio_tokenizer.py:28,32,36,158,190-200):

    .eval() .to(device) .sample_rate .codebook_size
    .pad_audio(f32[B,T]) -> f32[B,T']                      (:190)
    .encoder(f32[B,T']) -> z_e[B,F,dq]                     (:191)
    .quantizer.inference(z_e) -> (z_q, i64[B,F])           (:192)
    .quantizer.codebook.weight, .quantizer.codebook_proj   (:158,198)
    .decoder(z_q[B,F,dq]) -> [B,1,F*hop]                   (:200)

The published algorithms it restates:
  * sliding-window attention: keys j in [i-wl, i+wr] inclusive — flash-attn
    ``window_size=(left,right)`` semantics (flash_attn_interface.py, "local
    attention" docstring of flash_attn_func in the installed 2.8.3 wheel);
  * rotary: non-interleaved "rotate half" (flash-attn csrc/rotary):
    out[:h] = x[:h]*cos - x[h:]*sin ; out[h:] = x[:h]*sin + x[h:]*cos;
  * RMSNorm with fp32 statistics (flash-attn csrc/layer_norm, is_rms_norm);
  * SimVQ-style quantiser: nearest neighbour (squared L2, first index on
    ties) over ``codebook_proj(codebook.weight)``.
The golden vectors under tests/golden/ were produced by the UNMODIFIED
reference wrapper driving this module (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


def gelu_tanh(x: torch.Tensor) -> torch.Tensor:
    return F.gelu(x, approximate="tanh")


class _RMSNorm(nn.Module):
    def __init__(self, d: int, eps: float):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xf = x.float()
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        return (y * self.weight.float()).to(x.dtype)


def rope_tables(frames: int, head_dim: int, base: float, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    half = head_dim // 2
    inv_freq = 1.0 / (base ** (torch.arange(0, half, dtype=torch.float32, device=device) * 2.0 / head_dim))
    t = torch.arange(frames, dtype=torch.float32, device=device)
    freqs = torch.outer(t, inv_freq)
    return torch.cos(freqs), torch.sin(freqs)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: [B,F,H,dh]; cos/sin: [F,dh/2] — non-interleaved rotation."""
    half = x.shape[-1] // 2
    x1, x2 = x[..., :half].float(), x[..., half:].float()
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    return torch.cat((x1 * c - x2 * s, x1 * s + x2 * c), dim=-1).to(x.dtype)


def band_mask(frames: int, wl: int, wr: int, device=None) -> torch.Tensor:
    i = torch.arange(frames, device=device)[:, None]
    j = torch.arange(frames, device=device)[None, :]
    return (j >= i - wl) & (j <= i + wr)


class _Attention(nn.Module):
    def __init__(self, spec):
        super().__init__()
        d = spec.d_model
        self.h, self.dh = spec.n_heads, spec.head_dim
        self.wl, self.wr, self.base = spec.window_left, spec.window_right, spec.rope_base
        self.wqkv = nn.Linear(d, 3 * d)
        self.wo = nn.Linear(d, d)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, Fr, d = x.shape
        qkv = self.wqkv(x).view(B, Fr, 3, self.h, self.dh)
        cos, sin = rope_tables(Fr, self.dh, self.base, x.device)
        q = apply_rope(qkv[:, :, 0], cos, sin).transpose(1, 2)
        k = apply_rope(qkv[:, :, 1], cos, sin).transpose(1, 2)
        v = qkv[:, :, 2].transpose(1, 2)
        mask = band_mask(Fr, self.wl, self.wr, x.device)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=1.0 / math.sqrt(self.dh))
        return self.wo(o.transpose(1, 2).reshape(B, Fr, d))


class _MLP(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.w1 = nn.Linear(spec.d_model, spec.ffn_dim)
        self.w2 = nn.Linear(spec.ffn_dim, spec.d_model)

    def forward(self, x):
        return self.w2(gelu_tanh(self.w1(x)))


class _Block(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.norm1 = _RMSNorm(spec.d_model, spec.norm_eps)
        self.attn = _Attention(spec)
        self.norm2 = _RMSNorm(spec.d_model, spec.norm_eps)
        self.mlp = _MLP(spec)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _Encoder(nn.Module):
    def __init__(self, spec):
        super().__init__()
        ch = spec.enc_channels
        self.strides = tuple(spec.conv_strides)
        for i, s in enumerate(self.strides):
            self.add_module(f"conv{i}", nn.Conv1d(ch[i], ch[i + 1], 2 * s, stride=s))
        self.layers = nn.ModuleList(_Block(spec) for _ in range(spec.enc_layers))
        self.norm_f = _RMSNorm(spec.d_model, spec.norm_eps)
        self.proj = nn.Linear(spec.d_model, spec.codebook_dim)

    def frontend(self, wav: torch.Tensor) -> torch.Tensor:
        x = wav[:, None, :]
        n = len(self.strides)
        for i, s in enumerate(self.strides):
            conv = getattr(self, f"conv{i}")
            x = conv(F.pad(x, (s, 0)))          # causal: kernel 2s, left pad s  ->  T/s outputs
            if i < n - 1:
                x = gelu_tanh(x)
        return x.transpose(1, 2)                # [B,F,d]

    def forward(self, wav: torch.Tensor) -> torch.Tensor:
        x = self.frontend(wav)
        for blk in self.layers:
            x = blk(x)
        return self.proj(self.norm_f(x))        # z_e [B,F,dq]


class _Decoder(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.in_proj = nn.Linear(spec.codebook_dim, spec.d_model)
        self.layers = nn.ModuleList(_Block(spec) for _ in range(spec.dec_layers))
        self.norm_f = _RMSNorm(spec.d_model, spec.norm_eps)
        dch = spec.dec_channels
        self.strides = tuple(spec.dec_strides)
        for i, s in enumerate(self.strides):
            self.add_module(f"up{i}", nn.ConvTranspose1d(dch[i], dch[i + 1], 2 * s, stride=s))

    def forward(self, z_q: torch.Tensor) -> torch.Tensor:
        x = self.in_proj(z_q)
        for blk in self.layers:
            x = blk(x)
        x = self.norm_f(x).transpose(1, 2)      # [B,d,F]
        n = len(self.strides)
        for i, s in enumerate(self.strides):
            t = x.shape[-1]
            x = getattr(self, f"up{i}")(x)[..., : t * s]   # causal: drop the trailing s samples
            if i < n - 1:
                x = gelu_tanh(x)
        return x                                 # [B,1,F*hop]


class _Quantizer(nn.Module):
    def __init__(self, spec):
        super().__init__()
        self.codebook = nn.Embedding(spec.codebook_size, spec.codebook_dim)
        self.codebook_proj = nn.Linear(spec.codebook_dim, spec.codebook_dim)

    def projected(self) -> torch.Tensor:
        return self.codebook_proj(self.codebook.weight)

    @torch.no_grad()
    def inference(self, z_e: torch.Tensor, return_margin: bool = False, block: int = 1024):
        """Nearest neighbour in squared L2; ties -> lowest index (torch.argmin semantics on CPU)."""
        cb = self.projected().float()
        z = z_e.float().reshape(-1, z_e.shape[-1])
        c2 = cb.pow(2).sum(-1)
        idx = torch.empty(z.shape[0], dtype=torch.long, device=z.device)
        margin = torch.empty(z.shape[0], dtype=torch.float32, device=z.device)
        for s in range(0, z.shape[0], block):
            zz = z[s:s + block]
            dist = zz.pow(2).sum(-1, keepdim=True) - 2.0 * zz @ cb.t() + c2[None, :]
            if return_margin:
                top2 = torch.topk(dist, 2, dim=-1, largest=False).values
                margin[s:s + block] = top2[:, 1] - top2[:, 0]
            idx[s:s + block] = torch.argmin(dist, dim=-1)
        idx = idx.view(z_e.shape[:-1])
        z_q = F.embedding(idx, cb).to(z_e.dtype)
        if return_margin:
            return z_q, idx, margin.view(z_e.shape[:-1])
        return z_q, idx


class OracleGenerator(nn.Module):
    """Duck-typed stand-in for the MagiCodec generator the reference wrapper expects."""

    def __init__(self, spec, weights: Optional[Dict[str, torch.Tensor]] = None):
        super().__init__()
        self.spec = spec
        self.sample_rate = spec.sample_rate
        self.codebook_size = spec.codebook_size
        self.hop = spec.hop
        self.encoder = _Encoder(spec)
        self.quantizer = _Quantizer(spec)
        self.decoder = _Decoder(spec)
        if weights is not None:
            self.load_flat(weights)
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    def load_flat(self, weights: Dict[str, torch.Tensor]) -> None:
        remap = {}
        for k, v in weights.items():
            k2 = k.replace("enc.", "encoder.", 1) if k.startswith("enc.") else k
            k2 = k2.replace("dec.", "decoder.", 1) if k2.startswith("dec.") else k2
            remap[k2] = v
        missing, unexpected = self.load_state_dict(remap, strict=True), None
        del missing, unexpected

    def pad_audio(self, x: torch.Tensor) -> torch.Tensor:
        """Right-pad to a multiple of the hop (SURVEY.md §7.2 'pad_audio semantics')."""
        rem = (-x.shape[-1]) % self.hop
        return F.pad(x, (0, rem)) if rem else x
