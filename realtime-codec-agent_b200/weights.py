"""Seeded random-init weights and the checkpoint format of the B200 engine.

No MagiCodec checkpoint is reachable offline, so BASELINE.json
prescribes "identical random-init weights" for parity and benchmarking.  This
module is the single producer of those weights: a flat ``dict[str, Tensor]``
(fp32, CPU) whose keys are the parameter names listed in ``param_shapes``.
The oracle (tests only) loads the same dict into its nn.Module; the engine
packs it to bf16 / split-bf16 device buffers (see engine.py).

Key layout
----------
enc.conv{i}.weight [Cout,Cin,k]   enc.conv{i}.bias [Cout]          (torch Conv1d layout)
{enc,dec}.layers.{l}.norm1.weight [d]
{enc,dec}.layers.{l}.attn.wqkv.weight [3d,d] / .bias [3d]         (q | k | v, head-major inside each)
{enc,dec}.layers.{l}.attn.wo.weight [d,d] / .bias [d]
{enc,dec}.layers.{l}.norm2.weight [d]
{enc,dec}.layers.{l}.mlp.w1.weight [f,d] / .bias [f]
{enc,dec}.layers.{l}.mlp.w2.weight [d,f] / .bias [d]
enc.norm_f.weight [d]   enc.proj.weight [dq,d] / .bias [dq]
quantizer.codebook.weight [K,dq]   quantizer.codebook_proj.weight [dq,dq] / .bias [dq]
dec.in_proj.weight [d,dq] / .bias [d]   dec.norm_f.weight [d]
dec.up{i}.weight [Cin,Cout,k]   dec.up{i}.bias [Cout]               (torch ConvTranspose1d layout)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict

import torch

from .spec import MagiCodecSpec


def param_shapes(spec: MagiCodecSpec) -> "OrderedDict[str, tuple]":
    spec.validate()
    d, f, dq = spec.d_model, spec.ffn_dim, spec.codebook_dim
    shapes: "OrderedDict[str, tuple]" = OrderedDict()
    ch = spec.enc_channels
    for i, s in enumerate(spec.conv_strides):
        shapes[f"enc.conv{i}.weight"] = (ch[i + 1], ch[i], 2 * s)
        shapes[f"enc.conv{i}.bias"] = (ch[i + 1],)

    def blocks(prefix: str, n: int) -> None:
        for l in range(n):
            p = f"{prefix}.layers.{l}"
            shapes[f"{p}.norm1.weight"] = (d,)
            shapes[f"{p}.attn.wqkv.weight"] = (3 * d, d)
            shapes[f"{p}.attn.wqkv.bias"] = (3 * d,)
            shapes[f"{p}.attn.wo.weight"] = (d, d)
            shapes[f"{p}.attn.wo.bias"] = (d,)
            shapes[f"{p}.norm2.weight"] = (d,)
            shapes[f"{p}.mlp.w1.weight"] = (f, d)
            shapes[f"{p}.mlp.w1.bias"] = (f,)
            shapes[f"{p}.mlp.w2.weight"] = (d, f)
            shapes[f"{p}.mlp.w2.bias"] = (d,)

    blocks("enc", spec.enc_layers)
    shapes["enc.norm_f.weight"] = (d,)
    shapes["enc.proj.weight"] = (dq, d)
    shapes["enc.proj.bias"] = (dq,)
    shapes["quantizer.codebook.weight"] = (spec.codebook_size, dq)
    shapes["quantizer.codebook_proj.weight"] = (dq, dq)
    shapes["quantizer.codebook_proj.bias"] = (dq,)
    shapes["dec.in_proj.weight"] = (d, dq)
    shapes["dec.in_proj.bias"] = (d,)
    blocks("dec", spec.dec_layers)
    shapes["dec.norm_f.weight"] = (d,)
    dch = spec.dec_channels
    for i, s in enumerate(spec.dec_strides):
        shapes[f"dec.up{i}.weight"] = (dch[i], dch[i + 1], 2 * s)
        shapes[f"dec.up{i}.bias"] = (dch[i + 1],)
    return shapes


def init_random_weights(spec: MagiCodecSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic fp32 CPU weights; variance-preserving so activations stay O(1)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    n_res = 2 * max(spec.enc_layers, spec.dec_layers, 1)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(spec).items():
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("norm_f.weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif name == "quantizer.codebook.weight":
            t = torch.randn(shape, generator=g)
        elif name == "quantizer.codebook_proj.weight":
            t = torch.eye(shape[0]) + 0.1 * torch.randn(shape, generator=g)
        elif ".conv" in name:           # Conv1d [Cout,Cin,k]
            fan_in = shape[1] * shape[2]
            t = torch.randn(shape, generator=g) * (1.7 / math.sqrt(fan_in))
        elif ".up" in name:             # ConvTranspose1d [Cin,Cout,k]; two taps land on each output
            fan_in = shape[0] * 2
            t = torch.randn(shape, generator=g) * (1.7 / math.sqrt(fan_in))
        else:
            fan_in = shape[1]
            std = 1.0 / math.sqrt(fan_in)
            if name.endswith("attn.wo.weight") or name.endswith("mlp.w2.weight"):
                std /= math.sqrt(n_res) / 2.0
            t = torch.randn(shape, generator=g) * std
        out[name] = t.to(torch.float32).contiguous()
    # the very first conv sees a waveform in [-1,1]: lift it to O(1)
    out["enc.conv0.weight"] *= 4.0
    # ... and the last transposed conv lands back in a waveform-like range
    out[f"dec.up{len(spec.dec_strides) - 1}.weight"] *= 0.1
    return out


def save_checkpoint(path: str, spec: MagiCodecSpec, weights: Dict[str, torch.Tensor]) -> None:
    """A plain dict of python scalars / lists and tensors: loadable with ``weights_only=True``."""
    import dataclasses
    s = {k: (list(v) if isinstance(v, tuple) else v) for k, v in dataclasses.asdict(spec).items()}
    torch.save({"spec": s, "weights": {k: v.detach().cpu() for k, v in weights.items()}}, path)


def load_checkpoint(path: str):
    # weights_only: a checkpoint path comes from the caller or from $MAGICODEC_B200_CHECKPOINT; never unpickle code
    blob = torch.load(path, map_location="cpu", weights_only=True)
    s = dict(blob["spec"])
    for k in ("conv_channels", "conv_strides"):
        s[k] = tuple(s[k])
    spec = MagiCodecSpec(**s)
    shapes = param_shapes(spec)
    weights = blob["weights"]
    missing = [k for k in shapes if k not in weights]
    if missing:
        raise KeyError(f"checkpoint {path} lacks parameters: {missing[:5]}...")
    for k, shp in shapes.items():
        if tuple(weights[k].shape) != tuple(shp):
            raise ValueError(f"{k}: checkpoint shape {tuple(weights[k].shape)} != spec shape {shp}")
    return spec, {k: weights[k].to(torch.float32).contiguous() for k in shapes}
