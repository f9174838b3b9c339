"""Test helper: the engine's dataflow (engine.cu) replayed with plain torch ops on the PACKED
weights — same buffer layouts (left-pad rows, row-block views, junk rows, tap ordering, zero-padded
K) but fp32 math.  With ``gemm_dtype=float32`` packing it must reproduce the oracle to fp32
round-off, which pins every packed layout on the CPU before a kernel ever runs."""
import math

import torch
import torch.nn.functional as F


def gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def gemm_rowblock(A_flat, a_k_wrap, W, bias, M, K, grp_in, grp_valid):
    """Rows g of the row-block view: A[g, k] = A_flat[g*a_k_wrap + k], k < K (overlapping rows)."""
    total = A_flat.numel()
    idx = torch.arange(M)[:, None] * a_k_wrap + torch.arange(K)[None, :]
    A = torch.where(idx < total, A_flat[idx.clamp(max=total - 1)], torch.zeros(()))
    out = A.float() @ W.float().t() + (bias if bias is not None else 0.0)
    r = torch.arange(M) % grp_in
    return out, r < grp_valid


def rmsnorm(x, g, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * g


def layers(spec, p, prefix, n_layers, x, B, Fr):
    d, H = spec.d_model, spec.n_heads
    cos, sin = p["rope.cos"].t()[:Fr], p["rope.sin"].t()[:Fr]
    i = torch.arange(Fr)[:, None]
    j = torch.arange(Fr)[None, :]
    mask = (j >= i - spec.window_left) & (j <= i + spec.window_right)
    for l in range(n_layers):
        q = f"{prefix}.layers.{l}"
        h = rmsnorm(x, p[f"{q}.norm1"], spec.norm_eps)
        qkv = h @ p[f"{q}.wqkv"].float().t() + p[f"{q}.bqkv"]
        qkv = qkv.view(B, Fr, 3, H, 64)

        def rope(t):
            x1, x2 = t[..., :32], t[..., 32:]
            c, s = cos[None, :, None, :], sin[None, :, None, :]
            return torch.cat((x1 * c - x2 * s, x1 * s + x2 * c), -1)

        qq, kk, vv = rope(qkv[:, :, 0]).transpose(1, 2), rope(qkv[:, :, 1]).transpose(1, 2), qkv[:, :, 2].transpose(1, 2)
        att = (qq @ kk.transpose(-1, -2)) * 0.125
        att = att.masked_fill(~mask, float("-inf")).softmax(-1) @ vv
        att = att.transpose(1, 2).reshape(B * Fr, d)
        x = x + att @ p[f"{q}.wo"].float().t() + p[f"{q}.bo"]
        h = rmsnorm(x, p[f"{q}.norm2"], spec.norm_eps)
        u = gelu_tanh(h @ p[f"{q}.w1"].float().t() + p[f"{q}.b1"])
        x = x + u @ p[f"{q}.w2"].float().t() + p[f"{q}.b2"]
    return x


def encode(spec, p, wav):
    """wav [B,T] -> z_e [B,F,16] following engine.cu::encode_impl."""
    B, T = wav.shape
    hop = spec.hop
    Fr = -(-T // hop)
    Tp = Fr * hop
    n = len(spec.conv_strides)
    ch = list(spec.conv_channels) + [spec.d_model]
    Tl, t = [], Tp
    for s in spec.conv_strides:
        t //= s
        Tl.append(t)
    # conv0 (tap-major fp32 weights), output into left-padded channels-last buffer
    s0 = spec.conv_strides[0]
    xpad = F.pad(wav, (s0, Tp - T))
    win = xpad.unfold(1, 2 * s0, s0)[:, : Tl[0]]                      # [B, T0, 2*s0]
    y = gelu_tanh(win @ p["enc.conv0.w"] + p["enc.conv0.b"])         # [B, T0, C0]
    buf = F.pad(y, (0, 0, spec.conv_strides[1], 0))                    # pad rows in front
    for i in range(1, n):
        si = spec.conv_strides[i]
        a_k_wrap = si * ch[i - 1]
        M = B * (1 + Tl[i])
        out, ok = gemm_rowblock(buf.reshape(-1), a_k_wrap, p[f"enc.conv{i}.w"], p[f"enc.conv{i}.b"], M, 2 * a_k_wrap,
                                1 + Tl[i], Tl[i])
        out = out[ok].view(B, Tl[i], ch[i])
        if i + 1 < n:
            buf = F.pad(gelu_tanh(out), (0, 0, spec.conv_strides[i + 1], 0))
        else:
            x = out.reshape(B * Fr, spec.d_model)
    x = layers(spec, p, "enc", spec.enc_layers, x, B, Fr)
    h = rmsnorm(x, p["enc.norm_f"], spec.norm_eps)
    z = h @ p["enc.proj.w"].float().t() + p["enc.proj.b"]
    return z.view(B, Fr, -1)


def vq_scores_packed(p, z):
    """The tensor-core VQ arithmetic: split-bf16 A rows x packed B rows -> -2*score (distance scale)."""
    z = z.reshape(-1, 16).float()
    zh = z.to(torch.bfloat16)
    zl = (z - zh.float()).to(torch.bfloat16)
    A = torch.zeros(z.shape[0], 64)
    A[:, 0:16], A[:, 16:32], A[:, 32:48] = zh.float(), zh.float(), zl.float()
    A[:, 48:51] = 1.0
    return -2.0 * (A.double() @ p["vq.packed"].double().t())


def decode(spec, p, codes):
    """codes [B,F] -> wav [B, F*hop] following engine.cu::decode_impl."""
    B, Fr = codes.shape
    n = len(spec.conv_strides)
    d = spec.d_model
    a0 = torch.zeros(B * Fr, 64)
    a0[:, :16] = p["vq.codebook"][codes.reshape(-1)].to(p["dec.in_proj.w"].dtype).float()
    x = a0 @ p["dec.in_proj.w"].float().t() + p["dec.in_proj.b"]
    x = layers(spec, p, "dec", spec.dec_layers, x, B, Fr)
    h = rmsnorm(x, p["dec.norm_f"], spec.norm_eps).view(B, Fr, d)
    buf = F.pad(h, (0, 0, 1, 0))                                       # [B, 1+F, d]
    dch, ds = list(spec.dec_channels), list(spec.dec_strides)
    Tin = Fr
    for i in range(n - 1):
        M = B * (1 + Tin)
        out, ok = gemm_rowblock(buf.reshape(-1), dch[i], p[f"dec.up{i}.w"], p[f"dec.up{i}.b"], M, 2 * dch[i], 1 + Tin, Tin)
        out = gelu_tanh(out[ok]).view(B, Tin * ds[i], dch[i + 1])      # [rows, s*Cout] == channels-last [T*s, Cout]
        buf = F.pad(out, (0, 0, 1, 0))
        Tin *= ds[i]
    s = ds[-1]
    w = p[f"dec.up{n - 1}.w"]                                          # [Cin, 2s]
    cur, prev = buf[:, 1:], buf[:, :-1]
    y = cur @ w[:, :s] + prev @ w[:, s:] + p[f"dec.up{n - 1}.b"]
    return y.reshape(B, Tin * s)
