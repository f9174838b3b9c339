"""GPU (B200): pins the network's attention / rotary / RMSNorm semantics to the reference's NAMED kernel
dependency, flash-attention (/root/reference/magicodec_build.sh:4-16 builds flash-attn + csrc/rotary +
csrc/layer_norm; the installed wheel is flash_attn 2.8.3, whose Python entry points are the same ones the
@92dd570 checkout exposes).  Two things are checked against it on identical bf16 inputs:

  * the ORACLE's restatement (SDPA + explicit band mask, rotate-half RoPE, fp32-statistics RMSNorm —
    oracle/magicodec_oracle.py), so the oracle's semantics are no longer pinned only to themselves;
  * the ENGINE's kernels (mc_op_attention, the QKV GEMM's RoPE epilogue, mc_op_rmsnorm) through the C ABI.

Tolerances: flash-attn returns bf16; its own distance to an fp64 evaluation of the same formula is measured
in the test and the engine / oracle are held to a small multiple of it.
"""
import math

import pytest
import torch

import realtime_codec_agent_b200 as pkg
from oracle import magicodec_oracle as orc

pytestmark = pytest.mark.gpu

flash_attn = pytest.importorskip("flash_attn", reason="flash-attn (the reference's kernel dependency) is not installed")


@pytest.fixture(scope="module")
def gen():
    spec = pkg.MID_SPEC                      # 8 heads x 64, causal window 32 — the kernels' instantiated shapes
    return pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda", max_positions=1024)


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def _attention_fp64(q, k, v, wl, wr):
    """softmax(q k^T / sqrt(dh) + band mask) v in fp64; q,k,v [B,F,H,dh]."""
    B, Fr, H, dh = q.shape
    qd, kd, vd = (t.double().transpose(1, 2) for t in (q, k, v))
    s = qd @ kd.transpose(-1, -2) / math.sqrt(dh)
    mask = orc.band_mask(Fr, wl, wr, q.device)
    s = s.masked_fill(~mask, float("-inf"))
    return (torch.softmax(s, -1) @ vd).transpose(1, 2)            # [B,F,H,dh]


@pytest.mark.parametrize("Fr", [37, 100, 128, 300])
def test_attention_engine_and_oracle_match_flash_attn(gen, Fr):
    """flash_attn_func(causal=True, window_size=(wl, 0)) == the oracle's SDPA + band mask == mc_op_attention."""
    spec = gen.spec
    B, H, dh, d = 3, spec.n_heads, spec.head_dim, spec.d_model
    wl, wr = spec.window_left, spec.window_right
    assert wr == 0
    qkv = _rand((B * Fr, 3 * d), seed=100 + Fr).to(torch.bfloat16)            # q | k | v, head-major inside each
    q, k, v = (qkv[:, i * d:(i + 1) * d].reshape(B, Fr, H, dh) for i in range(3))

    ref_fa = flash_attn.flash_attn_func(q.contiguous(), k.contiguous(), v.contiguous(), softmax_scale=1.0 / math.sqrt(dh),
                                        causal=True, window_size=(wl, 0))                    # [B,F,H,dh] bf16
    exact = _attention_fp64(q, k, v, wl, wr)
    fa_err = (ref_fa.double() - exact).abs().max().item()

    # the oracle's restatement (what _Attention.forward runs after RoPE), fp32 on the same bf16 values
    mask = orc.band_mask(Fr, wl, wr, q.device)
    o_oracle = torch.nn.functional.scaled_dot_product_attention(
        q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2), attn_mask=mask,
        scale=1.0 / math.sqrt(dh)).transpose(1, 2)
    oracle_vs_fa = (o_oracle.double() - ref_fa.double()).abs().max().item()
    oracle_err = (o_oracle.double() - exact).abs().max().item()

    o_eng = gen.op_attention(qkv, B, Fr).reshape(B, Fr, H, dh)
    eng_vs_fa = (o_eng.double() - ref_fa.double()).abs().max().item()
    eng_err = (o_eng.double() - exact).abs().max().item()
    print(f"[flash-attn pin] F={Fr}: |flash - fp64| {fa_err:.4f}  |oracle - fp64| {oracle_err:.2e}  |oracle - flash| {oracle_vs_fa:.4f}  "
          f"|engine - fp64| {eng_err:.4f}  |engine - flash| {eng_vs_fa:.4f}")
    assert oracle_err < 1e-5                              # same function
    assert oracle_vs_fa <= fa_err + 1e-5                  # flash-attn deviates from it only by its own bf16 rounding
    assert eng_err <= 2.0 * fa_err + 4e-3                 # the engine is as close to the exact value as flash-attn is
    assert eng_vs_fa <= 3.0 * fa_err + 4e-3
    # the SIMT cross-check kernel too
    o_simt = gen.op_attention(qkv, B, Fr, impl=1).reshape(B, Fr, H, dh)
    assert (o_simt.double() - exact).abs().max().item() <= 2.0 * fa_err + 4e-3


def test_band_mask_edges_match_flash_attn_window_semantics(gen):
    """One-hot values make the attended key set readable: key j contributes to query i iff j in [i - wl, i]
    (flash-attn's inclusive window, flash_attn_interface.py 'local attention' docstring)."""
    spec = gen.spec
    B, Fr, H, dh, d = 1, 100, spec.n_heads, spec.head_dim, spec.d_model
    wl = spec.window_left
    q = torch.zeros(B, Fr, H, dh, device="cuda", dtype=torch.bfloat16)       # zero scores -> uniform weights over the window
    k = torch.zeros_like(q)
    v = torch.zeros_like(q)
    for j in range(Fr):
        v[0, j, :, j % dh] = 1.0 + (j // dh)                                  # key j marks column j % dh with 1 or 2
    fa = flash_attn.flash_attn_func(q, k, v, softmax_scale=0.125, causal=True, window_size=(wl, 0)).float()
    qkv = torch.cat([t.reshape(B * Fr, d) for t in (q, k, v)], dim=1).contiguous()
    eng = gen.op_attention(qkv, B, Fr).reshape(B, Fr, H, dh).float()
    for i in (0, 1, wl - 1, wl, wl + 1, 64, Fr - 1):
        n = min(i, wl) + 1
        want = torch.zeros(dh, device="cuda")
        for j in range(i - n + 1, i + 1):
            want[j % dh] += (1.0 + (j // dh)) / n
        assert (fa[0, i, 0] - want).abs().max().item() < 2e-2, f"flash-attn row {i}"
        assert (eng[0, i, 0] - want).abs().max().item() < 2e-2, f"engine row {i}"
        mask_row = orc.band_mask(Fr, wl, 0)[i]
        assert int(mask_row.sum()) == n and bool(mask_row[i]) and bool(mask_row[i - n + 1])


def test_rope_oracle_and_engine_match_flash_attn_rotary(gen):
    """flash_attn.layers.rotary.apply_rotary_emb(interleaved=False) == oracle.apply_rope == the QKV epilogue."""
    from flash_attn.layers.rotary import apply_rotary_emb

    spec = gen.spec
    B, Fr, H, dh, d = 3, 100, spec.n_heads, spec.head_dim, spec.d_model
    cos, sin = orc.rope_tables(Fr, dh, spec.rope_base, device="cuda")          # [F, dh/2] fp32
    x = _rand((B, Fr, H, dh), seed=5)
    fa = apply_rotary_emb(x, cos, sin, interleaved=False)
    ours = orc.apply_rope(x, cos, sin)
    err = (fa - ours).abs().max().item()
    print(f"[flash-attn pin] rotary fp32: |oracle - flash| {err:.2e}")
    assert err < 2e-6
    # engine tables are the same angles
    assert torch.allclose(gen._dev["rope.cos"].t()[:Fr], cos, atol=1e-7) and torch.allclose(gen._dev["rope.sin"].t()[:Fr], sin, atol=1e-7)
    # engine: QKV GEMM with the RoPE epilogue vs (A W^T + b) rotated by flash-attn's kernel
    K = 256
    A = _rand((B * Fr, K), seed=6).to(torch.bfloat16)
    W = (_rand((3 * d, K), seed=7) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((3 * d,), seed=8, scale=0.1)
    out = gen.op_gemm(A, W, bias=bias, out_mode=1, rope_cols=2 * d, rope_period=Fr)           # fp32 out
    lin = (A.float() @ W.float().t() + bias).view(B, Fr, 3, H, dh)
    ref = torch.stack((apply_rotary_emb(lin[:, :, 0].contiguous(), cos, sin), apply_rotary_emb(lin[:, :, 1].contiguous(), cos, sin),
                       lin[:, :, 2]), dim=2).reshape(B * Fr, 3 * d)
    e2 = (out - ref).abs().max().item()
    print(f"[flash-attn pin] rotary QKV epilogue: |engine - flash| {e2:.2e}")
    assert e2 < 2e-3


def test_rmsnorm_oracle_and_engine_match_flash_attn_layer_norm(gen):
    """flash_attn.ops.triton.layer_norm.rms_norm_fn (the successor of csrc/layer_norm's is_rms_norm path; fp32
    statistics, y = x * rsqrt(mean(x^2) + eps) * w) == oracle._RMSNorm == mc_op_rmsnorm."""
    try:
        from flash_attn.ops.triton.layer_norm import rms_norm_fn
    except Exception as ex:                                    # triton not usable on this box
        pytest.skip(f"flash_attn triton layer_norm unavailable: {ex!r}")
    spec = gen.spec
    M, d = 777, spec.d_model
    x = _rand((M, d), seed=9, scale=3.0)
    gamma = 1.0 + 0.1 * _rand((d,), seed=10)
    try:
        fa = rms_norm_fn(x, gamma, None, eps=spec.norm_eps)
    except Exception as ex:
        pytest.skip(f"flash_attn triton rms_norm_fn failed to run here: {ex!r}")
    norm = orc._RMSNorm(d, spec.norm_eps).cuda()
    with torch.no_grad():
        norm.weight.copy_(gamma)
        ours = norm(x)
    e_oracle = (ours - fa).abs().max().item()
    eng = gen.op_rmsnorm(x, gamma)                             # bf16 out
    e_eng = (eng.float() - fa).abs().max().item()
    scale = fa.abs().max().item()
    print(f"[flash-attn pin] rmsnorm: |oracle - flash| {e_oracle:.2e}  |engine(bf16) - flash| {e_eng:.2e} (max |y| {scale:.2f})")
    assert e_oracle < 1e-5
    assert e_eng <= scale * 2.0 ** -8                          # one bf16 rounding of the output
