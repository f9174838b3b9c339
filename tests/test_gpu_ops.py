"""GPU (B200): every kernel of the engine, one launch at a time through the C ABI, against plain
PyTorch fp32 on the same device."""
import math

import numpy as np
import pytest
import torch

import realtime_codec_agent_b200 as pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gen():
    spec = pkg.MID_SPEC
    return pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda", max_positions=1024)


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(shape, generator=g, device="cuda") * scale


def gelu_tanh(x):
    return torch.nn.functional.gelu(x, approximate="tanh")


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 256, 64), (100, 256, 128), (300, 1024, 256), (1000, 3072, 512),
                                   (2560, 512, 2048), (77, 16, 512)])
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_gemm_fp32_out(gen, M, N, K, block_n):
    A = _rand((M, K), seed=1).to(torch.bfloat16)
    W = (_rand((N, K), seed=2) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=3)
    out = gen.op_gemm(A, W, bias=bias, out_mode=1, block_n=block_n)
    ref = A.float() @ W.float().t() + bias
    err = (out - ref).abs().max().item()
    assert err < 2e-3, f"max abs err {err}"


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (300, 256, 128), (1000, 3072, 512), (2560, 512, 2048), (25600, 1024, 1024)])
@pytest.mark.parametrize("mode", ["f32", "bf16_gelu", "residual"])
def test_gemm_cta_pair_kernel(gen, M, N, K, mode):
    """block_n=512 forces the cta_group::2 kernel (256 x 256 tiles on CTA pairs)."""
    A = _rand((M, K), seed=21).to(torch.bfloat16)
    W = (_rand((N, K), seed=22) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=23)
    ref = A.float() @ W.float().t() + bias
    if mode == "f32":
        out = gen.op_gemm(A, W, bias=bias, out_mode=1, block_n=512)
        tol = 2e-3
    elif mode == "bf16_gelu":
        out = gen.op_gemm(A, W, bias=bias, act=1, out_mode=0, block_n=512).float()
        ref, tol = gelu_tanh(ref), 0.03
    else:
        x0 = _rand((M, N), seed=24)
        out = x0.clone()
        gen.op_gemm(A, W, bias=bias, out_mode=2, out=out, block_n=512)
        ref, tol = x0 + ref, 2e-3
    err = (out - ref).abs().max().item()
    assert err < tol, f"max abs err {err}"


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (300, 512, 128), (1000, 3072, 512), (2560, 512, 2048), (25600, 1024, 1024)])
@pytest.mark.parametrize("mode", ["rope", "gelu", "residual"])
def test_specialised_pair_epilogues_are_bit_identical_to_generic(gen, M, N, K, mode):
    """The mode-specialised, software-pipelined epilogues of the CTA-pair GEMM reorder loads, not arithmetic."""
    A = _rand((M, K), seed=31).to(torch.bfloat16)
    W = (_rand((N, K), seed=32) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=33)
    x0 = _rand((M, N), seed=34)

    def run():
        if mode == "rope":
            return gen.op_gemm(A, W, bias=bias, out_mode=0, rope_cols=(N // 128) * 64, rope_period=100, block_n=512)
        if mode == "gelu":
            return gen.op_gemm(A, W, bias=bias, act=1, out_mode=0, block_n=512)
        out = x0.clone()
        gen.op_gemm(A, W, bias=bias, out_mode=2, out=out, block_n=512)
        return out

    try:
        gen.set_option("fast_epilogue", 1)
        fast = run()
        gen.set_option("fast_epilogue", 0)
        generic = run()
    finally:
        gen.set_option("fast_epilogue", 1)
    assert torch.equal(fast, generic)
    ref = A.float() @ W.float().t() + bias
    if mode == "gelu":
        assert (fast.float() - gelu_tanh(ref)).abs().max().item() < 0.03
    elif mode == "residual":
        assert (fast - (x0 + ref)).abs().max().item() < 2e-3


@pytest.mark.parametrize("act", [0, 1])
def test_gemm_bf16_out_bias_act(gen, act):
    M, N, K = 515, 1024, 512
    A = _rand((M, K), seed=4).to(torch.bfloat16)
    W = (_rand((N, K), seed=5) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=6)
    out = gen.op_gemm(A, W, bias=bias, act=act, out_mode=0)
    ref = A.float() @ W.float().t() + bias
    ref = gelu_tanh(ref) if act else ref
    assert out.dtype == torch.bfloat16
    assert (out.float() - ref).abs().max().item() < 0.03


def test_gemm_residual_accumulates_in_place(gen):
    M, N, K = 300, 512, 1024
    A = _rand((M, K), seed=7).to(torch.bfloat16)
    W = (_rand((N, K), seed=8) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=9)
    x0 = _rand((M, N), seed=10)
    x = x0.clone()
    gen.op_gemm(A, W, bias=bias, out_mode=2, out=x)
    ref = x0 + A.float() @ W.float().t() + bias
    assert (x - ref).abs().max().item() < 2e-3


def test_gemm_rope_epilogue(gen):
    Fr, B, d = 100, 3, 512
    M = B * Fr
    A = _rand((M, 256), seed=11).to(torch.bfloat16)
    W = (_rand((3 * d, 256), seed=12) / 16).to(torch.bfloat16)
    out = gen.op_gemm(A, W, out_mode=1, rope_cols=2 * d, rope_period=Fr)
    ref = (A.float() @ W.float().t()).view(B, Fr, 3, d // 64, 64)
    cos, sin = gen._dev["rope.cos"].t()[:Fr], gen._dev["rope.sin"].t()[:Fr]

    def rope(t):
        x1, x2 = t[..., :32], t[..., 32:]
        c, s = cos[None, :, None, :], sin[None, :, None, :]
        return torch.cat((x1 * c - x2 * s, x1 * s + x2 * c), -1)

    ref = torch.stack((rope(ref[:, :, 0]), rope(ref[:, :, 1]), ref[:, :, 2]), dim=2).reshape(M, 3 * d)
    assert (out - ref).abs().max().item() < 2e-3


def test_gemm_rowblock_conv_view_and_group_remap(gen):
    """Implicit-GEMM conv: K spans two consecutive row blocks, one junk row per item is dropped and
    the result lands behind the next layer's left padding."""
    Bn, Tout, blk, N, pad_next = 3, 50, 128, 64, 4
    rows = Bn * (1 + Tout)
    buf = _rand((rows, blk), seed=13).to(torch.bfloat16)            # [B, 1+Tout, blk] row blocks
    W = (_rand((N, 2 * blk), seed=14) / 16).to(torch.bfloat16)
    bias = _rand((N,), seed=15)
    out = torch.full((Bn, pad_next + Tout, N), 7.0, device="cuda", dtype=torch.bfloat16)
    gen.op_gemm(buf, W, bias=bias, act=1, out_mode=0, out=out, a_k_wrap=blk, M=rows, K=2 * blk, grp_in=1 + Tout,
                grp_valid=Tout, grp_stride=(pad_next + Tout) * N, grp_off=pad_next * N, ldo=N)
    flat = torch.cat([buf.reshape(-1), torch.zeros(blk, device="cuda", dtype=torch.bfloat16)])
    idx = torch.arange(rows, device="cuda")[:, None] * blk + torch.arange(2 * blk, device="cuda")[None, :]
    ref = gelu_tanh(flat[idx].float() @ W.float().t() + bias).view(Bn, 1 + Tout, N)[:, :Tout]
    assert (out[:, pad_next:].float() - ref).abs().max().item() < 0.03
    assert torch.all(out[:, :pad_next] == 7.0)                       # padding rows untouched


@pytest.mark.parametrize("S", [2, 4, 8])
@pytest.mark.parametrize("M,N,K", [(1, 1024, 1024), (77, 512, 512), (100, 3072, 1024), (100, 1024, 4096), (128, 16, 1024),
                                   (200, 1024, 1024), (256, 4096, 1024), (100, 72, 512)])
def test_splitk_gemm_every_epilogue(gen, S, M, N, K):
    """gemm_splitk_sm100_kernel (K split over a cluster of S CTAs, fixed-order DSMEM reduction): fp32 / bf16+GELU /
    residual / RoPE epilogues against fp32 torch, and bit-reproducible run to run."""
    A = _rand((M, K), seed=41).to(torch.bfloat16)
    W = (_rand((N, K), seed=42) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=43)
    ref = A.float() @ W.float().t() + bias
    bn = 1000 + S
    out = gen.op_gemm(A, W, bias=bias, out_mode=1, block_n=bn)
    assert (out - ref).abs().max().item() < 2e-3
    assert torch.equal(out, gen.op_gemm(A, W, bias=bias, out_mode=1, block_n=bn))
    nob = gen.op_gemm(A, W, out_mode=1, block_n=bn)
    assert (nob - (ref - bias)).abs().max().item() < 2e-3
    act = gen.op_gemm(A, W, bias=bias, act=1, out_mode=0, block_n=bn)
    assert act.dtype == torch.bfloat16 and (act.float() - gelu_tanh(ref)).abs().max().item() < 0.03
    x0 = _rand((M, N), seed=44)
    x = x0.clone()
    gen.op_gemm(A, W, bias=bias, out_mode=2, out=x, block_n=bn)
    assert (x - (x0 + ref)).abs().max().item() < 2e-3
    # against the single-accumulator kernel: same products, different fp32 summation order
    one = gen.op_gemm(A, W, bias=bias, out_mode=1, block_n=64)
    assert (out - one).abs().max().item() < 1e-4
    if N % 128 == 0:
        Fr = 50 if M % 50 == 0 else M
        r1 = gen.op_gemm(A, W, bias=bias, out_mode=1, rope_cols=N // 2, rope_period=Fr, block_n=bn)
        r0 = gen.op_gemm(A, W, bias=bias, out_mode=1, rope_cols=N // 2, rope_period=Fr, block_n=64)
        assert (r1 - r0).abs().max().item() < 1e-4
        rb = gen.op_gemm(A, W, bias=bias, out_mode=0, rope_cols=N // 2, rope_period=Fr, block_n=bn)
        assert (rb.float() - r0).abs().max().item() < 0.03


def test_splitk_gemm_conv_view_and_auto_selection(gen):
    """Row-block (implicit conv) A view + group remap through the split-K kernel, and the automatic choice: GEMMs of
    <= 256 rows take it unless small_m_split_k is switched off (then they equal the 128 x 64 single-CTA kernel bit for bit)."""
    Bn, Tout, blk, N, pad_next = 2, 100, 128, 256, 4
    rows = Bn * (1 + Tout)
    buf = _rand((rows, blk), seed=45).to(torch.bfloat16)
    W = (_rand((N, 2 * blk), seed=46) / 16).to(torch.bfloat16)
    bias = _rand((N,), seed=47)
    flat = torch.cat([buf.reshape(-1), torch.zeros(blk, device="cuda", dtype=torch.bfloat16)])
    idx = torch.arange(rows, device="cuda")[:, None] * blk + torch.arange(2 * blk, device="cuda")[None, :]
    ref = gelu_tanh(flat[idx].float() @ W.float().t() + bias).view(Bn, 1 + Tout, N)[:, :Tout]
    for bn in (1002, 1004):
        out = torch.full((Bn, pad_next + Tout, N), 7.0, device="cuda", dtype=torch.bfloat16)
        gen.op_gemm(buf, W, bias=bias, act=1, out_mode=0, out=out, a_k_wrap=blk, M=rows, K=2 * blk, grp_in=1 + Tout,
                    grp_valid=Tout, grp_stride=(pad_next + Tout) * N, grp_off=pad_next * N, ldo=N, block_n=bn)
        assert (out[:, pad_next:].float() - ref).abs().max().item() < 0.03
        assert torch.all(out[:, :pad_next] == 7.0)
    A = _rand((100, 1024), seed=48).to(torch.bfloat16)
    Wl = (_rand((1024, 1024), seed=49) / 32).to(torch.bfloat16)
    big = _rand((300, 1024), seed=50).to(torch.bfloat16)
    single = gen.op_gemm(A, Wl, out_mode=1, block_n=64)
    assert torch.equal(gen.op_gemm(A, Wl, out_mode=1), single)                 # stateless calls: batch-invariant kernels by default
    try:
        gen.set_option("small_m_split_k", 2)                                   # every GEMM of <= 256 rows
        auto = gen.op_gemm(A, Wl, out_mode=1)
        assert torch.equal(auto, gen.op_gemm(A, Wl, out_mode=1, block_n=1008))  # 16 n-tiles x 8 splits = 128 CTAs
        assert not torch.equal(auto, single) and (auto - single).abs().max().item() < 1e-4
        assert torch.equal(gen.op_gemm(big, Wl, out_mode=1), gen.op_gemm(big, Wl, out_mode=1, block_n=64))   # > 256 rows: never
        gen.set_option("small_m_split_k", 0)
        assert torch.equal(gen.op_gemm(A, Wl, out_mode=1), single)
    finally:
        gen.set_option("small_m_split_k", 1)


@pytest.mark.parametrize("M", [100, 300, 2560, 25600])
def test_fused_rmsnorm_producer_and_consumer_kernels_agree_bit_for_bit(gen, M):
    """RMSNorm folded into the GEMMs around it.  Producer (residual GEMM emitting x_new, bf16(x_new * gamma) and per-row
    sums of squares per 64 columns): the CTA-pair kernel's TMA epilogue (EPI_RESID_NORM), the generic per-thread epilogue
    (block_n 64 / 128 / 256), the split-K kernel and the standalone rowstats pass must agree BIT FOR BIT on x, xb and the
    statistics of every 64-column chunk they share, or results would depend on the batch size.  Consumer (QKV + RoPE, W1 +
    GELU): fast and generic epilogues bit-identical, and the fused pair of GEMMs equals RMSNorm-then-GEMM in fp32."""
    N, K = 1024, 1024
    A = _rand((M, K), seed=61).to(torch.bfloat16)
    W = (_rand((N, K), seed=62) / math.sqrt(K)).to(torch.bfloat16)
    bias = _rand((N,), seed=63)
    gamma = 1.0 + 0.1 * _rand((N,), seed=64)
    x0 = _rand((M, N), seed=65)
    outs = {}
    variants = [("generic64", 64), ("generic128", 128), ("generic256", 256)]
    if M >= 2560:
        variants.append(("pair", 512))
    if M <= 256:
        variants.append(("splitk8", 1008))
    for name, bn in variants:
        x = x0.clone()
        _, xb, st = gen.op_gemm_fused(A, W, bias=bias, out_mode=2, out=x, xb_gamma=gamma, block_n=bn)
        outs[name] = (x, xb, st)
    ref_x = x0 + A.float() @ W.float().t() + bias
    base = outs["generic64"]
    assert (base[0] - ref_x).abs().max().item() < 2e-3
    assert (base[1].float() - ref_x * gamma).abs().max().item() < 0.05
    assert torch.allclose(base[2].sum(1), ref_x.pow(2).sum(1), rtol=1e-4)
    for name, (x, xb, st) in outs.items():
        if name == "splitk8":                                      # other fp32 summation order: close, not identical
            assert (x - base[0]).abs().max().item() < 1e-4 and torch.allclose(st, base[2], rtol=1e-3)
            continue
        assert torch.equal(x, base[0]) and torch.equal(xb, base[1]) and torch.equal(st, base[2]), name
    xb_rs, st_rs = gen.op_rowstats(base[0], gamma)                 # the standalone first producer, same arithmetic
    assert torch.equal(xb_rs, base[1]) and torch.equal(st_rs, base[2])
    # consumers
    Wc = (_rand((3072, N), seed=66) / math.sqrt(N)).to(torch.bfloat16)
    bc = _rand((3072,), seed=67)
    eps = gen.spec.norm_eps
    normed = (base[0] * torch.rsqrt(base[0].pow(2).mean(-1, keepdim=True) + eps) * gamma)
    ref_q = normed @ Wc.float().t() + bc
    res = {}
    for name, bn in (("generic64", 64), ("generic256", 256)) + ((("pair", 512),) if M >= 2560 else ()) + ((("splitk", 1004),) if M <= 128 else ()):
        res[name] = gen.op_gemm_fused(base[1], Wc, bias=bc, out_mode=1, row_stats=base[2], block_n=bn)
    q64 = res["generic64"]
    assert (q64 - ref_q).abs().max().item() < 0.06                 # bf16(x * gamma) operand vs fp32 normalised operand
    for name, q in res.items():
        if name == "splitk":
            assert (q - q64).abs().max().item() < 1e-3
        else:
            assert torch.equal(q, q64), name
    if M >= 2560:                                                  # fast RoPE / GELU epilogues with the row scale == generic
        Fr = 100
        for kw in (dict(out_mode=0, rope_cols=2048, rope_period=Fr), dict(out_mode=0, act=1)):
            try:
                gen.set_option("fast_epilogue", 1)
                fast = gen.op_gemm_fused(base[1], Wc, bias=bc, row_stats=base[2], block_n=512, **kw)
                gen.set_option("fast_epilogue", 0)
                slow = gen.op_gemm_fused(base[1], Wc, bias=bc, row_stats=base[2], block_n=512, **kw)
            finally:
                gen.set_option("fast_epilogue", 1)
            assert torch.equal(fast, slow) and torch.equal(fast, gen.op_gemm_fused(base[1], Wc, bias=bc, row_stats=base[2], block_n=64, **kw))
        try:                                                       # and the TMA producer epilogue == the pair kernel's generic one
            gen.set_option("fast_epilogue", 0)
            x = x0.clone()
            _, xb, st = gen.op_gemm_fused(A, W, bias=bias, out_mode=2, out=x, xb_gamma=gamma, block_n=512)
        finally:
            gen.set_option("fast_epilogue", 1)
        assert torch.equal(x, base[0]) and torch.equal(xb, base[1]) and torch.equal(st, base[2])


def test_rmsnorm(gen):
    x = _rand((333, 512), scale=3.0, seed=16)
    g = 1.0 + 0.1 * _rand((512,), seed=17)
    out = gen.op_rmsnorm(x, g)
    ref = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + gen.spec.norm_eps) * g
    assert (out.float() - ref).abs().max().item() < 0.03


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("B,Fr", [(1, 100), (3, 100), (2, 37), (1, 128), (2, 300), (40, 100)])
def test_window_attention(gen, impl, B, Fr):
    d, H = gen.spec.d_model, gen.spec.n_heads
    qkv = _rand((B * Fr, 3 * d), seed=18).to(torch.bfloat16)
    out = gen.op_attention(qkv, B, Fr, impl=impl)
    t = qkv.float().view(B, Fr, 3, H, 64)
    q, k, v = (t[:, :, i].transpose(1, 2) for i in range(3))
    i = torch.arange(Fr, device="cuda")[:, None]
    j = torch.arange(Fr, device="cuda")[None, :]
    mask = (j >= i - gen.spec.window_left) & (j <= i + gen.spec.window_right)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=0.125)
    ref = ref.transpose(1, 2).reshape(B * Fr, d)
    err = (out.float() - ref).abs().max().item()
    assert err < 0.03, f"impl {impl}: max abs err {err}"


@pytest.mark.parametrize("B,Fr", [(1, 100), (5, 100), (300, 100), (2, 128), (3, 129), (2, 300), (1, 7)])
def test_attention_v4_is_bit_identical_to_v3(gen, B, Fr):
    """The default kernel (v4: staged loads, P in tensor memory, three compute slots) and the shared-memory-P kernel (v3)
    differ in scheduling and operand placement only: same masks, same ex2, same key-group sums."""
    d = gen.spec.d_model
    qkv = _rand((B * Fr, 3 * d), seed=19).to(torch.bfloat16)
    a = gen.op_attention(qkv, B, Fr, impl=0)
    assert torch.equal(a, gen.op_attention(qkv, B, Fr, impl=0))
    try:
        gen.set_option("attn_p_tmem", 0)
        c = gen.op_attention(qkv, B, Fr, impl=0)
    finally:
        gen.set_option("attn_p_tmem", 1)
    assert torch.equal(a, c)


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("M", [1, 5, 128, 1000])
def test_vq_search(gen, impl, M):
    gen.set_debug_impl(attention=0, vq=impl)
    try:
        z = _rand((M, 16), seed=19)
        codes, margin = gen.vq_search(z, return_margin=True)
    finally:
        gen.set_debug_impl(0, 0)
    cb = gen._dev["vq.codebook"].double()
    dist = cb.pow(2).sum(-1)[None] - 2.0 * z.double() @ cb.t()
    top2 = torch.topk(dist, 2, dim=-1, largest=False)
    ref_margin = (top2.values[:, 1] - top2.values[:, 0]).float()
    clear = ref_margin > 2e-3
    assert torch.equal(codes[clear], top2.indices[:, 0][clear])
    assert clear.float().mean().item() > 0.97
    assert (margin - ref_margin).abs().max().item() < 2e-3
