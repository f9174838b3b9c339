"""Bring-up diagnostics (GPU): python tools/diag.py {gemm|attn|vq|e2e} — prints error maps instead
of asserting, one process per stage so that a trapped kernel cannot hide the other stages."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg


def make_gen(spec=None):
    spec = spec or pkg.TINY_SPEC
    return pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda", max_positions=1024)


def summarize(name, out, ref):
    err = (out.float() - ref.float()).abs()
    print(f"{name}: max_err={err.max().item():.5f} mean_err={err.mean().item():.6f} ref_absmax={ref.abs().max().item():.3f} "
          f"nan={int(torch.isnan(out.float()).sum())}")
    if err.max().item() > 0.05:
        bad_rows = (err.max(dim=1).values > 0.05).nonzero().flatten()
        bad_cols = (err.max(dim=0).values > 0.05).nonzero().flatten()
        print(f"   bad rows {bad_rows.numel()}/{out.shape[0]} first {bad_rows[:12].tolist()} | bad cols {bad_cols.numel()}/{out.shape[1]} first {bad_cols[:12].tolist()}")
        print("   out[0,:8]", out[0, :8].float().tolist())
        print("   ref[0,:8]", ref[0, :8].float().tolist())


def diag_gemm():
    gen = make_gen()
    torch.manual_seed(0)
    # identity probe: W = I (N=64,K=64) -> out must equal A
    A = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
    W = torch.eye(64, device="cuda").to(torch.bfloat16)
    out = gen.op_gemm(A, W, out_mode=1, block_n=64)
    torch.cuda.synchronize()
    summarize("identity 128x64x64 bn64", out, A.float())
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256), (128, 64, 256, 64), (256, 256, 512, 256),
                          (100, 192, 128, 64), (1000, 3072, 512, 0), (25600, 1024, 1024, 0)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        out = gen.op_gemm(A, W, out_mode=1, block_n=bn)
        torch.cuda.synchronize()
        summarize(f"gemm M{M} N{N} K{K} bn{bn}", out, A.float() @ W.float().t())
    for (M, N, K) in [(256, 256, 64), (256, 256, 256), (300, 512, 128), (25600, 1024, 1024)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        out = gen.op_gemm(A, W, out_mode=1, block_n=512)
        torch.cuda.synchronize()
        summarize(f"PAIR gemm M{M} N{N} K{K}", out, A.float() @ W.float().t())
    # timing of the big one
    A = torch.randn(25600, 1024, device="cuda").to(torch.bfloat16)
    W = (torch.randn(4096, 1024, device="cuda") / 32).to(torch.bfloat16)
    out = torch.empty(25600, 4096, device="cuda", dtype=torch.bfloat16)
    for bn in (128, 256, 512):
        for _ in range(3):
            gen.op_gemm(A, W, out_mode=0, out=out, block_n=bn)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            gen.op_gemm(A, W, out_mode=0, out=out, block_n=bn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"gemm 25600x4096x1024 bn{bn}: {ms:.3f} ms  {2 * 25600 * 4096 * 1024 / ms / 1e9:.1f} TFLOP/s")
    ref = torch.empty_like(out)
    for _ in range(3):
        torch.matmul(A, W.t(), out=ref)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        torch.matmul(A, W.t(), out=ref)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"cublas same shape: {ms:.3f} ms  {2 * 25600 * 4096 * 1024 / ms / 1e9:.1f} TFLOP/s")


def diag_epi():
    """Isolated timing of the K=1024 GEMM shapes of one transformer block under each epilogue mode."""
    gen = make_gen(pkg.DEFAULT_SPEC)
    torch.manual_seed(0)
    M = 25600

    def timeit(fn, flops, label):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"{label:58s} {ms*1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s")

    for (N, K) in [(3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        of = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        fl = 2.0 * M * N * K
        timeit(lambda: gen.op_gemm(A, W, out_mode=0, out=ob), fl, f"N{N} K{K} bf16 out")
        timeit(lambda: gen.op_gemm(A, W, act=0x100, out_mode=0, out=ob), fl, f"N{N} K{K} EXPERIMENT tmem drain only, no stores")
        timeit(lambda: gen.op_gemm(A, W, act=0x200, out_mode=0, out=ob), fl, f"N{N} K{K} EXPERIMENT mainloop only")
        timeit(lambda: gen.op_gemm(A, W, bias=bias, act=0x101, out_mode=0, out=ob), fl, f"N{N} K{K} EXPERIMENT drain + bias + gelu, no stores")
        timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=0, out=ob), fl, f"N{N} K{K} bf16 out + bias")
        timeit(lambda: gen.op_gemm(A, W, bias=bias, act=1, out_mode=0, out=ob), fl, f"N{N} K{K} bf16 out + bias + gelu")
        if N == 3072:
            timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=0, out=ob, rope_cols=2048, rope_period=100), fl,
                   f"N{N} K{K} bf16 out + bias + rope")
        timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=1, out=of), fl, f"N{N} K{K} f32 out + bias")
        timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=2, out=of), fl, f"N{N} K{K} f32 residual + bias")
        timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=0, out=ob, block_n=256), fl, f"N{N} K{K} bf16 out + bias, single-CTA bn256")
        ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        timeit(lambda: torch.matmul(A, W.t(), out=ref), fl, f"N{N} K{K} cuBLAS")


def diag_epi_cycles():
    gen = make_gen(pkg.DEFAULT_SPEC)
    M, N, K = 25600, 4096, 1024
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") / 32).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    ob = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    of = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    for label, kw in [("mainloop only", dict(act=0x600)), ("drain only", dict(act=0x500)), ("bf16 store", dict(act=0x400)),
                      ("bias+gelu no store", dict(act=0x501, bias=bias)), ("bias+gelu bf16 store", dict(act=0x401, bias=bias)),
                      ("bias bf16 store", dict(act=0x400, bias=bias))]:
        for _ in range(2):
            gen.op_gemm(A, W, out_mode=0, out=ob, **{k: v for k, v in kw.items() if k != "act"}, act=kw["act"] & ~0x400)
        torch.cuda.synchronize()
        print("----", label, flush=True)
        gen.op_gemm(A, W, out_mode=0, out=ob, **kw)
        torch.cuda.synchronize()
    print("---- f32 residual + bias", flush=True)
    gen.op_gemm(A, W, bias=bias, out_mode=2, out=of, act=0x400)
    torch.cuda.synchronize()


def diag_attn():
    gen = make_gen()
    spec = gen.spec
    d, H = spec.d_model, spec.n_heads
    torch.manual_seed(1)
    for impl in (1, 0):
        for (B, Fr) in [(1, 100), (2, 37), (2, 300)]:
            qkv = torch.randn(B * Fr, 3 * d, device="cuda").to(torch.bfloat16)
            out = gen.op_attention(qkv, B, Fr, impl=impl)
            torch.cuda.synchronize()
            t = qkv.float().view(B, Fr, 3, H, 64)
            q, k, v = (t[:, :, i].transpose(1, 2) for i in range(3))
            i = torch.arange(Fr, device="cuda")[:, None]
            j = torch.arange(Fr, device="cuda")[None, :]
            mask = (j >= i - spec.window_left) & (j <= i + spec.window_right)
            ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, scale=0.125)
            summarize(f"attn impl{impl} B{B} F{Fr}", out, ref.transpose(1, 2).reshape(B * Fr, d))


def diag_vq():
    gen = make_gen()
    torch.manual_seed(2)
    z = torch.randn(300, 16, device="cuda")
    cb = gen._dev["vq.codebook"].double()
    dist = cb.pow(2).sum(-1)[None] - 2.0 * z.double() @ cb.t()
    top2 = torch.topk(dist, 2, dim=-1, largest=False)
    for impl in (1, 0):
        gen.set_debug_impl(0, impl)
        codes, margin = gen.vq_search(z, return_margin=True)
        torch.cuda.synchronize()
        agree = (codes == top2.indices[:, 0]).float().mean().item()
        merr = (margin.double() - (top2.values[:, 1] - top2.values[:, 0])).abs().max().item()
        print(f"vq impl{impl}: agree={agree:.4f} margin_err={merr:.6f} codes[:8]={codes[:8].tolist()} ref={top2.indices[:8, 0].tolist()}")


def diag_e2e():
    import __graft_entry__ as g
    g.smoke()


def diag_block_gemms():
    """The four GEMMs of one transformer block at M = 25 600 and 102 400: specialised (pipelined) epilogue vs the
    generic one vs cuBLAS (plain GEMM, no epilogue), 20 back-to-back launches each."""
    gen = make_gen(pkg.DEFAULT_SPEC)
    torch.manual_seed(0)

    def timeit(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    for M in (25600, 102400):
        for (name, N, K, kw) in [("QKV + bias + RoPE", 3072, 1024, dict(out_mode=0, rope_cols=2048, rope_period=100)),
                                 ("Wo  + bias, fp32 residual", 1024, 1024, dict(out_mode=2)),
                                 ("W1  + bias + GELU", 4096, 1024, dict(out_mode=0, act=1)),
                                 ("W2  + bias, fp32 residual", 1024, 4096, dict(out_mode=2))]:
            A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
            W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
            bias = torch.randn(N, device="cuda")
            out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if kw["out_mode"] == 0 else torch.float32)
            fl = 2.0 * M * N * K
            res = {}
            for tag, fast in (("specialised", 1), ("generic", 0)):
                gen.set_option("fast_epilogue", fast)
                ms = timeit(lambda: gen.op_gemm(A, W, bias=bias, out=out, **kw))
                res[tag] = ms
            gen.set_option("fast_epilogue", 1)
            ref = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            res["cuBLAS (no epilogue)"] = timeit(lambda: torch.matmul(A, W.t(), out=ref))
            print(f"M{M} {name:28s} " + "  ".join(f"{k}: {v * 1e3:7.1f} us {fl / v / 1e9:6.0f} TF/s" for k, v in res.items()))


def diag_skinny():
    """Batch-1 GEMM shapes (M = 100): time per launch for each tile width, 300 back-to-back launches."""
    gen = make_gen(pkg.DEFAULT_SPEC)
    torch.manual_seed(0)
    M = 100
    for (N, K) in [(3072, 1024), (1024, 1024), (4096, 1024), (1024, 4096), (1024, 5120)]:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
        res = []
        for bn in (64, 128, 256):
            for _ in range(10):
                gen.op_gemm(A, W, bias=bias, out=out, out_mode=0, block_n=bn)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(300):
                gen.op_gemm(A, W, bias=bias, out=out, out_mode=0, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            res.append(f"bn{bn}: {e0.elapsed_time(e1) / 300 * 1e3:6.1f} us")
        print(f"M{M} N{N} K{K}  weights {N * K * 2 / 1e6:5.1f} MB  " + "  ".join(res))


if __name__ == "__main__":
    {"skinny": diag_skinny, "block": diag_block_gemms, "gemm": diag_gemm, "attn": diag_attn, "vq": diag_vq, "e2e": diag_e2e, "epi": diag_epi, "cyc": diag_epi_cycles}[sys.argv[1]]()
