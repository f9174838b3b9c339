"""Chain cost (200 back-to-back launches from C, PDL on) of the fused-RMSNorm PRODUCER epilogue at the decoder's in_proj
shape (M = 100, N = 1024, K = 64) and at Wo's shape, against the plain fp32 / bf16 outputs and the standalone
rowstats kernel that the producer replaces."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rca_b200_loader  # noqa: F401
import realtime_codec_agent_b200 as pkg

spec = pkg.DEFAULT_SPEC.replace(enc_layers=1, dec_layers=1)
gen = pkg.B200Generator(spec, pkg.init_random_weights(spec, seed=0), device="cuda")
REP = 200


def timeit(fn, iters=5):
    gen.set_option("debug_repeat", REP)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / REP * 1e3)
    gen.set_option("debug_repeat", 1)
    return best


M = 100
for name, N, K in (("in_proj", 1024, 64), ("Wo", 1024, 1024)):
    A = (torch.randn((M, K), device="cuda")).to(torch.bfloat16)
    W = (torch.randn((N, K), device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn((N,), device="cuda")
    gamma = torch.rand((N,), device="cuda") + 0.5
    out32 = torch.zeros((M, N), dtype=torch.float32, device="cuda")
    out16 = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    res = {}
    for label, bn in (("1cta", 64),) + ((("splitk8", 1008),) if (K // 64) % 8 == 0 else ()):
        res[f"{label} bf16"] = timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=0, out=out16, block_n=bn))
        res[f"{label} f32"] = timeit(lambda: gen.op_gemm(A, W, bias=bias, out_mode=1, out=out32, block_n=bn))
        res[f"{label} f32+producer"] = timeit(lambda: gen.op_gemm_fused(A, W, bias, out_mode=1, out=out32, xb_gamma=gamma, block_n=bn))
    print(f"{name:8s} M={M} N={N} K={K}: " + "  ".join(f"{k} {v:.2f} us" for k, v in res.items()), flush=True)
x = torch.randn((M, 1024), device="cuda")
g = torch.rand((1024,), device="cuda")
print(f"rowstats_cast: {timeit(lambda: gen.op_rowstats(x, g)):.2f} us   rmsnorm: {timeit(lambda: gen.op_rmsnorm(x, g)):.2f} us")
