"""Per-launch cost of the batch-1 kernels in a dependent chain (same stream, PDL on): how much of a streaming step is
launch / prologue / dependency latency and how much is the K loop."""
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
    """fn launches its kernel REP times back to back from C (mc_set_option debug_repeat): device-side chain cost."""
    gen.set_option("debug_repeat", REP)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / REP * 1e3)
    gen.set_option("debug_repeat", 1)
    return best


def rand(shape, scale=1.0):
    return (torch.randn(shape, device="cuda") * scale)


M = 100
for name, N, K in (("QKV", 3072, 1024), ("Wo", 1024, 1024), ("W1", 4096, 1024), ("W2", 1024, 4096), ("tinyK", 1024, 64), ("proj", 16, 1024)):
    A = rand((M, K)).to(torch.bfloat16)
    Ws = [(rand((N, K)) / math.sqrt(K)).to(torch.bfloat16) for _ in range(1)]      # L2-resident weights: the chain's floor
    bias = rand((N,))
    out = torch.zeros((M, N), dtype=torch.bfloat16, device="cuda")
    res = {}
    ref = A.float() @ Ws[0].float().t() + bias
    base = gen.op_gemm(A, Ws[0], bias=bias, out_mode=1, block_n=64)
    if (K // 64) % 4 == 0:
        print(f"  kps4 bit-identical to bn64: {bool(torch.equal(gen.op_gemm(A, Ws[0], bias=bias, out_mode=1, block_n=4064), base))}; "
              f"max err vs fp32 torch {(base - ref).abs().max().item():.2e}")
    for label, bn in (("1cta_bn64", 64), ("kps4", 4064), ("splitk2", 1002), ("splitk4", 1004), ("splitk8", 1008)):
        if 1002 <= bn <= 1008 and (K // 64) % (bn - 1000) != 0:
            continue
        if 1002 <= bn <= 1008 and math.ceil(N / 64) * (bn - 1000) > 296:
            continue
        if bn == 4064 and (K // 64) % 4:
            continue
        i = [0]

        def fn():
            gen.op_gemm(A, Ws[0], bias=bias, out_mode=0, out=out, block_n=bn)
            i[0] += 1
        res[label] = timeit(fn)
    print(f"{name:6s} M={M} N={N} K={K}: " + "  ".join(f"{k} {v:.2f} us" for k, v in res.items()), flush=True)

x = rand((M, 1024))
g = rand((1024,))
print(f"rmsnorm: {timeit(lambda: gen.op_rmsnorm(x, g)):.2f} us")
qkv = rand((M, 3072)).to(torch.bfloat16)
print(f"attention: {timeit(lambda: gen.op_attention(qkv, 1, 100)):.2f} us")
gen.set_option("pdl", 0)
A = rand((M, 1024)).to(torch.bfloat16)
W = (rand((1024, 1024)) / 32).to(torch.bfloat16)
out = torch.zeros((M, 1024), dtype=torch.bfloat16, device="cuda")
print(f"Wo splitk8 without PDL: {timeit(lambda: gen.op_gemm(A, W, out_mode=0, out=out, block_n=1008)):.2f} us; "
      f"rmsnorm without PDL: {timeit(lambda: gen.op_rmsnorm(x, g)):.2f} us")
