// HBM-bound stages of the codec: first strided conv (Cin = 1), RMSNorm, code-embedding gather,
// last transposed conv (Cout = 1).  All are one pass over their input with 128-bit accesses;
// none has data reuse beyond what L1 gives for free, so no shared-memory tiling is used
// except for the tiny weight tables.
#pragma once
#include "gemm_sm100.cuh"

namespace mc {

// ---------------------------------------------------------------------------------------------
// First encoder conv: wav fp32 [B rows, row stride ld, T valid samples] -> bf16 channels-last
// [B, pad_rows + T0, C0], causal kernel 2*s0 / stride s0, bias + tanh-GELU fused.
// One thread = one output frame x 8 channels (one 16-byte store).  Samples >= T read as zero
// (this IS pad_audio: right padding to the hop multiple, audio_tokenizer.py:190).
// ---------------------------------------------------------------------------------------------
template <int KT>  // taps = 2 * stride
__global__ void __launch_bounds__(256)
conv_first_kernel(const float* __restrict__ wav, long long ld, int T, int B, int T0, int s0, int C0,
                  const float* __restrict__ w /*[KT, C0]*/, const float* __restrict__ bias,
                  __nv_bfloat16* __restrict__ out, int pad_rows) {
  pdl_launch_dependents();
  pdl_wait();
  // A thread keeps ONE group of 8 output channels for its whole life: its KT x 8 weights and 8 biases
  // live in registers, so the inner loop has no shared-memory traffic at all.
  const int cgroups = C0 / 8;
  const int cg = threadIdx.x % cgroups;
  const int slot = threadIdx.x / cgroups;
  const int frames_per_block = blockDim.x / cgroups;
  float wr[KT][8], br[8];
#pragma unroll
  for (int j = 0; j < KT; ++j) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + j * C0 + cg * 8));
    const float4 c = __ldg(reinterpret_cast<const float4*>(w + j * C0 + cg * 8 + 4));
    wr[j][0] = a.x; wr[j][1] = a.y; wr[j][2] = a.z; wr[j][3] = a.w;
    wr[j][4] = c.x; wr[j][5] = c.y; wr[j][6] = c.z; wr[j][7] = c.w;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) br[c] = __ldg(bias + cg * 8 + c);
  const long long total = static_cast<long long>(B) * T0;
  for (long long bt = blockIdx.x * static_cast<long long>(frames_per_block) + slot; bt < total;
       bt += static_cast<long long>(gridDim.x) * frames_per_block) {
    const int t = static_cast<int>(bt % T0);
    const int b = static_cast<int>(bt / T0);
    const float* x = wav + b * ld;
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = br[c];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const int n = t * s0 + j - s0;
      const float xv = (n >= 0 && n < T) ? __ldg(x + n) : 0.0f;
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(xv, wr[j][c], acc[c]);
    }
    uint4 o;
    o.x = pack_bf16x2(gelu_tanh_f(acc[0]), gelu_tanh_f(acc[1]));
    o.y = pack_bf16x2(gelu_tanh_f(acc[2]), gelu_tanh_f(acc[3]));
    o.z = pack_bf16x2(gelu_tanh_f(acc[4]), gelu_tanh_f(acc[5]));
    o.w = pack_bf16x2(gelu_tanh_f(acc[6]), gelu_tanh_f(acc[7]));
    __nv_bfloat16* dst = out + ((static_cast<long long>(b) * (pad_rows + T0) + pad_rows + t) * C0 + cg * 8);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm: x fp32 [M, d] -> bf16 rows, fp32 statistics.  One warp per row, float4 loads.
// Output row remap (grp_in/grp_out/grp_off) lets the decoder's final norm write straight into
// the left-padded input buffer of the first transposed conv.
// ---------------------------------------------------------------------------------------------
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                               __nv_bfloat16* __restrict__ out, int M, int d, float eps, int grp_in,
                               long long grp_stride, long long grp_off) {
  pdl_launch_dependents();
  pdl_wait();
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * d);
    const int n4 = d >> 2;
    float ss = 0.0f;
    for (int i = lane; i < n4; i += 32) {
      const float4 v = xr[i];
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(ss / static_cast<float>(d) + eps);
    const int grp = row / grp_in;
    const long long obase = static_cast<long long>(grp) * grp_stride + grp_off + static_cast<long long>(row - grp * grp_in) * d;
    uint2* orow_p = reinterpret_cast<uint2*>(out + obase);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    for (int i = lane; i < n4; i += 32) {
      const float4 v = xr[i];  // second read hits L1/L2
      const float4 g = __ldg(g4 + i);
      uint2 o;
      o.x = pack_bf16x2(v.x * inv * g.x, v.y * inv * g.y);
      o.y = pack_bf16x2(v.z * inv * g.z, v.w * inv * g.w);
      orow_p[i] = o;
    }
  }
}

// Register-resident variant for d = NV * 128: a lane keeps its NV float4 of the row, so x is read from HBM exactly
// once (all NV loads in flight together) and the statistics and the scaling both come from registers.  Summation
// order as in rmsnorm_kernel (per-lane partial sums over i = lane, lane+32, ..., then the xor tree).
template <int NV>
__global__ void __launch_bounds__(256)
rmsnorm_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ out, int M,
                    float eps, int grp_in, long long grp_stride, long long grp_off) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int d = NV * 128;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  float4 g[NV];
#pragma unroll
  for (int u = 0; u < NV; ++u) g[u] = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * u);
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < M; row += gridDim.x * warps_per_block) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * d);
    float4 v[NV];
#pragma unroll
    for (int u = 0; u < NV; ++u) v[u] = xr[lane + 32 * u];
    float ss = 0.0f;
#pragma unroll
    for (int u = 0; u < NV; ++u) ss += v[u].x * v[u].x + v[u].y * v[u].y + v[u].z * v[u].z + v[u].w * v[u].w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(ss / static_cast<float>(d) + eps);
    const int grp = row / grp_in;
    const long long obase = static_cast<long long>(grp) * grp_stride + grp_off + static_cast<long long>(row - grp * grp_in) * d;
    uint2* orow_p = reinterpret_cast<uint2*>(out + obase);
#pragma unroll
    for (int u = 0; u < NV; ++u) {
      uint2 o;
      o.x = pack_bf16x2(v[u].x * inv * g[u].x, v[u].y * inv * g[u].y);
      o.y = pack_bf16x2(v[u].z * inv * g[u].z, v[u].w * inv * g[u].w);
      orow_p[lane + 32 * u] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Keep the last r_out of r_in fp32 rows of every window (dead-output elimination between layers).
// ---------------------------------------------------------------------------------------------
__global__ void compact_rows_kernel(const float4* __restrict__ x, float4* __restrict__ y, int B, int r_in, int r_out,
                                    int d4) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * r_out * d4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4);
    const long long row = idx / d4;
    const int j = static_cast<int>(row % r_out);
    const long long b = row / r_out;
    y[idx] = x[(b * r_in + (r_in - r_out) + j) * d4 + c];
  }
}

// ---------------------------------------------------------------------------------------------
// Shared convolution stem: x[b, t, :] = stem[b*fstep + t, :] for t in [pre, F) — the frames of window b
// that do not depend on where the window starts, taken from the one pass over the whole span.
// ---------------------------------------------------------------------------------------------
__global__ void gather_stem_kernel(const float4* __restrict__ stem, float4* __restrict__ x, int B, int F, int pre,
                                   int fstep, int d4) {
  pdl_launch_dependents();
  pdl_wait();
  const int per = F - pre;
  const long long total = static_cast<long long>(B) * per * d4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4);
    const long long row = idx / d4;
    const int t = pre + static_cast<int>(row % per);
    const long long b = row / per;
    x[(b * F + t) * d4 + c] = __ldg(stem + (b * fstep + t) * d4 + c);
  }
}

// ---------------------------------------------------------------------------------------------
// Code embedding: codes int64 [M] -> bf16 [M, 64] = (projected codebook row, zero padded to the
// 64-wide K block of the decoder's input projection).  One thread = one 16-byte store.
// ---------------------------------------------------------------------------------------------
__global__ void embed_codes_kernel(const long long* __restrict__ codes, const float* __restrict__ table /*[K,16]*/,
                                   int K, __nv_bfloat16* __restrict__ out, long long M) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = M * 8;  // 8 x 16-byte pieces per 64-wide row
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = idx >> 3;
    const int piece = static_cast<int>(idx & 7);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (piece < 2) {
      long long code = codes[m];
      code = code < 0 ? 0 : (code >= K ? K - 1 : code);
      const float4* src = reinterpret_cast<const float4*>(table + code * 16 + piece * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      o.x = pack_bf16x2(a.x, a.y);
      o.y = pack_bf16x2(a.z, a.w);
      o.z = pack_bf16x2(b.x, b.y);
      o.w = pack_bf16x2(b.z, b.w);
    }
    *reinterpret_cast<uint4*>(out + m * 64 + piece * 8) = o;
  }
}

// Latent rows fp32 [M,16] -> bf16 [M,64] (decoder entered with z_q instead of codes).
__global__ void pack_latents_kernel(const float* __restrict__ z /*[M,16]*/, __nv_bfloat16* __restrict__ out,
                                    long long M) {
  const long long total = M * 8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = idx >> 3;
    const int piece = static_cast<int>(idx & 7);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (piece < 2) {
      const float4* src = reinterpret_cast<const float4*>(z + m * 16 + piece * 8);
      const float4 a = src[0], b = src[1];
      o.x = pack_bf16x2(a.x, a.y);
      o.y = pack_bf16x2(a.z, a.w);
      o.z = pack_bf16x2(b.x, b.y);
      o.w = pack_bf16x2(b.z, b.w);
    }
    *reinterpret_cast<uint4*>(out + m * 64 + piece * 8) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Last decoder transposed conv (Cout = 1): bf16 channels-last [B, 1 + Tin, Cin] (row 0 of each
// item is the zero left pad) -> fp32 waveform.  out[t*s + j] = b + x[t].w[:, j] + x[t-1].w[:, j+s].
// Only the last `keep` samples of each item are stored (audio_tokenizer.py:141-144), at
// out[b*keep + (n - (Tin*s - keep))].  One thread = one input frame = s output samples.
// ---------------------------------------------------------------------------------------------
template <int MAXS>
__global__ void tconv_last_kernel(const __nv_bfloat16* __restrict__ x, int B, int Tin, int Cin, int s,
                                  const float* __restrict__ w /*[Cin, 2*s]*/, const float* __restrict__ bias,
                                  float* __restrict__ out, int keep) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sw[];  // [Cin * 2*s]
  for (int i = threadIdx.x; i < Cin * 2 * s; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const float bias0 = __ldg(bias);
  const long long total = static_cast<long long>(B) * Tin;
  const int first_kept = Tin * s - keep;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(idx % Tin);
    const int b = static_cast<int>(idx / Tin);
    if ((t + 1) * s <= first_kept) continue;
    const __nv_bfloat16* cur = x + (static_cast<long long>(b) * (Tin + 1) + 1 + t) * Cin;
    const __nv_bfloat16* prev = cur - Cin;
    float acc[MAXS];
#pragma unroll
    for (int j = 0; j < MAXS; ++j) acc[j] = bias0;
    for (int c0 = 0; c0 < Cin; c0 += 8) {
      const uint4 cu = *reinterpret_cast<const uint4*>(cur + c0);
      const uint4 pu = *reinterpret_cast<const uint4*>(prev + c0);
      const __nv_bfloat162* c2 = reinterpret_cast<const __nv_bfloat162*>(&cu);
      const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&pu);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float2 cf = __bfloat1622float2(c2[u]);
        const float2 pf = __bfloat1622float2(p2[u]);
        const float* w0 = sw + (c0 + 2 * u) * 2 * s;
        const float* w1 = w0 + 2 * s;
#pragma unroll
        for (int j = 0; j < MAXS; ++j) {
          if (j < s) {
            acc[j] = fmaf(cf.x, w0[j], acc[j]);
            acc[j] = fmaf(pf.x, w0[j + s], acc[j]);
            acc[j] = fmaf(cf.y, w1[j], acc[j]);
            acc[j] = fmaf(pf.y, w1[j + s], acc[j]);
          }
        }
      }
    }
    float* o = out + static_cast<long long>(b) * keep;
#pragma unroll
    for (int j = 0; j < MAXS; ++j) {
      const int n = t * s + j - first_kept;
      if (j < s && n >= 0) o[n] = acc[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Left padding of the causal conv / transposed-conv inputs: `rows` zero rows of `width` bytes at the head of every
// item, for up to MC_MAX_CONVS buffers in ONE launch (four cudaMemset2D nodes cost ~2 us each on the batch-1 path).
// ---------------------------------------------------------------------------------------------
struct PadList {
  int n = 0;
  unsigned char* base[8];
  long long pitch[8];      // bytes between items
  int width[8];            // bytes to zero at the head of each item (multiple of 16)
};

__global__ void __launch_bounds__(256)
zero_pads_kernel(const PadList pl, int items) {
  pdl_launch_dependents();
  pdl_wait();                // the previous pass may still be reading these buffers
  for (int b = 0; b < pl.n; ++b) {
    const int chunks = pl.width[b] >> 4;
    const long long total = static_cast<long long>(items) * chunks;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const long long item = i / chunks;
      const int c = static_cast<int>(i - item * chunks);
      *reinterpret_cast<uint4*>(pl.base[b] + item * pl.pitch[b] + (static_cast<long long>(c) << 4)) = make_uint4(0, 0, 0, 0);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// First producer of a stack whose RMSNorms live in GEMM epilogues (GemmParams: fused RMSNorm): x fp32 [M, d] ->
// xb = bf16(x * gamma) [M, d] and stats[M, d/64] = per-row, per-64-column sums of x^2, with exactly the arithmetic of
// the GEMM producers (sequential fmaf chain over the chunk, __fmul_rn before the bf16 rounding), so that a stack entered
// through this kernel and one entered through a producing GEMM agree bit for bit.  One thread = one 64-column chunk.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rowstats_cast_kernel(const float* __restrict__ x, const float* __restrict__ gamma, __nv_bfloat16* __restrict__ xb,
                     float* __restrict__ stats, long long M, int d) {
  pdl_launch_dependents();
  pdl_wait();
  const int nst = d >> 6;
  const long long total = M * nst;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / nst;
    const int c = static_cast<int>(i - row * nst);
    const float4* src = reinterpret_cast<const float4*>(x + row * d + c * 64);
    const float4* gsrc = reinterpret_cast<const float4*>(gamma + c * 64);
    float v[64];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 a = src[j];
      v[4 * j] = a.x; v[4 * j + 1] = a.y; v[4 * j + 2] = a.z; v[4 * j + 3] = a.w;
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) ss = fmaf(v[j], v[j], ss);
    uint4* dst = reinterpret_cast<uint4*>(xb + row * d + c * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 g0 = __ldg(gsrc + 2 * j), g1 = __ldg(gsrc + 2 * j + 1);
      uint4 w;
      w.x = pack_bf16x2(__fmul_rn(v[8 * j], g0.x), __fmul_rn(v[8 * j + 1], g0.y));
      w.y = pack_bf16x2(__fmul_rn(v[8 * j + 2], g0.z), __fmul_rn(v[8 * j + 3], g0.w));
      w.z = pack_bf16x2(__fmul_rn(v[8 * j + 4], g1.x), __fmul_rn(v[8 * j + 5], g1.y));
      w.w = pack_bf16x2(__fmul_rn(v[8 * j + 6], g1.z), __fmul_rn(v[8 * j + 7], g1.w));
      dst[j] = w;
    }
    stats[i] = ss;
  }
}

}  // namespace mc
