// SIMT (CUDA-core) sliding-window attention and exact-fp32 nearest-neighbour search.
// These are the bring-up / cross-check implementations: every result of the tcgen05 kernels
// (attn_sm100.cuh, vq_sm100.cuh) is tested against them on the GPU, and the engine uses them
// for shapes the tensor-core kernels are not instantiated for.
#pragma once
#include "gemm_sm100.cuh"

namespace mc {

// ---------------------------------------------------------------------------------------------
// Sliding-window attention, head_dim 64.  qkv bf16 [B*F, 3*d] (RoPE already applied to q,k by the
// QKV GEMM epilogue); out bf16 [B*F, d].  Query i attends keys j in [i-wl, i+wr] ∩ [0,F) of its own
// window (flash-attn window_size semantics).  One warp per (row, head); lanes <-> keys for the
// scores, lanes <-> output dims for P*V.
// ---------------------------------------------------------------------------------------------
__global__ void attention_window_simt_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                             int B, int F, int H, int wl, int wr, int out_rows, float scale) {
  const int d = H * 64;
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long total = static_cast<long long>(B) * F * H;
  if (warp_global >= total) return;
  const int h = static_cast<int>(warp_global % H);
  const long long row = warp_global / H;
  const int i = static_cast<int>(row % F);
  const long long base_row = row - i;  // first row of this window
  const int first_out = F - out_rows;
  if (i < first_out) return;           // warp-uniform: this query is not kept
  const int j_lo = max(0, i - wl), j_hi = min(F - 1, i + wr);
  const int span = j_hi - j_lo + 1;

  // q: every lane holds the full 64-vector (same address across lanes -> broadcast loads)
  float q[64];
  {
    const uint4* qp = reinterpret_cast<const uint4*>(qkv + row * 3 * d + h * 64);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint4 w = __ldg(qp + u);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h2[e]);
        q[u * 8 + 2 * e] = f.x * scale;
        q[u * 8 + 2 * e + 1] = f.y * scale;
      }
    }
  }
  constexpr int MAXR = 5;  // up to 160 keys per query
  float s[MAXR];
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    const int jj = r * 32 + lane;
    s[r] = -INFINITY;
    if (jj < span) {
      const uint4* kp = reinterpret_cast<const uint4*>(qkv + (base_row + j_lo + jj) * 3 * d + d + h * 64);
      float acc = 0.0f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const uint4 w = __ldg(kp + u);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = __bfloat1622float2(h2[e]);
          acc = fmaf(q[u * 8 + 2 * e], f.x, acc);
          acc = fmaf(q[u * 8 + 2 * e + 1], f.y, acc);
        }
      }
      s[r] = acc;
    }
    mx = fmaxf(mx, s[r]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    s[r] = (r * 32 + lane < span) ? __expf(s[r] - mx) : 0.0f;
    sum += s[r];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;

  // P*V: lane owns output dims (2*lane, 2*lane+1)
  float o0 = 0.0f, o1 = 0.0f;
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if (r * 32 < span) {
      const int lim = min(32, span - r * 32);
      for (int t = 0; t < lim; ++t) {
        const float pj = __shfl_sync(0xffffffffu, s[r], t);
        const __nv_bfloat162 v2 = *reinterpret_cast<const __nv_bfloat162*>(
            qkv + (base_row + j_lo + r * 32 + t) * 3 * d + 2 * d + h * 64 + 2 * lane);
        const float2 vf = __bfloat1622float2(v2);
        o0 = fmaf(pj, vf.x, o0);
        o1 = fmaf(pj, vf.y, o1);
      }
    }
  }
  const long long orow = (base_row / F) * out_rows + (i - first_out);
  *reinterpret_cast<uint32_t*>(out + orow * d + h * 64 + 2 * lane) = pack_bf16x2(o0 * inv, o1 * inv);
}

// ---------------------------------------------------------------------------------------------
// Exact fp32 nearest neighbour.  z fp32 rows addressed through (row_map: query m -> source row
// m/keep*F + F-keep + m%keep) so only kept frames are searched.  Score = |c|^2 - 2 z.c (the
// |z|^2 term is constant per row).  Grid (row tiles of 128, codebook splits); each thread owns one
// query, codebook tiles are staged in smem and read as broadcasts.  Partials -> merge kernel.
// Ties resolve to the lowest index, like torch.argmin on the oracle.
// ---------------------------------------------------------------------------------------------
struct VqPartial {
  float best, second;
  int idx;
  int pad;
};

constexpr int VQ_SIMT_TILE = 256;

__global__ void __launch_bounds__(128)
vq_simt_kernel(const float* __restrict__ z, int Mq, int F, int keep, const float* __restrict__ cb /*[K,16]*/,
               const float* __restrict__ c2 /*[K]*/, int K, int splits, VqPartial* __restrict__ partial) {
  __shared__ float4 scb[VQ_SIMT_TILE * 4];
  __shared__ float sc2[VQ_SIMT_TILE];
  const int m = blockIdx.x * 128 + threadIdx.x;
  const int split = blockIdx.y;
  const int per_split = (K + splits - 1) / splits;
  const int k_begin = split * per_split, k_end = min(K, k_begin + per_split);
  float zr[16];
  {
    const int mm = min(m, Mq - 1);
    const long long src = static_cast<long long>(mm / keep) * F + (F - keep) + (mm % keep);
    const float4* zp = reinterpret_cast<const float4*>(z + src * 16);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 v = zp[u];
      zr[4 * u] = -2.0f * v.x; zr[4 * u + 1] = -2.0f * v.y; zr[4 * u + 2] = -2.0f * v.z; zr[4 * u + 3] = -2.0f * v.w;
    }
  }
  float best = INFINITY, second = INFINITY;
  int bidx = 0x7fffffff;
  for (int k0 = k_begin; k0 < k_end; k0 += VQ_SIMT_TILE) {
    const int n = min(VQ_SIMT_TILE, k_end - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < n * 4; i += 128) scb[i] = reinterpret_cast<const float4*>(cb + k0 * 16LL)[i];
    for (int i = threadIdx.x; i < n; i += 128) sc2[i] = c2[k0 + i];
    __syncthreads();
    for (int e = 0; e < n; ++e) {
      float acc = sc2[e];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 c = scb[e * 4 + u];
        acc = fmaf(zr[4 * u], c.x, acc);
        acc = fmaf(zr[4 * u + 1], c.y, acc);
        acc = fmaf(zr[4 * u + 2], c.z, acc);
        acc = fmaf(zr[4 * u + 3], c.w, acc);
      }
      if (acc < best) {
        second = best; best = acc; bidx = k0 + e;
      } else if (acc < second) {
        second = acc;
      }
    }
  }
  if (m < Mq) {
    VqPartial pr;
    pr.best = best; pr.second = second; pr.idx = bidx; pr.pad = 0;
    partial[static_cast<long long>(split) * Mq + m] = pr;
  }
}

// Merge per-split partials -> int64 code (+ optional top-2 margin).  Splits are ascending index
// ranges, so scanning them in order with a strict '<' keeps the lowest index on ties.
__global__ void vq_merge_kernel(const VqPartial* __restrict__ partial, int Mq, int splits,
                                long long* __restrict__ codes, float* __restrict__ margin) {
  pdl_launch_dependents();
  pdl_wait();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mq) return;
  float best = INFINITY, second = INFINITY;
  int bidx = 0;
  // eight independent loads in flight per step (a serial walk costs one L2 round trip per split: 11 us at batch 1)
  for (int s0 = 0; s0 < splits; s0 += 8) {
    VqPartial pr[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (s0 + u < splits) pr[u] = partial[static_cast<long long>(s0 + u) * Mq + m];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (s0 + u >= splits) break;
      if (pr[u].best < best) {
        second = fminf(best, pr[u].second);
        best = pr[u].best; bidx = pr[u].idx;
      } else {
        second = fminf(second, pr[u].best);
      }
    }
  }
  codes[m] = bidx;
  if (margin != nullptr) margin[m] = second - best;
}

}  // namespace mc
