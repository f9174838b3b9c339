// Weight-streaming GEMM for the batch-1 / few-rows shapes of the streaming path (M <= 256 rows: one or two 128-row
// tiles):  D[M,N] = epilogue(A[M,K] * W[N,K]^T)  with K split across a thread-block cluster.
//
// Why: a 100-row GEMM is bound by streaming W once from HBM, and with 128 x 64 tiles only N/64 CTAs (16 for N = 1024)
// exist to pull it, each walking the whole K serially at ~120 cycles per tcgen05.mma (profiles/r02_ncu_launches_stream_
// step_v0_baseline.csv: W2 23.5 us on 16 CTAs = 5 % of the HBM roofline).  Here a cluster of S CTAs shares one output
// tile: CTA r multiplies K-slice r into its own TMEM accumulator, spills the 128 x BN fp32 partial to its shared
// memory, and after one cluster barrier every CTA reduces 128/S rows of the tile over DISTRIBUTED shared memory in the
// fixed order r = 0..S-1 (bit-reproducible), applies the epilogue (bias / tanh-GELU / RoPE / residual / conv group
// remap) and stores.  Because weights do not depend on the predecessor kernel, their TMA loads are issued BEFORE
// griddepcontrol.wait (programmatic dependent launch): the weight stream of GEMM i+1 overlaps the tail of GEMM i, and
// each GEMM also pulls the NEXT GEMM's weight matrix into L2 (cp.async.bulk.prefetch).  `kps` 64-wide k-blocks share one
// full / empty barrier pair: at these tile sizes the loop is bound by the issuing thread's wait -> mma -> commit round
// trip, not by the tensor pipe (gemm_sm100.cuh, KPS).
//
// The K order of the summation differs from the single-accumulator kernels (S partial sums added in fp32), so a window
// encoded alone can differ from the same window inside a large batch in the last bit of a latent (and so in a
// near-tie code) — as with cuBLAS heuristics in the reference.  mc_set_option("small_m_split_k", 0) restores the
// batch-invariant kernels.
#pragma once
#include "gemm_sm100.cuh"

namespace mc {

template <int BN>
struct GemmSkCfg {
  static constexpr int kStages = 4;     // 97 KB: two CTAs per SM, so a dependent GEMM's CTAs can be resident (prologue done, weights
                                        // in flight) while the predecessor still runs; 8 stages / one CTA per SM measured 2x slower
  static constexpr int kStageBytesA = GEMM_BM * GEMM_BK * 2;
  static constexpr int kStageBytesB = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kPartBytes = GEMM_BM * BN * 4;          // fp32 partial tile, aliased over the (drained) stage ring
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;
  static_assert(kPartBytes <= kStages * kStageBytes, "partial tile must fit in the stage ring");
  static_assert(kSmemBytes <= 113 * 1024, "two CTAs per SM");
};

__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(cluster_addr));
  return v;
}

template <int BN, int S>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_splitk_sm100_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p,
                         const int kps /* k-blocks per barrier round trip: 1, 2 or 4; divides the CTA's k-blocks */) {
  using Cfg = GemmSkCfg<BN>;
  static_assert(BN % 64 == 0 && (S == 2 || S == 4 || S == 8), "tile / split shapes");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_full = empty_bar + Cfg::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* part = reinterpret_cast<float*>(smem);               // [128][BN] fp32, 16-byte chunks XOR-swizzled by (row & 7)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = static_cast<int>(cluster_ctarank());      // K slice of this CTA
  const int cid = blockIdx.x / S;                            // output tile of this cluster
  const int n_tiles = (p.N + BN - 1) / BN;
  const int m_blk = cid / n_tiles, n_blk = cid % n_tiles;
  const int nkb = (p.K / GEMM_BK) / S;                       // k-blocks per CTA (launcher guarantees divisibility, >= 1)
  const int kb0 = rank * nkb;
  const int n_groups = nkb / kps;                            // barrier round trips of this CTA
  const int slots = Cfg::kStages / kps;                      // groups resident in the ring

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 6) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  if (warp == 7 && lane == 0 && p.prefetch_bytes > 0) {
    // this CTA's share of the next GEMM's weight matrix -> L2 (HBM streams ahead of the dependency chain)
    constexpr long long kPiece = 32768;
    const long long pieces = (p.prefetch_bytes + kPiece - 1) / kPiece;
    for (long long i = blockIdx.x; i < pieces; i += gridDim.x) {
      const long long off = i * kPiece;
      const long long n = p.prefetch_bytes - off < kPiece ? p.prefetch_bytes - off : kPiece;
      l2_prefetch_bulk(reinterpret_cast<const uint8_t*>(p.prefetch_ptr) + off, static_cast<uint32_t>(n & ~15LL));
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // Reduce-phase identity of the epilogue threads (warps 0-3): CTA `rank` finishes rows [rank*128/S, (rank+1)*128/S) of
  // the tile, S threads per row, each owning kHalf columns in both halves of the 64-wide block (RoPE pairs j, j+32).
  static_assert(BN == 64, "one 64-column block per tile");
  constexpr int kRows = GEMM_BM / S;
  constexpr int kHalf = 32 / S;
  const int e_row = rank * kRows + (threadIdx.x & 127) / S;
  const int e_c0 = ((threadIdx.x & 127) % S) * kHalf;
  const int e_g = m_blk * GEMM_BM + e_row;                    // A row
  const int e_grp = e_g / p.grp_in;
  const int e_r = e_g - e_grp * p.grp_in;
  const bool e_live = e_g < p.M && e_r < p.grp_valid;
  const long long e_obase = static_cast<long long>(e_grp) * p.grp_stride + p.grp_off + static_cast<long long>(e_r) * p.ldo;
  const long long e_orow = (p.grp_in == INT_MAX) ? e_g : static_cast<long long>(e_grp) * p.grp_valid + e_r;
  float e_bias[2][kHalf], e_gamma[2][kHalf], e_xold[2][kHalf], e_cos[kHalf], e_sin[kHalf], e_rs = 1.0f;

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(Cfg::kStageBytes) * kps;
      const int n_pre = n_groups < slots ? n_groups : slots;
      // weights first: they do not depend on the predecessor kernel, so their loads run under its tail
      for (int g = 0; g < n_pre; ++g) {
        mbar_arrive_expect_tx(&full_bar[g], tx);
        for (int j = 0; j < kps; ++j) {
          uint8_t* sb = smem + (g * kps + j) * Cfg::kStageBytes + Cfg::kStageBytesA;
          tma_load_2d(sb, &map_b, &full_bar[g], (kb0 + g * kps + j) * GEMM_BK, n_blk * BN);
        }
      }
      pdl_wait();                                             // activations are the predecessor's output
      for (int g = 0; g < n_pre; ++g) {
        for (int j = 0; j < kps; ++j) {
          const int k_elem = (kb0 + g * kps + j) * GEMM_BK;
          const int row_off = k_elem / p.a_k_wrap;
          tma_load_2d(smem + (g * kps + j) * Cfg::kStageBytes, &map_a, &full_bar[g], k_elem - row_off * p.a_k_wrap,
                      m_blk * GEMM_BM + row_off);
        }
      }
      for (int g = n_pre; g < n_groups; ++g) {
        const int slot = g % slots;
        const uint32_t phase = (g / slots) & 1;
        mbar_wait(&empty_bar[slot], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[slot], tx);
        for (int j = 0; j < kps; ++j) {
          uint8_t* sa = smem + (slot * kps + j) * Cfg::kStageBytes;
          const int k_elem = (kb0 + g * kps + j) * GEMM_BK;
          const int row_off = k_elem / p.a_k_wrap;
          tma_load_2d(sa, &map_a, &full_bar[slot], k_elem - row_off * p.a_k_wrap, m_blk * GEMM_BM + row_off);
          tma_load_2d(sa + Cfg::kStageBytesA, &map_b, &full_bar[slot], k_elem, n_blk * BN);
        }
      }
    }
  } else if (warp == 5) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0, 0);
      for (int g = 0; g < n_groups; ++g) {
        const int slot = g % slots;
        const uint32_t phase = (g / slots) & 1;
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        for (int j = 0; j < kps; ++j) {
          const uint32_t sa = smem_u32(smem + (slot * kps + j) * Cfg::kStageBytes);
          const uint64_t adesc = umma_smem_desc_sw128(sa, 1024, 16);
          const uint64_t bdesc = umma_smem_desc_sw128(sa + Cfg::kStageBytesA, 1024, 16);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (g | j | k) != 0);
        }
        umma_commit(&empty_bar[slot]);
      }
      umma_commit(acc_full);                                  // every MMA retired: accumulator complete, stage ring drained
    }
  } else if (warp < 4) {
    pdl_wait();
    // Everything the epilogue needs from global memory is requested NOW, while the MMAs run: bias, RoPE angles, the
    // fused norm's gamma / row statistics and the old residual values do not depend on the accumulator, and an L2 round
    // trip after the reduction would sit on the dependency chain of the whole streaming step.
    if (e_live) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int q = 0; q < kHalf / 4; ++q) {
          const int n = n_blk * BN + hf * 32 + e_c0 + q * 4;
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f), gv = bv, xv = bv;
          if (n < p.N) {
            if (p.bias != nullptr) bv = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            if (p.xb_out != nullptr) gv = __ldg(reinterpret_cast<const float4*>(p.xb_gamma + n));
            if (p.out_mode == OUT_F32_RESIDUAL) xv = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + e_obase + n);
          }
          e_bias[hf][q * 4] = bv.x; e_bias[hf][q * 4 + 1] = bv.y; e_bias[hf][q * 4 + 2] = bv.z; e_bias[hf][q * 4 + 3] = bv.w;
          e_gamma[hf][q * 4] = gv.x; e_gamma[hf][q * 4 + 1] = gv.y; e_gamma[hf][q * 4 + 2] = gv.z; e_gamma[hf][q * 4 + 3] = gv.w;
          e_xold[hf][q * 4] = xv.x; e_xold[hf][q * 4 + 1] = xv.y; e_xold[hf][q * 4 + 2] = xv.z; e_xold[hf][q * 4 + 3] = xv.w;
        }
      if (p.rope_period > 0 && n_blk * BN < p.rope_cols) {
        const int pos = p.rope_offset + e_r % p.rope_period;
#pragma unroll
        for (int j = 0; j < kHalf; ++j) {
          e_cos[j] = __ldg(p.rope_cos + static_cast<long long>(e_c0 + j) * p.rope_ld + pos);
          e_sin[j] = __ldg(p.rope_sin + static_cast<long long>(e_c0 + j) * p.rope_ld + pos);
        }
      }
      if (p.row_stats != nullptr) e_rs = row_rs(p, e_g);
    }
    // ------------------------------------------- spill this CTA's partial tile
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;                         // TMEM lane = tile row
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    uint8_t* prow = reinterpret_cast<uint8_t*>(part) + row * (BN * 4);
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t raw[32];
      tmem_ld_32x32b_x32(taddr + c * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts_u4(prow + (((c * 8 + j) ^ sw) << 4), make_uint4(raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]));
    }
    tc_fence_before();
  }
  __syncwarp();
  cluster_sync_all();                                         // all S partial tiles are visible cluster-wide

  if (warp < 4) {
    // ------------------------------- reduce 128/S rows over DSMEM + epilogue
    float ss = 0.f;                                           // fused-RMSNorm producer: this thread's share of the row's sum of squares
    if (e_live) {
      const uint32_t prow = smem_u32(part) + e_row * (BN * 4);
      const int sw = e_row & 7;
      const int n0 = n_blk * BN;
      float v[2][kHalf];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int q = 0; q < kHalf / 4; ++q) {
          const int chunk = (hf * 32 + e_c0 + q * 4) >> 2;
          const uint32_t a = prow + ((chunk ^ sw) << 4);
          float4 parts[S];
#pragma unroll
          for (int s = 0; s < S; ++s) parts[s] = ld_dsmem_f4(mapa_shared(a, s));
          float4 acc = parts[0];
#pragma unroll
          for (int s = 1; s < S; ++s) { acc.x += parts[s].x; acc.y += parts[s].y; acc.z += parts[s].z; acc.w += parts[s].w; }
          v[hf][q * 4] = acc.x; v[hf][q * 4 + 1] = acc.y; v[hf][q * 4 + 2] = acc.z; v[hf][q * 4 + 3] = acc.w;
        }
      }
      if (p.row_stats != nullptr) {             // fmaf(acc, rs, bias): the consumer expression of every kernel (bias is 0 when absent)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int j = 0; j < kHalf; ++j) v[hf][j] = p.bias != nullptr ? fmaf(v[hf][j], e_rs, e_bias[hf][j]) : __fmul_rn(v[hf][j], e_rs);
      } else if (p.bias != nullptr) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int j = 0; j < kHalf; ++j) v[hf][j] += e_bias[hf][j];
      }
      if (p.act == ACT_GELU_TANH) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int j = 0; j < kHalf; ++j) v[hf][j] = gelu_tanh_f(v[hf][j]);
      }
      if (p.rope_period > 0 && n0 < p.rope_cols) {
#pragma unroll
        for (int j = 0; j < kHalf; ++j) {
          const float x1 = v[0][j], x2 = v[1][j];
          v[0][j] = x1 * e_cos[j] - x2 * e_sin[j];
          v[1][j] = x1 * e_sin[j] + x2 * e_cos[j];
        }
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
        for (int q = 0; q < kHalf / 4; ++q) {
          const int n = n0 + hf * 32 + e_c0 + q * 4;
          if (n >= p.N) continue;                             // N is a multiple of 8: whole float4 groups are in or out
          const float* vv = &v[hf][q * 4];
          if (p.out_mode == OUT_BF16) {
            uint2 w;
            w.x = pack_bf16x2(vv[0], vv[1]);
            w.y = pack_bf16x2(vv[2], vv[3]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + e_obase + n) = w;
          } else {
            float4 x = make_float4(vv[0], vv[1], vv[2], vv[3]);
            if (p.out_mode == OUT_F32_RESIDUAL) {
              const float* old = &e_xold[hf][q * 4];
              x.x += old[0]; x.y += old[1]; x.z += old[2]; x.w += old[3];
            }
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + e_obase + n) = x;
            if (p.xb_out != nullptr) {
              const float* ga = &e_gamma[hf][q * 4];
              ss = fmaf(x.x, x.x, ss); ss = fmaf(x.y, x.y, ss); ss = fmaf(x.z, x.z, ss); ss = fmaf(x.w, x.w, ss);
              uint2 w;
              w.x = pack_bf16x2(x.x * ga[0], x.y * ga[1]);
              w.y = pack_bf16x2(x.z * ga[2], x.w * ga[3]);
              *reinterpret_cast<uint2*>(p.xb_out + e_orow * p.N + n) = w;
            }
          }
        }
      }
    }
    if (p.xb_out != nullptr) {
      // the S threads of a row sit in consecutive lanes: fixed-order butterfly over them, lane cseg == 0 stores the chunk's sum
      __syncwarp();
#pragma unroll
      for (int o = S / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (e_live && (threadIdx.x % S) == 0 && n_blk * BN < p.N) p.stat_out[e_orow * ((p.N + 63) / 64) + n_blk * (BN / 64)] = ss;
    }
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                                         // peers have finished reading this CTA's partial tile
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mc
