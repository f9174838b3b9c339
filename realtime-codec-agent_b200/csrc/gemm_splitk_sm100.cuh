// Weight-streaming GEMM for the batch-1 / few-rows shapes of the streaming path (M <= 256 rows: one or two 128-row
// tiles):  D[M,N] = epilogue(A[M,K] * W[N,K]^T)  with K split across a thread-block cluster.
//
// Why: a 100-row GEMM is bound by streaming W once from HBM, and with 128 x 64 tiles only N/64 CTAs (16 for N = 1024)
// exist to pull it, each walking the whole K serially at ~120 cycles per tcgen05.mma (profiles/r02_ncu_launches_stream_
// step_v0_baseline.csv: W2 23.5 us on 16 CTAs = 5 % of the HBM roofline).  Here a cluster of S CTAs shares one output
// tile: CTA r multiplies K-slice r into its own TMEM accumulator, spills the 128 x BN fp32 partial to its shared
// memory, and after one cluster barrier every CTA reduces 128/S rows of the tile over DISTRIBUTED shared memory in the
// fixed order r = 0..S-1 (bit-reproducible), applies the epilogue (bias / tanh-GELU / RoPE / residual / conv group
// remap) and stores.  N/64 x S CTAs of 97 KB shared memory each: two fit on an SM, so with programmatic dependent
// launch the NEXT GEMM's CTAs are already resident while this one finishes — and because weights do not depend on the
// predecessor, their TMA loads are issued BEFORE griddepcontrol.wait: the weight stream of GEMM i+1 overlaps the tail
// of GEMM i.
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
  static constexpr int kStages = 4;
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
gemm_splitk_sm100_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmParams p) {
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

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int n_pre = nkb < Cfg::kStages ? nkb : Cfg::kStages;
      // weights first: they do not depend on the predecessor kernel, so their loads run under its tail
      for (int i = 0; i < n_pre; ++i) {
        uint8_t* sb = smem + i * Cfg::kStageBytes + Cfg::kStageBytesA;
        mbar_arrive_expect_tx(&full_bar[i], Cfg::kStageBytes);
        tma_load_2d(sb, &map_b, &full_bar[i], (kb0 + i) * GEMM_BK, n_blk * BN);
      }
      pdl_wait();                                             // activations are the predecessor's output
      for (int i = 0; i < n_pre; ++i) {
        const int k_elem = (kb0 + i) * GEMM_BK;
        const int row_off = k_elem / p.a_k_wrap;
        tma_load_2d(smem + i * Cfg::kStageBytes, &map_a, &full_bar[i], k_elem - row_off * p.a_k_wrap, m_blk * GEMM_BM + row_off);
      }
      for (int i = n_pre; i < nkb; ++i) {
        const int stage = i % Cfg::kStages;
        const uint32_t phase = (i / Cfg::kStages) & 1;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        const int k_elem = (kb0 + i) * GEMM_BK;
        const int row_off = k_elem / p.a_k_wrap;
        tma_load_2d(sa, &map_a, &full_bar[stage], k_elem - row_off * p.a_k_wrap, m_blk * GEMM_BM + row_off);
        tma_load_2d(sa + Cfg::kStageBytesA, &map_b, &full_bar[stage], k_elem, n_blk * BN);
      }
    }
  } else if (warp == 5) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, BN, 0, 0);
      for (int i = 0; i < nkb; ++i) {
        const int stage = i % Cfg::kStages;
        const uint32_t phase = (i / Cfg::kStages) & 1;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint64_t adesc = umma_smem_desc_sw128(sa, 1024, 16);
        const uint64_t bdesc = umma_smem_desc_sw128(sa + Cfg::kStageBytesA, 1024, 16);
#pragma unroll
        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
        umma_commit(&empty_bar[stage]);
      }
      umma_commit(acc_full);                                  // every MMA retired: accumulator complete, stage ring drained
    }
  } else if (warp < 4) {
    // ------------------------------------------- spill this CTA's partial tile
    pdl_wait();
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
    constexpr int kRows = GEMM_BM / S;                        // rows of the tile this CTA finishes
    constexpr int kHalf = 32 / S;                             // columns per thread in each half of a 64-wide block (4, 8, 16)
    const int t = threadIdx.x;                                // 0..127
    const int row = rank * kRows + t / S;
    const int cseg = t % S;
    const int g = m_blk * GEMM_BM + row;                      // A row
    const int grp = g / p.grp_in;
    const int r = g - grp * p.grp_in;
    if (g < p.M && r < p.grp_valid) {
      const long long obase = static_cast<long long>(grp) * p.grp_stride + p.grp_off + static_cast<long long>(r) * p.ldo;
      const uint32_t prow = smem_u32(part) + row * (BN * 4);
      const int sw = row & 7;
      const bool do_rope = p.rope_period > 0;
      const int pos = do_rope ? p.rope_offset + r % p.rope_period : 0;
#pragma unroll
      for (int blk = 0; blk < BN / 64; ++blk) {
        const int n0 = n_blk * BN + blk * 64;
        if (n0 >= p.N) break;
        float v[2][kHalf];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int q = 0; q < kHalf / 4; ++q) {
            const int chunk = (blk * 64 + hf * 32 + cseg * kHalf + q * 4) >> 2;
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
        const int c0 = cseg * kHalf;                          // first column (within a half) owned by this thread
        if (p.bias != nullptr) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int j = 0; j < kHalf; ++j) {
              const int n = n0 + hf * 32 + c0 + j;
              if (n < p.N) v[hf][j] += __ldg(p.bias + n);
            }
        }
        if (p.act == ACT_GELU_TANH) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int j = 0; j < kHalf; ++j) v[hf][j] = gelu_tanh_f(v[hf][j]);
        }
        if (do_rope && n0 < p.rope_cols) {
#pragma unroll
          for (int j = 0; j < kHalf; ++j) {
            const float cs = __ldg(p.rope_cos + static_cast<long long>(c0 + j) * p.rope_ld + pos);
            const float sn = __ldg(p.rope_sin + static_cast<long long>(c0 + j) * p.rope_ld + pos);
            const float x1 = v[0][j], x2 = v[1][j];
            v[0][j] = x1 * cs - x2 * sn;
            v[1][j] = x1 * sn + x2 * cs;
          }
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
          for (int q = 0; q < kHalf / 4; ++q) {
            const int n = n0 + hf * 32 + c0 + q * 4;
            if (n >= p.N) continue;                           // N is a multiple of 8: whole float4 groups are in or out
            const float* vv = &v[hf][q * 4];
            if (p.out_mode == OUT_BF16) {
              uint2 w;
              w.x = pack_bf16x2(vv[0], vv[1]);
              w.y = pack_bf16x2(vv[2], vv[3]);
              *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + obase + n) = w;
            } else {
              float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + obase + n);
              float4 x = make_float4(vv[0], vv[1], vv[2], vv[3]);
              if (p.out_mode == OUT_F32_RESIDUAL) {
                const float4 old = *o;
                x.x += old.x; x.y += old.y; x.z += old.z; x.w += old.w;
              }
              *o = x;
            }
          }
        }
      }
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
