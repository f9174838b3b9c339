// Codebook nearest-neighbour search as a tcgen05 "distance GEMM" fused with the argmin: the
// [queries x 131072] distance matrix lives only in TMEM / registers.
//
// argmin_k |z - c_k|^2 = argmax_k ( z.c_k - |c_k|^2 / 2 ).  Both operands are split into bf16
// hi + lo parts so the tensor core reproduces the fp32 dot product to ~2^-17 relative:
//     A row (64 bf16) = [ z_hi(16) | z_hi(16) | z_lo(16) | 1 1 1 0 ... ]
//     B row (64 bf16) = [ c_hi(16) | c_lo(16) | c_hi(16) | n1 n2 n3 0 ... ],  n1+n2+n3 = -|c|^2/2
// so one K=64 block (4 UMMA K-steps) yields score = z_hi.c_hi + z_hi.c_lo + z_lo.c_hi - |c|^2/2 in
// fp32.  B ("vq.packed", [K,64] bf16, 16 MiB -> L2 resident) is built once at load time.
//
// CTA = 128 query rows x one slice of the codebook.  warp 4: TMA producer of 256-entry B tiles;
// warp 5: MMA issuer, 128x256x64 per tile into one of two TMEM accumulators (both above the epilogue
// warps in the sub-partition arbiter's priority order); warps 0-3: epilogue,
// thread = query row, running (best, second, index) over the tile's 256 scores — a chunk of 32
// scores is skipped after one max-reduce unless it can change the top two.  This stage is
// epilogue-bound (K = 64 only), not tensor-bound; see DESIGN.md.
#pragma once
#include "attn_vq_simt.cuh"
#include "engine_common.cuh"
#include "gemm_sm100.cuh"

namespace mc {

constexpr int VQ_BN = 256;
constexpr int VQ_STAGES = 4;
constexpr int VQ_THREADS = 256;
constexpr int VQ_SMEM_A = 128 * 128;           // 16 KiB query tile
constexpr int VQ_SMEM_B = VQ_BN * 128;         // 32 KiB codebook tile
constexpr int VQ_SMEM_BYTES = VQ_SMEM_A + VQ_STAGES * VQ_SMEM_B + 256 + 1024;
constexpr int VQ_TMEM_COLS = 2 * VQ_BN;

__global__ void __launch_bounds__(VQ_THREADS, 1)
vq_argmin_sm100_kernel(const __grid_constant__ CUtensorMap map_b, const float* __restrict__ z, int Mq, int F, int keep,
                       int tiles_total, int splits, VqPartial* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + VQ_SMEM_A;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + VQ_STAGES * VQ_SMEM_B);
  uint64_t* empty_bar = full_bar + VQ_STAGES;
  uint64_t* tmem_full = empty_bar + VQ_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 128;
  const int split = blockIdx.y;
  const int per = (tiles_total + splits - 1) / splits;
  const int t_begin = split * per;
  const int t_end = min(tiles_total, t_begin + per);
  const int n_tiles = max(0, t_end - t_begin);

  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < VQ_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_mbar_init();
  }
  if (warp == 6) {
    tmem_alloc(tmem_ptr, VQ_TMEM_COLS);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  pdl_wait();   // the queries below are the previous kernel's output
  // ---- build the A tile (split-bf16 queries) directly in the swizzled K-major layout
  if (threadIdx.x < 128) {
    const int r = threadIdx.x;
    const int m = min(row0 + r, Mq - 1);
    const long long src = static_cast<long long>(m / keep) * F + (F - keep) + (m % keep);
    const float4* zp = reinterpret_cast<const float4*>(z + src * 16);
    float zv[16];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 v = zp[u];
      zv[4 * u] = v.x; zv[4 * u + 1] = v.y; zv[4 * u + 2] = v.z; zv[4 * u + 3] = v.w;
    }
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const __nv_bfloat16 h0 = __float2bfloat16_rn(zv[2 * e]), h1 = __float2bfloat16_rn(zv[2 * e + 1]);
      const float l0 = zv[2 * e] - __bfloat162float(h0), l1 = zv[2 * e + 1] - __bfloat162float(h1);
      __nv_bfloat162 hh; hh.x = h0; hh.y = h1;
      hi[e] = *reinterpret_cast<uint32_t*>(&hh);
      lo[e] = pack_bf16x2(l0, l1);
    }
    uint8_t* rowp = sA + r * 128;
    const int sw = r & 7;
    const uint4 hi0 = make_uint4(hi[0], hi[1], hi[2], hi[3]), hi1 = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    const uint4 lo0 = make_uint4(lo[0], lo[1], lo[2], lo[3]), lo1 = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    const uint32_t one2 = pack_bf16x2(1.0f, 1.0f), one1 = pack_bf16x2(1.0f, 0.0f);
    *reinterpret_cast<uint4*>(rowp + ((0 ^ sw) << 4)) = hi0;
    *reinterpret_cast<uint4*>(rowp + ((1 ^ sw) << 4)) = hi1;
    *reinterpret_cast<uint4*>(rowp + ((2 ^ sw) << 4)) = hi0;
    *reinterpret_cast<uint4*>(rowp + ((3 ^ sw) << 4)) = hi1;
    *reinterpret_cast<uint4*>(rowp + ((4 ^ sw) << 4)) = lo0;
    *reinterpret_cast<uint4*>(rowp + ((5 ^ sw) << 4)) = lo1;
    *reinterpret_cast<uint4*>(rowp + ((6 ^ sw) << 4)) = make_uint4(one2, one1, 0u, 0u);
    *reinterpret_cast<uint4*>(rowp + ((7 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], VQ_SMEM_B);
        tma_load_2d(sB + stage * VQ_SMEM_B, &map_b, &full_bar[stage], 0, (t_begin + t) * VQ_BN);
        if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, VQ_BN, 0, 0);
      const uint64_t adesc = umma_smem_desc_sw128(smem_u32(sA), 1024, 16);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (t >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint64_t bdesc = umma_smem_desc_sw128(smem_u32(sB + stage * VQ_SMEM_B), 1024, 16);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_base + acc * VQ_BN, adesc + 2 * k, bdesc + 2 * k, idesc, k != 0);
        umma_commit(&empty_bar[stage]);
        umma_commit(&tmem_full[acc]);
        if (++stage == VQ_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 4) {
    const int q = warp;
    const int m = row0 + q * 32 + lane;
    float best = -INFINITY, second = -INFINITY;
    int bidx = t_begin * VQ_BN;
    for (int t = 0; t < n_tiles; ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (t >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * VQ_BN;
      const int k_base = (t_begin + t) * VQ_BN;
#pragma unroll 1
      for (int c0 = 0; c0 < VQ_BN; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(taddr + c0, raw);
        tmem_ld_wait();
        float cm = fmaxf(__uint_as_float(raw[0]), __uint_as_float(raw[1]));
#pragma unroll
        for (int j = 2; j < 32; j += 2) cm = fmax3(cm, __uint_as_float(raw[j]), __uint_as_float(raw[j + 1]));
        if (cm > second) {  // this chunk can change the top two: exact ascending scan
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(raw[j]);
            if (v > best) {
              second = best; best = v; bidx = k_base + c0 + j;
            } else if (v > second) {
              second = v;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (m < Mq) {
      VqPartial pr;  // convert to the distance scale of the SIMT kernel: d = |c|^2 - 2 z.c = -2 * score
      pr.best = -2.0f * best; pr.second = -2.0f * second; pr.idx = bidx; pr.pad = 0;
      partial[static_cast<long long>(split) * Mq + m] = pr;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc(tmem_base, VQ_TMEM_COLS);
  }
}

inline size_t vq_sm100_scratch_bytes(int num_sms, int Mq) { return (size_t)148 * Mq * sizeof(VqPartial); }

inline int launch_vq_sm100(mc_handle* h, const float* z, int n_items, int F, int keep, int64_t* codes, float* margin,
                           void* scratch, cudaStream_t stream) {
  const mc_spec& s = h->spec;
  const int Mq = n_items * keep;
  const int tiles_total = s.codebook_size / VQ_BN;
  const int row_tiles = (Mq + 127) / 128;
  int splits = (h->num_sms + row_tiles - 1) / row_tiles;
  splits = std::max(1, std::min(std::min(148, tiles_total), splits));   // one row tile (a streaming frame): the whole chip scans the codebook
  splits = (tiles_total + (tiles_total + splits - 1) / splits - 1) / ((tiles_total + splits - 1) / splits);   // no empty split
  const CUtensorMap* mb = nullptr;
  MC_TRY(mc_internal::get_map_2d_bf16(h, h->ptr<bf16>("vq.packed"), 64, (uint64_t)s.codebook_size, 64, VQ_BN, &mb));
  MC_TRY(mc_allow_smem(h, vq_argmin_sm100_kernel, VQ_SMEM_BYTES));
  mc_launch(h, vq_argmin_sm100_kernel, dim3(row_tiles, splits), dim3(VQ_THREADS), VQ_SMEM_BYTES, stream, *mb, z, Mq, F, keep,
            tiles_total, splits, reinterpret_cast<VqPartial*>(scratch));
  MC_LAUNCH_CHECK(h, "vq_argmin_sm100_kernel");
  mc_launch(h, vq_merge_kernel, dim3((Mq + 255) / 256), dim3(256), 0, stream, reinterpret_cast<const VqPartial*>(scratch), Mq,
            splits, reinterpret_cast<long long*>(codes), margin);
  MC_LAUNCH_CHECK(h, "vq_merge_kernel");
  return MC_OK;
}

}  // namespace mc
