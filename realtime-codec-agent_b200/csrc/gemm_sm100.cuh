// Persistent warp-specialised bf16 GEMM for sm_100a:  D[M,N] = epilogue(A[M,K] * W[N,K]^T)
//
//   warps 0..3  epilogue         (tcgen05.ld -> bias / tanh-GELU / RoPE / residual -> smem -> TMA store)
//   warp 4      TMA producer     (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 5      MMA issuer       (one thread: tcgen05.mma.kind::f16, 128 x BN x 16 per CTA)
//   warp 6      TMEM allocator   (2 x BN fp32 accumulator columns, double buffered)
// The producer and the MMA issuer carry the HIGHEST warp ids on purpose: the SM sub-partition
// arbiter prefers higher warp ids, and measured with clock64 an ALU-busy epilogue warp sharing a
// sub-partition with a lower-numbered MMA thread delayed its tcgen05.mma issue by ~25 % per tile.
//
// The accumulator of tile i drains while the tensor pipe already works on tile i+1.
//
// Implicit-GEMM convolutions use the same kernel: A is addressed as a 2-D "row block" view
// [rows, a_k_wrap]; K runs over `K / a_k_wrap` consecutive row blocks, i.e. k-block kb loads
// box (k % a_k_wrap, row + k / a_k_wrap).  A causal conv with kernel 2*stride over a
// channels-last buffer [.., stride*Cin] is then K = 2*stride*Cin with a_k_wrap = stride*Cin, and
// a causal transposed conv is K = 2*Cin with a_k_wrap = Cin — no im2col buffer exists anywhere.
// Rows are grouped (grp_in rows per batch item, of which grp_valid are real) so that the one
// row per item that straddles the next item's left padding is computed but never stored.
#pragma once
#include <type_traits>

#include "ptx_sm100.cuh"

namespace mc {

enum : int { ACT_NONE = 0, ACT_GELU_TANH = 1 };
enum : int { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_RESIDUAL = 2 };

struct GemmParams {
  int M, N, K;
  int a_k_wrap;
  const float* bias;  // [N] or nullptr
  int act;
  int out_mode;
  void* out;
  long long ldo;  // elements between consecutive output rows
  int grp_in, grp_valid;          // rows per batch item in A / how many of them are real
  long long grp_stride, grp_off;  // output element offset = grp*grp_stride + grp_off + r*ldo
  int tma_store;  // 1: epilogue stages rows in smem and writes with TMA (store / reduce-add); 0: per-thread stores
  const float* rope_cos;  // transposed table [32][rope_ld] or nullptr
  int rope_ld;            // positions per table row (max_positions)
  const float* rope_sin;
  int rope_cols;    // RoPE applies to output columns [0, rope_cols), 64-wide heads
  int rope_period;  // position = rope_offset + (row within group) % rope_period
  int rope_offset;
  int one = 1;      // always 1; opaque to the compiler (see gemm_epilogue_slab_fast)
  // split-K kernel only: the NEXT GEMM's weights, pulled into L2 while this kernel runs (they do not depend on it)
  const void* prefetch_ptr = nullptr;
  long long prefetch_bytes = 0;
  // RMSNorm folded into the GEMMs around it (few-rows path; single-CTA generic epilogue and split-K kernel):
  //   consumer  y = acc * rs(row) + bias,  rs = rsqrt(sum_c row_stats[row][c] / norm_dim + eps)   (A = bf16(x * gamma))
  //   producer  (fp32 outputs) also emits xb = bf16(x_new * gamma) as dense rows [rows, N] and, per row and per
  //             64-column chunk, the sum of x_new^2 -> stat_out[row][N / 64]
  const float* row_stats = nullptr;
  int row_stats_n = 0;
  float norm_eps = 0.f, inv_norm_dim = 0.f;
  __nv_bfloat16* xb_out = nullptr;
  const float* xb_gamma = nullptr;
  float* stat_out = nullptr;
};

// rs of one row from its per-chunk partial sums (fixed order)
__device__ __forceinline__ float row_rs(const GemmParams& p, int row) {
  const float* st = p.row_stats + static_cast<long long>(row) * p.row_stats_n;
  float s = 0.f;
  // eight loads in flight, then added in index order: a load -> add -> load chain is one L2 round trip per chunk (measured
  // ~2 us for the 16 chunks of d = 1024, profiles/r02_stream_fusion_experiments.log); the sum is bit-identical
  for (int c0 = 0; c0 < p.row_stats_n; c0 += 8) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = (c0 + j < p.row_stats_n) ? __ldg(st + c0 + j) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j < p.row_stats_n) s += t[j];
  }
  return rsqrtf(fmaf(s, p.inv_norm_dim, p.norm_eps));
}

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 256;
constexpr int kEpiBufBytes = 4096;  // 32 rows x 128 B: one TMA store box per epilogue warp and buffer

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = GEMM_BM * GEMM_BK * 2;
  static constexpr int kStageBytesB = BN * GEMM_BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kEpiBytes = 4 * 2 * kEpiBufBytes;     // 4 epilogue warps x double buffer
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + 1024;  // + alignment slack
};

// tanh-GELU as 5 FP32 ops + one MUFU.TANH:  u = x * (k0 + k0*k1*x^2),  y = hx + hx*tanh(u), hx = x/2.
// Every kernel (GEMM epilogues of all variants, the first conv) uses this one function, so an activation
// computed through different tilings is bit-identical.
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float x2 = x * x;
  const float inner = fmaf(x2, 0.7978845608028654f * 0.044715f, 0.7978845608028654f);
  const float u = x * inner;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Per-thread epilogue state that lives in registers across tiles.
struct EpiRegs {
  float bias[64];          // bias of the chunk about to be drained (prefetched one chunk ahead)
  float rc[32], rs[32];    // cos / sin of this thread's row position (reloaded only when the row changes)
  int rope_pos = -1;
  int buf_sel = 0;
};

template <int BN>
__device__ __forceinline__ void epi_load_bias(const GemmParams& p, int n0, float (&b)[64]) {
#pragma unroll
  for (int j = 0; j < 64; j += 4) {
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + j < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
    b[j] = b4.x; b[j + 1] = b4.y; b[j + 2] = b4.z; b[j + 3] = b4.w;
  }
}

// Drains one 128-row x BN accumulator slab: rows row_base + [0,128), columns n_base + [0,BN).
// Called by the 4 epilogue warps (q = TMEM lane quarter); shared by the 1-CTA and 2-CTA kernels.
// Everything that needs global memory (bias of the first chunk, RoPE angles of the row) is requested
// BEFORE waiting for the accumulator, and each later chunk's bias is requested one chunk ahead, so no
// L2 round trip sits between tcgen05.ld and the store (L1 is ~0 KB here: smem takes the whole 228 KB).
template <int BN, int RPW = 32>   // RPW: accumulator rows per epilogue warp (32 for 128-row tiles, 16 for 64-row tiles: lanes 0..15)
__device__ __forceinline__ void gemm_epilogue_slab(const GemmParams& p, const CUtensorMap* map_out_p, uint32_t tmem_acc,
                                                   int row_base, int n_base, int q, int lane, uint8_t* my_bufs,
                                                   EpiRegs& st, uint64_t* ready_bar, uint32_t ready_parity) {
  const int g = row_base + q * RPW + (lane < RPW ? lane : 0);  // A row handled by this thread
  const int grp = g / p.grp_in;
  const int r = g - grp * p.grp_in;
  const bool row_ok = (lane < RPW) && (g < p.M) && (r < p.grp_valid);
  const long long obase = static_cast<long long>(grp) * p.grp_stride + p.grp_off + static_cast<long long>(r) * p.ldo;
  const bool has_bias = p.bias != nullptr;
  if (has_bias) epi_load_bias<BN>(p, n_base, st.bias);
  const float rs = (p.row_stats != nullptr && g < p.M) ? row_rs(p, g) : 1.0f;
  if (p.rope_period > 0) {
    const int pos = p.rope_offset + r % p.rope_period;
    if (pos != st.rope_pos) {
      // transposed tables [32][rope_ld]: consecutive lanes = consecutive positions -> coalesced
      st.rope_pos = pos;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        st.rc[j] = __ldg(p.rope_cos + static_cast<long long>(j) * p.rope_ld + pos);
        st.rs[j] = __ldg(p.rope_sin + static_cast<long long>(j) * p.rope_ld + pos);
      }
    }
  }
  mbar_wait(ready_bar, ready_parity);  // accumulator complete
  tc_fence_after();

#pragma unroll 1
  for (int c = 0; c < BN / 64; ++c) {
    const int n0 = n_base + c * 64;
    if (n0 >= p.N) break;  // uniform across the warp
    uint32_t raw0[32], raw1[32];
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + c * 64;
    tmem_ld_32x32b_x32(taddr, raw0);
    tmem_ld_32x32b_x32(taddr + 32, raw1);
    tmem_ld_wait();
    float v[64];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = __uint_as_float(raw0[j]);
      v[32 + j] = __uint_as_float(raw1[j]);
    }
    // fused-RMSNorm consumer: ONE expression in every kernel (fmaf(acc, rs, bias)) so that results do not depend on
    // which kernel a batch size selects
    if (p.row_stats != nullptr && has_bias) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = fmaf(v[j], rs, st.bias[j]);
      if (c + 1 < BN / 64 && n0 + 64 < p.N) epi_load_bias<BN>(p, n0 + 64, st.bias);
    } else if (p.row_stats != nullptr) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = __fmul_rn(v[j], rs);
    } else if (has_bias) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] += st.bias[j];
      if (c + 1 < BN / 64 && n0 + 64 < p.N) epi_load_bias<BN>(p, n0 + 64, st.bias);  // next chunk, in flight during this one
    }
    if (p.act == ACT_GELU_TANH) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = gelu_tanh_f(v[j]);
    }
    if (p.rope_period > 0 && n0 < p.rope_cols) {
      // one 64-wide head per chunk: rotate (j, j+32) by the angle of (row position, j)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x1 = v[j], x2 = v[32 + j];
        v[j] = x1 * st.rc[j] - x2 * st.rs[j];
        v[32 + j] = x1 * st.rs[j] + x2 * st.rc[j];
      }
    }
    if (p.tma_store) {
      // Stage this warp's 32 rows in smem (128-byte rows, 16-byte chunks XOR-swizzled by row so
      // the row-per-thread writes are bank-conflict free), then one TMA store / reduce-add per
      // box: global writes are full 128-byte lines and the fp32 residual is never read back.
      const int row0 = row_base + q * 32;
      const int sw = lane & 7;
      if (p.out_mode == OUT_BF16) {
        uint8_t* buf = my_bufs + st.buf_sel * kEpiBufBytes;
        if (lane == 0) tma_wait_group_read<1>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 w;
          w.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
          w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
          w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
          w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ sw) << 4)) = w;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(map_out_p, buf, n0, row0);
          tma_commit_group();
        }
        st.buf_sel ^= 1;
      } else {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (n0 + 32 * half < p.N) {  // uniform
            uint8_t* buf = my_bufs + st.buf_sel * kEpiBufBytes;
            if (lane == 0) tma_wait_group_read<1>();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float* vv = v + 32 * half + 4 * j;
              *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ sw) << 4)) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (p.out_mode == OUT_F32_RESIDUAL) tma_reduce_add_2d(map_out_p, buf, n0 + 32 * half, row0);
              else tma_store_2d(map_out_p, buf, n0 + 32 * half, row0);
              tma_commit_group();
            }
            st.buf_sel ^= 1;
          }
        }
      }
    } else if (row_ok) {
      if (p.out_mode == OUT_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + obase + n0;
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          if (n0 + j < p.N) {
            uint4 w;
            w.x = pack_bf16x2(v[j], v[j + 1]);
            w.y = pack_bf16x2(v[j + 2], v[j + 3]);
            w.z = pack_bf16x2(v[j + 4], v[j + 5]);
            w.w = pack_bf16x2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(o + j) = w;
          }
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out) + obase + n0;
        if (p.xb_out != nullptr) {
          // producer of a fused RMSNorm: x_new (fp32) + bf16(x_new * gamma) + this chunk's sum of squares.  All loads
          // first (gamma, old residual): interleaved with the stores they would serialise on one L2 round trip each.
          const long long orow = (p.grp_in == INT_MAX) ? g : static_cast<long long>(grp) * p.grp_valid + r;
          __nv_bfloat16* xb = p.xb_out + orow * p.N + n0;
          float gam[64];
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            float4 ga = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + j < p.N) ga = __ldg(reinterpret_cast<const float4*>(p.xb_gamma + n0 + j));
            gam[j] = ga.x; gam[j + 1] = ga.y; gam[j + 2] = ga.z; gam[j + 3] = ga.w;
          }
          if (p.out_mode == OUT_F32_RESIDUAL) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              if (n0 + j < p.N) {
                const float4 a = *reinterpret_cast<const float4*>(o + j);
                v[j] = __fadd_rn(v[j], a.x); v[j + 1] = __fadd_rn(v[j + 1], a.y); v[j + 2] = __fadd_rn(v[j + 2], a.z); v[j + 3] = __fadd_rn(v[j + 3], a.w);
              }
            }
          }
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 64; j += 8) {
            if (n0 + j < p.N) {
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              *reinterpret_cast<float4*>(o + j + 4) = make_float4(v[j + 4], v[j + 5], v[j + 6], v[j + 7]);
#pragma unroll
              for (int e = 0; e < 8; ++e) ss = fmaf(v[j + e], v[j + e], ss);
              uint4 w;
              w.x = pack_bf16x2(__fmul_rn(v[j], gam[j]), __fmul_rn(v[j + 1], gam[j + 1]));
              w.y = pack_bf16x2(__fmul_rn(v[j + 2], gam[j + 2]), __fmul_rn(v[j + 3], gam[j + 3]));
              w.z = pack_bf16x2(__fmul_rn(v[j + 4], gam[j + 4]), __fmul_rn(v[j + 5], gam[j + 5]));
              w.w = pack_bf16x2(__fmul_rn(v[j + 6], gam[j + 6]), __fmul_rn(v[j + 7], gam[j + 7]));
              *reinterpret_cast<uint4*>(xb + j) = w;
            }
          }
          p.stat_out[orow * ((p.N + 63) / 64) + (n0 >> 6)] = ss;
        } else if (p.out_mode == OUT_F32_RESIDUAL) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            if (n0 + j < p.N) {
              float4 x = *reinterpret_cast<const float4*>(o + j);
              x.x += v[j]; x.y += v[j + 1]; x.z += v[j + 2]; x.w += v[j + 3];
              *reinterpret_cast<float4*>(o + j) = x;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            if (n0 + j < p.N) {
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Specialised epilogues for the three GEMMs that carry the transformer (QKV + RoPE, W1 + GELU, Wo / W2
// residual): same arithmetic as gemm_epilogue_slab, restructured so that the epilogue of a K = 1024 tile
// fits under its main loop:
//   * the mode is a template parameter (no dead register ranges: the generic version keeps bias, RoPE
//     angles and both raw halves live at once and sits at the 255-register cap);
//   * the tile's 256 bias values are staged once per tile in shared memory (1 KB, shared by the four
//     epilogue warps, guarded by a 128-thread named barrier) instead of 64 registers refilled from L2
//     per chunk;
//   * TMEM loads are software-pipelined: the tcgen05.ld of chunk c+1 is in flight while chunk c is
//     computed and stored (measured with clock64: a drain-only epilogue spent ~380 cycles per
//     tcgen05.ld.x32 + wait, none of it overlapped).
// Requirements (checked by the launcher): TMA-store output, bias present, N a multiple of BN.
// ---------------------------------------------------------------------------------------------
enum : int { EPI_GENERIC = 0, EPI_ROPE_BF16 = 1, EPI_GELU_BF16 = 2, EPI_RESID_F32 = 3, EPI_RESID_NORM = 4 };

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Orders every later use of v[] after the preceding tcgen05.wait::ld (asm volatile statements keep their order;
// this one "modifies" the registers, so the compiler cannot hoist a consumer above the wait).
__device__ __forceinline__ void reg_fence32(uint32_t (&v)[32]) {
  asm volatile(""
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                 "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                 "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                 "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}

struct EpiFastRegs {
  float rc[32], rs[32];    // RoPE variant only (dead otherwise)
  int rope_pos = -1;
  int buf_sel = 0;
  float row_scale = 1.0f;  // fused-RMSNorm consumer: rs of this thread's row for the current tile
  uint32_t xphase = 0;     // EPI_RESID_NORM: parity bits of the two x_old staging barriers
};

__device__ __forceinline__ float4 lds_f4(const float* smem_ptr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(smem_ptr)));
  return v;
}
__device__ __forceinline__ void sts_u4(void* smem_ptr, uint4 w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(smem_ptr)), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}

template <int EPI>
__device__ __forceinline__ void epi_fast_chunk(const GemmParams& p, const CUtensorMap* map_out_p, uint32_t (&lo)[32],
                                               uint32_t (&hi)[32], const float* sb /*this chunk's 64 biases (smem)*/,
                                               int n0, int row0, int lane, uint8_t* my_bufs, EpiFastRegs& st) {
  float v[64];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    v[j] = __uint_as_float(lo[j]);
    v[32 + j] = __uint_as_float(hi[j]);
  }
  // `p.one` is always 1, but the compiler cannot know: the branch puts the arithmetic in its own basic block, so
  // ptxas cannot hoist it above the tcgen05.ld of the NEXT chunk that the caller has just issued (without it the
  // load was sunk below all 64 activations and its latency overlapped only the 8 shared-memory stores).
  if (p.one != 0) {
#pragma unroll
    if (p.row_stats != nullptr) {     // fused-RMSNorm consumer (uniform): fmaf(acc, rs, bias), the expression of every kernel
      const float rsc = st.row_scale;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b0 = lds_f4(sb + j);
        const float4 b1 = lds_f4(sb + 32 + j);
        v[j] = fmaf(v[j], rsc, b0.x); v[j + 1] = fmaf(v[j + 1], rsc, b0.y); v[j + 2] = fmaf(v[j + 2], rsc, b0.z); v[j + 3] = fmaf(v[j + 3], rsc, b0.w);
        v[32 + j] = fmaf(v[32 + j], rsc, b1.x); v[32 + j + 1] = fmaf(v[32 + j + 1], rsc, b1.y);
        v[32 + j + 2] = fmaf(v[32 + j + 2], rsc, b1.z); v[32 + j + 3] = fmaf(v[32 + j + 3], rsc, b1.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b0 = lds_f4(sb + j);
        const float4 b1 = lds_f4(sb + 32 + j);
        v[j] += b0.x; v[j + 1] += b0.y; v[j + 2] += b0.z; v[j + 3] += b0.w;
        v[32 + j] += b1.x; v[32 + j + 1] += b1.y; v[32 + j + 2] += b1.z; v[32 + j + 3] += b1.w;
      }
    }
    if (EPI == EPI_GELU_BF16) {
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = gelu_tanh_f(v[j]);
    }
    if (EPI == EPI_ROPE_BF16) {
      if (n0 < p.rope_cols) {   // warp-uniform: q and k heads rotate, v heads do not
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x1 = v[j], x2 = v[32 + j];
          v[j] = x1 * st.rc[j] - x2 * st.rs[j];
          v[32 + j] = x1 * st.rs[j] + x2 * st.rc[j];
        }
      }
    }
  }
  const int sw = lane & 7;
  if (EPI == EPI_RESID_F32) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint8_t* buf = my_bufs + st.buf_sel * kEpiBufBytes;
      if (lane == 0) tma_wait_group_read<1>();
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* vv = v + 32 * half + 4 * j;
        sts_u4(buf + lane * 128 + ((j ^ sw) << 4), make_uint4(__float_as_uint(vv[0]), __float_as_uint(vv[1]), __float_as_uint(vv[2]),
                                                               __float_as_uint(vv[3])));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(map_out_p, buf, n0 + 32 * half, row0);
        tma_commit_group();
      }
      st.buf_sel ^= 1;
    }
  } else {
    uint8_t* buf = my_bufs + st.buf_sel * kEpiBufBytes;
    if (lane == 0) tma_wait_group_read<1>();
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 w;
      w.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
      w.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
      w.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
      w.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
      sts_u4(buf + lane * 128 + ((j ^ sw) << 4), w);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(map_out_p, buf, n0, row0);
      tma_commit_group();
    }
    st.buf_sel ^= 1;
  }
}

template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_slab_fast(const GemmParams& p, const CUtensorMap* map_out_p, uint32_t tmem_acc,
                                                        int row_base, int n_base, int q, int lane, uint8_t* my_bufs,
                                                        float* sbias /*[BN] shared by the 4 epilogue warps*/, EpiFastRegs& st,
                                                        uint64_t* ready_bar, uint32_t ready_parity) {
  static_assert(BN % 128 == 0, "fast epilogue: chunk pairs");
  // stage the tile's bias: all four warps are past their reads of the previous tile's values, then each writes a quarter
  named_bar_sync(1, 128);
  {
    const int i = q * 32 + lane;                      // 128 threads x 2 floats (BN = 256) or 1 float (BN = 128)
#pragma unroll
    for (int u = 0; u < BN / 128; ++u) {
      const float bv = __ldg(p.bias + n_base + i * (BN / 128) + u);
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(sbias + i * (BN / 128) + u)), "f"(bv) : "memory");
    }
  }
  if (p.row_stats != nullptr) {
    const int g = row_base + q * 32 + lane;
    st.row_scale = g < p.M ? row_rs(p, g) : 1.0f;
  }
  if (EPI == EPI_ROPE_BF16) {
    if (n_base < p.rope_cols) {
      const int g = row_base + q * 32 + lane;
      const int r = g % p.grp_in;                     // grp_in = INT_MAX for plain GEMMs: r = g
      const int pos = p.rope_offset + r % p.rope_period;
      if (pos != st.rope_pos) {
        st.rope_pos = pos;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          st.rc[j] = __ldg(p.rope_cos + static_cast<long long>(j) * p.rope_ld + pos);
          st.rs[j] = __ldg(p.rope_sin + static_cast<long long>(j) * p.rope_ld + pos);
        }
      }
    }
  }
  named_bar_sync(1, 128);
  mbar_wait(ready_bar, ready_parity);  // accumulator complete
  tc_fence_after();

  const int row0 = row_base + q * 32;
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  uint32_t a_lo[32], a_hi[32], b_lo[32], b_hi[32];
  tmem_ld_32x32b_x32(tbase, a_lo);
  tmem_ld_32x32b_x32(tbase + 32, a_hi);
  tmem_ld_wait();
  reg_fence32(a_lo);
  reg_fence32(a_hi);
#pragma unroll
  for (int c = 0; c < BN / 64; c += 2) {
    // chunk c from the A registers while chunk c+1 streams into B
    tmem_ld_32x32b_x32(tbase + (c + 1) * 64, b_lo);
    tmem_ld_32x32b_x32(tbase + (c + 1) * 64 + 32, b_hi);
    reg_fence32(a_lo);   // pins the math on chunk c BEHIND the loads just issued (the compiler would otherwise
    reg_fence32(a_hi);   // hoist it above them, and the tcgen05.ld latency would be exposed again)
    epi_fast_chunk<EPI>(p, map_out_p, a_lo, a_hi, sbias + c * 64, n_base + c * 64, row0, lane, my_bufs, st);
    tmem_ld_wait();
    reg_fence32(b_lo);
    reg_fence32(b_hi);
    if (c + 2 < BN / 64) {
      tmem_ld_32x32b_x32(tbase + (c + 2) * 64, a_lo);
      tmem_ld_32x32b_x32(tbase + (c + 2) * 64 + 32, a_hi);
    }
    reg_fence32(b_lo);
    reg_fence32(b_hi);
    epi_fast_chunk<EPI>(p, map_out_p, b_lo, b_hi, sbias + (c + 1) * 64, n_base + (c + 1) * 64, row0, lane, my_bufs, st);
    if (c + 2 < BN / 64) {
      tmem_ld_wait();
      reg_fence32(a_lo);
      reg_fence32(a_hi);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// EPI_RESID_NORM: the residual GEMMs (Wo, W2) as PRODUCERS of the next RMSNorm (offline path).  The plain residual
// epilogue adds into x with cp.reduce.async.bulk and never sees x; here the old residual tile comes IN through TMA
// (fp32 boxes of 32 x 32, two per 64-column chunk, double-buffered per epilogue warp), the thread forms
// x_new = (acc + bias) + x_old for its row, accumulates the chunk's sum of squares, writes x_new back IN PLACE and
// bf16(x_new * gamma) beside it, and both leave through TMA stores; stat_out[row][N/64] gets the chunk sum.  The
// consumer GEMM (QKV / W1) then reads bf16(x * gamma) as its A operand and scales its accumulator rows by
// rsqrt(sum / d + eps): the 110 us / 629 MB rmsnorm pass between the two GEMMs disappears.  Arithmetic is expression
// for expression that of gemm_epilogue_slab's producer branch (tested bit-identical), so results do not depend on which
// kernel a batch size selects.  The epilogue of tile i runs under the main loop of tile i+1 (two TMEM accumulators), so
// its TMA round trips are hidden as long as it is shorter than a main loop (Wo: ~9 us per tile, W2: ~17 us).
// Shared memory per warp: 4 x 4 KB x_old / x_new boxes + one 4 KB bf16 box; the kernel runs 4 pipeline stages.
// ---------------------------------------------------------------------------------------------
template <int BN>
__device__ __forceinline__ void gemm_epilogue_slab_resid_norm(const GemmParams& p, const CUtensorMap* map_x, const CUtensorMap* map_xb,
                                                              uint32_t tmem_acc, int row_base, int n_base, int q, int lane,
                                                              uint8_t* wbuf /*20 KB of this warp*/, uint64_t* xbar /*[2] of this warp*/,
                                                              float* sbias /*[2*BN]: bias | gamma*/, EpiFastRegs& st,
                                                              uint64_t* ready_bar, uint32_t ready_parity) {
  static_assert(BN % 128 == 0, "chunk pairs");
  named_bar_sync(1, 128);
  {
    const int i = q * 32 + lane;
#pragma unroll
    for (int u = 0; u < BN / 128; ++u) {
      const int n = i * (BN / 128) + u;
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(sbias + n)), "f"(__ldg(p.bias + n_base + n)) : "memory");
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_u32(sbias + BN + n)), "f"(__ldg(p.xb_gamma + n_base + n)) : "memory");
    }
  }
  const int row0 = row_base + q * 32;
  const int g = row0 + lane;
  uint8_t* xb0 = wbuf;                    // x boxes: [set][half] 4 KB each
  uint8_t* bf0 = wbuf + 4 * kEpiBufBytes; // one bf16 box
  if (lane == 0) {
    tma_wait_group_read<0>();             // the previous tile's stores have been read out of these buffers
#pragma unroll
    for (int set = 0; set < 2; ++set) {
      mbar_arrive_expect_tx(&xbar[set], 2 * kEpiBufBytes);
      tma_load_2d(xb0 + (set * 2) * kEpiBufBytes, map_x, &xbar[set], n_base + set * 64, row0);
      tma_load_2d(xb0 + (set * 2 + 1) * kEpiBufBytes, map_x, &xbar[set], n_base + set * 64 + 32, row0);
    }
  }
  named_bar_sync(1, 128);
  mbar_wait(ready_bar, ready_parity);     // accumulator complete
  tc_fence_after();
  const uint32_t tbase = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
  const int sw = lane & 7;
  const int nst = p.N >> 6;
#pragma unroll 1
  for (int c = 0; c < BN / 64; ++c) {
    const int set = c & 1;
    const int n0 = n_base + c * 64;
    uint32_t lo[32], hi[32];
    tmem_ld_32x32b_x32(tbase + c * 64, lo);
    tmem_ld_32x32b_x32(tbase + c * 64 + 32, hi);
    tmem_ld_wait();
    mbar_wait(&xbar[set], (st.xphase >> set) & 1u);
    st.xphase ^= (1u << set);
    float v[64];
    const float* sb = sbias + c * 64;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint8_t* xbox = xb0 + (set * 2 + hf) * kEpiBufBytes + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 xo = lds_f4(reinterpret_cast<const float*>(xbox + ((j ^ sw) << 4)));
        const float4 bb = lds_f4(sb + hf * 32 + 4 * j);
        const uint32_t* a = (hf == 0 ? lo : hi) + 4 * j;
        float* vv = v + hf * 32 + 4 * j;
        vv[0] = __fadd_rn(__fadd_rn(__uint_as_float(a[0]), bb.x), xo.x);
        vv[1] = __fadd_rn(__fadd_rn(__uint_as_float(a[1]), bb.y), xo.y);
        vv[2] = __fadd_rn(__fadd_rn(__uint_as_float(a[2]), bb.z), xo.z);
        vv[3] = __fadd_rn(__fadd_rn(__uint_as_float(a[3]), bb.w), xo.w);
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) ss = fmaf(v[j], v[j], ss);
    if (g < p.M) p.stat_out[static_cast<long long>(g) * nst + (n0 >> 6)] = ss;
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      uint8_t* xbox = xb0 + (set * 2 + hf) * kEpiBufBytes + lane * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float* vv = v + hf * 32 + 4 * j;
        sts_u4(xbox + ((j ^ sw) << 4), make_uint4(__float_as_uint(vv[0]), __float_as_uint(vv[1]), __float_as_uint(vv[2]), __float_as_uint(vv[3])));
      }
    }
    if (lane == 0) tma_wait_group_read<0>();   // the previous chunk's bf16 store has left the (single) bf16 box
    __syncwarp();
    {
      uint8_t* bbox = bf0 + lane * 128;
      const float* sg = sbias + BN + c * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 g0 = lds_f4(sg + 8 * j), g1 = lds_f4(sg + 8 * j + 4);
        uint4 w;
        w.x = pack_bf16x2(__fmul_rn(v[8 * j], g0.x), __fmul_rn(v[8 * j + 1], g0.y));
        w.y = pack_bf16x2(__fmul_rn(v[8 * j + 2], g0.z), __fmul_rn(v[8 * j + 3], g0.w));
        w.z = pack_bf16x2(__fmul_rn(v[8 * j + 4], g1.x), __fmul_rn(v[8 * j + 5], g1.y));
        w.w = pack_bf16x2(__fmul_rn(v[8 * j + 6], g1.z), __fmul_rn(v[8 * j + 7], g1.w));
        sts_u4(bbox + ((j ^ sw) << 4), w);
      }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(map_x, xb0 + (set * 2) * kEpiBufBytes, n0, row0);
      tma_store_2d(map_x, xb0 + (set * 2 + 1) * kEpiBufBytes, n0 + 32, row0);
      tma_store_2d(map_xb, bf0, n0, row0);
      tma_commit_group();
      if (c + 2 < BN / 64) {              // this set's next use: chunk c + 2 — its stores must have left the buffers first
        tma_wait_group_read<0>();
        mbar_arrive_expect_tx(&xbar[set], 2 * kEpiBufBytes);
        tma_load_2d(xb0 + (set * 2) * kEpiBufBytes, map_x, &xbar[set], n0 + 128, row0);
        tma_load_2d(xb0 + (set * 2 + 1) * kEpiBufBytes, map_x, &xbar[set], n0 + 128 + 32, row0);
      }
    }
    __syncwarp();
  }
}

// BM: rows per UMMA tile (only 128 is instantiated; 64 — accumulator rows in TMEM lanes 32q .. 32q+15, drained by lanes
// 0..15 of epilogue warp q — computes bit-identical results but was not faster, see below).
//
// KPS = 64-wide k-blocks per pipeline stage (per full/empty barrier round trip).  With small tiles the loop is not bound
// by the tensor pipe or by operand traffic but by the issuing thread's own round trip (mbarrier try_wait, fence, four
// tcgen05.mma, tcgen05.commit: ~550 cycles per k-block whatever M, N or the smem layout — tools/diag_small_gemm.py);
// KPS > 1 puts several k-blocks behind ONE wait and ONE commit.
// What does NOT move the remaining ~90 cycles per tcgen05.mma at these tile sizes (all built, measured in a dependent
// chain and removed again; profiles/r02_small_gemm_chain_experiments.log): 64-row UMMA tiles on twice the CTAs (BM = 64:
// 7.9 -> 7.8 us, with KPS = 4 6.5 -> 6.3 us), the A operand in 32-byte-swizzled K = 16 slices (slower: four TMA
// boxes per k-block), two issuing threads on alternate k-groups with two accumulators (6.51 -> 6.49 us), N = 16
// instead of 64.  It is a per-instruction cost of the tensor pipe for shared-memory operands, so the only lever left
// is FEWER instructions per CTA: K split over a cluster (gemm_splitk_sm100.cuh) where the tile count allows.
template <int BN, int BM = GEMM_BM, int KPS = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_out, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStagesK = Cfg::kStages / KPS;                 // pipeline depth in units of KPS k-blocks
  constexpr int kStageBytesK = Cfg::kStageBytes * KPS;
  static_assert(Cfg::kStages % KPS == 0 && kStagesK >= 2, "stage grouping");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + Cfg::kStages * Cfg::kStageBytes;   // 1024-aligned: stage sizes are multiples of 1024
  uint8_t* bar_base = epi_base + Cfg::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;                // (barrier slots sized for KPS = 1; kStagesK of them used)
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / GEMM_BK;
  constexpr uint32_t kTxBytes = (BM * GEMM_BK * 2 + Cfg::kStageBytesB) * KPS;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_out);
  }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < kStagesK; ++s) {
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
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();   // everything above overlapped the previous kernel's tail; nothing below runs before it has finished

  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; kb += KPS) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], kTxBytes);
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            uint8_t* sa = smem + stage * kStageBytesK + j * Cfg::kStageBytes;
            uint8_t* sb = sa + Cfg::kStageBytesA;
            const int k_elem = (kb + j) * GEMM_BK;
            const int row_off = k_elem / p.a_k_wrap;
            const int a_k = k_elem - row_off * p.a_k_wrap;
            tma_load_2d(sa, &map_a, &full_bar[stage], a_k, m_blk * BM + row_off);
            tma_load_2d(sb, &map_b, &full_bar[stage], k_elem, n_blk * BN);
          }
          if (++stage == kStagesK) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // -------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; kb += KPS) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < KPS; ++j) {
            const uint32_t sa = smem_u32(smem + stage * kStageBytesK + j * Cfg::kStageBytes);
            const uint32_t sb = sa + Cfg::kStageBytesA;
            const uint64_t adesc = umma_smem_desc_sw128(sa, 1024, 16);
            const uint64_t bdesc = umma_smem_desc_sw128(sb, 1024, 16);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // +32 bytes per K=16 step inside the 128-byte swizzled row (encoded >>4 -> +2)
              umma_bf16_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | j | k) != 0);
            }
          }
          umma_commit(&empty_bar[stage]);  // smem slots reusable once these MMAs retire
          if (++stage == kStagesK) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);      // accumulator complete -> epilogue
      }
    }
  } else if (warp < 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp;  // TMEM lane quarter this warp may read (warp id % 4)
    uint8_t* my_bufs = epi_base + q * 2 * kEpiBufBytes;
    EpiRegs st;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      gemm_epilogue_slab<BN, BM / 4>(p, &map_out, tmem_base + acc * BN, m_blk * BM, n_blk * BN, q, lane, my_bufs, st,
                                     &tmem_full[acc], acc_phase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (p.tma_store && lane == 0) tma_wait_group_read<0>();  // smem of every bulk store has been read; the writes land by grid end
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}


// =============================================================================================
// CTA-pair GEMM (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x BN tile.
// Each CTA loads its own 128 rows of A and HALF of the B tile (BN/2 weight rows); the leader's
// single MMA thread issues 256 x BN x 16 UMMAs that read both CTAs' shared memory, and each CTA's
// TMEM receives its own 128 accumulator rows.  Per 64-deep K block the pair moves 32+32 KiB for
// 256x256x64 MACs — 1.5x the arithmetic intensity of the 128 x 256 single-CTA tile, which is what
// the L2 -> SM path needs (the single-CTA kernel saturates L2 bandwidth near 1.0 PFLOP/s).
//   full barrier   : leader's, count 2 (leader arrive.expect_tx for both CTAs' bytes + peer arrive)
//   empty barrier  : per CTA, released by the leader's multicast tcgen05.commit
//   tmem_full      : per CTA, multicast commit;  tmem_empty: leader's, 8 arrivals (4 warps x 2 CTAs)
// =============================================================================================
template <int BN, int EPI = EPI_GENERIC>
struct Gemm2Cfg {
  static constexpr int kStageBytesA = GEMM_BM * GEMM_BK * 2;
  static constexpr int kStageBytesB = (BN / 2) * GEMM_BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  // EPI_RESID_NORM stages the old residual tile through shared memory (6 boxes per epilogue warp): 4 stages instead of 6
  // (measured: all pair GEMMs at 4 stages lose 1-4 %; only Wo / W2 run this variant)
  static constexpr int kEpiBytes = (EPI == EPI_RESID_NORM ? 4 * 5 : 4 * 2) * kEpiBufBytes;
  static constexpr int kStages = (EPI == EPI_RESID_NORM ? 131072 : 196608) / kStageBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kBarBytes = 256;
  static constexpr int kBiasBytes = (EPI == EPI_RESID_NORM ? 2 : 1) * BN * 4;   // per-tile bias (and gamma) staging of the specialised epilogues
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kBarBytes + kBiasBytes + 1024;
  static_assert(kSmemBytes <= 232448, "CTA-pair GEMM exceeds 227 KB of shared memory");
};

template <int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16_sm100_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out2, const GemmParams p) {
  using Cfg = Gemm2Cfg<BN, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint8_t* bar_base = epi_base + Cfg::kEpiBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* xbar = tmem_empty + 2;                                       // [4 warps][2]: x_old staging of EPI_RESID_NORM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xbar + 8);
  float* sbias = reinterpret_cast<float*>(bar_base + Cfg::kBarBytes);   // [2*BN], fast epilogues only

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / GEMM_BK;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (p.tma_store) tma_prefetch_desc(&map_out);
    if (EPI == EPI_RESID_NORM) tma_prefetch_desc(&map_out2);
  }
  if (warp == 5 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);
    }
    for (int s = 0; s < 8; ++s) mbar_init(&xbar[s], 1);
    fence_mbar_init();
  }
  if (warp == 6) {
    tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kStageBytesA;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          else mbar_arrive_cluster(leader_full);
          const int k_elem = kb * GEMM_BK;
          const int row_off = k_elem / p.a_k_wrap;
          const int a_k = k_elem - row_off * p.a_k_wrap;
          tma_load_2d_pair(sa, &map_a, leader_full, a_k, m_blk * 2 * GEMM_BM + rank * GEMM_BM + row_off);
          tma_load_2d_pair(sb, &map_b, leader_full, k_elem, n_blk * BN + rank * (BN / 2));
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kStageBytesA;
          const uint64_t adesc = umma_smem_desc_sw128(sa, 1024, 16);
          const uint64_t bdesc = umma_smem_desc_sw128(sb, 1024, 16);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16_ss_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit_pair(&empty_bar[stage], 3);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[acc], 3);
      }
    }
  } else if (warp < 4) {
    const int q = warp;
    uint8_t* my_bufs = epi_base + q * (EPI == EPI_RESID_NORM ? 5 : 2) * kEpiBufBytes;
    typename std::conditional<EPI == EPI_GENERIC, EpiRegs, EpiFastRegs>::type st;
    int it = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++it) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      if constexpr (EPI == EPI_GENERIC) {
        gemm_epilogue_slab<BN>(p, &map_out, tmem_base + acc * BN, m_blk * 2 * GEMM_BM + rank * GEMM_BM, n_blk * BN, q, lane,
                               my_bufs, st, &tmem_full[acc], acc_phase);
      } else if constexpr (EPI == EPI_RESID_NORM) {
        gemm_epilogue_slab_resid_norm<BN>(p, &map_out, &map_out2, tmem_base + acc * BN, m_blk * 2 * GEMM_BM + rank * GEMM_BM,
                                          n_blk * BN, q, lane, my_bufs, xbar + 2 * q, sbias, st, &tmem_full[acc], acc_phase);
      } else {
        gemm_epilogue_slab_fast<BN, EPI>(p, &map_out, tmem_base + acc * BN, m_blk * 2 * GEMM_BM + rank * GEMM_BM, n_blk * BN, q,
                                         lane, my_bufs, sbias, st, &tmem_full[acc], acc_phase);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
    }
    if (p.tma_store && lane == 0) tma_wait_group_read<0>();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's smem / barriers stay alive until the leader's last MMA and commit retired
  if (warp == 6) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mc
