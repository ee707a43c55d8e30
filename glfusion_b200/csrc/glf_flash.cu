// glf_flash.cu — mode='embedded' (softmax) attention of TPAVIModule (R/models/ours.py:878-902 with
// f_div_C = softmax(f, dim=-1)):   Y_b = softmax(Theta_b Phi_b^T) G_b      (no 1/sqrt(d) scale in the reference).
//
// Version 1 (this file): exact attention with the N x N score matrix materialised for a bounded CHUNK of sequences
// at a time (scratch sized by attn_chunk()), every product on the tcgen05 GEMM of glf_gemm.cu:
//   fwd : S = Theta Phi^T (fp32) -> row softmax (P bf16, lse fp32) -> Y = P G
//   bwd : recompute S ; dPm = dY G^T ; P = exp(S - lse) ; dS = P o (dPm - delta) ; delta = rowsum(dY o Y)
//         dTheta = dS Phi ; dPhi = dS^T Theta ; dG = P^T dY        (transposes are MN-major operand reads, no copies)
// Memory is O(chunk * N^2), independent of the batch size; HBM traffic is O(N^2) per sequence, which is what a
// single-kernel flash version (scores kept in TMEM, online softmax in registers) removes — same entry points.
#include <cstdlib>

#include "glf_internal.h"
#include "glf_ptx.cuh"

namespace glf {

namespace {

__device__ __forceinline__ float block_reduce_max(float v, float* sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = sm[0];
  for (int i = 1; i < nw; ++i) r = fmaxf(r, sm[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* sm) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) sm[w] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < nw; ++i) r += sm[i];
  __syncthreads();
  return r;
}

// one block per score row: P = softmax(S) in bf16 (padding columns [N, ld) zeroed), lse = max + log(sum)
__global__ void __launch_bounds__(256)
    softmax_rows_kernel(const float* __restrict__ S, bf16* __restrict__ P, float* __restrict__ lse, int N, int ld) {
  __shared__ float sm[8];
  const long long row = blockIdx.x;
  const float* s = S + row * ld;
  bf16* p = P + row * ld;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) m = fmaxf(m, s[j]);
  m = block_reduce_max(m, sm);
  float sum = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) sum += __expf(s[j] - m);
  sum = block_reduce_sum(sum, sm);
  const float inv = 1.f / sum;
  for (int j = threadIdx.x; j < ld; j += blockDim.x) p[j] = __float2bfloat16(j < N ? __expf(s[j] - m) * inv : 0.f);
  if (threadIdx.x == 0) lse[row] = m + __logf(sum);
}

// delta[row] = sum_j dY[row, j] * Y[row, j]     (one warp per row)
__global__ void rowdot_kernel(const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ out,
                              long long rows, int d) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc = 0.f;
  for (int j = lane * 2; j < d; j += 64) {
    const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(A + row * d + j));
    const float2 b = unpack_bf16(*reinterpret_cast<const uint32_t*>(B + row * d + j));
    acc = fmaf(a.x, b.x, fmaf(a.y, b.y, acc));
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

// P = exp(S - lse[row]) ; dS = P * (dPm - delta[row]) ; both bf16, padding columns zeroed.  grid.y = row
__global__ void __launch_bounds__(256)
    softmax_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dPm, const float* __restrict__ lse,
                       const float* __restrict__ delta, bf16* __restrict__ P, bf16* __restrict__ dS, int N, int ld) {
  const long long row = blockIdx.y;
  const float l = lse[row], dl = delta[row];
  const long long off = row * ld;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ld; j += gridDim.x * blockDim.x) {
    float p = 0.f, ds = 0.f;
    if (j < N) {
      p = __expf(S[off + j] - l);
      ds = p * (dPm[off + j] - dl);
    }
    P[off + j] = __float2bfloat16(p);
    dS[off + j] = __float2bfloat16(ds);
  }
}

GemmOperand op(const void* p, int mn, long long ld, long long bs) {
  GemmOperand o;
  o.ptr = p; o.mn_major = mn; o.ld = ld; o.batch_stride = bs;
  return o;
}


}  // namespace

// ------------------------------------------------------------------------------------------------ flash forward
// One CTA = one 128-query tile of one sequence; it streams 128-key tiles of K = Phi and V = G.
//   warp 0      TMA producer: Q once, K / V tiles through 2-deep rings
//   warp 1      MMA issuer  : S_j = Q K_j^T (smem x smem -> TMEM, two S buffers), O += P_j V_j with P read from TMEM
//                              (tcgen05.mma A-from-TMEM) and V as an MN-major smem operand
//   warps 2..5  softmax     : one query row per thread: tcgen05.ld the scores, online max / sum in registers, rescale O in
//                              TMEM when the running max moved, write P as packed bf16 back over the consumed score
//                              columns (tcgen05.st), finally O / l -> Y (bf16) and lse = m + log l
// TMEM: columns [0,128) S0/P0, [128,256) S1/P1, [256,256+D) O.
template <int D>
struct FlashCfg {
  static constexpr int BQ = 128, BKV = 128;
  static constexpr uint32_t TILE = BQ * D * 2;               // bytes of a 128 x D bf16 tile (D/64 swizzle atoms)
  static constexpr uint32_t SMEM = TILE * 5 + 1024;          // Q + 2 K + 2 V
  static constexpr uint32_t COL_S0 = 0, COL_S1 = 128, COL_O = 256;
};

template <int D>
__global__ void __launch_bounds__(192, 1)
    flash_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ Y, float* __restrict__ lse, int N,
                     int q_tiles) {
  using Cfg = FlashCfg<D>;
  extern __shared__ uint8_t fsm_raw[];
  __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], o_done;
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(fsm_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sK = base + Cfg::TILE, sV = base + 3 * Cfg::TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / q_tiles, qt = blockIdx.x % q_tiles;
  const int q0 = qt * Cfg::BQ;
  const int T = (N + Cfg::BKV - 1) / Cfg::BKV;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&q_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&k_full[i]), 1);
      mbar_init(smem_u32(&k_empty[i]), 1);
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&v_empty[i]), 1);
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&p_full[i]), 128);
    }
    mbar_init(smem_u32(&o_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&q_full), Cfg::TILE);
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb) tma_load_4d(&tmQ, smem_u32(&q_full), sQ + kb * 16384, kb * 64, q0, b, 0);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = static_cast<uint32_t>(j >> 1) & 1u;
        mbar_wait(smem_u32(&k_empty[st]), ph ^ 1u);
        mbar_expect_tx(smem_u32(&k_full[st]), Cfg::TILE);
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb)
          tma_load_4d(&tmK, smem_u32(&k_full[st]), sK + st * Cfg::TILE + kb * 16384, kb * 64, j * Cfg::BKV, b, 0);
        mbar_wait(smem_u32(&v_empty[st]), ph ^ 1u);
        mbar_expect_tx(smem_u32(&v_full[st]), Cfg::TILE);
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb)
          tma_load_4d(&tmV, smem_u32(&v_full[st]), sV + st * Cfg::TILE + kb * 16384, kb * 64, j * Cfg::BKV, b, 0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);   // S = Q K^T : both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, D, false, true);      // O += P V  : A in TMEM, B MN-major
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        mbar_wait(smem_u32(&k_full[st]), static_cast<uint32_t>(j >> 1) & 1u);
        tc_fence_after();
        const uint32_t kbase = sK + st * Cfg::TILE;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16(tmem + (st ? Cfg::COL_S1 : Cfg::COL_S0), make_sdesc(sQ + off, 16, 1024),
                   make_sdesc(kbase + off, 16, 1024), idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&k_empty[st]));
        umma_commit(smem_u32(&s_full[st]));
      };
      mbar_wait(smem_u32(&q_full), 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = static_cast<uint32_t>(j >> 1) & 1u;
        if (j + 1 < T) issue_qk(j + 1);          // next scores are computed while the softmax of tile j runs
        mbar_wait(smem_u32(&p_full[st]), ph);    // P_j is in TMEM, O has been rescaled
        mbar_wait(smem_u32(&v_full[st]), ph);
        tc_fence_after();
        const uint32_t vbase = sV + st * Cfg::TILE;
        const uint32_t pcol = tmem + (st ? Cfg::COL_S1 : Cfg::COL_S0);
#pragma unroll
        for (int k = 0; k < Cfg::BKV / 16; ++k) {
          umma_f16_ts(tmem + Cfg::COL_O, pcol + k * 8, make_sdesc(vbase + k * 2048, 16384, 1024), idesc_pv,
                      (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&v_empty[st]));
        umma_commit(smem_u32(&o_done));
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < T; ++j) {
      const int st = j & 1;
      const uint32_t ph = static_cast<uint32_t>(j >> 1) & 1u;
      const uint32_t scol = tmem + lane_addr + (st ? Cfg::COL_S1 : Cfg::COL_S0);
      const int key0 = j * Cfg::BKV;
      mbar_wait(smem_u32(&s_full[st]), ph);
      tc_fence_after();
      // pass 1: running max
      float mx = m;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(scol + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float sv = (key0 + c * 32 + i < N) ? __uint_as_float(v[i]) : -INFINITY;
          mx = fmaxf(mx, sv);
        }
      }
      const float alpha = exp2f((m - mx) * LOG2E);   // exp(m_old - m_new); 0 on the first tile (m = -inf)
      if (j > 0) {
        mbar_wait(smem_u32(&o_done), static_cast<uint32_t>(j - 1) & 1u);   // P_{j-1} V_{j-1} has landed in O
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.f)) {
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(tmem + lane_addr + Cfg::COL_O + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            uint32_t lo[16], hi[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { lo[i] = v[i]; hi[i] = v[16 + i]; }
            tmem_st_32x32_x16(tmem + lane_addr + Cfg::COL_O + c * 32, lo);
            tmem_st_32x32_x16(tmem + lane_addr + Cfg::COL_O + c * 32 + 16, hi);
          }
          tmem_st_wait();
        }
      }
      l *= alpha;
      // pass 2: P = exp(S - m_new) as packed bf16 over the already consumed score columns
      const float mb = mx * LOG2E;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(scol + c * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = (key0 + c * 32 + i < N) ? exp2f(fmaf(__uint_as_float(v[i]), LOG2E, -mb)) : 0.f;
          const float p1 = (key0 + c * 32 + i + 1 < N) ? exp2f(fmaf(__uint_as_float(v[i + 1]), LOG2E, -mb)) : 0.f;
          l += p0 + p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
        tmem_st_32x32_x16(scol + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full[st]));
      m = mx;
    }
    // epilogue: O / l -> Y, lse
    mbar_wait(smem_u32(&o_done), static_cast<uint32_t>(T - 1) & 1u);
    tc_fence_after();
    const float inv = 1.f / l;
    const int grow = q0 + row;
    bf16* yrow = Y + (static_cast<long long>(b) * N + grow) * D;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + lane_addr + Cfg::COL_O + c * 32, v);
      tmem_ld_wait();
      if (grow < N) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk = make_uint4(pack_bf16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv),
                                pack_bf16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv),
                                pack_bf16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv),
                                pack_bf16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv));
          *reinterpret_cast<uint4*>(yrow + c * 32 + i) = pk;
        }
      }
    }
    if (grow < N) lse[static_cast<long long>(b) * N + grow] = m + __logf(l);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ flash backward
// Recompute-based (nothing N x N is stored): with lse from the forward and delta = rowsum(dY o Y),
//   P = exp(S - lse) ,  dS = P o (dY V^T - delta)
//   dQ kernel   (CTA = 128 queries, streams K/V tiles):   S = Q K^T, dP = dY V^T -> dS (bf16, TMEM) -> dQ += dS K
//   dK/dV kernel(CTA = 128 keys, streams Q/dY tiles):     S^T = K Q^T, dP^T = V dY^T -> P^T, dS^T (bf16, TMEM)
//                                                          -> dV += P^T dY , dK += dS^T Q
// Every product is a tcgen05.mma; P / dS never leave TMEM (they are the A operand of the second product); transposed
// uses of a tile (Q as [q,d] and as [d,q]) are the same shared-memory bytes read through K-major / MN-major descriptors.
// Deterministic: no atomics.  Scores are recomputed once per kernel (7 products per tile pair instead of 5).
template <int D>
__global__ void __launch_bounds__(192, 1)
    flash_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdY,
                        const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dP3, int N,
                        int q_tiles) {
  using Cfg = FlashCfg<D>;
  extern __shared__ uint8_t fsm_raw[];
  __shared__ uint64_t q_full, k_full[2], v_full[2], kv_empty[2], s_full, ds_full, dq_done;
  __shared__ uint32_t tmem_holder;
  const uint32_t base = (smem_u32(fsm_raw) + 1023u) & ~1023u;
  const uint32_t sQ = base, sdY = base + Cfg::TILE, sK = base + 2 * Cfg::TILE, sV = base + 4 * Cfg::TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / q_tiles, qt = blockIdx.x % q_tiles;
  const int q0 = qt * 128;
  const int T = (N + 127) / 128;
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DQ = 256;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdY);
    mbar_init(smem_u32(&q_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&k_full[i]), 1);
      mbar_init(smem_u32(&v_full[i]), 1);
      mbar_init(smem_u32(&kv_empty[i]), 1);
    }
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&ds_full), 128);
    mbar_init(smem_u32(&dq_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&q_full), 2 * Cfg::TILE);
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb) {
        tma_load_4d(&tmQ, smem_u32(&q_full), sQ + kb * 16384, kb * 64, q0, b, 0);
        tma_load_4d(&tmdY, smem_u32(&q_full), sdY + kb * 16384, kb * 64, q0, b, 0);
      }
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = static_cast<uint32_t>(j >> 1) & 1u;
        mbar_wait(smem_u32(&kv_empty[st]), ph ^ 1u);
        mbar_expect_tx(smem_u32(&k_full[st]), Cfg::TILE);
        mbar_expect_tx(smem_u32(&v_full[st]), Cfg::TILE);
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb) {
          tma_load_4d(&tmK, smem_u32(&k_full[st]), sK + st * Cfg::TILE + kb * 16384, kb * 64, j * 128, b, 0);
          tma_load_4d(&tmV, smem_u32(&v_full[st]), sV + st * Cfg::TILE + kb * 16384, kb * 64, j * 128, b, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, false, true);
      mbar_wait(smem_u32(&q_full), 0);
      tc_fence_after();
      for (int j = 0; j < T; ++j) {
        const int st = j & 1;
        const uint32_t ph = static_cast<uint32_t>(j >> 1) & 1u;
        mbar_wait(smem_u32(&k_full[st]), ph);
        mbar_wait(smem_u32(&v_full[st]), ph);
        tc_fence_after();
        const uint32_t kbase = sK + st * Cfg::TILE, vbase = sV + st * Cfg::TILE;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {   // S = Q K^T
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16(tmem + COL_S, make_sdesc(sQ + off, 16, 1024), make_sdesc(kbase + off, 16, 1024), idesc_kk,
                   k != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {   // dP = dY V^T
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16(tmem + COL_DP, make_sdesc(sdY + off, 16, 1024), make_sdesc(vbase + off, 16, 1024), idesc_kk,
                   k != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&s_full));
        mbar_wait(smem_u32(&ds_full), static_cast<uint32_t>(j) & 1u);   // dS (bf16) sits in TMEM columns [0,64)
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {        // dQ += dS K   (A from TMEM, K tile read MN-major)
          umma_f16_ts(tmem + COL_DQ, tmem + COL_S + k * 8, make_sdesc(kbase + k * 2048, 16384, 1024), idesc_dq,
                      (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&kv_empty[st]));
        if (j == T - 1) umma_commit(smem_u32(&dq_done));
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int grow = q0 + row;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    constexpr float LOG2E = 1.4426950408889634f;
    const bool rvalid = grow < N;
    const float l2 = rvalid ? lse[static_cast<long long>(b) * N + grow] * LOG2E : 0.f;
    const float dl = rvalid ? delta[static_cast<long long>(b) * N + grow] : 0.f;
    for (int j = 0; j < T; ++j) {
      mbar_wait(smem_u32(&s_full), static_cast<uint32_t>(j) & 1u);
      tc_fence_after();
      const int key0 = j * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem + lane_addr + COL_S + c * 32, sv);
        tmem_ld_32x32(tmem + lane_addr + COL_DP + c * 32, dv);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float d0 = 0.f, d1 = 0.f;
          if (rvalid && key0 + c * 32 + i < N)
            d0 = exp2f(fmaf(__uint_as_float(sv[i]), LOG2E, -l2)) * (__uint_as_float(dv[i]) - dl);
          if (rvalid && key0 + c * 32 + i + 1 < N)
            d1 = exp2f(fmaf(__uint_as_float(sv[i + 1]), LOG2E, -l2)) * (__uint_as_float(dv[i + 1]) - dl);
          pk[i >> 1] = pack_bf16(d0, d1);
        }
        tmem_st_32x32_x16(tmem + lane_addr + COL_S + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&ds_full));
    }
    mbar_wait(smem_u32(&dq_done), 0);
    tc_fence_after();
    bf16* orow = dP3 + (static_cast<long long>(b) * N + grow) * 3 * D;   // dTheta slot: columns [0, D)
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + lane_addr + COL_DQ + c * 32, v);
      tmem_ld_wait();
      if (rvalid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk = make_uint4(pack_bf16(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                pack_bf16(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])),
                                pack_bf16(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])),
                                pack_bf16(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
          *reinterpret_cast<uint4*>(orow + c * 32 + i) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int D>
__global__ void __launch_bounds__(192, 1)
    flash_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdY,
                         const float* __restrict__ lse, const float* __restrict__ delta, bf16* __restrict__ dP3,
                         int N, int kv_tiles) {
  using Cfg = FlashCfg<D>;
  extern __shared__ uint8_t fsm_raw[];
  __shared__ uint64_t kv_full, q_full[2], q_empty[2], s_full, p_full, acc_done;
  __shared__ uint32_t tmem_holder;
  __shared__ float lse_sm[2][128], del_sm[2][128];
  const uint32_t base = (smem_u32(fsm_raw) + 1023u) & ~1023u;
  const uint32_t sK = base, sV = base + Cfg::TILE, sQ = base + 2 * Cfg::TILE, sdY = base + 4 * Cfg::TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / kv_tiles, kt = blockIdx.x % kv_tiles;
  const int k0 = kt * 128;
  const int T = (N + 127) / 128;   // query tiles
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 384;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdY);
    mbar_init(smem_u32(&kv_full), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&q_full[i]), 1);
      mbar_init(smem_u32(&q_empty[i]), 1);
    }
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&p_full), 128);
    mbar_init(smem_u32(&acc_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_holder), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_holder;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&kv_full), 2 * Cfg::TILE);
#pragma unroll
      for (int kb = 0; kb < D / 64; ++kb) {
        tma_load_4d(&tmK, smem_u32(&kv_full), sK + kb * 16384, kb * 64, k0, b, 0);
        tma_load_4d(&tmV, smem_u32(&kv_full), sV + kb * 16384, kb * 64, k0, b, 0);
      }
      for (int i = 0; i < T; ++i) {
        const int st = i & 1;
        const uint32_t ph = static_cast<uint32_t>(i >> 1) & 1u;
        mbar_wait(smem_u32(&q_empty[st]), ph ^ 1u);
        mbar_expect_tx(smem_u32(&q_full[st]), 2 * Cfg::TILE);
#pragma unroll
        for (int kb = 0; kb < D / 64; ++kb) {
          tma_load_4d(&tmQ, smem_u32(&q_full[st]), sQ + st * Cfg::TILE + kb * 16384, kb * 64, i * 128, b, 0);
          tma_load_4d(&tmdY, smem_u32(&q_full[st]), sdY + st * Cfg::TILE + kb * 16384, kb * 64, i * 128, b, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, false, false);
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, D, false, true);
      mbar_wait(smem_u32(&kv_full), 0);
      tc_fence_after();
      for (int i = 0; i < T; ++i) {
        const int st = i & 1;
        const uint32_t ph = static_cast<uint32_t>(i >> 1) & 1u;
        mbar_wait(smem_u32(&q_full[st]), ph);
        tc_fence_after();
        const uint32_t qbase = sQ + st * Cfg::TILE, ybase = sdY + st * Cfg::TILE;
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {   // S^T = K Q^T
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16(tmem + COL_S, make_sdesc(sK + off, 16, 1024), make_sdesc(qbase + off, 16, 1024), idesc_kk,
                   k != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < D / 16; ++k) {   // dP^T = V dY^T
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          umma_f16(tmem + COL_DP, make_sdesc(sV + off, 16, 1024), make_sdesc(ybase + off, 16, 1024), idesc_kk,
                   k != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&s_full));
        mbar_wait(smem_u32(&p_full), static_cast<uint32_t>(i) & 1u);   // P^T in [0,64), dS^T in [128,192) (bf16)
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {        // dV += P^T dY   (dY tile read MN-major)
          umma_f16_ts(tmem + COL_DV, tmem + COL_S + k * 8, make_sdesc(ybase + k * 2048, 16384, 1024), idesc_acc,
                      (i | k) != 0 ? 1u : 0u);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {        // dK += dS^T Q   (Q tile read MN-major)
          umma_f16_ts(tmem + COL_DK, tmem + COL_DP + k * 8, make_sdesc(qbase + k * 2048, 16384, 1024), idesc_acc,
                      (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&q_empty[st]));
        if (i == T - 1) umma_commit(smem_u32(&acc_done));
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;          // key row of this thread
    const int gkey = k0 + row;
    const bool kvalid = gkey < N;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int e = (warp - 2) * 32 + lane;   // 0..127
    constexpr float LOG2E = 1.4426950408889634f;
    for (int i = 0; i < T; ++i) {
      const int sb = i & 1;
      {  // per-query lse / delta of this q tile (columns of S^T); double-buffered, published by a named barrier
        const int gq = i * 128 + e;
        lse_sm[sb][e] = gq < N ? lse[static_cast<long long>(b) * N + gq] * LOG2E : 0.f;
        del_sm[sb][e] = gq < N ? delta[static_cast<long long>(b) * N + gq] : 0.f;
      }
      named_bar_sync(1, 128);
      mbar_wait(smem_u32(&s_full), static_cast<uint32_t>(i) & 1u);
      tc_fence_after();
      const int qv = min(128, N - i * 128);   // valid queries in this tile
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(tmem + lane_addr + COL_S + c * 32, sv);
        tmem_ld_32x32(tmem + lane_addr + COL_DP + c * 32, dv);
        tmem_ld_wait();
        uint32_t pp[16], pd[16];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
          const int col = c * 32 + t;
          if (kvalid && col < qv) {
            p0 = exp2f(fmaf(__uint_as_float(sv[t]), LOG2E, -lse_sm[sb][col]));
            d0 = p0 * (__uint_as_float(dv[t]) - del_sm[sb][col]);
          }
          if (kvalid && col + 1 < qv) {
            p1 = exp2f(fmaf(__uint_as_float(sv[t + 1]), LOG2E, -lse_sm[sb][col + 1]));
            d1 = p1 * (__uint_as_float(dv[t + 1]) - del_sm[sb][col + 1]);
          }
          pp[t >> 1] = pack_bf16(p0, p1);
          pd[t >> 1] = pack_bf16(d0, d1);
        }
        tmem_st_32x32_x16(tmem + lane_addr + COL_S + c * 16, pp);
        tmem_st_32x32_x16(tmem + lane_addr + COL_DP + c * 16, pd);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full));
    }
    mbar_wait(smem_u32(&acc_done), 0);
    tc_fence_after();
    bf16* orow = dP3 + (static_cast<long long>(b) * N + gkey) * 3 * D;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t vk[32], vv[32];
      tmem_ld_32x32(tmem + lane_addr + COL_DK + c * 32, vk);
      tmem_ld_32x32(tmem + lane_addr + COL_DV + c * 32, vv);
      tmem_ld_wait();
      if (kvalid) {
#pragma unroll
        for (int t = 0; t < 32; t += 8) {
          uint4 a = make_uint4(pack_bf16(__uint_as_float(vk[t]), __uint_as_float(vk[t + 1])),
                               pack_bf16(__uint_as_float(vk[t + 2]), __uint_as_float(vk[t + 3])),
                               pack_bf16(__uint_as_float(vk[t + 4]), __uint_as_float(vk[t + 5])),
                               pack_bf16(__uint_as_float(vk[t + 6]), __uint_as_float(vk[t + 7])));
          uint4 g = make_uint4(pack_bf16(__uint_as_float(vv[t]), __uint_as_float(vv[t + 1])),
                               pack_bf16(__uint_as_float(vv[t + 2]), __uint_as_float(vv[t + 3])),
                               pack_bf16(__uint_as_float(vv[t + 4]), __uint_as_float(vv[t + 5])),
                               pack_bf16(__uint_as_float(vv[t + 6]), __uint_as_float(vv[t + 7])));
          *reinterpret_cast<uint4*>(orow + D + c * 32 + t) = a;        // dPhi slot
          *reinterpret_cast<uint4*>(orow + 2 * D + c * 32 + t) = g;    // dG slot
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// per-CTA partial column sums of a bf16 [rows, Ccols] matrix (bias gradients = column sums of dP): part [grid][2][Ccols]
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const bf16* __restrict__ A, float* __restrict__ part,
                                                         long long rows, int Ccols) {
  // thread = one 8-column group per pass; rows strided over (blockIdx, row lane)
  const int groups = Ccols / 8;
  for (int g0 = 0; g0 < groups; g0 += 64) {
    const int g = g0 + (threadIdx.x & 63);
    const int ty = threadIdx.x >> 6;
    float s[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) s[t] = 0.f;
    if (g < groups) {
      for (long long r = static_cast<long long>(blockIdx.x) * 4 + ty; r < rows; r += static_cast<long long>(gridDim.x) * 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(A + r * Ccols + g * 8);
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16(u[t]);
          s[2 * t] += x.x;
          s[2 * t + 1] += x.y;
        }
      }
    }
    __shared__ float sm[4][64 * 8];
#pragma unroll
    for (int t = 0; t < 8; ++t) sm[ty][(threadIdx.x & 63) * 8 + t] = s[t];
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 8; idx += 256) {
      const int col = g0 * 8 + idx;
      if (col < Ccols)
        part[static_cast<long long>(blockIdx.x) * 2 * Ccols + col] = sm[0][idx] + sm[1][idx] + sm[2][idx] + sm[3][idx];
    }
    __syncthreads();
  }
}

template <int D>
static int launch_flash_bwd(const bf16* P3, const bf16* dY, const float* lse, const float* delta, bf16* dP3, int B,
                            int N, cudaStream_t stream) {
  using Cfg = FlashCfg<D>;
  CUtensorMap tq, tk, tv, ty;
  const long long seq = static_cast<long long>(N) * 3 * D;
  int rc;
  if ((rc = make_tmap_bf16(&tq, P3, D, N, B, 3 * D, seq, 128))) return rc;
  if ((rc = make_tmap_bf16(&tk, P3 + D, D, N, B, 3 * D, seq, 128))) return rc;
  if ((rc = make_tmap_bf16(&tv, P3 + 2 * D, D, N, B, 3 * D, seq, 128))) return rc;
  if ((rc = make_tmap_bf16(&ty, dY, D, N, B, D, static_cast<long long>(N) * D, 128))) return rc;
  const uint32_t smem = Cfg::TILE * 6 + 1024;
  const int tiles = (N + 127) / 128;
  {
    auto kern = flash_bwd_dq_kernel<D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(flash_bwd_dq)");
    kern<<<B * tiles, 192, smem, stream>>>(tq, tk, tv, ty, lse, delta, dP3, N, tiles);
    if ((rc = check_cuda(cudaGetLastError(), "flash_bwd_dq launch"))) return rc;
  }
  {
    auto kern = flash_bwd_dkv_kernel<D>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(flash_bwd_dkv)");
    kern<<<B * tiles, 192, smem, stream>>>(tq, tk, tv, ty, lse, delta, dP3, N, tiles);
    if ((rc = check_cuda(cudaGetLastError(), "flash_bwd_dkv launch"))) return rc;
  }
  return 0;
}

template <int D>
static int launch_flash_fwd(const bf16* P3, bf16* Y, float* lse, int B, int N, cudaStream_t stream) {
  using Cfg = FlashCfg<D>;
  CUtensorMap tq, tk, tv;
  const long long seq = static_cast<long long>(N) * 3 * D;
  int rc;
  if ((rc = make_tmap_bf16(&tq, P3, D, N, B, 3 * D, seq, 128))) return rc;
  if ((rc = make_tmap_bf16(&tk, P3 + D, D, N, B, 3 * D, seq, 128))) return rc;
  if ((rc = make_tmap_bf16(&tv, P3 + 2 * D, D, N, B, 3 * D, seq, 128))) return rc;
  auto kern = flash_fwd_kernel<D>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
  if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(flash_fwd)");
  const int q_tiles = (N + 127) / 128;
  kern<<<B * q_tiles, 192, Cfg::SMEM, stream>>>(tq, tk, tv, Y, lse, N, q_tiles);
  return check_cuda(cudaGetLastError(), "flash_fwd launch");
}

namespace {
}  // namespace

int attn_ld(int N) { return (N + 7) & ~7; }

// sequences per chunk so that the backward scratch (12 bytes per score) stays under ~2 GiB (at least 1)
int attn_chunk(long long B, long long N) {
  const long long per_seq = static_cast<long long>(attn_ld(static_cast<int>(N))) * N * 12;
  long long nb = (2LL << 30) / (per_seq > 0 ? per_seq : 1);
  if (nb < 1) nb = 1;
  if (nb > B) nb = B;
  return static_cast<int>(nb);
}
size_t attn_scratch_bytes(long long B, long long N, bool backward) {
  const size_t nb = attn_chunk(B, N);
  const size_t per = static_cast<size_t>(attn_ld(static_cast<int>(N))) * N;
  return nb * per * (backward ? 12 : 6) + 1024;
}

static int attn_fwd_materialized(const bf16* P3, bf16* Y, float* lse, int B, int N, int Ci, void* scratch,
                                 cudaStream_t stream) {
  const int ld = attn_ld(N);
  const int nb = attn_chunk(B, N);
  const size_t per = static_cast<size_t>(ld) * N;
  float* S = reinterpret_cast<float*>(scratch);
  bf16* Pm = reinterpret_cast<bf16*>(S + nb * per);
  const long long seqP = static_cast<long long>(N) * 3 * Ci;
  for (int b0 = 0; b0 < B; b0 += nb) {
    const int cb = (B - b0 < nb) ? (B - b0) : nb;
    const bf16* Pb = P3 + b0 * seqP;
    {  // S = Theta Phi^T   (columns [N, ld) come out as zeros: Phi rows beyond N are TMA out-of-bounds)
      GemmArgs g;
      g.A = op(Pb, 0, 3 * Ci, seqP);
      g.B = op(Pb + Ci, 0, 3 * Ci, seqP);
      g.B.rows = N;
      g.M = N; g.N = ld; g.K = Ci; g.batch = cb;
      g.out_kind = 1;
      g.D = S; g.ldd = ld; g.strideD = static_cast<long long>(per);
      int rc = gemm(g, stream);
      if (rc) return rc;
    }
    softmax_rows_kernel<<<static_cast<unsigned>(static_cast<long long>(cb) * N), 256, 0, stream>>>(
        S, Pm, lse + static_cast<long long>(b0) * N, N, ld);
    int rc = check_cuda(cudaGetLastError(), "softmax_rows launch");
    if (rc) return rc;
    {  // Y = P G
      GemmArgs g;
      g.A = op(Pm, 0, ld, static_cast<long long>(per));
      g.B = op(Pb + 2 * Ci, 1, 3 * Ci, seqP);
      g.M = N; g.N = Ci; g.K = N; g.batch = cb;
      g.D = Y + static_cast<long long>(b0) * N * Ci; g.ldd = Ci; g.strideD = static_cast<long long>(N) * Ci;
      rc = gemm(g, stream);
      if (rc) return rc;
    }
  }
  return 0;
}

int flash_fwd(const bf16* P3, bf16* Y, float* lse, int B, int N, int Ci, void* scratch, cudaStream_t stream) {
  if (const char* e = getenv("GLF_DEBUG_ATTN_MATERIALIZED")) {
    if (e[0] == '1') return attn_fwd_materialized(P3, Y, lse, B, N, Ci, scratch, stream);
  }
  if (Ci == 128) return launch_flash_fwd<128>(P3, Y, lse, B, N, stream);
  if (Ci == 64) return launch_flash_fwd<64>(P3, Y, lse, B, N, stream);
  return attn_fwd_materialized(P3, Y, lse, B, N, Ci, scratch, stream);   // other head widths: exact chunked path
}

static int attn_bwd_materialized(const bf16* P3, const bf16* Y, const bf16* dY, const float* lse, bf16* dP3,
                                 float* delta, float* cs_t, float* cs_p, float* cs_g, int* cs_rows_out, int B, int N,
                                 int Ci, void* scratch, cudaStream_t stream) {
  const int ld = attn_ld(N);
  const int nb = attn_chunk(B, N);
  const size_t per = static_cast<size_t>(ld) * N;
  float* S = reinterpret_cast<float*>(scratch);
  float* dPm = S + nb * per;
  bf16* Pm = reinterpret_cast<bf16*>(dPm + nb * per);
  bf16* dS = Pm + nb * per;
  const long long seqP = static_cast<long long>(N) * 3 * Ci;
  const long long rows = static_cast<long long>(B) * N;
  int cs_rows_total = 0;   // rows of the three column-stat tables written so far (same for all three)
  {
    const long long threads = rows * 32;
    rowdot_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(dY, Y, delta, rows, Ci);
    int rc = check_cuda(cudaGetLastError(), "rowdot launch");
    if (rc) return rc;
  }
  for (int b0 = 0; b0 < B; b0 += nb) {
    const int cb = (B - b0 < nb) ? (B - b0) : nb;
    const bf16* Pb = P3 + b0 * seqP;
    const bf16* dYb = dY + static_cast<long long>(b0) * N * Ci;
    bf16* dPb = dP3 + b0 * seqP;
    const long long cs_off = static_cast<long long>(cs_rows_total) * 2 * Ci;
    int r_t = 0, r_p = 0, r_g = 0;
    int rc;
    {  // S = Theta Phi^T (recompute)
      GemmArgs g;
      g.A = op(Pb, 0, 3 * Ci, seqP);
      g.B = op(Pb + Ci, 0, 3 * Ci, seqP);
      g.B.rows = N;
      g.M = N; g.N = ld; g.K = Ci; g.batch = cb;
      g.out_kind = 1;
      g.D = S; g.ldd = ld; g.strideD = static_cast<long long>(per);
      if ((rc = gemm(g, stream))) return rc;
    }
    {  // dPm = dY G^T
      GemmArgs g;
      g.A = op(dYb, 0, Ci, static_cast<long long>(N) * Ci);
      g.B = op(Pb + 2 * Ci, 0, 3 * Ci, seqP);
      g.B.rows = N;
      g.M = N; g.N = ld; g.K = Ci; g.batch = cb;
      g.out_kind = 1;
      g.D = dPm; g.ldd = ld; g.strideD = static_cast<long long>(per);
      if ((rc = gemm(g, stream))) return rc;
    }
    {
      dim3 grid((ld + 255) / 256 > 8 ? 8 : (ld + 255) / 256, static_cast<unsigned>(cb) * N);
      if (grid.y > 65535) {
        // split rows over several launches
        const long long total = static_cast<long long>(cb) * N;
        for (long long r0 = 0; r0 < total; r0 += 65535) {
          dim3 g2(grid.x, static_cast<unsigned>(total - r0 < 65535 ? total - r0 : 65535));
          softmax_bwd_kernel<<<g2, 256, 0, stream>>>(S + r0 * ld, dPm + r0 * ld, lse + static_cast<long long>(b0) * N + r0,
                                                     delta + static_cast<long long>(b0) * N + r0, Pm + r0 * ld,
                                                     dS + r0 * ld, N, ld);
        }
      } else {
        softmax_bwd_kernel<<<grid, 256, 0, stream>>>(S, dPm, lse + static_cast<long long>(b0) * N,
                                                     delta + static_cast<long long>(b0) * N, Pm, dS, N, ld);
      }
      if ((rc = check_cuda(cudaGetLastError(), "softmax_bwd launch"))) return rc;
    }
    {  // dTheta = dS Phi
      GemmArgs g;
      g.A = op(dS, 0, ld, static_cast<long long>(per));
      g.B = op(Pb + Ci, 1, 3 * Ci, seqP);
      g.M = N; g.N = Ci; g.K = N; g.batch = cb;
      g.D = dPb; g.ldd = 3 * Ci; g.strideD = seqP;
      g.colstats = cs_t + cs_off;
      g.colstats_rows = &r_t;
      if ((rc = gemm(g, stream))) return rc;
    }
    {  // dPhi = dS^T Theta
      GemmArgs g;
      g.A = op(dS, 1, ld, static_cast<long long>(per));
      g.B = op(Pb, 1, 3 * Ci, seqP);
      g.M = N; g.N = Ci; g.K = N; g.batch = cb;
      g.D = dPb + Ci; g.ldd = 3 * Ci; g.strideD = seqP;
      g.colstats = cs_p + cs_off;
      g.colstats_rows = &r_p;
      if ((rc = gemm(g, stream))) return rc;
    }
    {  // dG = P^T dY
      GemmArgs g;
      g.A = op(Pm, 1, ld, static_cast<long long>(per));
      g.B = op(dYb, 1, Ci, static_cast<long long>(N) * Ci);
      g.M = N; g.N = Ci; g.K = N; g.batch = cb;
      g.D = dPb + 2 * Ci; g.ldd = 3 * Ci; g.strideD = seqP;
      g.colstats = cs_g + cs_off;
      g.colstats_rows = &r_g;
      if ((rc = gemm(g, stream))) return rc;
    }
    if (r_p != r_t || r_g != r_t) return set_error(GLF_ERR_INVALID, "internal: column-stat tables disagree");
    cs_rows_total += r_t;
  }
  *cs_rows_out = cs_rows_total;
  return 0;
}

int flash_bwd_colsum_blocks(long long rows) {
  long long b = (rows + 3) / 4;
  return static_cast<int>(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}

int flash_bwd(const bf16* P3, const bf16* Y, const bf16* dY, const float* lse, bf16* dP3, float* delta, float* cs_t,
              float* cs_p, float* cs_g, int cs_cap_rows, int* cs_rows_out, int B, int N, int Ci, void* scratch,
              cudaStream_t stream) {
  bool materialized = !(Ci == 128 || Ci == 64);
  if (const char* e = getenv("GLF_DEBUG_ATTN_MATERIALIZED")) materialized = materialized || e[0] == '1';
  if (materialized)
    return attn_bwd_materialized(P3, Y, dY, lse, dP3, delta, cs_t, cs_p, cs_g, cs_rows_out, B, N, Ci, scratch, stream);
  const long long rows = static_cast<long long>(B) * N;
  {
    const long long threads = rows * 32;
    rowdot_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(dY, Y, delta, rows, Ci);
    int rc = check_cuda(cudaGetLastError(), "rowdot launch");
    if (rc) return rc;
  }
  int rc = (Ci == 128) ? launch_flash_bwd<128>(P3, dY, lse, delta, dP3, B, N, stream)
                       : launch_flash_bwd<64>(P3, dY, lse, delta, dP3, B, N, stream);
  if (rc) return rc;
  // bias gradients: column sums of dP3 = [dTheta | dPhi | dG]; one table [grid][2][3Ci] written into cs_t, the caller's
  // three reductions address it with a row stride of 2*3Ci through the offsets below (cs_p / cs_g are unused here)
  int grid = flash_bwd_colsum_blocks(rows);
  if (grid > cs_cap_rows / 3) grid = cs_cap_rows / 3 > 0 ? cs_cap_rows / 3 : 1;   // cs_t holds cs_cap_rows x 2 x Ci floats
  colsum_bf16_kernel<<<grid, 256, 0, stream>>>(dP3, cs_t, rows, 3 * Ci);
  if ((rc = check_cuda(cudaGetLastError(), "colsum launch"))) return rc;
  *cs_rows_out = -grid;   // negative: "single table of width 3Ci" (see glf_tpavi_bwd)
  return 0;
}

}  // namespace glf
