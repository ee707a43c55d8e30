// glf_eltwise.cu — the HBM-bound kernels of the fusion path: layout packing, weight preparation, BatchNorm
// statistics finalisation, fused BN-normalise + residual + LayerNorm forward/backward, BN-backward apply,
// small fp32 batched products on C' x C' matrices, partial-sum reductions.
// All row kernels use 128-bit loads/stores, one warp per (row, 256-channel slice), warp-shuffle reductions, and
// per-CTA partials + fixed-order finalisation for every cross-row (per-channel) statistic, so results are
// deterministic (SURVEY.md §7 "hard parts").
#include "glf_internal.h"
#include "glf_ptx.cuh"

namespace glf {

namespace {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16(u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
  uint4 v = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  *reinterpret_cast<uint4*>(p) = v;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}

// Sum over the S warps that share one row (slices of 256 channels). red: [2][ROW_WARPS] floats, `buf` alternates.
template <int NV>
__device__ __forceinline__ void row_reduce(float (&v)[NV], int S, int rslot, int slice, float* red, int& buf) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (S > 1) {
    float* r = red + buf * (ROW_WARPS * NV);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) r[(rslot * S + slice) * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float s = 0.f;
      for (int j = 0; j < S; ++j) s += r[(rslot * S + j) * NV + i];
      v[i] = s;
    }
    buf ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------ transpose + cast
template <typename TI, typename TO>
__global__ void transpose_kernel(const TI* __restrict__ in, TO* __restrict__ out, int R, int S) {
  __shared__ float tile[64][65];
  const long long boff = static_cast<long long>(blockIdx.z) * R * S;
  const int r0 = blockIdx.y * 64, s0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  for (int i = ty; i < 64; i += 4) {
    const int r = r0 + i, s = s0 + tx;
    if (r < R && s < S) tile[i][tx] = static_cast<float>(in[boff + static_cast<long long>(r) * S + s]);
  }
  __syncthreads();
  for (int j = ty; j < 64; j += 4) {
    const int s = s0 + j, r = r0 + tx;
    if (r < R && s < S) out[boff + static_cast<long long>(s) * R + r] = static_cast<TO>(tile[tx][j]);
  }
}

// ------------------------------------------------------------------------------------------------ weights
// fp32 masters -> bf16 operands: Wcat [3Ci, C] (theta | phi | g), WcatT [C, 3Ci], bcat [3Ci] fp32, Wz [C, Ci], WzT [Ci, C]
__global__ void prep_weights_kernel(const float* __restrict__ tw, const float* __restrict__ pw,
                                    const float* __restrict__ gw, const float* __restrict__ tb,
                                    const float* __restrict__ pb, const float* __restrict__ gb,
                                    const float* __restrict__ wz, bf16* __restrict__ wcat, bf16* __restrict__ wcatT,
                                    float* __restrict__ bcat, bf16* __restrict__ wzb, bf16* __restrict__ wzT, int C,
                                    int Ci) {
  const int total = 3 * Ci * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / C, c = i % C;
    const int which = r / Ci, rr = r % Ci;
    const float* src = which == 0 ? tw : (which == 1 ? pw : gw);
    const float v = src[rr * C + c];
    wcat[i] = __float2bfloat16(v);
    wcatT[static_cast<long long>(c) * 3 * Ci + r] = __float2bfloat16(v);
    if (i < Ci * C) {  // W_z is [C, Ci]: same element count as one projection
      const int zc = i / Ci, zk = i % Ci;
      const float z = wz[i];
      wzb[i] = __float2bfloat16(z);
      wzT[static_cast<long long>(zk) * C + zc] = __float2bfloat16(z);
    }
    if (i < 3 * Ci) {
      const int w2 = i / Ci, k = i % Ci;
      bcat[i] = (w2 == 0 ? tb : (w2 == 1 ? pb : gb))[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ BN statistics
// partials: [np][2][C] (sum, sum of squares) -> mean, rstd, affine a = gamma*rstd, b = beta - mean*a; running stats.
__global__ void bn_finalize_kernel(const float* __restrict__ part, int np, int C, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, int training, int bn_layer, float* __restrict__ rm,
                                   float* __restrict__ rv, long long* __restrict__ nbt, float* __restrict__ mean_o,
                                   float* __restrict__ rstd_o, float* __restrict__ a_o, float* __restrict__ b_o) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && bn_layer && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  if (!bn_layer) {
    mean_o[c] = 0.f; rstd_o[c] = 1.f; a_o[c] = 1.f; b_o[c] = 0.f;
    return;
  }
  double mean, var;
  if (training) {
    double s = 0.0, s2 = 0.0;
    for (int i = 0; i < np; ++i) {
      s += static_cast<double>(part[(static_cast<long long>(i) * 2) * C + c]);
      s2 += static_cast<double>(part[(static_cast<long long>(i) * 2 + 1) * C + c]);
    }
    mean = s / count;
    var = s2 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    rm[c] = static_cast<float>((1.0 - momentum) * rm[c] + momentum * mean);
    rv[c] = static_cast<float>((1.0 - momentum) * rv[c] + momentum * unbiased);
  } else {
    mean = rm[c];
    var = rv[c];
  }
  const double rstd = 1.0 / sqrt(var + static_cast<double>(eps));
  const double a = gamma[c] * rstd;
  mean_o[c] = static_cast<float>(mean);
  rstd_o[c] = static_cast<float>(rstd);
  a_o[c] = static_cast<float>(a);
  b_o[c] = static_cast<float>(beta[c] - mean * a);
}

// ------------------------------------------------------------------------------------------------ BN + residual + LN fwd
// Z = LayerNorm_C(a*U + b + X) * lw + lb     (ours.py:908-915 after the W_z GEMM).  U may be nullptr (V = 0).
template <typename TO>
__global__ void __launch_bounds__(ROW_THREADS)
    bn_res_ln_fwd_kernel(const bf16* __restrict__ U, const bf16* __restrict__ X, const float* __restrict__ bn_a,
                         const float* __restrict__ bn_b, const float* __restrict__ lw, const float* __restrict__ lb,
                         TO* __restrict__ Z, float* __restrict__ mu_o, float* __restrict__ r_o, long long rows, int C,
                         int S, float eps, int accumulate) {
  __shared__ float red[2 * ROW_WARPS * 2];
  int buf = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int RPB = ROW_WARPS / S;
  const int rslot = warp / S, slice = warp % S;
  const int c0 = slice * 256 + lane * 8;
  const bool cact = c0 < C;
  float a[8], b[8], w[8], bb[8];
  if (cact) {
    load8(bn_a + c0, a); load8(bn_b + c0, b); load8(lw + c0, w); load8(lb + c0, bb);
  }
  const float invC = 1.f / static_cast<float>(C);
  for (long long base = static_cast<long long>(blockIdx.x) * RPB; base < rows; base += static_cast<long long>(gridDim.x) * RPB) {
    const long long row = base + rslot;
    const bool act = cact && row < rows;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    if (act) {
      float x[8];
      load8(X + row * C + c0, x);
      if (U != nullptr) {
        float u[8];
        load8(U + row * C + c0, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(a[i], u[i], b[i]) + x[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = x[i];
      }
    }
    float s[1] = {0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) s[0] += v[i];
    row_reduce<1>(s, S, rslot, slice, red, buf);
    const float mu = s[0] * invC;
    float q[1] = {0.f};
    if (act) {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[0] = fmaf(v[i] - mu, v[i] - mu, q[0]);
    }
    row_reduce<1>(q, S, rslot, slice, red, buf);
    const float r = rsqrtf(q[0] * invC + eps);
    if (act) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf((v[i] - mu) * r, w[i], bb[i]);
      if (accumulate) {
        float z0[8];
        load8(Z + row * C + c0, z0);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += z0[i];
      }
      store8(Z + row * C + c0, o);
      if (slice == 0 && lane == 0 && mu_o != nullptr) {
        mu_o[row] = mu;
        r_o[row] = r;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ BN + residual + LN bwd
// From dZ: dV = dZp (gradient of the pre-LayerNorm sum, also the residual part of dX), and per-CTA partials of the
// four per-channel reductions: d ln_w = sum dZ*xhat, d ln_b = sum dZ, d gamma = sum dV*uhat, d beta = sum dV.
template <typename TI>
__global__ void __launch_bounds__(ROW_THREADS)
    bn_res_ln_bwd_kernel(const TI* __restrict__ dZ, const bf16* __restrict__ U, const bf16* __restrict__ X,
                         const float* __restrict__ bn_a, const float* __restrict__ bn_b,
                         const float* __restrict__ bn_mean, const float* __restrict__ bn_rstd,
                         const float* __restrict__ lw, const float* __restrict__ mu_i, const float* __restrict__ r_i,
                         bf16* __restrict__ dV, float* __restrict__ part, long long rows, int C, int S) {
  __shared__ float red[2 * ROW_WARPS * 2];
  __shared__ float acc_sm[ROW_WARPS][4][32 * 8 + 8];
  int buf = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int RPB = ROW_WARPS / S;
  const int rslot = warp / S, slice = warp % S;
  const int c0 = slice * 256 + lane * 8;
  const bool cact = c0 < C;
  float a[8], b[8], w[8], bm[8], br[8];
  if (cact) {
    load8(bn_a + c0, a); load8(bn_b + c0, b); load8(lw + c0, w); load8(bn_mean + c0, bm); load8(bn_rstd + c0, br);
  }
  float g_lw[8], g_lb[8], g_ga[8], g_be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) g_lw[i] = g_lb[i] = g_ga[i] = g_be[i] = 0.f;
  const float invC = 1.f / static_cast<float>(C);
  for (long long base = static_cast<long long>(blockIdx.x) * RPB; base < rows; base += static_cast<long long>(gridDim.x) * RPB) {
    const long long row = base + rslot;
    const bool act = cact && row < rows;
    float xh[8], dxh[8], dz[8], uh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) xh[i] = dxh[i] = dz[i] = uh[i] = 0.f;
    float r = 0.f;
    if (act) {
      const float mu = mu_i[row];
      r = r_i[row];
      float x[8];
      load8(X + row * C + c0, x);
      load8(dZ + row * C + c0, dz);
      if (U != nullptr) {
        float u[8];
        load8(U + row * C + c0, u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          xh[i] = (fmaf(a[i], u[i], b[i]) + x[i] - mu) * r;
          uh[i] = (u[i] - bm[i]) * br[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) xh[i] = (x[i] - mu) * r;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) dxh[i] = dz[i] * w[i];
    }
    float s[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[0] += dxh[i];
      s[1] = fmaf(dxh[i], xh[i], s[1]);
    }
    row_reduce<2>(s, S, rslot, slice, red, buf);
    if (act) {
      const float m1 = s[0] * invC, m2 = s[1] * invC;
      float dv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dv[i] = r * (dxh[i] - m1 - xh[i] * m2);
        g_lw[i] = fmaf(dz[i], xh[i], g_lw[i]);
        g_lb[i] += dz[i];
        g_ga[i] = fmaf(dv[i], uh[i], g_ga[i]);
        g_be[i] += dv[i];
      }
      store8(dV + row * C + c0, dv);
    }
  }
  // reduce the accumulators over the row slots of this CTA (fixed order), write one partial per CTA
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc_sm[warp][0][lane * 8 + i] = g_lw[i];
    acc_sm[warp][1][lane * 8 + i] = g_lb[i];
    acc_sm[warp][2][lane * 8 + i] = g_ga[i];
    acc_sm[warp][3][lane * 8 + i] = g_be[i];
  }
  __syncthreads();
  // output index space: 4 stats x (S*256) channels
  for (int idx = threadIdx.x; idx < 4 * S * 256; idx += ROW_THREADS) {
    const int st = idx / (S * 256), cc = idx % (S * 256);
    const int sl = cc / 256, ci = cc % 256;
    if (cc < C) {
      float t = 0.f;
      for (int rs = 0; rs < RPB; ++rs) t += acc_sm[rs * S + sl][st][ci];
      part[(static_cast<long long>(blockIdx.x) * 4 + st) * C + cc] = t;
    }
  }
}

// partials [np][4][C] -> parameter gradients + dU coefficient vectors: dU = k1*dV + k2*U + k3   
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int np, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ bn_mean,
                                       const float* __restrict__ bn_rstd, int training, int bn_layer,
                                       float* __restrict__ d_lnw, float* __restrict__ d_lnb,
                                       float* __restrict__ d_gamma, float* __restrict__ d_beta,
                                       float* __restrict__ d_bz, float* __restrict__ k1, float* __restrict__ k2,
                                       float* __restrict__ k3) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s[4] = {0, 0, 0, 0};
  for (int i = 0; i < np; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] += static_cast<double>(part[(static_cast<long long>(i) * 4 + j) * C + c]);
  }
  d_lnw[c] = static_cast<float>(s[0]);
  d_lnb[c] = static_cast<float>(s[1]);
  if (!bn_layer) {
    if (d_gamma) d_gamma[c] = 0.f;
    if (d_beta) d_beta[c] = 0.f;
    d_bz[c] = static_cast<float>(s[3]);  // dU = dV
    k1[c] = 1.f; k2[c] = 0.f; k3[c] = 0.f;
    return;
  }
  d_gamma[c] = static_cast<float>(s[2]);
  d_beta[c] = static_cast<float>(s[3]);
  const double g = gamma[c], rstd = bn_rstd[c], mean = bn_mean[c];
  if (training) {
    const double mdv = s[3] / count, mdvu = s[2] / count;
    k1[c] = static_cast<float>(g * rstd);
    k2[c] = static_cast<float>(-g * rstd * rstd * mdvu);
    k3[c] = static_cast<float>(-g * rstd * mdv + g * rstd * rstd * mean * mdvu);
    d_bz[c] = 0.f;  // train-mode BN cancels any per-channel shift of U: the gradient is analytically zero
  } else {
    k1[c] = static_cast<float>(g * rstd);
    k2[c] = 0.f;
    k3[c] = 0.f;
    d_bz[c] = static_cast<float>(g * rstd * s[3]);
  }
}

// dU = k1*dV + k2*U + k3 (per channel), bf16 in/out, 8 channels per thread
__global__ void bn_bwd_apply_kernel(const bf16* __restrict__ dV, const bf16* __restrict__ U,
                                    const float* __restrict__ k1, const float* __restrict__ k2,
                                    const float* __restrict__ k3, bf16* __restrict__ dU, long long nvec, int C) {
  const int cv = C / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(i % cv) * 8;
    float a[8], b[8], c[8], dv[8], u[8], o[8];
    load8(k1 + c0, a); load8(k2 + c0, b); load8(k3 + c0, c);
    load8(dV + i * 8, dv);
    load8(U + i * 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(a[j], dv[j], fmaf(b[j], u[j], c[j]));
    store8(dU + i * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------ small fp32 GEMM
// C[b][m][n] = alpha * sum_{rb} sum_k A[b,rb][m][k] * B[b,rb][k][n]     generic strides, fp32 SIMT, 64x64 tiles.
struct SmallGemmP {
  const float* A; const float* B;
  long long a_rs, a_cs, a_bs, a_rbs;
  long long b_rs, b_cs, b_bs, b_rbs;
  int M, N, K, RB;
  float alpha;
  float* Cf; bf16* Cb; bf16* CbT;   // any subset; Cb [b][M][N], CbT [b][N][M]
  long long c_bs;
};
__global__ void __launch_bounds__(256) small_gemm_kernel(const SmallGemmP p) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16, each thread 4x4
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int rb = 0; rb < p.RB; ++rb) {
    const float* A = p.A + b * p.a_bs + rb * p.a_rbs;
    const float* B = p.B + b * p.b_bs + rb * p.b_rbs;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
      for (int i = threadIdx.x; i < 16 * 64; i += 256) {
        int kk, mm;
        if (p.a_cs == 1) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
        const int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < p.M && gk < p.K) ? A[gm * p.a_rs + gk * p.a_cs] : 0.f;
        int kb, nn;
        if (p.b_cs == 1) { nn = i & 63; kb = i >> 6; } else { kb = i & 15; nn = i >> 4; }
        const int gn = n0 + nn, gk2 = k0 + kb;
        Bs[kb][nn] = (gn < p.N && gk2 < p.K) ? B[gk2 * p.b_rs + gn * p.b_cs] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < p.M && gn < p.N) {
        const float v = acc[i][j] * p.alpha;
        const long long o = b * p.c_bs + static_cast<long long>(gm) * p.N + gn;
        if (p.Cf) p.Cf[o] = v;
        if (p.Cb) p.Cb[o] = __float2bfloat16(v);
        if (p.CbT) p.CbT[b * p.c_bs + static_cast<long long>(gn) * p.M + gm] = __float2bfloat16(v);
      }
    }
  }
}

// out[c] = alpha * sum_i part[i*stride + c]  (fixed order)
__global__ void reduce_partials_kernel(const float* __restrict__ part, int np, long long stride, int n, float alpha,
                                       float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  double s = 0.0;
  for (int i = 0; i < np; ++i) s += static_cast<double>(part[i * stride + c]);
  out[c] = static_cast<float>(s * alpha);
}

__global__ void copy_f32_kernel(const float* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = in[i];
}

int row_grid(long long rows, int S) {
  const int RPB = ROW_WARPS / S;
  long long blocks = (rows + RPB - 1) / RPB;
  const long long cap = 148 * 8;
  return static_cast<int>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host launchers
int transpose_cast(const void* in, void* out, int batch, int R, int S, int in_dtype, int out_dtype,
                   cudaStream_t stream) {
  if (batch <= 0 || R <= 0 || S <= 0) return set_error(GLF_ERR_INVALID, "transpose: empty");
  dim3 grid((S + 63) / 64, (R + 63) / 64, batch);
  if (grid.y > 65535 || grid.z > 65535) return set_error(GLF_ERR_INVALID, "transpose: grid too large");
  if (in_dtype == GLF_DTYPE_F32 && out_dtype == GLF_DTYPE_BF16)
    transpose_kernel<float, bf16><<<grid, 256, 0, stream>>>((const float*)in, (bf16*)out, R, S);
  else if (in_dtype == GLF_DTYPE_BF16 && out_dtype == GLF_DTYPE_BF16)
    transpose_kernel<bf16, bf16><<<grid, 256, 0, stream>>>((const bf16*)in, (bf16*)out, R, S);
  else if (in_dtype == GLF_DTYPE_BF16 && out_dtype == GLF_DTYPE_F32)
    transpose_kernel<bf16, float><<<grid, 256, 0, stream>>>((const bf16*)in, (float*)out, R, S);
  else
    transpose_kernel<float, float><<<grid, 256, 0, stream>>>((const float*)in, (float*)out, R, S);
  return check_cuda(cudaGetLastError(), "transpose launch");
}

int prep_weights(const glf_weights* w, int C, int Ci, bf16* wcat, bf16* wcatT, float* bcat, bf16* wz, bf16* wzT,
                 cudaStream_t stream) {
  const int total = 3 * Ci * C;
  int blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  prep_weights_kernel<<<blocks, 256, 0, stream>>>(w->theta_w, w->phi_w, w->g_w, w->theta_b, w->phi_b, w->g_b, w->wz_w,
                                                  wcat, wcatT, bcat, wz, wzT, C, Ci);
  return check_cuda(cudaGetLastError(), "prep_weights launch");
}

int bn_finalize(const float* part, int np, int C, double count, const glf_desc* d, const glf_weights* w, float* mean,
                float* rstd, float* a, float* b, cudaStream_t stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, np, C, count, w->bn_w, w->bn_b, d->eps_bn, d->momentum,
                                                          d->training, d->bn_layer, w->bn_running_mean,
                                                          w->bn_running_var,
                                                          reinterpret_cast<long long*>(w->bn_num_batches_tracked), mean,
                                                          rstd, a, b);
  return check_cuda(cudaGetLastError(), "bn_finalize launch");
}

int bn_res_ln_fwd(const bf16* U, const bf16* X, const float* a, const float* b, const float* lw, const float* lb,
                  void* Z, int z_dtype, float* mu, float* r, long long rows, int C, float eps, int accumulate,
                  cudaStream_t stream) {
  if (C % 8 != 0 || C > 2048) return set_error(GLF_ERR_INVALID, "LayerNorm kernel needs C %% 8 == 0 and C <= 2048");
  int S = (C + 255) / 256;
  while (ROW_WARPS % S != 0) ++S;
  const int grid = row_grid(rows, S);
  if (z_dtype == GLF_DTYPE_BF16)
    bn_res_ln_fwd_kernel<bf16><<<grid, ROW_THREADS, 0, stream>>>(U, X, a, b, lw, lb, (bf16*)Z, mu, r, rows, C, S, eps, accumulate);
  else
    bn_res_ln_fwd_kernel<float><<<grid, ROW_THREADS, 0, stream>>>(U, X, a, b, lw, lb, (float*)Z, mu, r, rows, C, S, eps, accumulate);
  return check_cuda(cudaGetLastError(), "bn_res_ln_fwd launch");
}

int bn_res_ln_bwd_blocks(long long rows, int C) {
  int S = (C + 255) / 256;
  while (ROW_WARPS % S != 0) ++S;
  const int g = row_grid(rows, S);
  return g < 148 * 2 ? g : 148 * 2;
}

int bn_res_ln_bwd(const void* dZ, int dz_dtype, const bf16* U, const bf16* X, const float* a, const float* b,
                  const float* mean, const float* rstd, const float* lw, const float* mu, const float* r, bf16* dV,
                  float* part, long long rows, int C, cudaStream_t stream) {
  if (C % 8 != 0 || C > 2048) return set_error(GLF_ERR_INVALID, "LayerNorm kernel needs C %% 8 == 0 and C <= 2048");
  int S = (C + 255) / 256;
  while (ROW_WARPS % S != 0) ++S;
  const int grid = bn_res_ln_bwd_blocks(rows, C);
  if (dz_dtype == GLF_DTYPE_BF16)
    bn_res_ln_bwd_kernel<bf16><<<grid, ROW_THREADS, 0, stream>>>((const bf16*)dZ, U, X, a, b, mean, rstd, lw, mu, r, dV, part, rows, C, S);
  else
    bn_res_ln_bwd_kernel<float><<<grid, ROW_THREADS, 0, stream>>>((const float*)dZ, U, X, a, b, mean, rstd, lw, mu, r, dV, part, rows, C, S);
  return check_cuda(cudaGetLastError(), "bn_res_ln_bwd launch");
}

int bn_bwd_finalize(const float* part, int np, int C, double count, const glf_desc* d, const glf_weights* w,
                    const float* mean, const float* rstd, const glf_grads* g, float* k1, float* k2, float* k3,
                    cudaStream_t stream) {
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, np, C, count, w->bn_w, mean, rstd, d->training,
                                                              d->bn_layer, g->ln_w, g->ln_b, g->bn_w, g->bn_b, g->wz_b,
                                                              k1, k2, k3);
  return check_cuda(cudaGetLastError(), "bn_bwd_finalize launch");
}

int bn_bwd_apply(const bf16* dV, const bf16* U, const float* k1, const float* k2, const float* k3, bf16* dU,
                 long long rows, int C, cudaStream_t stream) {
  const long long nvec = rows * C / 8;
  long long blocks = (nvec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_bwd_apply_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(dV, U, k1, k2, k3, dU, nvec, C);
  return check_cuda(cudaGetLastError(), "bn_bwd_apply launch");
}

int small_gemm(const float* A, long long a_rs, long long a_cs, long long a_bs, long long a_rbs, const float* B,
               long long b_rs, long long b_cs, long long b_bs, long long b_rbs, int batch, int RB, int M, int N, int K,
               float alpha, float* Cf, bf16* Cb, bf16* CbT, cudaStream_t stream) {
  SmallGemmP p;
  p.A = A; p.B = B;
  p.a_rs = a_rs; p.a_cs = a_cs; p.a_bs = a_bs; p.a_rbs = a_rbs;
  p.b_rs = b_rs; p.b_cs = b_cs; p.b_bs = b_bs; p.b_rbs = b_rbs;
  p.M = M; p.N = N; p.K = K; p.RB = RB; p.alpha = alpha;
  p.Cf = Cf; p.Cb = Cb; p.CbT = CbT;
  p.c_bs = static_cast<long long>(M) * N;
  dim3 grid((N + 63) / 64, (M + 63) / 64, batch);
  small_gemm_kernel<<<grid, 256, 0, stream>>>(p);
  return check_cuda(cudaGetLastError(), "small_gemm launch");
}

int reduce_partials(const float* part, int np, long long stride, int n, float alpha, float* out, cudaStream_t stream) {
  reduce_partials_kernel<<<(n + 127) / 128, 128, 0, stream>>>(part, np, stride, n, alpha, out);
  return check_cuda(cudaGetLastError(), "reduce_partials launch");
}

int copy_f32(const float* in, float* out, long long n, cudaStream_t stream) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  copy_f32_kernel<<<static_cast<int>(blocks < 1 ? 1 : blocks), 256, 0, stream>>>(in, out, n);
  return check_cuda(cudaGetLastError(), "copy launch");
}

}  // namespace glf
