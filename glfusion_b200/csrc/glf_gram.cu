// glf_gram.cu — the small per-sequence kernels of the Gram form of mode='dot' (oracle/tpavi_oracle.py:
// tpavi_dot_gram_form; R/models/ours.py:866-908 reassociated a second time).
//
// With the homogeneous token x~ = [x, 1] and W~ = [W | b], theta / phi / g are never formed per token:
//     S~_b = X~_b^T X~_b            (tcgen05 GEMM over the tokens, the row sums ride along as a side product)
//     M_b  = W~phi S~_b W~g^T / N ,  W'_b = Wz M_b^T ,  Q~_b = W'_b W~theta ,  U_b = X~_b Q~_b^T
// Everything between the two token-sized products is [C x C]-sized per sequence.  The tcgen05 GEMM does the products;
// the kernels here assemble / convert / combine those small matrices (augmented width Ca = C + 8, column C = the
// homogeneous coordinate, columns C+1.. are zero padding so that rows stay 16-byte aligned).
#include "glf_internal.h"
#include "glf_ptx.cuh"

namespace glf {

namespace {

// W~{theta,phi,g} = [W | b | 0] as bf16 [3][Ci][Ca]; Wz as bf16 [C][Ci]
__global__ void gram_prep_weights_kernel(const float* __restrict__ tw, const float* __restrict__ tb,
                                         const float* __restrict__ pw, const float* __restrict__ pb,
                                         const float* __restrict__ gw, const float* __restrict__ gb,
                                         const float* __restrict__ wz, bf16* __restrict__ waug, bf16* __restrict__ wzb,
                                         int C, int Ci, int Ca) {
  const int naug = 3 * Ci * Ca, nz = C * Ci;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < naug + nz; i += gridDim.x * blockDim.x) {
    if (i < naug) {
      const int m = i / (Ci * Ca), r = (i / Ca) % Ci, c = i % Ca;
      const float* W = m == 0 ? tw : (m == 1 ? pw : gw);
      const float* bb = m == 0 ? tb : (m == 1 ? pb : gb);
      const float v = c < C ? W[static_cast<size_t>(r) * C + c] : (c == C ? bb[r] : 0.f);
      waug[i] = __float2bfloat16(v);
    } else {
      wzb[i - naug] = __float2bfloat16(wz[i - naug]);
    }
  }
}

// Augmented [B][Ca][Ca] bf16 matrix [[M, colv], [rowv^T, corner]] (zero padded) from M [B][C][C] fp32 (split-K path)
__global__ void gram_assemble_aug_kernel(const float* __restrict__ Mf, const float* __restrict__ colv,
                                         const float* __restrict__ rowv, const float* __restrict__ rowscale,
                                         bf16* __restrict__ out, long long total, int C, int Ca, float corner) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % Ca);
    const int r = static_cast<int>((i / Ca) % Ca);
    const long long b = i / (static_cast<long long>(Ca) * Ca);
    float v = 0.f;
    if (r < C && j < C) v = Mf[(b * C + r) * C + j];
    else if (r < C && j == C) v = colv[b * C + r];
    else if (r == C && j < C) v = rowv[b * C + j];
    else if (r == C && j == C) v = corner;
    if (rowscale != nullptr && r < C) v *= rowscale[r];
    out[i] = __float2bfloat16(v);
  }
}

// Border of an augmented matrix whose [C x C] block a GEMM already wrote in bf16: column C = colv, row C = rowv,
// corner, zero padding.  grid = (batch, 8).
__global__ void gram_border_kernel(const float* __restrict__ colv, const float* __restrict__ rowv,
                                   const float* __restrict__ rowscale, bf16* __restrict__ out, int C, int Ca,
                                   float corner) {
  const long long b = blockIdx.x;
  bf16* M = out + b * static_cast<long long>(Ca) * Ca;
  const int pad = Ca - C;
  const int tid = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;
  for (int i = tid; i < C * pad; i += nth) {        // columns C.. of rows < C
    const int r = i / pad, j = C + i % pad;
    M[static_cast<long long>(r) * Ca + j] =
        __float2bfloat16(j == C ? colv[b * C + r] * (rowscale != nullptr ? rowscale[r] : 1.f) : 0.f);
  }
  for (int i = tid; i < pad * Ca; i += nth) {       // rows C..
    const int r = C + i / Ca, j = i % Ca;
    float v = 0.f;
    if (r == C) v = j < C ? rowv[b * C + j] : (j == C ? corner : 0.f);
    M[static_cast<long long>(r) * Ca + j] = __float2bfloat16(v);
  }
}

// c_b = W'_b b_theta in fp32 (the per-sequence bias of U): one warp per (b, r) row of W'
__global__ void gram_cvec_kernel(const bf16* __restrict__ Wp, const float* __restrict__ tb, float* __restrict__ cvec,
                                 long long rows, int Ci) {
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* w = Wp + row * Ci;
  float acc = 0.f;
  for (int i = lane * 2; i < Ci; i += 64) {
    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(w + i));
    acc = fmaf(v.x, tb[i], fmaf(v.y, tb[i + 1], acc));
  }
  acc = warp_sum(acc);
  if (lane == 0) cvec[row] = acc;
}

// Per-sequence operands of the backward products, from Q~ (bf16), c and the BatchNorm-backward coefficients
// (dU = k1 dV + k2 U + k3 per output channel = row r of Q~):
//   Qk[b]    = k2 Q~  with column C = k2 c + k3            (dQ~ = Qk S~ + k1 [R | rv] ; H = Q~^T Qk)
//   EF[b][0] = E  = k1 Q                                    (dX = dV E + ...)
// 8 columns per thread.
__global__ void gram_kprep_kernel(const bf16* __restrict__ Qb, const float* __restrict__ cvec,
                                  const float* __restrict__ k1, const float* __restrict__ k2,
                                  const float* __restrict__ k3, bf16* __restrict__ AK, bf16* __restrict__ EF,
                                  long long total8, int C, int Ca) {
  const int cv = Ca / 8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j0 = static_cast<int>(i % cv) * 8;
    const int r = static_cast<int>((i / cv) % C);
    const long long b = i / (static_cast<long long>(cv) * C);
    const float a1 = k1[r], a2 = k2[r];
    const uint4 qv = *reinterpret_cast<const uint4*>(Qb + i * 8);
    const uint32_t* q32 = reinterpret_cast<const uint32_t*>(&qv);
    uint32_t qk[4], ee[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 q = unpack_bf16(q32[t]);
      const int j = j0 + 2 * t;
      float x0 = j < C ? a2 * q.x : 0.f, x1 = j + 1 < C ? a2 * q.y : 0.f;
      if (j == C) x0 = fmaf(a2, cvec[b * C + r], k3[r]);
      if (j + 1 == C) x1 = fmaf(a2, cvec[b * C + r], k3[r]);
      qk[t] = pack_bf16(x0, x1);
      ee[t] = pack_bf16(a1 * q.x, a1 * q.y);
    }
    *reinterpret_cast<uint4*>(AK + i * 8) = make_uint4(qk[0], qk[1], qk[2], qk[3]);
    if (j0 < C)   // C % 8 == 0: a vector is entirely inside or outside the first C columns
      *reinterpret_cast<uint4*>(EF + (b * 2 * C + r) * static_cast<long long>(C) + j0) = make_uint4(ee[0], ee[1], ee[2], ee[3]);
  }
}

// F = (dS + dS^T + H)[:C, :C] -> EF[b][1]   (64 x 64 tiles, the transposed tile goes through shared memory; every
// thread moves bf16 pairs);  e = (dS[:, C] + dS[C, :] + H[:, C])[:C]  (blocks of the first tile row also emit their 64
// entries of e; only rows < C of dS and H are computed, the missing row dS[C, :] = b_phi^T dT is formed here)
__global__ void __launch_bounds__(256)
    gram_assemble_F_kernel(const bf16* __restrict__ G0, const bf16* __restrict__ Hf, const bf16* __restrict__ dT,
                           const bf16* __restrict__ wphi, bf16* __restrict__ EF, float* __restrict__ evec, int C,
                           int Ci, int Ca) {
  __shared__ float t[64][65];
  const long long b = blockIdx.z;
  const bf16* G = G0 + b * static_cast<long long>(C) * Ca;      // [C rows][Ca]
  const bf16* H = Hf != nullptr ? Hf + b * static_cast<long long>(C) * Ca : nullptr;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8: thread = column pair 2 tx, rows ty + 8 i
  // stage G[c0 + a][r0 + bb] at t[a][bb]  (C % 8 == 0: a pair is inside or outside the matrix together)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = ty + 8 * i, rr = c0 + a, cc = r0 + 2 * tx;
    float2 v = make_float2(0.f, 0.f);
    if (rr < C && cc < C) v = unpack_bf16(*reinterpret_cast<const uint32_t*>(G + static_cast<long long>(rr) * Ca + cc));
    t[a][2 * tx] = v.x;
    t[a][2 * tx + 1] = v.y;
  }
  __syncthreads();
  bf16* F = EF + (b * 2 + 1) * static_cast<long long>(C) * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = ty + 8 * i, rr = r0 + a, cc = c0 + 2 * tx;
    if (rr < C && cc < C) {
      float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(G + static_cast<long long>(rr) * Ca + cc));
      v.x += t[2 * tx][a];          // G[cc][rr]
      v.y += t[2 * tx + 1][a];      // G[cc + 1][rr]
      if (H != nullptr) {
        const float2 h = unpack_bf16(*reinterpret_cast<const uint32_t*>(H + static_cast<long long>(rr) * Ca + cc));
        v.x += h.x; v.y += h.y;
      }
      *reinterpret_cast<uint32_t*>(F + static_cast<long long>(rr) * C + cc) = pack_bf16(v.x, v.y);
    }
  }
  if (blockIdx.y == 0) {     // block-uniform: the first tile row also emits e for its 64 columns
    __syncthreads();
    float* part = &t[0][0];  // reuse: [8][64]
    const int cA = c0 + 2 * tx;
    float a0 = 0.f, a1 = 0.f;
    if (cA < C) {
      const bf16* dt = dT + b * static_cast<long long>(Ci) * Ca + cA;
      for (int i = ty; i < Ci; i += 8) {
        const float w = __bfloat162float(wphi[static_cast<long long>(i) * Ca + C]);
        const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(dt + static_cast<long long>(i) * Ca));
        a0 = fmaf(w, d.x, a0);
        a1 = fmaf(w, d.y, a1);
      }
    }
    part[ty * 64 + 2 * tx] = a0;
    part[ty * 64 + 2 * tx + 1] = a1;
    __syncthreads();
    const int tid = ty * 32 + tx;
    if (tid < 64 && c0 + tid < C) {
      const int c = c0 + tid;
      float v = __bfloat162float(G[static_cast<long long>(c) * Ca + C]);
      if (H != nullptr) v += __bfloat162float(H[static_cast<long long>(c) * Ca + C]);
#pragma unroll
      for (int i = 0; i < 8; ++i) v += part[i * 64 + tid];
      evec[b * C + c] = v;
    }
  }
}

// dW~ [3][Ci][Ca] fp32 -> the six caller-visible gradients (weight [Ci][C], bias [Ci])
__global__ void gram_unpack_grads_kernel(const float* __restrict__ dwaug, float* __restrict__ tw, float* __restrict__ tb,
                                         float* __restrict__ pw, float* __restrict__ pb, float* __restrict__ gw,
                                         float* __restrict__ gb, int C, int Ci, int Ca) {
  const int total = 3 * Ci * Ca;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int m = i / (Ci * Ca), r = (i / Ca) % Ci, c = i % Ca;
    float* W = m == 0 ? tw : (m == 1 ? pw : gw);
    float* bb = m == 0 ? tb : (m == 1 ? pb : gb);
    if (c < C) W[static_cast<size_t>(r) * C + c] = dwaug[i];
    else if (c == C) bb[r] = dwaug[i];
  }
}

int blocks_for(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  return static_cast<int>(b < 1 ? 1 : b);
}

}  // namespace

int gram_prep_weights(const glf_weights* w, int C, int Ci, int Ca, bf16* waug, bf16* wzb, cudaStream_t stream) {
  gram_prep_weights_kernel<<<blocks_for(3LL * Ci * Ca + static_cast<long long>(C) * Ci, 256), 256, 0, stream>>>(
      w->theta_w, w->theta_b, w->phi_w, w->phi_b, w->g_w, w->g_b, w->wz_w, waug, wzb, C, Ci, Ca);
  return check_cuda(cudaGetLastError(), "gram_prep_weights launch");
}

int gram_assemble_aug(const float* Mf, const float* colv, const float* rowv, const float* rowscale, bf16* out, int B,
                      int C, int Ca, float corner, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * Ca * Ca;
  gram_assemble_aug_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(Mf, colv, rowv, rowscale, out, total, C, Ca,
                                                                       corner);
  return check_cuda(cudaGetLastError(), "gram_assemble_aug launch");
}

int gram_border(const float* colv, const float* rowv, const float* rowscale, bf16* out, int B, int C, int Ca,
                float corner, cudaStream_t stream) {
  gram_border_kernel<<<dim3(B, 8), 256, 0, stream>>>(colv, rowv, rowscale, out, C, Ca, corner);
  return check_cuda(cudaGetLastError(), "gram_border launch");
}

int gram_cvec(const bf16* Wp, const float* theta_b, float* cvec, int B, int C, int Ci, cudaStream_t stream) {
  const long long rows = static_cast<long long>(B) * C;
  gram_cvec_kernel<<<static_cast<int>((rows * 32 + 255) / 256), 256, 0, stream>>>(Wp, theta_b, cvec, rows, Ci);
  return check_cuda(cudaGetLastError(), "gram_cvec launch");
}

int gram_kprep(const bf16* Qb, const float* cvec, const float* k1, const float* k2, const float* k3, bf16* AK, bf16* EF,
               int B, int C, int Ca, cudaStream_t stream) {
  const long long total8 = static_cast<long long>(B) * C * (Ca / 8);
  gram_kprep_kernel<<<blocks_for(total8, 256), 256, 0, stream>>>(Qb, cvec, k1, k2, k3, AK, EF, total8, C, Ca);
  return check_cuda(cudaGetLastError(), "gram_kprep launch");
}

int gram_assemble_F(const bf16* G0, const bf16* Hf, const bf16* dT, const bf16* wphi, bf16* EF, float* evec, int B,
                    int C, int Ci, int Ca, cudaStream_t stream) {
  if (B > 65535) return set_error(GLF_ERR_INVALID, "gram form: more than 65535 sequences per call");
  dim3 grid((C + 63) / 64, (C + 63) / 64, B);
  gram_assemble_F_kernel<<<grid, dim3(32, 8), 0, stream>>>(G0, Hf, dT, wphi, EF, evec, C, Ci, Ca);
  return check_cuda(cudaGetLastError(), "gram_assemble_F launch");
}

int gram_unpack_grads(const float* dwaug, const glf_grads* g, int C, int Ci, int Ca, cudaStream_t stream) {
  gram_unpack_grads_kernel<<<blocks_for(3LL * Ci * Ca, 256), 256, 0, stream>>>(dwaug, g->theta_w, g->theta_b, g->phi_w,
                                                                               g->phi_b, g->g_w, g->g_b, C, Ci, Ca);
  return check_cuda(cudaGetLastError(), "gram_unpack_grads launch");
}

}  // namespace glf
