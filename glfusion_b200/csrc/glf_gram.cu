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

// S~ [B][Ca][Ca] bf16 from S [B][C][C] fp32 and s [B][C] fp32:  [[S, s], [s^T, N]], zero padded
__global__ void gram_assemble_S_kernel(const float* __restrict__ Sf, const float* __restrict__ sf,
                                       bf16* __restrict__ Sa, long long total, int C, int Ca, float ntok) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % Ca);
    const int r = static_cast<int>((i / Ca) % Ca);
    const long long b = i / (static_cast<long long>(Ca) * Ca);
    float v = 0.f;
    if (r < C && j < C) v = Sf[(b * C + r) * C + j];
    else if (r < C && j == C) v = sf[b * C + r];
    else if (r == C && j < C) v = sf[b * C + j];
    else if (r == C && j == C) v = ntok;
    Sa[i] = __float2bfloat16(v);
  }
}

// Q~ fp32 [B][C][Ca] -> bf16 copy (GEMM operand) + c = Q~[:, :, C] kept in fp32 (the per-sequence bias of U)
__global__ void gram_convert_Q_kernel(const float* __restrict__ Qf, bf16* __restrict__ Qb, float* __restrict__ cvec,
                                      long long total, int C, int Ca) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % Ca);
    const float v = Qf[i];
    Qb[i] = __float2bfloat16(j <= C ? v : 0.f);
    if (j == C) cvec[i / Ca] = v;
  }
}

// dQ~ = k1 * [R | rv] + k2 * (Q~ S~) + k3 * [s | N]   ;   Qk = k2 * Q~   ;   E = k1 * Q  (-> EF[b][0])
// (k1, k2, k3: per output channel = row of Q~).  QSf may be null when k2 == k3 == 0 (eval-mode BN / no BN).
__global__ void gram_combine_dQ_kernel(const float* __restrict__ Rf, const float* __restrict__ rv,
                                       const float* __restrict__ QSf, const float* __restrict__ sf,
                                       const bf16* __restrict__ Qb, const float* __restrict__ k1,
                                       const float* __restrict__ k2, const float* __restrict__ k3,
                                       bf16* __restrict__ dQa, bf16* __restrict__ Qk, bf16* __restrict__ EF,
                                       long long total, int C, int Ca, float ntok) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % Ca);
    const int r = static_cast<int>((i / Ca) % C);
    const long long b = i / (static_cast<long long>(Ca) * C);
    const float a1 = k1[r];
    const float q = __bfloat162float(Qb[i]);
    float v = 0.f;
    if (j < C) v = a1 * Rf[(b * C + r) * C + j];
    else if (j == C) v = a1 * rv[b * C + r];
    if (QSf != nullptr && j <= C) {
      const float sa = j < C ? sf[b * C + j] : ntok;
      v = fmaf(k2[r], QSf[i], fmaf(k3[r], sa, v));
      Qk[i] = __float2bfloat16(k2[r] * q);
    }
    dQa[i] = __float2bfloat16(v);
    if (j < C) EF[(b * 2 * C + r) * C + j] = __float2bfloat16(a1 * q);
  }
}

// F = (G0 + G0^T)[:C, :C] -> EF[b][1]   (32 x 32 tiles, the transposed tile goes through shared memory)
__global__ void __launch_bounds__(256)
    gram_assemble_F_kernel(const float* __restrict__ G0, bf16* __restrict__ EF, int C, int Ca) {
  __shared__ float t[32][33];
  const long long b = blockIdx.z;
  const float* G = G0 + b * static_cast<long long>(Ca) * Ca;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = c0 + ty + 8 * i, cc = r0 + tx;   // element (rr, cc) of G lands at t[ty + 8 i][tx]
    t[ty + 8 * i][tx] = (rr < C && cc < C) ? G[static_cast<long long>(rr) * Ca + cc] : 0.f;
  }
  __syncthreads();
  bf16* F = EF + (b * 2 + 1) * static_cast<long long>(C) * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int rr = r0 + ty + 8 * i, cc = c0 + tx;
    if (rr < C && cc < C)
      F[static_cast<long long>(rr) * C + cc] = __float2bfloat16(G[static_cast<long long>(rr) * Ca + cc] + t[tx][ty + 8 * i]);
  }
}

// e[b][c] = G0[c][C] + G0[C][c] + sum_c' Q[c'][c] k3[c']
__global__ void gram_evec_kernel(const float* __restrict__ G0, const bf16* __restrict__ Qb,
                                 const float* __restrict__ k3, float* __restrict__ evec, int B, int C, int Ca,
                                 int use_k3) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<long long>(B) * C) return;
  const int c = static_cast<int>(i % C);
  const long long b = i / C;
  const float* G = G0 + b * static_cast<long long>(Ca) * Ca;
  float acc = G[static_cast<long long>(c) * Ca + C] + G[static_cast<long long>(C) * Ca + c];
  if (use_k3) {
    const bf16* Q = Qb + b * static_cast<long long>(C) * Ca;
    float a0 = 0.f, a1 = 0.f;
    int r = 0;
    for (; r + 1 < C; r += 2) {
      a0 = fmaf(__bfloat162float(Q[static_cast<long long>(r) * Ca + c]), k3[r], a0);
      a1 = fmaf(__bfloat162float(Q[static_cast<long long>(r + 1) * Ca + c]), k3[r + 1], a1);
    }
    if (r < C) a0 = fmaf(__bfloat162float(Q[static_cast<long long>(r) * Ca + c]), k3[r], a0);
    acc += a0 + a1;
  }
  evec[i] = acc;
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

int gram_assemble_S(const float* Sf, const float* sf, bf16* Sa, int B, int C, int Ca, float ntok, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * Ca * Ca;
  gram_assemble_S_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(Sf, sf, Sa, total, C, Ca, ntok);
  return check_cuda(cudaGetLastError(), "gram_assemble_S launch");
}

int gram_convert_Q(const float* Qf, bf16* Qb, float* cvec, int B, int C, int Ca, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * C * Ca;
  gram_convert_Q_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(Qf, Qb, cvec, total, C, Ca);
  return check_cuda(cudaGetLastError(), "gram_convert_Q launch");
}

int gram_combine_dQ(const float* Rf, const float* rv, const float* QSf, const float* sf, const bf16* Qb, const float* k1,
                    const float* k2, const float* k3, bf16* dQa, bf16* Qk, bf16* EF, int B, int C, int Ca, float ntok,
                    cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * C * Ca;
  gram_combine_dQ_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(Rf, rv, QSf, sf, Qb, k1, k2, k3, dQa, Qk, EF, total,
                                                                     C, Ca, ntok);
  return check_cuda(cudaGetLastError(), "gram_combine_dQ launch");
}

int gram_assemble_F(const float* G0, const bf16* Qb, const float* k3, int use_k3, bf16* EF, float* evec, int B, int C,
                    int Ca, cudaStream_t stream) {
  if (B > 65535) return set_error(GLF_ERR_INVALID, "gram form: more than 65535 sequences per call");
  dim3 grid((C + 31) / 32, (C + 31) / 32, B);
  gram_assemble_F_kernel<<<grid, dim3(32, 8), 0, stream>>>(G0, EF, C, Ca);
  int rc = check_cuda(cudaGetLastError(), "gram_assemble_F launch");
  if (rc) return rc;
  const long long n = static_cast<long long>(B) * C;
  gram_evec_kernel<<<static_cast<int>((n + 127) / 128), 128, 0, stream>>>(G0, Qb, k3, evec, B, C, Ca, use_k3);
  return check_cuda(cudaGetLastError(), "gram_evec launch");
}

int gram_unpack_grads(const float* dwaug, const glf_grads* g, int C, int Ci, int Ca, cudaStream_t stream) {
  gram_unpack_grads_kernel<<<blocks_for(3LL * Ci * Ca, 256), 256, 0, stream>>>(dwaug, g->theta_w, g->theta_b, g->phi_w,
                                                                               g->phi_b, g->g_w, g->g_b, C, Ci, Ca);
  return check_cuda(cudaGetLastError(), "gram_unpack_grads launch");
}

}  // namespace glf
