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

int flash_fwd(const bf16* P3, bf16* Y, float* lse, int B, int N, int Ci, void* scratch, cudaStream_t stream) {
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

int flash_bwd(const bf16* P3, const bf16* Y, const bf16* dY, const float* lse, bf16* dP3, float* delta, float* cs_t,
              float* cs_p, float* cs_g, int* cs_rows_out, int B, int N, int Ci, void* scratch, cudaStream_t stream) {
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

}  // namespace glf
