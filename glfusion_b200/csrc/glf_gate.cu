// glf_gate.cu — the MLFM-specific work of the reference call site (R/models/ours.py:1802-1820, 1826-1827):
// centre-aware gate a = sigmoid(w * max_c sigmoid(cls) * sigmoid(ctr)), f4_local = f4 * a, and the view concat
// cat([f4_v.unsqueeze(2)], dim=2) for both the global (ungated) and local (gated) inputs.  In the reference this is
// ~10 element-wise / copy kernels per view; here one HBM-bound pass reads each per-view NCHW feature map once,
// transposes it through shared memory and writes both token-major [B, V*h*w, C] operands with 128-bit stores.
// The backward is the mirror pass: df4 = dXg + a * dXl (back to NCHW) plus the gate gradient da = sum_c f4 * dXl
// chained through the two sigmoids / the class max to the logits.
#include "glf_internal.h"
#include "glf_ptx.cuh"

namespace glf {

namespace {

constexpr int MAXV = 8;
struct ViewPtrs {
  const void* f4[MAXV];
  const float* cls[MAXV];
  const float* ctr[MAXV];
  void* df4[MAXV];
  float* dcls[MAXV];
  float* dctr[MAXV];
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return static_cast<float>(*p); }
// two adjacent spatial positions of one channel (NCHW side), 4-byte / 8-byte accesses
__device__ __forceinline__ void ld2(const bf16* p, float& a, float& b) {
  const float2 t = unpack_bf16(*reinterpret_cast<const uint32_t*>(p));
  a = t.x; b = t.y;
}
__device__ __forceinline__ void ld2(const float* p, float& a, float& b) {
  const float2 t = *reinterpret_cast<const float2*>(p);
  a = t.x; b = t.y;
}
__device__ __forceinline__ void st2(bf16* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack_bf16(a, b); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
// eight consecutive channels of one token (token-major side)
__device__ __forceinline__ void st8(bf16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ void st8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void ld8(const bf16* p, float (&f)[8]) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = unpack_bf16(u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void ld8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// grid: (ceil(hw/64), C/64 rounded up, B*V) ; block 256
template <typename TIO, int VEC, typename TX>
__global__ void __launch_bounds__(256)
    gate_concat_fwd_kernel(const ViewPtrs vp, TX* __restrict__ xg, TX* __restrict__ xl, float* __restrict__ gate,
                           int C, int V, int hw, int ncls, float weight) {
  __shared__ float tile[64][65];
  __shared__ float a_sm[64];
  const int bv = blockIdx.z, b = bv / V, v = bv % V;
  const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const TIO* f4 = reinterpret_cast<const TIO*>(vp.f4[v]) + static_cast<long long>(b) * C * hw;
  if (threadIdx.x < 64) {
    const int p = p0 + threadIdx.x;
    float a = 0.f;
    if (p < hw) {
      const float* cl = vp.cls[v] + static_cast<long long>(b) * ncls * hw + p;
      float lmax = cl[0];
      for (int k = 1; k < ncls; ++k) lmax = fmaxf(lmax, cl[static_cast<long long>(k) * hw]);
      const float m = sigmoidf_(lmax);  // max_c sigmoid(l_c) == sigmoid(max_c l_c)
      const float c = sigmoidf_(vp.ctr[v][static_cast<long long>(b) * hw + p]);
      a = sigmoidf_(weight * m * c);
      if (blockIdx.y == 0) gate[static_cast<long long>(bv) * hw + p] = a;
    }
    a_sm[threadIdx.x] = a;
  }
  if (VEC == 2) {  // hw even: every (c, even p) address is 4-byte aligned
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = ty; i < 64; i += 8) {
      const int c = c0 + i, p = p0 + 2 * tx;
      float a = 0.f, b = 0.f;
      if (c < C && p < hw) ld2(f4 + static_cast<long long>(c) * hw + p, a, b);
      tile[i][2 * tx] = a;
      tile[i][2 * tx + 1] = b;
    }
  } else {
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    for (int i = ty; i < 64; i += 4) {
      const int c = c0 + i, p = p0 + tx;
      tile[i][tx] = (c < C && p < hw) ? ldf(f4 + static_cast<long long>(c) * hw + p) : 0.f;
    }
  }
  __syncthreads();
  // write: 8 lanes x 16 B cover the 64 channels of one token row
  const int chunk = threadIdx.x & 7, pr = threadIdx.x >> 3;  // 32 rows per pass
  for (int pp = pr; pp < 64; pp += 32) {
    const int p = p0 + pp, c = c0 + chunk * 8;
    if (p < hw && c < C) {
      float f[8], g[8];
      const float a = a_sm[pp];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = tile[chunk * 8 + i][pp];
        g[i] = f[i] * a;
      }
      const long long o = (static_cast<long long>(b) * V * hw + static_cast<long long>(v) * hw + p) * C + c;
      st8(xg + o, f);
      st8(xl + o, g);
    }
  }
}

// grid: (ceil(hw/64), ceil(C/64), B*V) ; block 256.  Each block handles one 64-position x 64-channel tile:
// df4 = dXg + a * dXl back to NCHW, and its partial of the gate gradient da[p] = sum_c f4[c,p] * dXl[p,c], written to
// da_part[c_tile][b,v,p] (reduced in fixed order by gate_finish_kernel: deterministic, no atomics).
template <typename TIO, int VEC, typename TX>
__global__ void __launch_bounds__(256)
    gate_concat_bwd_kernel(const ViewPtrs vp, const float* __restrict__ gate, const TX* __restrict__ dxg,
                           const TX* __restrict__ dxl, float* __restrict__ da_part, int C, int V, int hw) {
  __shared__ float tg[64][65];
  __shared__ float tl[64][65];
  __shared__ float a_sm[64];
  __shared__ float da_sm[8][64];
  const int bv = blockIdx.z, b = bv / V, v = bv % V;
  const int p0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const TIO* f4 = reinterpret_cast<const TIO*>(vp.f4[v]) + static_cast<long long>(b) * C * hw;
  TIO* df4 = reinterpret_cast<TIO*>(vp.df4[v]) + static_cast<long long>(b) * C * hw;
  if (threadIdx.x < 64) {
    const int p = p0 + threadIdx.x;
    a_sm[threadIdx.x] = p < hw ? gate[static_cast<long long>(bv) * hw + p] : 0.f;
  }
  // NCHW side mapping: VEC==2 -> 32 position pairs x 8 channel rows ; VEC==1 -> 64 positions x 4 channel rows
  const int tx = VEC == 2 ? (threadIdx.x & 31) : (threadIdx.x & 63);
  const int ty = VEC == 2 ? (threadIdx.x >> 5) : (threadIdx.x >> 6);
  constexpr int NY = VEC == 2 ? 8 : 4;
  // issue the NCHW-side f4 loads first: they do not depend on shared memory
  float fv0[64 / NY], fv1[64 / NY];
#pragma unroll
  for (int k = 0; k < 64 / NY; ++k) {
    const int i = ty + k * NY;
    const int c = c0 + i;
    fv0[k] = fv1[k] = 0.f;
    if (VEC == 2) {
      const int p = p0 + 2 * tx;
      if (c < C && p < hw) ld2(f4 + static_cast<long long>(c) * hw + p, fv0[k], fv1[k]);
    } else {
      const int p = p0 + tx;
      if (c < C && p < hw) fv0[k] = ldf(f4 + static_cast<long long>(c) * hw + p);
    }
  }
  {
    const int chunk = threadIdx.x & 7, pr = threadIdx.x >> 3;
#pragma unroll
    for (int pp = pr; pp < 64; pp += 32) {
      const int p = p0 + pp, c = c0 + chunk * 8;
      float g8[8], l8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g8[i] = l8[i] = 0.f;
      if (p < hw && c < C) {
        const long long o = (static_cast<long long>(b) * V * hw + static_cast<long long>(v) * hw + p) * C + c;
        ld8(dxg + o, g8);
        ld8(dxl + o, l8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        tg[pp][chunk * 8 + i] = g8[i];
        tl[pp][chunk * 8 + i] = l8[i];
      }
    }
  }
  __syncthreads();
  float da0 = 0.f, da1 = 0.f;
  if (VEC == 2) {
    const int pl = 2 * tx;
    const float a0 = a_sm[pl], a1 = a_sm[pl + 1];
#pragma unroll
    for (int k = 0; k < 64 / NY; ++k) {
      const int i = ty + k * NY;
      const int c = c0 + i, p = p0 + pl;
      if (c < C && p < hw) {
        const float dl0 = tl[pl][i], dl1 = tl[pl + 1][i];
        st2(df4 + static_cast<long long>(c) * hw + p, tg[pl][i] + a0 * dl0, tg[pl + 1][i] + a1 * dl1);
        da0 = fmaf(fv0[k], dl0, da0);
        da1 = fmaf(fv1[k], dl1, da1);
      }
    }
    da_sm[ty][pl] = da0;
    da_sm[ty][pl + 1] = da1;
  } else {
    const float a = a_sm[tx];
#pragma unroll
    for (int k = 0; k < 64 / NY; ++k) {
      const int i = ty + k * NY;
      const int c = c0 + i, p = p0 + tx;
      if (c < C && p < hw) {
        const float dl = tl[tx][i];
        df4[static_cast<long long>(c) * hw + p] = static_cast<TIO>(tg[tx][i] + a * dl);
        da0 = fmaf(fv0[k], dl, da0);
      }
    }
    da_sm[ty][tx] = da0;
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int p = p0 + threadIdx.x;
    if (p < hw) {
      float dA = 0.f;
#pragma unroll
      for (int y = 0; y < NY; ++y) dA += da_sm[y][threadIdx.x];
      da_part[(static_cast<long long>(blockIdx.y) * gridDim.z + bv) * hw + p] = dA;
    }
  }
}

// da = sum over channel tiles (fixed order), then the chain a = sigmoid(w*m*c), m = sigmoid(max_k cls), c = sigmoid(ctr)
__global__ void gate_finish_kernel(const ViewPtrs vp, const float* __restrict__ gate, const float* __restrict__ da_part,
                                   int nct, int BV, int V, int hw, int ncls, float weight) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<long long>(BV) * hw) return;
  const int bv = static_cast<int>(idx / hw), p = static_cast<int>(idx % hw);
  const int b = bv / V, v = bv % V;
  float dA = 0.f;
  for (int t = 0; t < nct; ++t) dA += da_part[(static_cast<long long>(t) * BV + bv) * hw + p];
  const float* cl = vp.cls[v] + static_cast<long long>(b) * ncls * hw + p;
  float lmax = cl[0];
  int arg = 0;
  for (int k = 1; k < ncls; ++k) {
    const float l = cl[static_cast<long long>(k) * hw];
    if (l > lmax) { lmax = l; arg = k; }
  }
  const float m = sigmoidf_(lmax);
  const float c = sigmoidf_(vp.ctr[v][static_cast<long long>(b) * hw + p]);
  const float a = gate[idx];
  const float dt = dA * a * (1.f - a);
  const float dm = dt * weight * c, dc = dt * weight * m;
  vp.dctr[v][static_cast<long long>(b) * hw + p] = dc * c * (1.f - c);
  float* dcl = vp.dcls[v] + static_cast<long long>(b) * ncls * hw + p;
  for (int k = 0; k < ncls; ++k) dcl[static_cast<long long>(k) * hw] = (k == arg) ? dm * m * (1.f - m) : 0.f;
}

template <typename TX>
int launch_gate_fwd(dim3 grid, bool vec2, int io_dtype, const ViewPtrs& vp, void* xg, void* xl, float* gate, int C, int V,
                    int hw, int ncls, float weight, cudaStream_t stream) {
  if (io_dtype == GLF_DTYPE_BF16) {
    if (vec2) gate_concat_fwd_kernel<bf16, 2, TX><<<grid, 256, 0, stream>>>(vp, (TX*)xg, (TX*)xl, gate, C, V, hw, ncls, weight);
    else gate_concat_fwd_kernel<bf16, 1, TX><<<grid, 256, 0, stream>>>(vp, (TX*)xg, (TX*)xl, gate, C, V, hw, ncls, weight);
  } else {
    if (vec2) gate_concat_fwd_kernel<float, 2, TX><<<grid, 256, 0, stream>>>(vp, (TX*)xg, (TX*)xl, gate, C, V, hw, ncls, weight);
    else gate_concat_fwd_kernel<float, 1, TX><<<grid, 256, 0, stream>>>(vp, (TX*)xg, (TX*)xl, gate, C, V, hw, ncls, weight);
  }
  return check_cuda(cudaGetLastError(), "gate_concat_fwd launch");
}
template <typename TX>
int launch_gate_bwd(dim3 grid, bool vec2, int io_dtype, const ViewPtrs& vp, const float* gate, const void* dxg,
                    const void* dxl, float* da_part, int C, int V, int hw, cudaStream_t stream) {
  if (io_dtype == GLF_DTYPE_BF16) {
    if (vec2) gate_concat_bwd_kernel<bf16, 2, TX><<<grid, 256, 0, stream>>>(vp, gate, (const TX*)dxg, (const TX*)dxl, da_part, C, V, hw);
    else gate_concat_bwd_kernel<bf16, 1, TX><<<grid, 256, 0, stream>>>(vp, gate, (const TX*)dxg, (const TX*)dxl, da_part, C, V, hw);
  } else {
    if (vec2) gate_concat_bwd_kernel<float, 2, TX><<<grid, 256, 0, stream>>>(vp, gate, (const TX*)dxg, (const TX*)dxl, da_part, C, V, hw);
    else gate_concat_bwd_kernel<float, 1, TX><<<grid, 256, 0, stream>>>(vp, gate, (const TX*)dxg, (const TX*)dxl, da_part, C, V, hw);
  }
  return check_cuda(cudaGetLastError(), "gate_concat_bwd launch");
}

}  // namespace

int gate_concat_fwd(int B, int C, int V, int h, int w, int ncls, float weight, int io_dtype, int x_dtype,
                    const void* const* f4, const float* const* cls, const float* const* ctr, void* xg, void* xl,
                    float* gate, cudaStream_t stream) {
  if (V < 1 || V > MAXV) return set_error(GLF_ERR_INVALID, "gate_concat: 1 <= V <= %d", MAXV);
  if (C % 8 != 0) return set_error(GLF_ERR_INVALID, "gate_concat: C %% 8 != 0");
  ViewPtrs vp{};
  for (int v = 0; v < V; ++v) { vp.f4[v] = f4[v]; vp.cls[v] = cls[v]; vp.ctr[v] = ctr[v]; }
  const int hw = h * w;
  dim3 grid((hw + 63) / 64, (C + 63) / 64, B * V);
  if (grid.y > 65535 || grid.z > 65535) return set_error(GLF_ERR_INVALID, "gate_concat: grid too large");
  const bool vec2 = (hw % 2 == 0);
  if (x_dtype == GLF_DTYPE_BF16) return launch_gate_fwd<bf16>(grid, vec2, io_dtype, vp, xg, xl, gate, C, V, hw, ncls, weight, stream);
  return launch_gate_fwd<float>(grid, vec2, io_dtype, vp, xg, xl, gate, C, V, hw, ncls, weight, stream);
}

size_t gate_bwd_scratch_bytes(int B, int C, int V, int h, int w) {
  return static_cast<size_t>((C + 63) / 64) * B * V * h * w * sizeof(float);
}

int gate_concat_bwd(int B, int C, int V, int h, int w, int ncls, float weight, int io_dtype, int x_dtype,
                    const void* const* f4, const float* const* cls, const float* const* ctr, const float* gate,
                    const void* dxg, const void* dxl, void* const* df4, float* const* dcls, float* const* dctr,
                    float* da_part, cudaStream_t stream) {
  if (V < 1 || V > MAXV) return set_error(GLF_ERR_INVALID, "gate_concat: 1 <= V <= %d", MAXV);
  if (C % 8 != 0) return set_error(GLF_ERR_INVALID, "gate_concat: C %% 8 != 0");
  if (da_part == nullptr) return set_error(GLF_ERR_WORKSPACE, "gate_concat_bwd: scratch is NULL");
  ViewPtrs vp{};
  for (int v = 0; v < V; ++v) {
    vp.f4[v] = f4[v]; vp.cls[v] = cls[v]; vp.ctr[v] = ctr[v];
    vp.df4[v] = df4[v]; vp.dcls[v] = dcls[v]; vp.dctr[v] = dctr[v];
  }
  const int hw = h * w;
  const int nct = (C + 63) / 64;
  dim3 grid((hw + 63) / 64, nct, B * V);
  if (grid.y > 65535 || grid.z > 65535) return set_error(GLF_ERR_INVALID, "gate_concat: grid too large");
  const bool vec2 = (hw % 2 == 0);
  int rc = (x_dtype == GLF_DTYPE_BF16)
               ? launch_gate_bwd<bf16>(grid, vec2, io_dtype, vp, gate, dxg, dxl, da_part, C, V, hw, stream)
               : launch_gate_bwd<float>(grid, vec2, io_dtype, vp, gate, dxg, dxl, da_part, C, V, hw, stream);
  if (rc) return rc;
  const long long n = static_cast<long long>(B) * V * hw;
  gate_finish_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(vp, gate, da_part, nct, B * V, V, hw, ncls,
                                                                                weight);
  return check_cuda(cudaGetLastError(), "gate_finish launch");
}

}  // namespace glf
