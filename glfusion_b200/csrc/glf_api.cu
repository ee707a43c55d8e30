// glf_api.cu — the C ABI (include/glfusion.h) and the host-side orchestration of one TPAVIModule forward/backward
// (R/models/ours.py:845-917) as a fixed sequence of kernel launches on the caller's stream.  No allocation, no
// synchronisation: the whole sequence is CUDA-graph capturable.
//
// mode='dot' is executed in its reassociated O(N C'^2) form, token-major throughout:
//   fwd:  P=[Theta|Phi|G] = X Wcat^T + b         tcgen05 GEMM (x read once, three projections concatenated)
//         M_b = Phi_b^T G_b / N                   tcgen05 GEMM, MN-major operands, split-K, fp32 red.add
//         W'_b = W_z M_b^T                        small fp32 SIMT product (C x C' per sequence)
//         U_b = Theta_b W'_b^T + b_z              tcgen05 GEMM + per-tile BatchNorm column statistics
//         BN finalise, then  Z = LN(BN(U) + X)    HBM-bound fused epilogue
//   bwd:  LN/BN backward (two HBM-bound passes), then six tcgen05 GEMMs (dTheta, dW', dPhi, dG, dWcat, dX) and two
//         small fp32 products (dW_z, dM).
// When the sequences are long against the channel count (N >= 5 C) mode='dot' runs in its Gram form instead
// (tpavi_fwd_gram / tpavi_bwd_gram below; oracle/tpavi_oracle.py: tpavi_dot_gram_form): theta / phi / g and dU are never
// materialised, the only token-sized products are S = X^T X, U = X Q^T, R = dV^T X and dX = [dV | X] [E ; F].
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "glf_internal.h"

namespace glf {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return set_error(GLF_ERR_DEVICE, "%s: %s", what, cudaGetErrorString(e));
}
int check_device_sm100() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) return set_error(GLF_ERR_DEVICE, "libglf_sm100a needs an sm_100 device (Blackwell B200); found sm_%d%d and there is no fallback path", major, minor);
  return 0;
}

namespace {

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Dims {
  long long B, N, rows;
  int C, Ci;
  int tiles_seq;   // m-tiles per sequence (128 tokens)
  int tiles_all;   // m-tiles over all rows
  bool pack_x;     // x must be repacked to token-major bf16
  bool pack_dz;
  bool dot;
  bool gram;       // mode='dot' in its Gram form (channel-space products only)
  int Ca;          // augmented width of the per-sequence matrices (gram)
};

int make_dims(const glf_desc* d, Dims* o) {
  if (d == nullptr) return set_error(GLF_ERR_INVALID, "null descriptor");
  if (d->B <= 0 || d->T <= 0 || d->H <= 0 || d->W <= 0) return set_error(GLF_ERR_INVALID, "empty input (B,T,H,W must be > 0)");
  if (d->C <= 0 || d->Ci <= 0 || d->C % 8 != 0 || d->Ci % 8 != 0)
    return set_error(GLF_ERR_INVALID, "C and inter_channels must be positive multiples of 8 (got %d, %d)", d->C, d->Ci);
  if (d->C > 2048) return set_error(GLF_ERR_INVALID, "C <= 2048 supported (got %d)", d->C);
  if (d->mode != GLF_MODE_DOT && d->mode != GLF_MODE_EMBEDDED)
    return set_error(GLF_ERR_UNSUPPORTED, "mode must be dot or embedded ('gaussian'/'concatenate' are not built by the reference network)");
  if (d->precision != GLF_PRECISION_BF16 && d->precision != GLF_PRECISION_F32X3)
    return set_error(GLF_ERR_INVALID, "bad precision (GLF_PRECISION_BF16 or GLF_PRECISION_F32X3)");
  if (d->precision == GLF_PRECISION_F32X3 && d->mode != GLF_MODE_DOT)
    return set_error(GLF_ERR_UNSUPPORTED, "GLF_PRECISION_F32X3 is implemented for mode='dot' only");
  if (d->io_dtype != GLF_DTYPE_BF16 && d->io_dtype != GLF_DTYPE_F32) return set_error(GLF_ERR_INVALID, "bad io_dtype");
  o->B = d->B;
  o->N = static_cast<long long>(d->T) * d->H * d->W;
  o->rows = o->B * o->N;
  if (o->rows * static_cast<long long>(d->C) >= (1LL << 40)) return set_error(GLF_ERR_INVALID, "problem too large");
  if (o->N > (1LL << 30)) return set_error(GLF_ERR_INVALID, "sequence too long");
  o->C = d->C;
  o->Ci = d->Ci;
  o->tiles_seq = gemm_tiles_m(static_cast<int>(o->N));
  o->tiles_all = gemm_tiles_m(static_cast<int>(o->rows > 0x7fffffff ? 0x7fffffff : o->rows));
  if (o->rows > 0x7fffff00LL) return set_error(GLF_ERR_INVALID, "too many tokens");
  o->pack_x = !(d->x_layout == GLF_LAYOUT_TOKEN && d->io_dtype == GLF_DTYPE_BF16);
  if (d->x_layout == GLF_LAYOUT_TOKEN && d->io_dtype == GLF_DTYPE_F32 && o->rows * d->C > 0x7fffffffLL)
    return set_error(GLF_ERR_INVALID, "token-major fp32 input too large for the cast kernel");
  o->pack_dz = d->dz_layout != GLF_LAYOUT_TOKEN;
  o->dot = d->mode == GLF_MODE_DOT;
  // reserved[1]: 0 = choose, 1 = token-space form (theta/phi/g per token), 2 = Gram form
  if (d->reserved[1] < 0 || d->reserved[1] > 2) return set_error(GLF_ERR_INVALID, "reserved[1] (dot algorithm) must be 0, 1 or 2");
  const bool gram_ok = o->dot && d->precision == GLF_PRECISION_BF16 && o->B <= 65535;
  if (d->reserved[1] == 2 && !gram_ok)
    return set_error(GLF_ERR_UNSUPPORTED, "the Gram form needs mode='dot', GLF_PRECISION_BF16 and B <= 65535");
  // per-sequence [C x C] products cost ~15 C^3 against 3.5 N C^2 of saved token-space FLOPs and 11 saved activation
  // passes.  Measured crossover on B200 (profiles/algo_crossover.py, r02_algo_crossover.txt): with the one-CTA-per-
  // sequence chain kernels (C = 256, C' = 128) the Gram form wins from N/C = 3 (1.00 vs 1.11 ms at N = 784, B = 512);
  // at C = 128 between 3 and 24; above C = 256 the token-space form (every GEMM on CTA pairs) still wins at N/C = 6
  // (C = 512: 1.37 vs 1.46 ms; C = 1024: 1.85 vs 1.96 ms).
  const long long thr = gram_chain_supported(d->C, d->Ci) ? 3 : (d->C > 256 ? 8 : 5);
  o->gram = gram_ok && (d->reserved[1] == 2 || (d->reserved[1] == 0 && o->N >= thr * d->C));
  o->Ca = gram_ca(d->C);
  return 0;
}

struct Saved {
  bf16 *xtok, *P, *Y, *U, *Mb, *Wp, *wcat, *wcatT, *wz, *wzT;
  float *lse, *bcat, *bn_mean, *bn_rstd, *bn_a, *bn_b, *ln_mu, *ln_r;
  bf16 *Sa, *T, *Qb, *waug;   // Gram form: S~ [B][Ca][Ca], T = W~phi S~ [B][Ci][Ca], Q~ [B][C][Ca], W~ [3][Ci][Ca]
  float *sfv, *cvec;          //            s = column sums of X [B][C], c = Q~[:, :, C] [B][C]
  float *tv;                  //            t = T~[:, :, C] [B][Ci] in fp32 (chain kernels)
};
size_t carve_saved(const Dims& m, void* base, Saved* s) {
  Carver c(base);
  const size_t rows = m.rows, C = m.C, Ci = m.Ci, B = m.B;
  s->xtok = m.pack_x ? c.take<bf16>(rows * C) : nullptr;
  s->Sa = s->T = s->Qb = s->waug = nullptr;
  s->sfv = s->cvec = s->tv = nullptr;
  if (m.gram) {
    const size_t Ca = m.Ca;
    s->P = s->Y = s->wcat = s->wcatT = s->wzT = nullptr;
    s->lse = s->bcat = nullptr;
    s->U = c.take<bf16>(rows * C);
    s->Sa = c.take<bf16>(B * Ca * Ca);
    s->T = c.take<bf16>(B * Ci * Ca);
    s->Mb = c.take<bf16>(B * Ci * Ci);
    s->Wp = c.take<bf16>(B * C * Ci);
    s->Qb = c.take<bf16>(B * C * Ca);
    s->waug = c.take<bf16>(3 * Ci * Ca);
    s->wz = c.take<bf16>(C * Ci);
    s->sfv = c.take<float>(B * C);
    s->cvec = c.take<float>(B * C);
    s->tv = c.take<float>(B * Ci);
    s->bn_mean = c.take<float>(C);
    s->bn_rstd = c.take<float>(C);
    s->bn_a = c.take<float>(C);
    s->bn_b = c.take<float>(C);
    s->ln_mu = c.take<float>(rows);
    s->ln_r = c.take<float>(rows);
    return (c.off + 255) & ~static_cast<size_t>(255);
  }
  s->P = c.take<bf16>(rows * 3 * Ci);
  s->U = c.take<bf16>(rows * C);
  if (m.dot) {
    s->Y = nullptr; s->lse = nullptr;
    s->Mb = c.take<bf16>(B * Ci * Ci);
    s->Wp = c.take<bf16>(B * C * Ci);
  } else {
    s->Y = c.take<bf16>(rows * Ci);
    s->lse = c.take<float>(rows);
    s->Mb = nullptr; s->Wp = nullptr;
  }
  s->wcat = c.take<bf16>(3 * Ci * C);
  s->wcatT = c.take<bf16>(3 * Ci * C);
  s->bcat = c.take<float>(3 * Ci);
  s->wz = c.take<bf16>(C * Ci);
  s->wzT = c.take<bf16>(C * Ci);
  s->bn_mean = c.take<float>(C);
  s->bn_rstd = c.take<float>(C);
  s->bn_a = c.take<float>(C);
  s->bn_b = c.take<float>(C);
  s->ln_mu = c.take<float>(rows);
  s->ln_r = c.take<float>(rows);
  return (c.off + 255) & ~static_cast<size_t>(255);
}

struct WsFwd { float* colstats; float* red1; float* Mf; void* attn; void* saved_fallback; float* Sf; };
size_t carve_ws_fwd(const Dims& m, void* base, WsFwd* w, size_t saved_bytes) {
  Carver c(base);
  const size_t np = 4 * (m.dot ? static_cast<size_t>(m.B) * m.tiles_seq : static_cast<size_t>(m.tiles_all));
  w->colstats = c.take<float>(np * 2 * m.C);   // one partial per (tile, 32-row quarter)
  w->red1 = c.take<float>(static_cast<size_t>(REDUCE_STAGE1_ROWS) * 2 * m.C);
  w->Sf = nullptr;
  if (m.gram) {
    w->Mf = nullptr; w->attn = nullptr;
    w->Sf = c.take<float>(static_cast<size_t>(m.B) * m.C * m.C);   // split-K accumulation target of S = X^T X
    w->saved_fallback = c.take<uint8_t>(saved_bytes);
    return (c.off + 255) & ~static_cast<size_t>(255);
  }
  w->Mf = m.dot ? c.take<float>(static_cast<size_t>(m.B) * m.Ci * m.Ci) : nullptr;  // split-K accumulation target
  w->attn = m.dot ? nullptr : c.take<uint8_t>(attn_scratch_bytes(m.B, m.N, false));
  w->saved_fallback = c.take<uint8_t>(saved_bytes);  // used when the caller passes saved == NULL (inference)
  return (c.off + 255) & ~static_cast<size_t>(255);
}

struct WsBwd {
  bf16 *dztok, *dV, *dU, *dP, *dY, *dxtok, *dM, *dWpb;
  float *dWpf, *part_ln, *k1, *k2, *k3, *cs_t, *cs_p, *cs_g, *red1, *dwcat, *delta;
  void* attn;
  // Gram form
  float *Rf, *rv, *evec, *dwaug, *dcv, *dtv, *wpart;
  bf16 *Rb, *AK, *dQa, *dT, *EF, *G0, *Hf;
};
size_t carve_ws_bwd(const glf_desc* d, const Dims& m, void* base, WsBwd* w) {
  Carver c(base);
  const size_t rows = m.rows, C = m.C, Ci = m.Ci, B = m.B;
  w->dztok = m.pack_dz ? c.take<bf16>(rows * C) : nullptr;
  w->dV = c.take<bf16>(rows * C);
  w->Rf = w->rv = w->evec = w->dwaug = w->dcv = w->dtv = w->wpart = nullptr;
  w->Rb = w->AK = w->dQa = w->dT = w->EF = w->G0 = w->Hf = nullptr;
  if (m.gram) {
    const size_t Ca = m.Ca;
    w->dU = w->dP = w->dY = nullptr;
    w->dWpf = w->cs_t = w->cs_p = w->cs_g = w->red1 = w->dwcat = w->delta = nullptr;
    w->attn = nullptr;
    w->dxtok = m.pack_x ? c.take<bf16>(rows * C) : nullptr;
    w->part_ln = c.take<float>(static_cast<size_t>(bn_res_ln_bwd_blocks(m.rows, m.C)) * 4 * C);
    w->k1 = c.take<float>(C);
    w->k2 = c.take<float>(C);
    w->k3 = c.take<float>(C);
    w->Rf = c.take<float>(B * C * C);        // split-K accumulation target of R = dV^T X
    w->rv = c.take<float>(B * C);
    w->Rb = c.take<bf16>(B * Ca * Ca);       // k1 [R | rv] (rows < C)
    w->AK = c.take<bf16>(B * C * Ca);        // Qk = k2 Q~ (column C: k2 c + k3) per sequence
    w->dQa = c.take<bf16>(B * C * Ca);
    w->dWpb = c.take<bf16>(B * C * Ci);
    w->dM = c.take<bf16>(B * Ci * Ci);
    w->dT = c.take<bf16>(B * Ci * Ca);
    w->G0 = c.take<bf16>(B * C * Ca);        // dS~[:C, :] (the three terms of F are rounded to bf16, summed in fp32)
    w->Hf = c.take<bf16>(B * C * Ca);
    w->EF = c.take<bf16>(B * 2 * C * C);
    w->evec = c.take<float>(B * C);
    w->dwaug = c.take<float>(3 * Ci * Ca);
    w->dcv = c.take<float>(B * C);
    w->dtv = c.take<float>(B * Ci);
    w->wpart = gram_chain_supported(m.C, m.Ci) ? c.take<float>(gram_wgrad_scratch_floats()) : nullptr;
    return (c.off + 255) & ~static_cast<size_t>(255);
  }
  w->dU = d->bn_layer ? c.take<bf16>(rows * C) : nullptr;
  w->dP = c.take<bf16>(rows * 3 * Ci);
  w->dxtok = m.pack_x ? c.take<bf16>(rows * C) : nullptr;
  if (m.dot) {
    w->dY = nullptr; w->delta = nullptr;
    w->dWpf = c.take<float>(B * C * Ci);
    w->dWpb = c.take<bf16>(B * C * Ci);
    w->dM = c.take<bf16>(B * Ci * Ci);
  } else {
    w->dY = c.take<bf16>(rows * Ci);
    w->delta = c.take<float>(rows);
    w->dWpf = nullptr; w->dWpb = nullptr; w->dM = nullptr;
  }
  w->attn = m.dot ? nullptr : c.take<uint8_t>(attn_scratch_bytes(m.B, m.N, true));
  w->part_ln = c.take<float>(static_cast<size_t>(bn_res_ln_bwd_blocks(m.rows, m.C)) * 4 * C);
  w->k1 = c.take<float>(C);
  w->k2 = c.take<float>(C);
  w->k3 = c.take<float>(C);
  const size_t np = 4 * B * m.tiles_seq;       // one partial per (tile, 32-row quarter)
  w->cs_t = c.take<float>(np * 2 * Ci);
  w->cs_p = c.take<float>(np * 2 * Ci);
  w->cs_g = c.take<float>(np * 2 * Ci);
  w->red1 = c.take<float>(static_cast<size_t>(REDUCE_STAGE1_ROWS) * 3 * Ci);
  w->dwcat = c.take<float>(3 * Ci * C);
  return (c.off + 255) & ~static_cast<size_t>(255);
}

int pick_split(long long tiles, int K) {
  const int kb = (K + 63) / 64;
  if (tiles >= 96) return 1;  // enough CTAs already: write the result directly, no atomics
  long long s = (2 * 148) / tiles;   // floor: tiles * s work items fill at most two full rounds of the 148 persistent CTAs
  if (s < 1) s = 1;
  if (s > kb) s = kb;
  return static_cast<int>(s);
}

// Split of a long token contraction (K = all rows) whose output tiles alone leave the last round of the 148 persistent
// CTAs mostly empty: e.g. dWcat at C = 2048 is 24 x 8 = 192 tiles of 128 x 256 = 1.3 rounds (0.65 of the machine);
// three K-slices make 576 work items = 3.9 rounds of a third the length (0.97).  Mirrors gemm()'s tile choice.
int pick_split_rounds(long long M, long long N, long long batch, int K) {
  const long long BNt = N <= 64 ? 64 : ((K >= 512 && N % 256 == 0) ? 256 : 128);
  const long long tiles = ((M + 127) / 128) * ((N + BNt - 1) / BNt) * batch;
  if (tiles < 96) return pick_split(tiles, K);
  const int kb = (K + 63) / 64;
  int best = 1;
  double best_eff = 0.0;
  for (int sp = 1; sp <= 4 && kb / sp >= 16; ++sp) {
    const long long items = tiles * sp;
    const double eff = static_cast<double>(items) / (148.0 * static_cast<double>((items + 147) / 148));
    if (eff > best_eff + 0.08) { best_eff = eff; best = sp; }     // a further slice has to buy 8 % of the machine
  }
  return best;
}

GemmOperand opnd(const void* p, int mn, long long ld, long long bs) {
  GemmOperand o;
  o.ptr = p; o.mn_major = mn; o.ld = ld; o.batch_stride = bs;
  return o;
}

#define GLF_TRY(expr)        \
  do {                       \
    int rc__ = (expr);       \
    if (rc__ != 0) return rc__; \
  } while (0)

// reserved[0] = 1: the LayerNorm stage (forward epilogue / first backward kernel) is run for MGFM and MLFM together
// by glf_fusion_ln_fwd / glf_fusion_ln_bwd instead of per block
bool defer_ln(const glf_desc* d) { return d->reserved[0] == 1; }

int check_pair(const glf_desc* d, const Dims& m) {
  if (d->precision != GLF_PRECISION_BF16 || d->io_dtype != GLF_DTYPE_BF16 || m.pack_x || d->dz_layout != GLF_LAYOUT_TOKEN)
    return set_error(GLF_ERR_UNSUPPORTED, "fused LayerNorm pair needs bf16 token-major activations");
  if (!ln_tma_supported(m.C) && !ln_pair_wide_supported(m.C))
    return set_error(GLF_ERR_UNSUPPORTED, "fused LayerNorm pair needs C %% 8 == 0 and C <= 2048");
  return 0;
}

int check_ptr(const void* p, const char* name) {
  if (p == nullptr) return set_error(GLF_ERR_INVALID, "%s is NULL", name);
  if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) return set_error(GLF_ERR_INVALID, "%s must be 16-byte aligned", name);
  return 0;
}


// ------------------------------------------------------------------------------------------------ Gram form of 'dot'
// Tuning aids (GLF_GRAM_WIDE / GLF_GRAM_BIG = 128 | 256): N tile of the Ca-wide small products / of U and dX.
// Measured on B200 at cfg2 (profiles/README.md): 128 / 128 is fastest (the 256-wide tile has only three ring stages
// for four k-blocks per tile), so that is the default.
int gram_wide_tile() {
  const char* e = getenv("GLF_GRAM_WIDE");
  const int v = e ? atoi(e) : 0;
  return (v == 128 || v == 256) ? v : 128;
}
int gram_big_tile() {
  const char* e = getenv("GLF_GRAM_BIG");
  const int v = e ? atoi(e) : 0;
  return (v == 128 || v == 256) ? v : 128;
}

// Token contraction of a sequence, S = A^T X (A = X or dV): both operands MN-major views of token-major
// activations, the column sums of A as the GEMM's row-sum side product; result = the augmented bf16 matrix
// [[S, colsum(A)], [rowv^T, corner]] of width Ca.  Enough sequences: the GEMM writes bf16 straight into it and a border
// kernel adds the homogeneous row / column; few long sequences: split-K into fp32, then one assembling pass.
int gram_token_contraction(const bf16* A, const bf16* X, bf16* out_aug, float* scratch, float* rowsum, const float* rowv,
                           const float* rowscale, float corner, int B, int N, int C, int Ca, cudaStream_t stream,
                           const GramPrep* prep = nullptr, bool* prep_done = nullptr) {
  if (prep_done != nullptr) *prep_done = false;
  const char* ek = getenv("GLF_GRAM_KERNEL");   // tuning aid: 0 = always the generic tile GEMM
  if (gram_contraction_supported(C) && !(ek && ek[0] == '0')) {
    // one CTA per sequence; few long sequences are split along the tokens to fill two rounds of the SMs
    int ks = 1;
    if (B < 96) {
      ks = (2 * 148) / B;
      if (ks < 1) ks = 1;
    }
    GLF_TRY(gram_contraction(A, X, out_aug, scratch, rowsum, rowscale, rowv, corner, 1, B, N, C, Ca, ks, stream, prep));
    if (prep != nullptr && prep_done != nullptr) *prep_done = true;
    const int kb = (N + 63) / 64;
    if ((ks > kb ? kb : ks) > 1)
      return gram_assemble_aug(scratch, rowsum, rowv ? rowv : rowsum, rowscale, out_aug, B, C, Ca, corner, stream);
    return 0;   // one CTA per sequence: the kernel wrote the border as well
  }
  GemmArgs g;
  g.A = opnd(A, 1, C, static_cast<long long>(N) * C);
  g.B = opnd(X, 1, C, static_cast<long long>(N) * C);
  g.M = C; g.N = C; g.K = N; g.batch = B;
  g.bn_hint = 128;
  g.rowsum = rowsum; g.rowsum_stride = C;
  g.split_k = pick_split(static_cast<long long>(B) * ((C + 127) / 128) * ((C + 127) / 128), N);
  if (g.split_k > 1 || rowscale != nullptr) {   // (a per-row scale needs the fp32 result: the GEMM's alpha is a scalar)
    if (g.split_k > 1) {
      GLF_TRY(check_cuda(cudaMemsetAsync(scratch, 0, sizeof(float) * B * C * C, stream), "memset S"));
      GLF_TRY(check_cuda(cudaMemsetAsync(rowsum, 0, sizeof(float) * B * C, stream), "memset s"));
    }
    g.out_kind = g.split_k > 1 ? 2 : 1;
    g.D = scratch; g.ldd = C; g.strideD = static_cast<long long>(C) * C;
    GLF_TRY(gemm(g, stream));
    return gram_assemble_aug(scratch, rowsum, rowv ? rowv : rowsum, rowscale, out_aug, B, C, Ca, corner, stream);
  }
  g.out_kind = 0;
  g.D = out_aug; g.ldd = Ca; g.strideD = static_cast<long long>(Ca) * Ca;
  GLF_TRY(gemm(g, stream));
  return gram_border(rowsum, rowv ? rowv : rowsum, nullptr, out_aug, B, C, Ca, corner, stream);
}

// D = A0 B0^T + A1 B1^T (+ bias + addend), bf16 output: both products accumulate into one TMEM tile when each operand
// pair can be described as two "limbs" of one tensor map (their distance in memory is the limb stride); otherwise two
// passes, the second adding onto the first in place.  g.A / g.B carry layout, leading dimension and batch stride.
int gemm_pair2(GemmArgs g, const bf16* A0, const bf16* A1, const bf16* B0, const bf16* B1, cudaStream_t stream) {
  const long long dA = A1 - A0, dB = B1 - B0;
  auto ok = [](long long d) {
    const long long a = d < 0 ? -d : d;
    return a != 0 && a % 8 == 0 && a < (1LL << 38);
  };
  if (ok(dA) && ok(dB)) {
    g.A.ptr = dA > 0 ? A0 : A1; g.A.limb_stride = dA > 0 ? dA : -dA;
    g.B.ptr = dB > 0 ? B0 : B1; g.B.limb_stride = dB > 0 ? dB : -dB;
    g.npairs = 2;
    g.pairA[0] = dA > 0 ? 0 : 1; g.pairA[1] = 1 - g.pairA[0];
    g.pairB[0] = dB > 0 ? 0 : 1; g.pairB[1] = 1 - g.pairB[0];
    return gemm(g, stream);
  }
  g.A.ptr = A0; g.B.ptr = B0;
  GLF_TRY(gemm(g, stream));
  g.A.ptr = A1; g.B.ptr = B1;
  g.bias = nullptr;
  g.addend = reinterpret_cast<const bf16*>(g.D); g.ld_add = g.ldd; g.stride_add = g.strideD;
  return gemm(g, stream);
}

int tpavi_fwd_gram(const glf_desc* d, const Dims& m, const void* x, const glf_weights* w, void* z, const Saved& s,
                   const WsFwd& wf, cudaStream_t stream) {
  const int C = m.C, Ci = m.Ci, Ca = m.Ca, C1 = m.C + 1;
  const int N = static_cast<int>(m.N), B = static_cast<int>(m.B);
  const long long CaCa = static_cast<long long>(Ca) * Ca, CiCa = static_cast<long long>(Ci) * Ca;
  const long long CCa = static_cast<long long>(C) * Ca, CiCi = static_cast<long long>(Ci) * Ci;
  const long long CCi = static_cast<long long>(C) * Ci;
  const int wide = gram_wide_tile();
  const bf16* X = reinterpret_cast<const bf16*>(x);
  if (m.pack_x) {
    if (d->x_layout == GLF_LAYOUT_NCTHW)
      GLF_TRY(transpose_cast(x, s.xtok, B, C, N, d->io_dtype, GLF_DTYPE_BF16, stream));
    else
      GLF_TRY(transpose_cast(x, s.xtok, 1, 1, static_cast<int>(m.rows * C), d->io_dtype, GLF_DTYPE_BF16, stream));
    X = s.xtok;
  }
  // S_b = X_b^T X_b, s_b = X_b^T 1   ->   S~_b
  //   (the fp32 -> bf16 weight preparation for the chain that follows rides along as extra CTAs of the same launch)
  GramPrep prep;
  prep.tw = w->theta_w; prep.tb = w->theta_b; prep.pw = w->phi_w; prep.pb = w->phi_b; prep.gw = w->g_w; prep.gb = w->g_b;
  prep.wz = w->wz_w; prep.waug = s.waug; prep.wzb = s.wz; prep.C = C; prep.Ci = Ci; prep.Ca = Ca;
  bool prep_done = false;
  GLF_TRY(gram_token_contraction(X, X, s.Sa, wf.Sf, s.sfv, nullptr, nullptr, static_cast<float>(N), B, N, C, Ca, stream,
                                 &prep, &prep_done));
  if (!prep_done) GLF_TRY(gram_prep_weights(w, C, Ci, Ca, s.waug, s.wz, stream));
  if (gram_chain_supported(C, Ci)) {
    // T~, M, W', Q~ and c of every sequence in one launch, one CTA per sequence (glf_chain.cu)
    GLF_TRY(gram_chain_fwd(s.Sa, s.sfv, s.waug, s.wz, w->phi_b, w->g_b, w->theta_b, s.T, s.Mb, s.Wp, s.Qb, s.cvec, s.tv, B,
                           N, stream));
  } else {
    {  // T_b = W~phi S~_b                      [Ci x Ca]   (= Phi_b^T X~_b; S~ is symmetric)
      GemmArgs g;
      g.A = opnd(s.waug + CiCa, 0, Ca, 0);
      g.B = opnd(s.Sa, 0, Ca, CaCa);
      g.B.rows = C1;
      g.M = Ci; g.N = Ca; g.K = C1; g.batch = B;
      g.bn_hint = wide;
      g.D = s.T; g.ldd = Ca; g.strideD = CiCa;
      GLF_TRY(gemm(g, stream));
    }
    {  // M_b = T_b W~g^T / N                   [Ci x Ci]   (= Phi_b^T G_b / N)
      GemmArgs g;
      g.A = opnd(s.T, 0, Ca, CiCa);
      g.B = opnd(s.waug + 2 * CiCa, 0, Ca, 0);
      g.M = Ci; g.N = Ci; g.K = C1; g.batch = B;
      g.alpha = 1.f / static_cast<float>(N);
      g.D = s.Mb; g.ldd = Ci; g.strideD = CiCi;
      GLF_TRY(gemm(g, stream));
    }
    {  // W'_b = Wz M_b^T                       [C x Ci]
      GemmArgs g;
      g.A = opnd(s.wz, 0, Ci, 0);
      g.B = opnd(s.Mb, 0, Ci, CiCi);
      g.M = C; g.N = Ci; g.K = Ci; g.batch = B;
      g.D = s.Wp; g.ldd = Ci; g.strideD = CCi;
      GLF_TRY(gemm(g, stream));
    }
    {  // Q~_b = W'_b W~theta                   [C x Ca]    (B operand = W~theta read MN-major: stored [K = i][rows = a])
      GemmArgs g;
      g.A = opnd(s.Wp, 0, Ci, CCi);
      g.B = opnd(s.waug, 1, Ca, 0);
      g.B.rows = C1;
      g.M = C; g.N = Ca; g.K = Ci; g.batch = B;
      g.bn_hint = wide;
      g.D = s.Qb; g.ldd = Ca; g.strideD = CCa;
      GLF_TRY(gemm(g, stream));
    }
    GLF_TRY(gram_cvec(s.Wp, w->theta_b, s.cvec, B, C, Ci, stream));   // c_b = W'_b b_theta, kept in fp32
  }
  int np = 0;
  {  // U_b = X_b Q_b^T + c_b   (+ BatchNorm column statistics); bz stays folded into the BN affine
    GemmArgs g;
    g.A = opnd(X, 0, C, static_cast<long long>(N) * C);
    g.B = opnd(s.Qb, 0, Ca, CCa);
    g.M = N; g.N = C; g.K = C; g.batch = B;
    if (C % 256 == 0) g.bn_hint = gram_big_tile();
    g.bias = s.cvec; g.bias_stride = C;
    g.D = s.U; g.ldd = C; g.strideD = static_cast<long long>(N) * C;
    g.colstats = (d->training && d->bn_layer) ? wf.colstats : nullptr;
    g.colstats_rows = &np;
    GLF_TRY(gemm(g, stream));
  }
  const float* bn_part = wf.colstats;
  if (d->training && d->bn_layer && np > 4 * REDUCE_STAGE1_ROWS) {
    long long rs = 2LL * C;
    const int np_in[1] = {np};
    GLF_TRY(reduce_stage1(wf.colstats, nullptr, nullptr, 1, np_in, &np, &rs, C, 2, C, wf.red1, stream));
    bn_part = wf.red1;
  }
  GLF_TRY(bn_finalize(bn_part, np, C, static_cast<double>(m.rows), d, w, w->wz_b, s.bn_mean, s.bn_rstd, s.bn_a, s.bn_b,
                      stream));
  if (defer_ln(d)) return 0;
  GLF_TRY(bn_res_ln_fwd(s.U, X, GLF_DTYPE_BF16, s.bn_a, s.bn_b, w->ln_w, w->ln_b, z, d->io_dtype, s.ln_mu, s.ln_r,
                        m.rows, C, d->eps_ln, d->accumulate, stream));
  return 0;
}

// Continues after the LayerNorm backward and bn_bwd_finalize: wb.dV, wb.k1..k3 are valid.
int tpavi_bwd_gram(const glf_desc* d, const Dims& m, const bf16* X, const glf_weights* w, const Saved& s, const WsBwd& wb,
                   void* dx, const glf_grads* g_, cudaStream_t stream) {
  const int C = m.C, Ci = m.Ci, Ca = m.Ca, C1 = m.C + 1;
  const int N = static_cast<int>(m.N), B = static_cast<int>(m.B);
  const long long CaCa = static_cast<long long>(Ca) * Ca, CiCa = static_cast<long long>(Ci) * Ca;
  const long long CCa = static_cast<long long>(C) * Ca, CiCi = static_cast<long long>(Ci) * Ci;
  const long long CCi = static_cast<long long>(C) * Ci, CC = static_cast<long long>(C) * C;
  const float invN = 1.f / static_cast<float>(N);
  const int wide = gram_wide_tile();   // N tile of the products whose output is Ca (= C + 8) columns wide
  const bool bn_train = d->bn_layer && d->training;   // k2, k3 != 0 only then
  // k1 * [R_b | rv_b]  with R_b = dV_b^T X_b, rv_b = dV_b^T 1: the BatchNorm-backward scale k1 (per channel of dV =
  // row of R) is applied in fp32 by the contraction's epilogue
  GLF_TRY(gram_token_contraction(wb.dV, X, wb.Rb, wb.Rf, wb.rv, s.sfv, wb.k1, static_cast<float>(N), B, N, C, Ca, stream));
  // One CTA per sequence runs dQ~ -> dW' -> dM -> dT~ -> F (and E, e) in a single launch (glf_chain.cu); the four
  // weight-gradient sums over the sequences below then read its dQ~, dW', dM / N and dT~.
  const bool chain = gram_chain_supported(C, Ci);
  if (chain)
    GLF_TRY(gram_chain_bwd(s.Sa, s.Qb, s.waug, s.wz, wb.Rb, s.sfv, s.cvec, wb.rv, wb.k1, wb.k2, wb.k3, w->theta_b,
                           w->phi_b, w->g_b, bn_train ? 1 : 0, wb.dQa, wb.dWpb, wb.dM, wb.dT, wb.EF, wb.evec, wb.dcv,
                           wb.dtv, B, N, stream));
  if (chain)   // dW~theta, dWz, dW~g, dW~phi: one launch of K-concatenated products + a fixed-order reduction
    GLF_TRY(gram_wgrad(s.Wp, wb.dQa, wb.dWpb, s.Mb, wb.dM, s.T, wb.dT, s.Sa, wb.dcv, s.tv, wb.dtv, s.sfv, wb.wpart, g_, B, N,
                       stream));
  // Qk = k2 Q~ (column C: k2 c + k3), E = k1 Q
  if (!chain) GLF_TRY(gram_kprep(s.Qb, s.cvec, wb.k1, wb.k2, wb.k3, wb.AK, wb.EF, B, C, Ca, stream));
  if (!chain) {  // dQ~_b = dU_b^T X~_b = Qk_b S~_b + k1 [R_b | rv_b]     [C x Ca]   (S~ is symmetric; the second term is the addend)
    GemmArgs g;
    g.A = opnd(wb.AK, 0, Ca, CCa);
    g.B = opnd(s.Sa, 0, Ca, CaCa);
    g.B.rows = C1;
    g.M = C; g.N = Ca; g.K = C1; g.batch = B;
    g.bn_hint = wide;
    g.addend = wb.Rb; g.ld_add = Ca; g.stride_add = CaCa;
    g.D = wb.dQa; g.ldd = Ca; g.strideD = CCa;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) {  // dW'_b = dQ~_b W~theta^T               [C x Ci]
    GemmArgs g;
    g.A = opnd(wb.dQa, 0, Ca, CCa);
    g.B = opnd(s.waug, 0, Ca, 0);
    g.M = C; g.N = Ci; g.K = C1; g.batch = B;
    g.D = wb.dWpb; g.ldd = Ci; g.strideD = CCi;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) GLF_TRY(check_cuda(cudaMemsetAsync(wb.dwaug, 0, sizeof(float) * 3 * CiCa, stream), "memset dW~"));
  if (!chain) {  // dW~theta = sum_b W'_b^T dQ~_b         [Ci x Ca]   (batch reduced by fp32 red.add)
    GemmArgs g;
    g.A = opnd(s.Wp, 1, Ci, CCi);
    g.B = opnd(wb.dQa, 1, Ca, CCa);
    g.B.rows = C1;
    g.M = Ci; g.N = Ca; g.K = C; g.batch = B;
    g.bn_hint = wide;
    g.out_kind = 2;
    g.D = wb.dwaug; g.ldd = Ca; g.strideD = 0;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) GLF_TRY(check_cuda(cudaMemsetAsync(g_->wz_w, 0, sizeof(float) * CCi, stream), "memset dWz"));
  if (!chain) {  // dWz = sum_b dW'_b M_b
    GemmArgs g;
    g.A = opnd(wb.dWpb, 0, Ci, CCi);
    g.B = opnd(s.Mb, 1, Ci, CiCi);
    g.M = C; g.N = Ci; g.K = Ci; g.batch = B;
    g.out_kind = 2;
    g.D = g_->wz_w; g.ldd = Ci; g.strideD = 0;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) {  // dM_b = dW'_b^T Wz                     [Ci x Ci]
    GemmArgs g;
    g.A = opnd(wb.dWpb, 1, Ci, CCi);
    g.B = opnd(s.wz, 1, Ci, 0);
    g.M = Ci; g.N = Ci; g.K = C; g.batch = B;
    g.D = wb.dM; g.ldd = Ci; g.strideD = CiCi;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) {  // dW~g = sum_b (dM_b / N)^T T_b         [Ci x Ca]
    GemmArgs g;
    g.A = opnd(wb.dM, 1, Ci, CiCi);
    g.B = opnd(s.T, 1, Ca, CiCa);
    g.B.rows = C1;
    g.M = Ci; g.N = Ca; g.K = Ci; g.batch = B;
    g.bn_hint = wide;
    g.alpha = invN;
    g.out_kind = 2;
    g.D = wb.dwaug + 2 * CiCa; g.ldd = Ca; g.strideD = 0;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) {  // dT_b = (dM_b / N) W~g                 [Ci x Ca]
    GemmArgs g;
    g.A = opnd(wb.dM, 0, Ci, CiCi);
    g.B = opnd(s.waug + 2 * CiCa, 1, Ca, 0);
    g.B.rows = C1;
    g.M = Ci; g.N = Ca; g.K = Ci; g.batch = B;
    g.bn_hint = wide;
    g.alpha = invN;
    g.D = wb.dT; g.ldd = Ca; g.strideD = CiCa;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) {  // dW~phi = sum_b dT_b S~_b              [Ci x Ca]
    GemmArgs g;
    g.A = opnd(wb.dT, 0, Ca, CiCa);
    g.B = opnd(s.Sa, 0, Ca, CaCa);
    g.B.rows = C1;
    g.M = Ci; g.N = Ca; g.K = C1; g.batch = B;
    g.bn_hint = wide;
    g.out_kind = 2;
    g.D = wb.dwaug + CiCa; g.ldd = Ca; g.strideD = 0;
    GLF_TRY(gemm(g, stream));
  }
  if (!chain) {  // G0_b = dS~_b[:C, :] = W_phi^T dT_b     [C x Ca]   (row C is only needed for e: formed in gram_assemble_F)
    GemmArgs g;
    g.A = opnd(s.waug + CiCa, 1, Ca, 0);
    g.B = opnd(wb.dT, 1, Ca, CiCa);
    g.B.rows = C1;
    g.M = C; g.N = Ca; g.K = Ci; g.batch = B;
    g.bn_hint = wide;
    g.D = wb.G0; g.ldd = Ca; g.strideD = CCa;
    GLF_TRY(gemm(g, stream));
  }
  if (bn_train && !chain) {  // H_b = Q_b^T Qk_b    [C x Ca]  (dU Q = dV E + X H[:C,:C] + 1 H[:C,C]^T: the k2 U + k3 part of dU)
    GemmArgs g;
    g.A = opnd(s.Qb, 1, Ca, CCa);
    g.B = opnd(wb.AK, 1, Ca, CCa);
    g.B.rows = C1;
    g.M = C; g.N = Ca; g.K = C; g.batch = B;
    g.bn_hint = wide;
    g.D = wb.Hf; g.ldd = Ca; g.strideD = CCa;
    GLF_TRY(gemm(g, stream));
  }
  // F = (G0 + G0^T + H)[:C, :C] ;  e = (G0[:, C] + G0[C, :] + H[:, C])[:C]
  if (!chain) GLF_TRY(gram_assemble_F(wb.G0, bn_train ? wb.Hf : nullptr, wb.dT, s.waug + CiCa, wb.EF, wb.evec, B, C, Ci, Ca, stream));
  {  // dX_b = dV_b E_b + X_b F_b + 1 e_b^T + dV_b
    GemmArgs g;
    g.A = opnd(nullptr, 0, C, static_cast<long long>(N) * C);
    g.B = opnd(nullptr, 1, C, 2 * CC);
    g.M = N; g.N = C; g.K = C; g.batch = B;
    if (C % 256 == 0) g.bn_hint = gram_big_tile();
    g.bias = wb.evec; g.bias_stride = C;
    if (!(chain && gram_residual_in_E())) { g.addend = wb.dV; g.ld_add = C; g.stride_add = static_cast<long long>(N) * C; }
    g.D = m.pack_x ? static_cast<void*>(wb.dxtok) : dx; g.ldd = C; g.strideD = static_cast<long long>(N) * C;
    GLF_TRY(gemm_pair2(g, wb.dV, X, wb.EF, wb.EF + CC, stream));
  }
  if (!chain) GLF_TRY(gram_unpack_grads(wb.dwaug, g_, C, Ci, Ca, stream));
  if (m.pack_x) {
    if (d->x_layout == GLF_LAYOUT_NCTHW)
      GLF_TRY(transpose_cast(wb.dxtok, dx, B, N, C, GLF_DTYPE_BF16, d->io_dtype, stream));
    else
      GLF_TRY(transpose_cast(wb.dxtok, dx, 1, 1, static_cast<int>(m.rows * C), GLF_DTYPE_BF16, d->io_dtype, stream));
  }
  return 0;
}

}  // namespace
}  // namespace glf

using namespace glf;

extern "C" {

GLF_API int glf_version(void) { return GLF_VERSION; }
GLF_API const char* glf_last_error(void) { return g_err; }

GLF_API int glf_tpavi_sizes(const glf_desc* d, glf_sizes* out) {
  Dims m;
  GLF_TRY(make_dims(d, &m));
  if (out == nullptr) return set_error(GLF_ERR_INVALID, "out is NULL");
  if (d->precision == GLF_PRECISION_F32X3) return tpavi_sizes_f32x3(d, out);
  Saved s; WsFwd wf; WsBwd wb;
  out->saved_bytes = carve_saved(m, nullptr, &s);
  out->ws_fwd_bytes = carve_ws_fwd(m, nullptr, &wf, out->saved_bytes);
  out->ws_bwd_bytes = carve_ws_bwd(d, m, nullptr, &wb);
  return 0;
}

GLF_API int glf_tpavi_fwd(const glf_desc* d, const void* x, const glf_weights* w, void* z, void* saved, void* ws,
                  glf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Dims m;
  GLF_TRY(make_dims(d, &m));
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(x, "x"));
  GLF_TRY(check_ptr(z, "z"));
  GLF_TRY(check_ptr(ws, "ws"));
  if (w == nullptr) return set_error(GLF_ERR_INVALID, "weights is NULL");
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0 || (saved && (reinterpret_cast<uintptr_t>(saved) & 255) != 0))
    return set_error(GLF_ERR_WORKSPACE, "saved/ws blobs must be 256-byte aligned");
  if (d->precision == GLF_PRECISION_F32X3) return tpavi_fwd_f32x3(d, x, w, z, saved, ws, stream);
  Saved s; WsFwd wf;
  const size_t saved_bytes = carve_saved(m, nullptr, &s);
  carve_ws_fwd(m, ws, &wf, saved_bytes);
  carve_saved(m, saved ? saved : wf.saved_fallback, &s);
  if (m.gram) return tpavi_fwd_gram(d, m, x, w, z, s, wf, stream);
  const int C = m.C, Ci = m.Ci;
  const int N = static_cast<int>(m.N), rows = static_cast<int>(m.rows), B = static_cast<int>(m.B);

  GLF_TRY(prep_weights(w, C, Ci, s.wcat, s.wcatT, s.bcat, s.wz, s.wzT, stream));
  const bf16* X = reinterpret_cast<const bf16*>(x);
  if (m.pack_x) {
    if (d->x_layout == GLF_LAYOUT_NCTHW)
      GLF_TRY(transpose_cast(x, s.xtok, B, C, N, d->io_dtype, GLF_DTYPE_BF16, stream));
    else
      GLF_TRY(transpose_cast(x, s.xtok, 1, 1, static_cast<int>(m.rows * C), d->io_dtype, GLF_DTYPE_BF16, stream));  // cast only
    X = s.xtok;
  }
  {  // P = X Wcat^T + bcat
    GemmArgs g;
    g.A = opnd(X, 0, C, 0);
    g.B = opnd(s.wcat, 0, C, 0);
    g.M = rows; g.N = 3 * Ci; g.K = C;
    g.bias = s.bcat;
    g.D = s.P; g.ldd = 3 * Ci;
    GLF_TRY(gemm(g, stream));
  }
  int np = 0;
  if (m.dot) {
    const long long CiCi = static_cast<long long>(Ci) * Ci;
    {  // M_b[i,j] = sum_n Phi[n,i] G[n,j] / N      (token contraction: both operands MN-major views of P)
      GemmArgs g;
      g.A = opnd(s.P + Ci, 1, 3 * Ci, static_cast<long long>(N) * 3 * Ci);
      g.B = opnd(s.P + 2 * Ci, 1, 3 * Ci, static_cast<long long>(N) * 3 * Ci);
      g.M = Ci; g.N = Ci; g.K = N; g.batch = B;
      g.alpha = 1.f / static_cast<float>(N);
      g.ldd = Ci; g.strideD = CiCi;
      g.split_k = pick_split(static_cast<long long>(B) * ((Ci + 127) / 128) * ((Ci + 127) / 128), N);
      if (g.split_k > 1) {
        GLF_TRY(check_cuda(cudaMemsetAsync(wf.Mf, 0, sizeof(float) * B * CiCi, stream), "memset M"));
        g.out_kind = 2;
        g.D = wf.Mf;
        GLF_TRY(gemm(g, stream));
        GLF_TRY(cast_bf16(wf.Mf, s.Mb, B * CiCi, stream));
      } else {
        g.out_kind = 0;
        g.D = s.Mb;
        GLF_TRY(gemm(g, stream));
      }
    }
    {  // W'_b[c,i] = sum_j Wz[c,j] M_b[i,j]        (folds Y = Theta M and U = Y Wz^T into one product per token)
      GemmArgs g;
      g.A = opnd(s.wz, 0, Ci, 0);
      g.B = opnd(s.Mb, 0, Ci, CiCi);
      g.M = C; g.N = Ci; g.K = Ci; g.batch = B;
      g.D = s.Wp; g.ldd = Ci; g.strideD = static_cast<long long>(C) * Ci;
      GLF_TRY(gemm(g, stream));
    }
    {  // U_b = Theta_b W'_b^T   (+ BatchNorm column statistics); the bias bz is folded into the BN affine (bn_finalize)
      GemmArgs g;
      g.A = opnd(s.P, 0, 3 * Ci, static_cast<long long>(N) * 3 * Ci);
      g.B = opnd(s.Wp, 0, Ci, static_cast<long long>(C) * Ci);
      g.M = N; g.N = C; g.K = Ci; g.batch = B;
      g.D = s.U; g.ldd = C; g.strideD = static_cast<long long>(N) * C;
      g.colstats = (d->training && d->bn_layer) ? wf.colstats : nullptr;
      g.colstats_rows = &np;
      GLF_TRY(gemm(g, stream));
    }
  } else {
    GLF_TRY(flash_fwd(s.P, s.Y, s.lse, B, N, Ci, wf.attn, stream));
    {  // U = Y Wz^T   (bz folded into the BN affine)
      GemmArgs g;
      g.A = opnd(s.Y, 0, Ci, 0);
      g.B = opnd(s.wz, 0, Ci, 0);
      g.M = rows; g.N = C; g.K = Ci;
      g.D = s.U; g.ldd = C;
      g.colstats = (d->training && d->bn_layer) ? wf.colstats : nullptr;
      g.colstats_rows = &np;
      GLF_TRY(gemm(g, stream));
    }
  }
  const float* bn_part = wf.colstats;
  if (d->training && d->bn_layer && np > 4 * REDUCE_STAGE1_ROWS) {
    long long rs = 2LL * C;
    const int np_in[1] = {np};
    GLF_TRY(reduce_stage1(wf.colstats, nullptr, nullptr, 1, np_in, &np, &rs, C, 2, C, wf.red1, stream));
    bn_part = wf.red1;
  }
  GLF_TRY(bn_finalize(bn_part, np, C, static_cast<double>(m.rows), d, w, w->wz_b, s.bn_mean, s.bn_rstd, s.bn_a, s.bn_b,
                      stream));
  if (defer_ln(d)) return 0;   // the pair entry point glf_fusion_ln_fwd finishes both blocks in one pass
  GLF_TRY(bn_res_ln_fwd(s.U, X, GLF_DTYPE_BF16, s.bn_a, s.bn_b, w->ln_w, w->ln_b, z, d->io_dtype, s.ln_mu, s.ln_r,
                        m.rows, C, d->eps_ln, d->accumulate, stream));
  return 0;
}

GLF_API int glf_tpavi_bwd(const glf_desc* d, const void* dz, const void* x, const glf_weights* w, const void* saved, void* dx,
                  const glf_grads* g_, void* ws, glf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Dims m;
  GLF_TRY(make_dims(d, &m));
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(dz, "dz"));
  GLF_TRY(check_ptr(x, "x"));
  GLF_TRY(check_ptr(dx, "dx"));
  GLF_TRY(check_ptr(saved, "saved"));
  GLF_TRY(check_ptr(ws, "ws"));
  if (w == nullptr || g_ == nullptr) return set_error(GLF_ERR_INVALID, "weights/grads is NULL");
  if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0 || (reinterpret_cast<uintptr_t>(saved) & 255) != 0)
    return set_error(GLF_ERR_WORKSPACE, "saved/ws blobs must be 256-byte aligned");
  if (d->precision == GLF_PRECISION_F32X3) return tpavi_bwd_f32x3(d, dz, x, w, saved, dx, g_, ws, stream);
  Saved s; WsBwd wb;
  carve_saved(m, const_cast<void*>(saved), &s);
  carve_ws_bwd(d, m, ws, &wb);
  const int C = m.C, Ci = m.Ci;
  const int N = static_cast<int>(m.N), rows = static_cast<int>(m.rows), B = static_cast<int>(m.B);
  const bf16* X = m.pack_x ? s.xtok : reinterpret_cast<const bf16*>(x);
  const long long seqP = static_cast<long long>(N) * 3 * Ci;

  const void* dZ = dz;
  int dz_dtype = d->io_dtype;
  if (m.pack_dz && !defer_ln(d)) {
    GLF_TRY(transpose_cast(dz, wb.dztok, B, C, N, d->io_dtype, GLF_DTYPE_BF16, stream));
    dZ = wb.dztok;
    dz_dtype = GLF_DTYPE_BF16;
  }
  int nb = 0;
  if (defer_ln(d)) {
    // dV and the partials were written by glf_fusion_ln_bwd
    nb = ln_tma_supported(C) ? ln_bwd_tma_blocks(m.rows) : bn_res_ln_bwd_blocks(m.rows, C);
  } else {
    GLF_TRY(bn_res_ln_bwd(dZ, dz_dtype, s.U, X, GLF_DTYPE_BF16, s.bn_a, s.bn_b, s.bn_mean, s.bn_rstd, w->ln_w, s.ln_mu,
                          s.ln_r, wb.dV, wb.part_ln, m.rows, C, &nb, stream));
  }
  GLF_TRY(bn_bwd_finalize(wb.part_ln, nb, C, static_cast<double>(m.rows), d, w, s.bn_mean, s.bn_rstd, g_, wb.k1, wb.k2,
                          wb.k3, stream));
  if (m.gram) return tpavi_bwd_gram(d, m, X, w, s, wb, dx, g_, stream);
  const bf16* dU = wb.dV;
  if (d->bn_layer) {
    GLF_TRY(bn_bwd_apply(wb.dV, s.U, GLF_DTYPE_BF16, wb.k1, wb.k2, wb.k3, wb.dU, m.rows, C, stream));
    dU = wb.dU;
  }
  int np = 0, np_p = 0, np_g = 0;
  if (m.dot) {
    const long long CiCi = static_cast<long long>(Ci) * Ci;
    const long long CCi = static_cast<long long>(C) * Ci;
    {  // dTheta_b = dU_b W'_b          (B operand = W'_b read MN-major: stored [K=c][N=i])
      GemmArgs g;
      g.A = opnd(dU, 0, C, static_cast<long long>(N) * C);
      g.B = opnd(s.Wp, 1, Ci, CCi);
      g.M = N; g.N = Ci; g.K = C; g.batch = B;
      g.D = wb.dP; g.ldd = 3 * Ci; g.strideD = seqP;
      g.colstats = wb.cs_t;
      g.colstats_rows = &np;
      GLF_TRY(gemm(g, stream));
    }
    {  // dW'_b[c,i] = sum_n dU[n,c] Theta[n,i]
      GemmArgs g;
      g.A = opnd(dU, 1, C, static_cast<long long>(N) * C);
      g.B = opnd(s.P, 1, 3 * Ci, seqP);
      g.M = C; g.N = Ci; g.K = N; g.batch = B;
      g.ldd = Ci; g.strideD = CCi;
      g.split_k = pick_split(static_cast<long long>(B) * ((C + 127) / 128) * ((Ci + 127) / 128), N);
      if (g.split_k > 1) {
        GLF_TRY(check_cuda(cudaMemsetAsync(wb.dWpf, 0, sizeof(float) * B * CCi, stream), "memset dW'"));
        g.out_kind = 2;
        g.D = wb.dWpf;
        GLF_TRY(gemm(g, stream));
        GLF_TRY(cast_bf16(wb.dWpf, wb.dWpb, B * CCi, stream));
      } else {
        g.out_kind = 0;
        g.D = wb.dWpb;
        GLF_TRY(gemm(g, stream));
      }
    }
    GLF_TRY(check_cuda(cudaMemsetAsync(g_->wz_w, 0, sizeof(float) * CCi, stream), "memset dWz"));
    {  // dWz[c,j] = sum_b sum_i dW'_b[c,i] M_b[i,j]     (batch reduced by fp32 red.add into one output)
      GemmArgs g;
      g.A = opnd(wb.dWpb, 0, Ci, CCi);
      g.B = opnd(s.Mb, 1, Ci, CiCi);
      g.M = C; g.N = Ci; g.K = Ci; g.batch = B;
      g.out_kind = 2;
      g.D = g_->wz_w; g.ldd = Ci; g.strideD = 0;
      GLF_TRY(gemm(g, stream));
    }
    {  // dM_b[i,j] = sum_c dW'_b[c,i] Wz[c,j]
      GemmArgs g;
      g.A = opnd(wb.dWpb, 1, Ci, CCi);
      g.B = opnd(s.wz, 1, Ci, 0);
      g.M = Ci; g.N = Ci; g.K = C; g.batch = B;
      g.D = wb.dM; g.ldd = Ci; g.strideD = CiCi;
      GLF_TRY(gemm(g, stream));
    }
    {  // dPhi_b = G_b dM_b^T / N
      GemmArgs g;
      g.A = opnd(s.P + 2 * Ci, 0, 3 * Ci, seqP);
      g.B = opnd(wb.dM, 0, Ci, CiCi);
      g.M = N; g.N = Ci; g.K = Ci; g.batch = B;
      g.alpha = 1.f / static_cast<float>(N);
      g.D = wb.dP + Ci; g.ldd = 3 * Ci; g.strideD = seqP;
      g.colstats = wb.cs_p;
      g.colstats_rows = &np_p;
      GLF_TRY(gemm(g, stream));
    }
    {  // dG_b = Phi_b dM_b / N          (B operand = dM_b read MN-major)
      GemmArgs g;
      g.A = opnd(s.P + Ci, 0, 3 * Ci, seqP);
      g.B = opnd(wb.dM, 1, Ci, CiCi);
      g.M = N; g.N = Ci; g.K = Ci; g.batch = B;
      g.alpha = 1.f / static_cast<float>(N);
      g.D = wb.dP + 2 * Ci; g.ldd = 3 * Ci; g.strideD = seqP;
      g.colstats = wb.cs_g;
      g.colstats_rows = &np_g;
      GLF_TRY(gemm(g, stream));
    }
  } else {
    {  // dY = dU Wz
      GemmArgs g;
      g.A = opnd(dU, 0, C, 0);
      g.B = opnd(s.wzT, 0, C, 0);
      g.M = rows; g.N = Ci; g.K = C;
      g.D = wb.dY; g.ldd = Ci;
      GLF_TRY(gemm(g, stream));
    }
    GLF_TRY(check_cuda(cudaMemsetAsync(g_->wz_w, 0, sizeof(float) * C * Ci, stream), "memset dWz"));
    {  // dWz[c,j] = sum_n dU[n,c] Y[n,j]
      GemmArgs g;
      g.A = opnd(dU, 1, C, 0);
      g.B = opnd(s.Y, 1, Ci, 0);
      g.M = C; g.N = Ci; g.K = rows;
      g.out_kind = 2;
      g.D = g_->wz_w; g.ldd = Ci;
      g.split_k = pick_split_rounds(C, Ci, 1, rows);
      GLF_TRY(gemm(g, stream));
    }
    GLF_TRY(flash_bwd(s.P, s.Y, wb.dY, s.lse, wb.dP, wb.delta, wb.cs_t, wb.cs_p, wb.cs_g, 4 * B * m.tiles_seq, &np, B, N,
                      Ci, wb.attn, stream));
    np_p = np_g = np;
  }
  GLF_TRY(check_cuda(cudaMemsetAsync(wb.dwcat, 0, sizeof(float) * 3 * Ci * C, stream), "memset dWcat"));
  {  // dWcat[r,c] = sum_n dP[n,r] X[n,c]
    GemmArgs g;
    g.A = opnd(wb.dP, 1, 3 * Ci, 0);
    g.B = opnd(X, 1, C, 0);
    g.M = 3 * Ci; g.N = C; g.K = rows;
    g.out_kind = 2;
    g.D = wb.dwcat; g.ldd = C;
    g.split_k = pick_split_rounds(3 * Ci, C, 1, rows);
    GLF_TRY(gemm(g, stream));
  }
  const size_t wbytes = sizeof(float) * Ci * C;
  GLF_TRY(check_cuda(cudaMemcpyAsync(g_->theta_w, wb.dwcat, wbytes, cudaMemcpyDeviceToDevice, stream), "copy dtheta"));
  GLF_TRY(check_cuda(cudaMemcpyAsync(g_->phi_w, wb.dwcat + static_cast<size_t>(Ci) * C, wbytes, cudaMemcpyDeviceToDevice, stream), "copy dphi"));
  GLF_TRY(check_cuda(cudaMemcpyAsync(g_->g_w, wb.dwcat + 2 * static_cast<size_t>(Ci) * C, wbytes, cudaMemcpyDeviceToDevice, stream), "copy dg"));
  {  // dX = dP Wcat + dV
    GemmArgs g;
    g.A = opnd(wb.dP, 0, 3 * Ci, 0);
    g.B = opnd(s.wcatT, 0, 3 * Ci, 0);
    g.M = rows; g.N = C; g.K = 3 * Ci;
    g.addend = wb.dV; g.ld_add = C;
    g.D = m.pack_x ? static_cast<void*>(wb.dxtok) : dx; g.ldd = C;
    GLF_TRY(gemm(g, stream));
  }
  {  // bias gradients = column sums of dTheta / dPhi / dG, from the per-sub-block partials of their GEMM epilogues
    // the three tables can have different row counts (tile shape and grid differ per product): stage 1 always
    long long rs = 2LL * Ci;
    const float *t0 = wb.cs_t, *t1 = wb.cs_p, *t2 = wb.cs_g;
    if (np < 0) {  // flash backward: one table of width 3Ci (column sums of dP = [dTheta | dPhi | dG])
      np = np_p = np_g = -np;
      rs = 2LL * 3 * Ci;
      t1 = wb.cs_t + Ci;
      t2 = wb.cs_t + 2 * Ci;
    }
    if (np == np_p && np == np_g && np <= 4 * REDUCE_STAGE1_ROWS) {
      // short tables (one row per CTA): a single fixed-order reduction
      GLF_TRY(reduce_partials3(t0, t1, t2, np, rs, Ci, g_->theta_b, g_->phi_b, g_->g_b, stream));
    } else {
      const int np_in[3] = {np, np_p, np_g};
      GLF_TRY(reduce_stage1(t0, t1, t2, 3, np_in, &np, &rs, 0, 1, Ci, wb.red1, stream));
      GLF_TRY(reduce_partials3(wb.red1, wb.red1 + static_cast<long long>(REDUCE_STAGE1_ROWS) * Ci,
                               wb.red1 + 2LL * REDUCE_STAGE1_ROWS * Ci, np, rs, Ci, g_->theta_b, g_->phi_b, g_->g_b,
                               stream));
    }
  }
  if (m.pack_x) {
    if (d->x_layout == GLF_LAYOUT_NCTHW)
      GLF_TRY(transpose_cast(wb.dxtok, dx, B, N, C, GLF_DTYPE_BF16, d->io_dtype, stream));
    else
      GLF_TRY(transpose_cast(wb.dxtok, dx, 1, 1, static_cast<int>(m.rows * C), GLF_DTYPE_BF16, d->io_dtype, stream));
  }
  return 0;
}

GLF_API int glf_fusion_ln_supported(const glf_desc* d) {
  Dims m;
  if (make_dims(d, &m) != 0) return 0;
  return check_pair(d, m) == 0 ? 1 : 0;
}

GLF_API int glf_fusion_ln_fwd(const glf_desc* d, const void* xg, const void* xl, const glf_weights* wg,
                              const glf_weights* wl, void* z, void* saved_g, void* saved_l, glf_stream_t stream_) {
  return glf_fusion_ln_fwd_parts(d, xg, xl, wg, wl, z, nullptr, saved_g, saved_l, stream_);
}

GLF_API int glf_fusion_ln_fwd_parts(const glf_desc* d, const void* xg, const void* xl, const glf_weights* wg,
                                    const glf_weights* wl, void* z, void* z_global, void* saved_g, void* saved_l,
                                    glf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (z_global != nullptr) {
    GLF_TRY(check_ptr(z_global, "z_global"));
    if (d != nullptr && d->accumulate) return set_error(GLF_ERR_INVALID, "glf_fusion_ln_fwd_parts: accumulate is not supported");
  }
  Dims m;
  GLF_TRY(make_dims(d, &m));
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_pair(d, m));
  GLF_TRY(check_ptr(xg, "xg"));
  GLF_TRY(check_ptr(xl, "xl"));
  GLF_TRY(check_ptr(z, "z"));
  GLF_TRY(check_ptr(saved_g, "saved_g"));
  GLF_TRY(check_ptr(saved_l, "saved_l"));
  if (wg == nullptr || wl == nullptr) return set_error(GLF_ERR_INVALID, "weights is NULL");
  Saved sg, sl;
  carve_saved(m, saved_g, &sg);
  carve_saved(m, saved_l, &sl);
  const bf16* U[2] = {sg.U, sl.U};
  const bf16* X[2] = {reinterpret_cast<const bf16*>(xg), reinterpret_cast<const bf16*>(xl)};
  const float* a[2] = {sg.bn_a, sl.bn_a};
  const float* b[2] = {sg.bn_b, sl.bn_b};
  const float* lw[2] = {wg->ln_w, wl->ln_w};
  const float* lb[2] = {wg->ln_b, wl->ln_b};
  float* mu[2] = {sg.ln_mu, sl.ln_mu};
  float* r[2] = {sg.ln_r, sl.ln_r};
  if (!ln_tma_supported(m.C))     // wide rows (256 < C <= 2048): the sliced-row ring kernel, both blocks in one pass
    return ln_pair_fwd_wide(U, X, a, b, lw, lb, mu, r, reinterpret_cast<bf16*>(z), m.rows, m.C, d->eps_ln, d->accumulate,
                            stream, reinterpret_cast<bf16*>(z_global));
  return ln_fwd_tma(2, U, X, a, b, lw, lb, mu, r, reinterpret_cast<bf16*>(z), m.rows, m.C, d->eps_ln, d->accumulate,
                    stream, reinterpret_cast<bf16*>(z_global));
}

GLF_API int glf_fusion_ln_bwd(const glf_desc* d, const void* dz, const void* xg, const void* xl, const glf_weights* wg,
                              const glf_weights* wl, const void* saved_g, const void* saved_l, void* ws_g, void* ws_l,
                              glf_stream_t stream_) {
  return glf_fusion_ln_bwd_views(d, dz, nullptr, nullptr, xg, xl, wg, wl, saved_g, saved_l, ws_g, ws_l, stream_);
}

GLF_API int glf_fusion_ln_bwd_views_supported(const glf_desc* d) {
  Dims m;
  if (d == nullptr || make_dims(d, &m) != 0 || check_pair(d, m) != 0 || !ln_tma_supported(m.C)) return 0;
  return (static_cast<long long>(d->H) * d->W) % ln_bwd_tma_tile_rows() == 0 && d->T <= 8 ? 1 : 0;
}

GLF_API int glf_fusion_ln_bwd_views(const glf_desc* d, const void* dz, const void* const* dz_views,
                                    const int64_t* dz_stride_b, const void* xg, const void* xl, const glf_weights* wg,
                                    const glf_weights* wl, const void* saved_g, const void* saved_l, void* ws_g,
                                    void* ws_l, glf_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  Dims m;
  GLF_TRY(make_dims(d, &m));
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_pair(d, m));
  const bool views = dz_views != nullptr;
  if (views && dz_stride_b == nullptr) return set_error(GLF_ERR_INVALID, "dz_stride_b is NULL");
  if (!views) GLF_TRY(check_ptr(dz, "dz"));
  GLF_TRY(check_ptr(xg, "xg"));
  GLF_TRY(check_ptr(xl, "xl"));
  GLF_TRY(check_ptr(saved_g, "saved_g"));
  GLF_TRY(check_ptr(saved_l, "saved_l"));
  GLF_TRY(check_ptr(ws_g, "ws_g"));
  GLF_TRY(check_ptr(ws_l, "ws_l"));
  if (wg == nullptr || wl == nullptr) return set_error(GLF_ERR_INVALID, "weights is NULL");
  Saved sg, sl;
  WsBwd bg, bl;
  carve_saved(m, const_cast<void*>(saved_g), &sg);
  carve_saved(m, const_cast<void*>(saved_l), &sl);
  carve_ws_bwd(d, m, ws_g, &bg);
  carve_ws_bwd(d, m, ws_l, &bl);
  const bf16* U[2] = {sg.U, sl.U};
  const bf16* X[2] = {reinterpret_cast<const bf16*>(xg), reinterpret_cast<const bf16*>(xl)};
  const float* a[2] = {sg.bn_a, sl.bn_a};
  const float* b[2] = {sg.bn_b, sl.bn_b};
  const float* lw[2] = {wg->ln_w, wl->ln_w};
  const float* mean[2] = {sg.bn_mean, sl.bn_mean};
  const float* rstd[2] = {sg.bn_rstd, sl.bn_rstd};
  const float* mu[2] = {sg.ln_mu, sl.ln_mu};
  const float* r[2] = {sg.ln_r, sl.ln_r};
  bf16* dV[2] = {bg.dV, bl.dV};
  float* part[2] = {bg.part_ln, bl.part_ln};
  int nb = 0;
  static_assert(sizeof(long long) == sizeof(int64_t), "stride tables are passed through unchanged");
  if (!ln_tma_supported(m.C)) {
    // wide rows: the backward is issue-bound, not bandwidth-bound, so sharing the dz read buys nothing;
    // the two blocks run the sliced-row ring kernel one after the other
    if (views) return set_error(GLF_ERR_UNSUPPORTED, "per-view dz needs C <= 256");
    for (int k = 0; k < 2; ++k)
      GLF_TRY(bn_res_ln_bwd(dz, GLF_DTYPE_BF16, U[k], X[k], GLF_DTYPE_BF16, a[k], b[k], mean[k], rstd[k], lw[k], mu[k], r[k],
                            dV[k], part[k], m.rows, m.C, &nb, stream));
    return 0;
  }
  return ln_bwd_tma(2, reinterpret_cast<const bf16*>(dz), U, X, a, b, lw, mean, rstd, mu, r, dV, part, m.rows, m.C, &nb,
                    stream, views ? d->T : 0, dz_views, reinterpret_cast<const long long*>(dz_stride_b), d->H * d->W);
}

GLF_API int glf_gate_concat_fwd(int B, int C, int V, int h, int w, int ncls, float weight, int io_dtype, int x_dtype,
                        const void* const* f4, const float* const* cls, const float* const* ctr, void* xg, void* xl,
                        float* gate, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || ncls <= 0) return set_error(GLF_ERR_INVALID, "gate_concat: empty input");
  if (f4 == nullptr || cls == nullptr || ctr == nullptr) return set_error(GLF_ERR_INVALID, "gate_concat: NULL pointer table");
  GLF_TRY(check_ptr(xg, "xg"));
  GLF_TRY(check_ptr(xl, "xl"));
  return gate_concat_fwd(B, C, V, h, w, ncls, weight, io_dtype, x_dtype, f4, cls, ctr, xg, xl, gate,
                         reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_gate_concat_bwd(int B, int C, int V, int h, int w, int ncls, float weight, int io_dtype, int x_dtype,
                        const void* const* f4, const float* const* cls, const float* const* ctr, const float* gate,
                        const void* dxg, const void* dxl, void* const* df4, float* const* dcls, float* const* dctr,
                        void* scratch, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || ncls <= 0) return set_error(GLF_ERR_INVALID, "gate_concat: empty input");
  if (f4 == nullptr || cls == nullptr || ctr == nullptr || df4 == nullptr || dcls == nullptr || dctr == nullptr)
    return set_error(GLF_ERR_INVALID, "gate_concat: NULL pointer table");
  GLF_TRY(check_ptr(dxg, "dxg"));
  GLF_TRY(check_ptr(dxl, "dxl"));
  return gate_concat_bwd(B, C, V, h, w, ncls, weight, io_dtype, x_dtype, f4, cls, ctr, gate, dxg, dxl, df4, dcls, dctr,
                         reinterpret_cast<float*>(scratch), reinterpret_cast<cudaStream_t>(stream));
}

GLF_API size_t glf_gate_concat_bwd_scratch_bytes(int B, int C, int V, int h, int w) {
  return gate_bwd_scratch_bytes(B, C, V, h, w);
}

GLF_API int glf_gate_concat_cl_fwd(int B, int C, int V, int h, int w, int ncls, float weight, int io_dtype,
                                   const void* const* f4, const int64_t* stride_b, const int64_t* stride_t,
                                   const float* const* cls, const float* const* ctr, void* xg, void* xl, float* gate,
                                   glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || ncls <= 0) return set_error(GLF_ERR_INVALID, "gate_concat: empty input");
  if (f4 == nullptr || cls == nullptr || ctr == nullptr || stride_b == nullptr || stride_t == nullptr)
    return set_error(GLF_ERR_INVALID, "gate_concat: NULL table");
  GLF_TRY(check_ptr(xg, "xg"));
  GLF_TRY(check_ptr(xl, "xl"));
  return gate_concat_cl_fwd(B, C, V, h, w, ncls, weight, io_dtype, f4, reinterpret_cast<const long long*>(stride_b),
                            reinterpret_cast<const long long*>(stride_t), cls, ctr, xg, xl, gate,
                            reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_gate_concat_cl_bwd(int B, int C, int V, int h, int w, int ncls, float weight, int io_dtype,
                                   const void* const* f4, const int64_t* stride_b, const int64_t* stride_t,
                                   const float* const* cls, const float* const* ctr, const float* gate, const void* dxg,
                                   const void* dxl, void* const* df4, const int64_t* dstride_b, const int64_t* dstride_t,
                                   float* const* dcls, float* const* dctr, void* scratch, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (B <= 0 || C <= 0 || h <= 0 || w <= 0 || ncls <= 0) return set_error(GLF_ERR_INVALID, "gate_concat: empty input");
  if (f4 == nullptr || cls == nullptr || ctr == nullptr || df4 == nullptr || dcls == nullptr || dctr == nullptr ||
      stride_b == nullptr || stride_t == nullptr || dstride_b == nullptr || dstride_t == nullptr)
    return set_error(GLF_ERR_INVALID, "gate_concat: NULL table");
  GLF_TRY(check_ptr(dxg, "dxg"));
  GLF_TRY(check_ptr(dxl, "dxl"));
  return gate_concat_cl_bwd(B, C, V, h, w, ncls, weight, io_dtype, f4, reinterpret_cast<const long long*>(stride_b),
                            reinterpret_cast<const long long*>(stride_t), cls, ctr, gate, dxg, dxl, df4,
                            reinterpret_cast<const long long*>(dstride_b), reinterpret_cast<const long long*>(dstride_t),
                            dcls, dctr, reinterpret_cast<float*>(scratch), reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_views_to_tokens(int B, int C, int V, int T, int src_dtype, const void* const* src,
                                const int64_t* stride_b, const int64_t* stride_c, const int64_t* stride_t, void* out,
                                glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (B <= 0 || C <= 0 || T <= 0) return set_error(GLF_ERR_INVALID, "views_to_tokens: empty input");
  if (src == nullptr || stride_b == nullptr || stride_c == nullptr || stride_t == nullptr)
    return set_error(GLF_ERR_INVALID, "views_to_tokens: NULL table");
  GLF_TRY(check_ptr(out, "out"));
  static_assert(sizeof(long long) == sizeof(int64_t), "stride tables are passed through unchanged");
  return views_to_tokens(B, C, V, T, src_dtype, src, reinterpret_cast<const long long*>(stride_b),
                         reinterpret_cast<const long long*>(stride_c), reinterpret_cast<const long long*>(stride_t), out,
                         reinterpret_cast<cudaStream_t>(stream));
}

GLF_API size_t glf_p2p_signal_bytes(int world) { return p2p_signal_bytes(world); }
GLF_API int64_t glf_p2p_max_floats(void) { return p2p_max_floats(); }
GLF_API int glf_p2p_export(const void* ptr, unsigned char handle[64], uint64_t* offset) {
  if (ptr == nullptr || handle == nullptr || offset == nullptr) return set_error(GLF_ERR_INVALID, "p2p_export: NULL argument");
  unsigned long long off = 0;
  GLF_TRY(p2p_export(ptr, handle, &off));
  *offset = off;
  return 0;
}
GLF_API int glf_p2p_open(const unsigned char handle[64], uint64_t offset, void** out) {
  if (handle == nullptr || out == nullptr) return set_error(GLF_ERR_INVALID, "p2p_open: NULL argument");
  return p2p_open(handle, offset, out);
}
GLF_API int glf_p2p_close(void* ptr, uint64_t offset) {
  if (ptr == nullptr) return set_error(GLF_ERR_INVALID, "p2p_close: NULL argument");
  return p2p_close(ptr, offset);
}
GLF_API int glf_p2p_allreduce(void* const* bufs, void* const* sigs, int rank, int world, int64_t n, float scale,
                      glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (bufs == nullptr || sigs == nullptr) return set_error(GLF_ERR_INVALID, "p2p_allreduce: NULL pointer table");
  return p2p_allreduce(bufs, sigs, rank, world, n, scale, reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int batch, int a_mn, int b_mn,
                  int64_t lda, int64_t ldb, int64_t ldd, int64_t strideA, int64_t strideB, int64_t strideD,
                  const float* bias, float alpha, const void* addend, int64_t ld_add, int64_t stride_add,
                  int out_kind, int split_k, float* colstats, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(A, "A"));
  GLF_TRY(check_ptr(B, "B"));
  GLF_TRY(check_ptr(D, "D"));
  GemmArgs g;
  g.A = opnd(A, a_mn, lda, strideA);
  g.B = opnd(B, b_mn, ldb, strideB);
  g.M = M; g.N = N; g.K = K; g.batch = batch;
  g.alpha = alpha; g.bias = bias;
  g.out_kind = out_kind;
  g.D = D; g.ldd = ldd; g.strideD = strideD;
  g.addend = reinterpret_cast<const bf16*>(addend); g.ld_add = ld_add; g.stride_add = stride_add;
  g.colstats = colstats;
  g.split_k = split_k;
  return gemm(g, reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_gemm_bf16_ex(const void* A, const void* B, void* D, int M, int N, int K, int batch, int a_mn, int b_mn,
                     int64_t lda, int64_t ldb, int64_t ldd, int64_t strideA, int64_t strideB, int64_t strideD,
                     const float* bias, int64_t bias_stride, float alpha, int out_kind, int split_k, float* rowsum,
                     glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(A, "A"));
  GLF_TRY(check_ptr(B, "B"));
  GLF_TRY(check_ptr(D, "D"));
  GemmArgs g;
  g.A = opnd(A, a_mn, lda, strideA);
  g.B = opnd(B, b_mn, ldb, strideB);
  g.M = M; g.N = N; g.K = K; g.batch = batch;
  g.alpha = alpha; g.bias = bias; g.bias_stride = bias_stride;
  g.out_kind = out_kind;
  g.D = D; g.ldd = ldd; g.strideD = strideD;
  g.split_k = split_k;
  g.rowsum = rowsum; g.rowsum_stride = M;
  return gemm(g, reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_gram_contraction(const void* A, const void* X, void* D, float* colsum, int B, int N, int C, int ldd,
                         glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(A, "A"));
  GLF_TRY(check_ptr(X, "X"));
  GLF_TRY(check_ptr(D, "D"));
  GLF_TRY(check_ptr(colsum, "colsum"));
  if (B <= 0 || N <= 0) return set_error(GLF_ERR_INVALID, "gram_contraction: empty input");
  if (ldd < C || ldd % 8 != 0) return set_error(GLF_ERR_INVALID, "gram_contraction: ldd must be >= C and a multiple of 8");
  return gram_contraction(reinterpret_cast<const bf16*>(A), reinterpret_cast<const bf16*>(X), reinterpret_cast<bf16*>(D),
                          nullptr, colsum, nullptr, nullptr, 0.f, 0, B, N, C, ldd, 1,
                          reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_bn_res_ln_fwd(int64_t rows, int C, const void* U, const void* X, const float* bn_a, const float* bn_b,
                              const float* ln_w, const float* ln_b, void* Z, int z_dtype, float* mu, float* r,
                              float eps, int accumulate, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(X, "X"));
  GLF_TRY(check_ptr(Z, "Z"));
  if (rows <= 0) return set_error(GLF_ERR_INVALID, "empty input");
  return bn_res_ln_fwd(U, X, GLF_DTYPE_BF16, bn_a, bn_b, ln_w, ln_b, Z, z_dtype, mu, r, rows, C, eps, accumulate,
                       reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_bn_res_ln_bwd(int64_t rows, int C, const void* dZ, int dz_dtype, const void* U, const void* X,
                              const float* bn_a, const float* bn_b, const float* bn_mean, const float* bn_rstd,
                              const float* ln_w, const float* mu, const float* r, void* dV, float* part,
                              int* nblocks_out, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(dZ, "dZ"));
  GLF_TRY(check_ptr(X, "X"));
  GLF_TRY(check_ptr(dV, "dV"));
  if (rows <= 0) return set_error(GLF_ERR_INVALID, "empty input");
  return bn_res_ln_bwd(dZ, dz_dtype, U, X, GLF_DTYPE_BF16, bn_a, bn_b, bn_mean, bn_rstd, ln_w, mu, r, dV, part, rows, C,
                       nblocks_out, reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_bn_res_ln_bwd_max_blocks(void) { return 148 * 2; }

// Pair forms with explicit operands (unit tests, roofline probe): index 0 = MGFM, 1 = MLFM.
GLF_API int glf_bn_res_ln_pair_fwd(int64_t rows, int C, const void* const* U, const void* const* X,
                                   const float* const* bn_a, const float* const* bn_b, const float* const* ln_w,
                                   const float* const* ln_b, void* Z, float* const* mu, float* const* r, float eps,
                                   int accumulate, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (rows <= 0) return set_error(GLF_ERR_INVALID, "empty input");
  if (!ln_tma_supported(C)) return set_error(GLF_ERR_UNSUPPORTED, "fused LayerNorm pair needs C <= 256, C %% 8 == 0");
  GLF_TRY(check_ptr(Z, "Z"));
  for (int m = 0; m < 2; ++m) {
    GLF_TRY(check_ptr(U[m], "U"));
    GLF_TRY(check_ptr(X[m], "X"));
  }
  return ln_fwd_tma(2, reinterpret_cast<const bf16* const*>(U), reinterpret_cast<const bf16* const*>(X), bn_a, bn_b, ln_w,
                    ln_b, mu, r, reinterpret_cast<bf16*>(Z), rows, C, eps, accumulate,
                    reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_bn_res_ln_pair_bwd(int64_t rows, int C, const void* dZ, const void* const* U, const void* const* X,
                                   const float* const* bn_a, const float* const* bn_b, const float* const* bn_mean,
                                   const float* const* bn_rstd, const float* const* ln_w, const float* const* mu,
                                   const float* const* r, void* const* dV, float* const* part, int* nblocks_out,
                                   glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (rows <= 0) return set_error(GLF_ERR_INVALID, "empty input");
  if (!ln_tma_supported(C)) return set_error(GLF_ERR_UNSUPPORTED, "fused LayerNorm pair needs C <= 256, C %% 8 == 0");
  GLF_TRY(check_ptr(dZ, "dZ"));
  for (int m = 0; m < 2; ++m) {
    GLF_TRY(check_ptr(U[m], "U"));
    GLF_TRY(check_ptr(X[m], "X"));
    GLF_TRY(check_ptr(dV[m], "dV"));
    GLF_TRY(check_ptr(mu[m], "mu"));
    GLF_TRY(check_ptr(r[m], "r"));
  }
  return ln_bwd_tma(2, reinterpret_cast<const bf16*>(dZ), reinterpret_cast<const bf16* const*>(U),
                    reinterpret_cast<const bf16* const*>(X), bn_a, bn_b, ln_w, bn_mean, bn_rstd, mu, r,
                    reinterpret_cast<bf16* const*>(dV), part, rows, C, nblocks_out,
                    reinterpret_cast<cudaStream_t>(stream));
}

GLF_API size_t glf_cycle_loss_scratch_bytes(int T, int C, int n_starts) {
  if (T <= 0 || C <= 0 || n_starts <= 0) return 0;
  return cycle_loss_scratch_bytes(T, C, n_starts);
}

GLF_API int glf_cycle_loss(const float* feat, int T, int C, int target_region, int cyc_off, int chunk_size,
                           float temperature, int start, int step, int n_starts, int soft_label, float scale,
                           float* loss, float* dfeat, void* scratch, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(feat, "feat"));
  GLF_TRY(check_ptr(loss, "loss"));
  GLF_TRY(check_ptr(dfeat, "dfeat"));
  GLF_TRY(check_ptr(scratch, "scratch"));
  return cycle_loss(feat, T, C, target_region, cyc_off, chunk_size, temperature, start, step, n_starts, soft_label,
                    scale, loss, dfeat, reinterpret_cast<float*>(scratch), reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_spatial_sums(const void* x, int dtype, int B, int C, int T, int64_t stride_b, int64_t stride_c,
                             int64_t stride_t, float* out, glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  if (x == nullptr || out == nullptr) return set_error(GLF_ERR_INVALID, "spatial_sums: NULL pointer");   // any element alignment
  return spatial_sums(x, dtype, B, C, T, stride_b, stride_c, stride_t, out, reinterpret_cast<cudaStream_t>(stream));
}

GLF_API int glf_transpose(const void* in, void* out, int batch, int R, int S, int in_dtype, int out_dtype,
                  glf_stream_t stream) {
  GLF_TRY(check_device_sm100());
  GLF_TRY(check_ptr(in, "in"));
  GLF_TRY(check_ptr(out, "out"));
  return transpose_cast(in, out, batch, R, S, in_dtype, out_dtype, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
