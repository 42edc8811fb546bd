// Kernels of the frozen EMG encoder's perceptual losses (ste_gan/models/emg_encoder.py:36-88,
// ste_gan/layers/transformer.py:63-306, ste_gan/losses/emg_encoder_loss.py:56-84) that are not convolutions / GEMMs:
//   * LayerNorm forward / input-gradient (post-norm transformer layers, transformer.py:54-60)
//   * multi-head self-attention with per-head learned relative positional logits, forward and input-gradient
//     (transformer.py:87-113 + LearnedRelativePositionalEmbedding, unmasked, :163-306) - sequence length T/16 = 100-128
//     frames, head dim 96: one CTA per (sample, head), operands in shared memory, fp32 arithmetic
//   * speech-unit (mean pairwise L2 distance, eps 1e-6) and phoneme (cross-entropy) losses with their gradients
// The projections (q/k/v, output, feed-forward, w_raw_in / w_out / w_aux) and the strided ResBlocks with their eval-mode
// BatchNorm folded into weights and bias run on the tcgen05 convolution engine as k = 1 / k = 3 convs (passes_encoder.py).
// The encoder is frozen (emg_encoder_loss.py:61): only the gradient w.r.t. the EMG input exists.
#include "common.cuh"

namespace stg {
namespace {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// one warp per row; stats[row] = (mean, rstd)
template <typename T>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int rows, int D, float eps,
                                                            T* __restrict__ y, float* __restrict__ stats) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xr = x + (int64_t)row * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += to_f(xr[i]);
  const float mean = warp_sum_f(s) / (float)D;
  float v = 0.f;
  for (int i = lane; i < D; i += 32) { const float d = to_f(xr[i]) - mean; v = fmaf(d, d, v); }
  const float rstd = rsqrtf(warp_sum_f(v) / (float)D + eps);
  T* yr = y + (int64_t)row * D;
  for (int i = lane; i < D; i += 32) yr[i] = from_f<T>((to_f(xr[i]) - mean) * rstd * gamma[i] + beta[i]);
  if (lane == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd      (+ add, if given)
template <typename T>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ stats, const float* __restrict__ gamma,
                                                            int rows, int D, T* __restrict__ dx) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  const T* xr = x + (int64_t)row * D;
  const T* gr = dy + (int64_t)row * D;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float g = to_f(gr[i]) * gamma[i], xh = (to_f(xr[i]) - mean) * rstd;
    s1 += g; s2 = fmaf(g, xh, s2);
  }
  s1 = warp_sum_f(s1) / (float)D; s2 = warp_sum_f(s2) / (float)D;
  T* o = dx + (int64_t)row * D;
  for (int i = lane; i < D; i += 32) {
    const float g = to_f(gr[i]) * gamma[i], xh = (to_f(xr[i]) - mean) * rstd;
    o[i] = from_f<T>(rstd * (g - s1 - xh * s2));
  }
}

// ---------------------------------------------------------------------------------------------- attention
// qkv [B][L][3*H*d] (q | k | v, each head-major), emb [H][2M-1][d] fp32, probs [B][H][L][L] fp32, o [B][L][H*d].
// logit(i, j) = q_i . k_j * scale + (|j - i| < M ? q_i . emb[h][j - i + M - 1] : -1e8)            (transformer.py:100-108,262-268)
struct AttnP { int B, L, H, d, M; float scale; };

template <typename T>
__device__ __forceinline__ void load_rows(const T* __restrict__ base, int64_t row_stride, int L, int d, int ds, float* __restrict__ dst) {
  for (int i = threadIdx.x; i < L * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    dst[r * ds + c] = to_f(base[(int64_t)r * row_stride + c]);
  }
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, fmaf(a.w, b.w, acc))));
}
__device__ __forceinline__ void axpy4(float s, const float4& v, float4& acc) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float4& v, float s) {
  p[0] = from_f<T>(v.x * s); p[1] = from_f<T>(v.y * s); p[2] = from_f<T>(v.z * s); p[3] = from_f<T>(v.w * s);
}

// Operand rows in shared memory are ds = d + 4 floats long: 16-byte aligned, and for d = 96 (ds = 100 = 4 mod 32 banks) the
// LDS.128 of 8 lanes that read 8 different rows hit 32 different banks - every dot product below reads float4s (3 LDS.128 per 8
// FMAs instead of 3 LDS.32 per 2; the first version of these kernels was shared-memory-issue-bound at 3 % of the FMA rate).
// shared memory (floats): K [L][ds] | V [L][ds] | E [nE][ds] | per warp: q [d] + p [Lp]      (d % 4 == 0, Lp = roundup4(L))
template <typename T>
__global__ void __launch_bounds__(512) relattn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ emb, AttnP p,
                                                          T* __restrict__ o, float* __restrict__ probs) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int L = p.L, d = p.d, ds = d + 4, d4 = d >> 2, M = p.M, HD = p.H * d, Lp = (L + 3) & ~3;
  const int e_lo = max(0, M - L), nE = 2 * M - 1 - 2 * e_lo;          // embedding rows any (i, j) of this length can touch
  float* Ks = sm; float* Vs = Ks + L * ds; float* Es = Vs + L * ds; float* wq = Es + nE * ds;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* myq = wq + warp * (d + Lp); float* myp = myq + d;
  const T* base = qkv + (int64_t)b * L * 3 * HD + h * d;
  load_rows(base + HD, 3 * HD, L, d, ds, Ks);
  load_rows(base + 2 * HD, 3 * HD, L, d, ds, Vs);
  for (int i = threadIdx.x; i < nE * d; i += blockDim.x) {
    const int r = i / d, c = i - r * d;
    Es[r * ds + c] = emb[((int64_t)h * (2 * M - 1) + e_lo + r) * d + c];
  }
  __syncthreads();
  for (int i = warp; i < L; i += nw) {
    for (int c = lane; c < d; c += 32) myq[c] = to_f(base[(int64_t)i * 3 * HD + c]);
    __syncwarp();
    const float4* q4 = reinterpret_cast<const float4*>(myq);
    float mx = -INFINITY;
    for (int j = lane; j < L; j += 32) {
      const int rel = j - i;
      const bool in = rel > -M && rel < M;
      const float4* k4 = reinterpret_cast<const float4*>(Ks + j * ds);
      const float4* e4 = reinterpret_cast<const float4*>(Es + (in ? (rel + M - 1 - e_lo) : 0) * ds);
      float qk = 0.f, qe = 0.f;
#pragma unroll 4
      for (int c = 0; c < d4; ++c) { const float4 q = q4[c]; qk = dot4(q, k4[c], qk); qe = dot4(q, e4[c], qe); }
      const float lg = qk * p.scale + (in ? qe : -1e8f);
      myp[j] = lg;
      mx = fmaxf(mx, lg);
    }
    mx = warp_max_f(mx);
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) { const float e = __expf(myp[j] - mx); myp[j] = e; sum += e; }
    sum = warp_sum_f(sum);
    const float inv = 1.f / sum;
    float* pr = probs + (((int64_t)b * p.H + h) * L + i) * L;
    for (int j = lane; j < L; j += 32) { const float pv = myp[j] * inv; myp[j] = pv; pr[j] = pv; }
    __syncwarp();
    for (int c = lane; c < d4; c += 32) {          // 4 output channels per lane
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int j = 0; j < L; ++j) axpy4(myp[j], *reinterpret_cast<const float4*>(Vs + j * ds + 4 * c), acc);
      store4(o + ((int64_t)b * L + i) * HD + h * d + 4 * c, acc, 1.f);
    }
    __syncwarp();
  }
}

// shared memory (floats): Q | K | V | dO [L][ds] each | Dr [Lp] | per warp: row [Lp] + prow [Lp]   (the embedding rows are read
// through L1: four fp32 operand tiles of 100-128 frames leave no room for the 76 KB table of a head)
template <typename T>
__global__ void __launch_bounds__(512) relattn_bwd_kernel(const T* __restrict__ qkv, const float* __restrict__ emb,
                                                          const float* __restrict__ probs, const T* __restrict__ dout, AttnP p,
                                                          T* __restrict__ dqkv) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int L = p.L, d = p.d, ds = d + 4, d4 = d >> 2, M = p.M, HD = p.H * d, Lp = (L + 3) & ~3;
  float* Qs = sm; float* Ks = Qs + L * ds; float* Vs = Ks + L * ds; float* Gs = Vs + L * ds;
  float* Dr = Gs + L * ds; float* wrow = Dr + Lp;
  const float* Eh = emb + (int64_t)h * (2 * M - 1) * d;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* row = wrow + warp * 2 * Lp; float* prow = row + Lp;
  const T* base = qkv + (int64_t)b * L * 3 * HD + h * d;
  load_rows(base, 3 * HD, L, d, ds, Qs);
  load_rows(base + HD, 3 * HD, L, d, ds, Ks);
  load_rows(base + 2 * HD, 3 * HD, L, d, ds, Vs);
  load_rows(dout + (int64_t)b * L * HD + h * d, HD, L, d, ds, Gs);
  __syncthreads();
  const float* P = probs + ((int64_t)b * p.H + h) * L * L;
  T* dbase = dqkv + (int64_t)b * L * 3 * HD + h * d;
  // pass 1, a warp per query row i: dP_ij = dO_i . V_j ; D_i = sum_j dP_ij P_ij ; dS_ij = P_ij (dP_ij - D_i) ;
  //                                 dq_i = sum_j dS_ij (scale K_j + E[j - i])
  for (int i = warp; i < L; i += nw) {
    const float4* g4 = reinterpret_cast<const float4*>(Gs + i * ds);
    float dsum = 0.f;
    for (int j = lane; j < L; j += 32) {
      const float4* v4 = reinterpret_cast<const float4*>(Vs + j * ds);
      float dp = 0.f;
#pragma unroll 4
      for (int c = 0; c < d4; ++c) dp = dot4(g4[c], v4[c], dp);
      const float pij = P[(int64_t)i * L + j];
      prow[j] = pij;
      row[j] = dp;
      dsum = fmaf(dp, pij, dsum);
    }
    dsum = warp_sum_f(dsum);
    if (lane == 0) Dr[i] = dsum;
    for (int j = lane; j < L; j += 32) row[j] = prow[j] * (row[j] - dsum);
    __syncwarp();
    for (int c = lane; c < d4; c += 32) {
      float4 ak = make_float4(0.f, 0.f, 0.f, 0.f), ae = ak;
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float w = row[j];
        axpy4(w, *reinterpret_cast<const float4*>(Ks + j * ds + 4 * c), ak);
        const int rel = j - i;
        if (rel > -M && rel < M) axpy4(w, *reinterpret_cast<const float4*>(Eh + (int64_t)(rel + M - 1) * d + 4 * c), ae);
      }
      ak.x = ak.x * p.scale + ae.x; ak.y = ak.y * p.scale + ae.y; ak.z = ak.z * p.scale + ae.z; ak.w = ak.w * p.scale + ae.w;
      store4(dbase + (int64_t)i * 3 * HD + 4 * c, ak, 1.f);
    }
    __syncwarp();
  }
  __syncthreads();
  // pass 2, a warp per key row j: dS_ij recomputed from D_i ; dk_j = scale sum_i dS_ij q_i ; dv_j = sum_i P_ij dO_i
  for (int j = warp; j < L; j += nw) {
    const float4* v4 = reinterpret_cast<const float4*>(Vs + j * ds);
    for (int i = lane; i < L; i += 32) {
      const float4* g4 = reinterpret_cast<const float4*>(Gs + i * ds);
      float dp = 0.f;
#pragma unroll 4
      for (int c = 0; c < d4; ++c) dp = dot4(g4[c], v4[c], dp);
      const float pij = P[(int64_t)i * L + j];
      prow[i] = pij;
      row[i] = pij * (dp - Dr[i]);
    }
    __syncwarp();
    for (int c = lane; c < d4; c += 32) {
      float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
#pragma unroll 2
      for (int i = 0; i < L; ++i) {
        axpy4(row[i], *reinterpret_cast<const float4*>(Qs + i * ds + 4 * c), dk);
        axpy4(prow[i], *reinterpret_cast<const float4*>(Gs + i * ds + 4 * c), dv);
      }
      store4(dbase + (int64_t)j * 3 * HD + HD + 4 * c, dk, p.scale);
      store4(dbase + (int64_t)j * 3 * HD + 2 * HD + 4 * c, dv, 1.f);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------- losses
// one warp per (b, t) row.  slots[0] += mean_r || tgt_r - pred_r + 1e-6 ||_2          (F.pairwise_distance, eps 1e-6)
//                           slots[1] += mean_r -log softmax(logits_r)[target_r]       (F.cross_entropy)
// d_units = gs[0] * d slots[0] / d pred ; d_logits = gs[1] * d slots[1] / d logits
template <typename T>
__global__ void __launch_bounds__(256) encoder_loss_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                           const float* __restrict__ logits, const int64_t* __restrict__ ph,
                                                           int N, int Du, int P, float* __restrict__ slots, float gs_u, float gs_p,
                                                           T* __restrict__ d_units, T* __restrict__ d_logits) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= N) return;
  const float inv_n = 1.f / (float)N;
  const float* pr = pred + (int64_t)r * Du; const float* tr = tgt + (int64_t)r * Du;
  float ss = 0.f;
  for (int i = lane; i < Du; i += 32) { const float df = tr[i] - pr[i] + 1e-6f; ss = fmaf(df, df, ss); }
  ss = warp_sum_f(ss);
  const float nrm = sqrtf(ss);
  if (d_units) {
    const float c = nrm > 0.f ? -gs_u * inv_n / nrm : 0.f;
    for (int i = lane; i < Du; i += 32) d_units[(int64_t)r * Du + i] = from_f<T>(c * (tr[i] - pr[i] + 1e-6f));
  }
  const float* lr = logits + (int64_t)r * P;
  float mx = -INFINITY;
  for (int i = lane; i < P; i += 32) mx = fmaxf(mx, lr[i]);
  mx = warp_max_f(mx);
  float se = 0.f;
  for (int i = lane; i < P; i += 32) se += __expf(lr[i] - mx);
  se = warp_sum_f(se);
  const int t = (int)ph[r];
  const float lse = mx + __logf(se);
  if (d_logits)
    for (int i = lane; i < P; i += 32)
      d_logits[(int64_t)r * P + i] = from_f<T>(gs_p * inv_n * (__expf(lr[i] - lse) - (i == t ? 1.f : 0.f)));
  if (lane == 0) {
    atomicAdd(slots, nrm * inv_n);
    atomicAdd(slots + 1, (lse - lr[t]) * inv_n);
  }
}

}  // namespace
}  // namespace stg

using namespace stg;
#define S_ static_cast<cudaStream_t>(stream)

extern "C" int stg_layernorm_fwd(const void* x, int dtype, const float* gamma, const float* beta, int rows, int D, float eps,
                                 void* y, float* stats, stg_stream_t stream) {
  if (!x || !gamma || !beta || !y || rows < 1 || D < 1) return STG_EINVAL;
  const int blocks = (rows + 7) / 8;
  if (dtype == STG_F32) layernorm_fwd_kernel<float><<<blocks, 256, 0, S_>>>((const float*)x, gamma, beta, rows, D, eps, (float*)y, stats);
  else if (dtype == STG_BF16) layernorm_fwd_kernel<bf16><<<blocks, 256, 0, S_>>>((const bf16*)x, gamma, beta, rows, D, eps, (bf16*)y, stats);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_layernorm_bwd(const void* dy, const void* x, int dtype, const float* stats, const float* gamma, int rows, int D,
                                 void* dx, stg_stream_t stream) {
  if (!dy || !x || !stats || !gamma || !dx || rows < 1 || D < 1) return STG_EINVAL;
  const int blocks = (rows + 7) / 8;
  if (dtype == STG_F32) layernorm_bwd_kernel<float><<<blocks, 256, 0, S_>>>((const float*)dy, (const float*)x, stats, gamma, rows, D, (float*)dx);
  else if (dtype == STG_BF16) layernorm_bwd_kernel<bf16><<<blocks, 256, 0, S_>>>((const bf16*)dy, (const bf16*)x, stats, gamma, rows, D, (bf16*)dx);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

static size_t attn_smem(int L, int d, int M, int n_tiles, bool with_emb, int per_warp, int extra, int n_warps) {
  const int ds = d + 4, e_lo = M - L > 0 ? M - L : 0, nE = with_emb ? 2 * M - 1 - 2 * e_lo : 0;
  return sizeof(float) * ((size_t)n_tiles * L * ds + (size_t)nE * ds + extra + (size_t)n_warps * per_warp);
}

extern "C" int stg_relattn_fwd(const void* qkv, int dtype, const float* emb, int B, int L, int H, int d, int max_rel, float scale,
                               void* o, float* probs, stg_stream_t stream) {
  if (!qkv || !emb || !o || !probs || B < 1 || L < 1 || H < 1 || d < 4 || (d & 3) || max_rel < 1) return STG_EINVAL;
  const int Lp = (L + 3) & ~3;
  int nw = 16;                                          // 16 warps when their scratch fits beside the operand tiles, else 8
  size_t smem = attn_smem(L, d, max_rel, 2, true, d + Lp, 0, nw);
  if (smem > 220 * 1024) { nw = 8; smem = attn_smem(L, d, max_rel, 2, true, d + Lp, 0, nw); }
  if (smem > 220 * 1024) return STG_EUNSUPPORTED;      // sequence too long for the one-CTA-per-head kernel
  AttnP p{B, L, H, d, max_rel, scale};
  if (dtype == STG_F32) {
    STG_CUDA_CHECK(cudaFuncSetAttribute(relattn_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    relattn_fwd_kernel<float><<<B * H, 32 * nw, smem, S_>>>((const float*)qkv, emb, p, (float*)o, probs);
  } else if (dtype == STG_BF16) {
    STG_CUDA_CHECK(cudaFuncSetAttribute(relattn_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    relattn_fwd_kernel<bf16><<<B * H, 32 * nw, smem, S_>>>((const bf16*)qkv, emb, p, (bf16*)o, probs);
  } else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_relattn_bwd(const void* qkv, int dtype, const float* emb, const float* probs, const void* dout, int B, int L,
                               int H, int d, int max_rel, float scale, void* dqkv, stg_stream_t stream) {
  if (!qkv || !emb || !probs || !dout || !dqkv || B < 1 || L < 1 || H < 1 || d < 4 || (d & 3) || max_rel < 1) return STG_EINVAL;
  const int Lp = (L + 3) & ~3;
  int nw = 16;
  size_t smem = attn_smem(L, d, max_rel, 4, false, 2 * Lp, Lp, nw);
  if (smem > 220 * 1024) { nw = 8; smem = attn_smem(L, d, max_rel, 4, false, 2 * Lp, Lp, nw); }
  if (smem > 220 * 1024) return STG_EUNSUPPORTED;
  AttnP p{B, L, H, d, max_rel, scale};
  if (dtype == STG_F32) {
    STG_CUDA_CHECK(cudaFuncSetAttribute(relattn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    relattn_bwd_kernel<float><<<B * H, 32 * nw, smem, S_>>>((const float*)qkv, emb, probs, (const float*)dout, p, (float*)dqkv);
  } else if (dtype == STG_BF16) {
    STG_CUDA_CHECK(cudaFuncSetAttribute(relattn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    relattn_bwd_kernel<bf16><<<B * H, 32 * nw, smem, S_>>>((const bf16*)qkv, emb, probs, (const bf16*)dout, p, (bf16*)dqkv);
  } else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_encoder_losses(const float* unit_pred, const float* unit_target, const float* phoneme_logits,
                                  const int64_t* phoneme_target, int N, int Du, int P, float* slots, float gs_units, float gs_phonemes,
                                  void* d_units, void* d_logits, int grad_dtype, stg_stream_t stream) {
  if (!unit_pred || !unit_target || !phoneme_logits || !phoneme_target || !slots || N < 1 || Du < 1 || P < 1) return STG_EINVAL;
  const int blocks = (N + 7) / 8;
  if (grad_dtype == STG_F32) encoder_loss_kernel<float><<<blocks, 256, 0, S_>>>(unit_pred, unit_target, phoneme_logits, phoneme_target, N, Du, P, slots, gs_units, gs_phonemes, (float*)d_units, (float*)d_logits);
  else if (grad_dtype == STG_BF16) encoder_loss_kernel<bf16><<<blocks, 256, 0, S_>>>(unit_pred, unit_target, phoneme_logits, phoneme_target, N, Du, P, slots, gs_units, gs_phonemes, (bf16*)d_units, (bf16*)d_logits);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}
