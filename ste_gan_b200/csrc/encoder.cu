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
// LDS.128 of 8 lanes that read 8 different rows hit 32 different banks.  Every product is register-blocked over AR = 4 rows of
// the side that is NOT spread over the lanes: a K / V / E / Q / dO row costs a lane one LDS.128 (4 shared-memory wavefronts)
// and is then used against 4 broadcast rows (1 wavefront each), i.e. 16 FMAs per ~8 wavefronts instead of 4 per ~9.
// History at B = 16, L = 100, d = 96 (per layer): scalar dots 203 us forward / 568 us backward -> float4 dots, 16 warps 115 / 241
// -> 4-row blocks: see DESIGN.md section 3.5.
constexpr int AR = 4;

// shared memory (floats): K [L][ds] | V [L][ds] | per warp: q [AR][d] + logits/probs [AR][Lp] + positional logits [AR][Lt]
// (d % 4 == 0, Lp = roundup4(L), Lt = roundup4(L + AR)); the head's embedding rows are read through L1.
template <typename T>
__global__ void __launch_bounds__(512, 1) relattn_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ emb, AttnP p,
                                                          T* __restrict__ o, float* __restrict__ probs) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int L = p.L, d = p.d, ds = d + 4, d4 = d >> 2, M = p.M, HD = p.H * d, Lp = (L + 3) & ~3, Lt = (L + AR + 3) & ~3;
  float* Ks = sm; float* Vs = Ks + L * ds; float* wsc = Vs + L * ds;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qs = wsc + warp * (AR * d + AR * Lp + AR * Lt); float* lg = qs + AR * d; float* ts = lg + AR * Lp;
  const float* Eh = emb + (int64_t)h * (2 * M - 1) * d;
  const T* base = qkv + (int64_t)b * L * 3 * HD + h * d;
  load_rows(base + HD, 3 * HD, L, d, ds, Ks);
  load_rows(base + 2 * HD, 3 * HD, L, d, ds, Vs);
  __syncthreads();
  for (int ib = warp * AR; ib < L; ib += nw * AR) {
    const int nr = min(AR, L - ib);
    for (int x = lane; x < AR * d; x += 32) {
      const int r = x / d, c = x - r * d;
      qs[x] = r < nr ? to_f(base[(int64_t)(ib + r) * 3 * HD + c]) : 0.f;
    }
    __syncwarp();
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    for (int j = lane; j < L; j += 32) {                          // q . k
      const float4* k4 = reinterpret_cast<const float4*>(Ks + j * ds);
      float acc[AR] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int c = 0; c < d4; ++c) {
        const float4 k = k4[c];
#pragma unroll
        for (int r = 0; r < AR; ++r) acc[r] = dot4(q4[r * d4 + c], k, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < AR; ++r) lg[r * Lp + j] = acc[r] * p.scale;
    }
    // q . emb[rel] for every relative position the block can meet: rel = mi - (ib + AR - 1), i.e. j - i = rel for row i = ib + r
    // at j = mi - (AR - 1) + r
    for (int mi = lane; mi < L + AR - 1; mi += 32) {
      const int rel = mi - (ib + AR - 1);
      const bool in = rel > -M && rel < M;
      float acc[AR] = {0.f, 0.f, 0.f, 0.f};
      if (in) {
        const float4* e4 = reinterpret_cast<const float4*>(Eh + (int64_t)(rel + M - 1) * d);
#pragma unroll 2
        for (int c = 0; c < d4; ++c) {
          const float4 e = e4[c];
#pragma unroll
          for (int r = 0; r < AR; ++r) acc[r] = dot4(q4[r * d4 + c], e, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < AR; ++r) ts[r * Lt + mi] = in ? acc[r] : -1e8f;
    }
    __syncwarp();
    for (int r = 0; r < nr; ++r) {                                // softmax of row ib + r
      float* row = lg + r * Lp;
      const float* tr = ts + r * Lt + (AR - 1 - r);               // tr[j] = positional logit of (ib + r, j)
      float mx = -INFINITY;
      for (int j = lane; j < L; j += 32) { const float v = row[j] + tr[j]; row[j] = v; mx = fmaxf(mx, v); }
      mx = warp_max_f(mx);
      float sum = 0.f;
      for (int j = lane; j < L; j += 32) { const float e = __expf(row[j] - mx); row[j] = e; sum += e; }
      sum = warp_sum_f(sum);
      const float inv = 1.f / sum;
      float* pr = probs + (((int64_t)b * p.H + h) * L + ib + r) * L;
      for (int j = lane; j < L; j += 32) { const float pv = row[j] * inv; row[j] = pv; pr[j] = pv; }
    }
    __syncwarp();
    for (int c = lane; c < d4; c += 32) {                         // probs . v, 4 output channels per lane
      float4 acc[AR];
#pragma unroll
      for (int r = 0; r < AR; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(Vs + j * ds + 4 * c);
#pragma unroll
        for (int r = 0; r < AR; ++r) axpy4(lg[r * Lp + j], v, acc[r]);
      }
      for (int r = 0; r < nr; ++r) store4(o + ((int64_t)b * L + ib + r) * HD + h * d + 4 * c, acc[r], 1.f);
    }
    __syncwarp();
  }
}

// shared memory (floats): Q | K | V | dO [L][ds] each | Dr [Lp] | per warp: dS block [AR][Lp]   (embedding rows through L1)
template <typename T>
__global__ void __launch_bounds__(512, 1) relattn_bwd_kernel(const T* __restrict__ qkv, const float* __restrict__ emb,
                                                          const float* __restrict__ probs, const T* __restrict__ dout, AttnP p,
                                                          T* __restrict__ dqkv) {
  extern __shared__ __align__(16) float sm[];
  const int b = blockIdx.x / p.H, h = blockIdx.x - b * p.H;
  const int L = p.L, d = p.d, ds = d + 4, d4 = d >> 2, M = p.M, HD = p.H * d, Lp = (L + 3) & ~3;
  float* Qs = sm; float* Ks = Qs + L * ds; float* Vs = Ks + L * ds; float* Gs = Vs + L * ds;
  float* Dr = Gs + L * ds; float* wsc = Dr + Lp;
  const float* Eh = emb + (int64_t)h * (2 * M - 1) * d;
  const int nw = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* dS = wsc + warp * AR * Lp;
  const T* base = qkv + (int64_t)b * L * 3 * HD + h * d;
  load_rows(base, 3 * HD, L, d, ds, Qs);
  load_rows(base + HD, 3 * HD, L, d, ds, Ks);
  load_rows(base + 2 * HD, 3 * HD, L, d, ds, Vs);
  load_rows(dout + (int64_t)b * L * HD + h * d, HD, L, d, ds, Gs);
  __syncthreads();
  const float* P = probs + ((int64_t)b * p.H + h) * L * L;
  T* dbase = dqkv + (int64_t)b * L * 3 * HD + h * d;
  // pass 1, a warp per block of AR query rows: dP_ij = dO_i . V_j ; D_i = sum_j dP_ij P_ij ; dS_ij = P_ij (dP_ij - D_i) ;
  //                                             dq_i = sum_j dS_ij (scale K_j + E[j - i])
  for (int ib = warp * AR; ib < L; ib += nw * AR) {
    const int nr = min(AR, L - ib);
    float dsum[AR] = {0.f, 0.f, 0.f, 0.f};
    for (int j = lane; j < L; j += 32) {
      const float4* v4 = reinterpret_cast<const float4*>(Vs + j * ds);
      float acc[AR] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int c = 0; c < d4; ++c) {
        const float4 v = v4[c];
#pragma unroll
        for (int r = 0; r < AR; ++r)
          acc[r] = dot4(*reinterpret_cast<const float4*>(Gs + min(ib + r, L - 1) * ds + 4 * c), v, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < AR; ++r) {
        dS[r * Lp + j] = acc[r];
        if (r < nr) dsum[r] = fmaf(acc[r], P[(int64_t)(ib + r) * L + j], dsum[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < AR; ++r) dsum[r] = warp_sum_f(dsum[r]);
    if (lane == 0)
      for (int r = 0; r < nr; ++r) Dr[ib + r] = dsum[r];
    for (int j = lane; j < L; j += 32)
#pragma unroll
      for (int r = 0; r < AR; ++r) dS[r * Lp + j] = r < nr ? P[(int64_t)(ib + r) * L + j] * (dS[r * Lp + j] - dsum[r]) : 0.f;
    __syncwarp();
    for (int c = lane; c < d4; c += 32) {
      float4 ak[AR], ae[AR];
#pragma unroll
      for (int r = 0; r < AR; ++r) { ak[r] = make_float4(0.f, 0.f, 0.f, 0.f); ae[r] = ak[r]; }
#pragma unroll 2
      for (int j = 0; j < L; ++j) {
        const float4 k = *reinterpret_cast<const float4*>(Ks + j * ds + 4 * c);
#pragma unroll
        for (int r = 0; r < AR; ++r) axpy4(dS[r * Lp + j], k, ak[r]);
      }
      for (int mi = 0; mi < L + AR - 1; ++mi) {                  // relative position rel meets row ib + r at j = mi - (AR - 1) + r
        const int rel = mi - (ib + AR - 1);
        if (rel <= -M || rel >= M) continue;
        const float4 e = *reinterpret_cast<const float4*>(Eh + (int64_t)(rel + M - 1) * d + 4 * c);
#pragma unroll
        for (int r = 0; r < AR; ++r) {
          const int j = mi - (AR - 1) + r;
          if (j >= 0 && j < L) axpy4(dS[r * Lp + j], e, ae[r]);
        }
      }
      for (int r = 0; r < nr; ++r) {
        float4 q = ak[r];
        q.x = q.x * p.scale + ae[r].x; q.y = q.y * p.scale + ae[r].y; q.z = q.z * p.scale + ae[r].z; q.w = q.w * p.scale + ae[r].w;
        store4(dbase + (int64_t)(ib + r) * 3 * HD + 4 * c, q, 1.f);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // pass 2, a warp per block of AR key rows: dS_ij recomputed from D_i ; dk_j = scale sum_i dS_ij q_i ; dv_j = sum_i P_ij dO_i
  for (int jb = warp * AR; jb < L; jb += nw * AR) {
    const int nr = min(AR, L - jb);
    for (int i = lane; i < L; i += 32) {
      const float4* g4 = reinterpret_cast<const float4*>(Gs + i * ds);
      float acc[AR] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
      for (int c = 0; c < d4; ++c) {
        const float4 g = g4[c];
#pragma unroll
        for (int r = 0; r < AR; ++r)
          acc[r] = dot4(g, *reinterpret_cast<const float4*>(Vs + min(jb + r, L - 1) * ds + 4 * c), acc[r]);
      }
      const float di = Dr[i];
#pragma unroll
      for (int r = 0; r < AR; ++r) dS[r * Lp + i] = r < nr ? P[(int64_t)i * L + jb + r] * (acc[r] - di) : 0.f;
    }
    __syncwarp();
    for (int c = lane; c < d4; c += 32) {
      float4 dk[AR], dv[AR];
#pragma unroll
      for (int r = 0; r < AR; ++r) { dk[r] = make_float4(0.f, 0.f, 0.f, 0.f); dv[r] = dk[r]; }
#pragma unroll 2
      for (int i = 0; i < L; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(Qs + i * ds + 4 * c);
        const float4 g = *reinterpret_cast<const float4*>(Gs + i * ds + 4 * c);
#pragma unroll
        for (int r = 0; r < AR; ++r) {
          axpy4(dS[r * Lp + i], q, dk[r]);
          axpy4(r < nr ? P[(int64_t)i * L + jb + r] : 0.f, g, dv[r]);
        }
      }
      for (int r = 0; r < nr; ++r) {
        store4(dbase + (int64_t)(jb + r) * 3 * HD + HD + 4 * c, dk[r], p.scale);
        store4(dbase + (int64_t)(jb + r) * 3 * HD + 2 * HD + 4 * c, dv[r], 1.f);
      }
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

static size_t attn_smem(int L, int d, int n_tiles, int per_warp, int extra, int n_warps) {
  const int ds = d + 4;
  return sizeof(float) * ((size_t)n_tiles * L * ds + extra + (size_t)n_warps * per_warp);
}
constexpr size_t ATTN_SMEM_MAX = 226 * 1024;

extern "C" int stg_relattn_fwd(const void* qkv, int dtype, const float* emb, int B, int L, int H, int d, int max_rel, float scale,
                               void* o, float* probs, stg_stream_t stream) {
  if (!qkv || !emb || !o || !probs || B < 1 || L < 1 || H < 1 || d < 4 || (d & 3) || max_rel < 1) return STG_EINVAL;
  const int Lp = (L + 3) & ~3;
  const int Lt = (L + 4 + 3) & ~3, pw = 4 * d + 4 * Lp + 4 * Lt;      // per-warp scratch (AR = 4 rows)
  int nw = 16;                                          // 16 warps when their scratch fits beside the operand tiles, else 8
  size_t smem = attn_smem(L, d, 2, pw, 0, nw);
  if (smem > ATTN_SMEM_MAX) { nw = 8; smem = attn_smem(L, d, 2, pw, 0, nw); }
  if (smem > ATTN_SMEM_MAX) return STG_EUNSUPPORTED;    // sequence too long for the one-CTA-per-head kernel
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
  size_t smem = attn_smem(L, d, 4, 4 * Lp, Lp, nw);
  if (smem > ATTN_SMEM_MAX) { nw = 8; smem = attn_smem(L, d, 4, 4 * Lp, Lp, nw); }
  if (smem > ATTN_SMEM_MAX) return STG_EUNSUPPORTED;
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
