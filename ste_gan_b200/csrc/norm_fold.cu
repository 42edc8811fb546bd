// weight_norm / spectral_norm re-parametrisation kernels: fold (v,g) or (W,u,v) into the
// packed operand layouts the conv engines read, and the matching backward.
//   forward pack  wf [k][c_out][cin_g]      (K-major rows for the forward GEMM)
//   dgrad pack    wd [k][c_in][cout_g]      (K-major rows for the data-gradient GEMM)
// All reductions in fp32.  HBM-bound: each launch reads v once and writes each pack once.
#include "common.cuh"

namespace stg {
namespace {

// scale[co] = g[co] / ||v[co]||   (one block per output channel)
__global__ void __launch_bounds__(256) wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g, int n,
                                                       float* __restrict__ scale) {
  __shared__ float red[32];
  const int co = blockIdx.x;
  const float* row = v + (int64_t)co * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const float x = row[i]; s = fmaf(x, x, s); }
  s = block_sum(s, red);
  if (threadIdx.x == 0) scale[co] = g[co] / sqrtf(s);
}

// Packed operand layouts.  `pg` (pack groups) divides `groups`: the packs describe the convolution as one with
// only pg groups whose per-group matrices are block-diagonal (zeros between the real groups).  This is how
// narrow groups (8..32 channels) are widened to the 64-channel K chunks of the tcgen05 engine.
//   wf[j][co][ci']  ci' in [0, c_in/pg):  input channel c = (co / (c_out/pg)) * (c_in/pg) + ci'
// grid: (ceil(cin_gp*k/256), c_out)
template <typename T>
__global__ void __launch_bounds__(256) pack_fwd_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                                       int c_out, int cin_g, int k, int groups, int pg,
                                                       T* __restrict__ wf) {
  const int co = blockIdx.y;
  const int c_in = cin_g * groups, cin_gp = c_in / pg, cout_gp = c_out / pg, cout_g = c_out / groups;
  const int idx = blockIdx.x * 256 + threadIdx.x;  // j*cin_gp + ci'
  if (idx >= cin_gp * k) return;
  const int j = idx / cin_gp, cip = idx - j * cin_gp;
  const int c = (co / cout_gp) * cin_gp + cip;
  const int gc = c / cin_g;
  float w = 0.f;
  if (gc == co / cout_g) w = v[((int64_t)co * cin_g + (c - gc * cin_g)) * k + j] * scale[co];
  wf[((int64_t)j * c_out + co) * cin_gp + cip] = from_f<T>(w);
}

//   wd[j][ci][co']  co' in [0, c_out/pg):  output channel co = (ci / (c_in/pg)) * (c_out/pg) + co'
// 32x32 smem transpose; grid: (ceil(cout_gp/32), ceil(cin_gp/32), k*pg)
template <typename T>
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                                         int c_out, int cin_g, int k, int groups, int pg,
                                                         T* __restrict__ wd) {
  __shared__ float tile[32][33];
  const int c_in = cin_g * groups, cin_gp = c_in / pg, cout_gp = c_out / pg, cout_g = c_out / groups;
  const int j = blockIdx.z % k, gp = blockIdx.z / k;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {  // r: co', tx: ci'
    const int cop = co0 + r, cip = ci0 + tx;
    float w = 0.f;
    if (cop < cout_gp && cip < cin_gp) {
      const int co = gp * cout_gp + cop, c = gp * cin_gp + cip;
      const int gc = c / cin_g;
      if (gc == co / cout_g) w = v[((int64_t)co * cin_g + (c - gc * cin_g)) * k + j] * scale[co];
    }
    tile[r][tx] = w;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {  // r: ci', tx: co'
    const int cip = ci0 + r, cop = co0 + tx;
    if (cip < cin_gp && cop < cout_gp)
      wd[((int64_t)j * c_in + gp * cin_gp + cip) * cout_gp + cop] = from_f<T>(tile[tx][r]);
  }
}

// "Unfolded" packs for tiny-channel first layers (groups == 1): taps and channels form ONE K axis of Kp >= k*c_in
// elements (zero padded), matching stg_unfold's im2col rows:  wf[co][j*c_in + c], wd[j*c_in + c][co].
template <typename T>
__global__ void __launch_bounds__(256) pack_unfold_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                                          int c_out, int c_in, int k, int Kp, T* __restrict__ wf,
                                                          T* __restrict__ wd) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= c_out * Kp) return;
  const int co = idx / Kp, q = idx - co * Kp;
  float w = 0.f;
  if (q < k * c_in) {
    const int j = q / c_in, c = q - j * c_in;
    w = v[((int64_t)co * c_in + c) * k + j] * scale[co];
  }
  if (wf) wf[(int64_t)co * Kp + q] = from_f<T>(w);
  if (wd) wd[(int64_t)q * c_out + co] = from_f<T>(w);
}

// Gradient layouts: element (co, j, ci) of dw lives at co*dw_ld + j*span + goff(co) + ci.  Compact: span = cin_g,
// goff = 0.  "Span" layout of the tcgen05 wgrad for grouped convs (wgrad_tc.cu): span = input channels met by a
// 128-row co tile, goff = offset of co's own group inside that span.
__device__ __forceinline__ int span_goff(int co, int cin_g, int cout_g, int span) {
  if (span == cin_g || cout_g >= 128) return 0;
  return ((co / cout_g) % (128 / cout_g)) * cin_g;
}

// dv[co][ci][j] (+)= scale*dw[co][j][ci] - (g*dot/norm^3) v ; dg[co] (+)= dot/norm
__global__ void __launch_bounds__(256) wn_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                                     const float* __restrict__ g, int cin_g, int k, int dw_ld, int span,
                                                     int cout_g, float* __restrict__ dv, float* __restrict__ dg,
                                                     int accumulate) {
  __shared__ float red[32];
  const int co = blockIdx.x, n = cin_g * k;
  const float* vr = v + (int64_t)co * n;
  const float* dr = dw + (int64_t)co * dw_ld + span_goff(co, cin_g, cout_g, span);
  float ss = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    const float x = vr[i];
    ss = fmaf(x, x, ss);
    dot = fmaf(x, dr[j * span + ci], dot);
  }
  ss = block_sum(ss, red);
  dot = block_sum(dot, red);
  const float norm = sqrtf(ss), gg = g[co];
  const float a = gg / norm, bcoef = gg * dot / (norm * ss);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    const float val = a * dr[j * span + ci] - bcoef * vr[i];
    float* o = dv + (int64_t)co * n + i;
    *o = accumulate ? (*o + val) : val;
  }
  if (threadIdx.x == 0) dg[co] = accumulate ? (dg[co] + dot / norm) : dot / norm;
}

// ---- spectral norm
// out[col] += sum_{row in chunk} W[row][col] * u[row]     grid (ceil(n/256), ceil(rows/32))
__global__ void __launch_bounds__(256) sn_wt_u_kernel(const float* __restrict__ W, const float* __restrict__ u, int rows,
                                                      int n, float* __restrict__ out) {
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col >= n) return;
  const int r0 = blockIdx.y * 32, r1 = min(rows, r0 + 32);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s = fmaf(W[(int64_t)r * n + col], u[r], s);
  atomicAdd(out + col, s);
}
// x <- x / max(||x||, eps) ; optionally dst = normalized, and sigma = <normalized, raw>
__global__ void __launch_bounds__(1024) sn_normalize_kernel(const float* __restrict__ raw, int n, float eps,
                                                            float* __restrict__ dst, float* __restrict__ sigma) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(raw[i], raw[i], s);
  s = block_sum(s, red);
  const float nrm = fmaxf(sqrtf(s), eps);
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = raw[i] / nrm;
  if (sigma && threadIdx.x == 0) *sigma = s / nrm;
}
// out[row] = sum_col W[row][col] * v[col]   (block per row)
__global__ void __launch_bounds__(256) sn_w_v_kernel(const float* __restrict__ W, const float* __restrict__ v, int n,
                                                     float* __restrict__ out) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(W[(int64_t)row * n + i], v[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[row] = s;
}
// Fused forms (the power iteration of a layer is a chain of dependent launches on the step's critical path: 7 -> 4):
// raw_u[row] = sum_col W[row][col] * v[col] with v = raw_v / max(||raw_v||, eps) normalised ON THE FLY (every block re-derives the
// norm of the <= 2560-element vector); block 0 also stores v (the module's buffer) and the copy the backward will use.
__global__ void __launch_bounds__(256) sn_w_v_norm_kernel(const float* __restrict__ W, const float* __restrict__ raw_v, int n,
                                                          float eps, float* __restrict__ out, float* __restrict__ v_dst,
                                                          float* __restrict__ v_copy) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) ss = fmaf(raw_v[i], raw_v[i], ss);
  ss = block_sum(ss, red);
  const float nrm = fmaxf(sqrtf(ss), eps);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float vi = raw_v[i] / nrm;
    s = fmaf(W[(int64_t)row * n + i], vi, s);
    if (row == 0) { v_dst[i] = vi; if (v_copy) v_copy[i] = vi; }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[row] = s;
}
// u = raw / max(||raw||, eps), sigma = ||raw||^2 / max(||raw||, eps), scale[0..c_out) = 1 / sigma, optional copy of u - one launch
// instead of normalise + fill-scale + a device-to-device copy.
__global__ void __launch_bounds__(1024) sn_finish_u_kernel(const float* __restrict__ raw, int n, float eps, float* __restrict__ u,
                                                           float* __restrict__ u_copy, float* __restrict__ sigma,
                                                           float* __restrict__ scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(raw[i], raw[i], s);
  s = block_sum(s, red);
  const float nrm = fmaxf(sqrtf(s), eps), sg = s / nrm;
  __syncthreads();      // every thread has read raw[] (scale aliases it) before anyone overwrites
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float ui = raw[i] / nrm;
    u[i] = ui;
    if (u_copy) u_copy[i] = ui;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) scale[i] = 1.f / sg;
  if (threadIdx.x == 0) *sigma = sg;
}
__global__ void __launch_bounds__(1024) sn_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int n,
                                                      float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(a[i], b[i], s);
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s;
}
__global__ void sn_fill_scale_kernel(const float* __restrict__ sigma, int c_out, float* __restrict__ scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c_out) scale[i] = 1.f / sigma[0];
}
// acc += sum dw[co][j][ci] * W[co][ci][j]
__global__ void __launch_bounds__(256) sn_bwd_dot_kernel(const float* __restrict__ dw, const float* __restrict__ W,
                                                         int cin_g, int k, int dw_ld, int span, int cout_g,
                                                         float* __restrict__ acc) {
  __shared__ float red[32];
  const int co = blockIdx.x, n = cin_g * k;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    s = fmaf(W[(int64_t)co * n + i], dw[(int64_t)co * dw_ld + span_goff(co, cin_g, cout_g, span) + j * span + ci], s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}
__global__ void __launch_bounds__(256) sn_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ u,
                                                     const float* __restrict__ v, const float* __restrict__ sigma,
                                                     const float* __restrict__ dot, int cin_g, int k, int dw_ld,
                                                     int span, int cout_g, float* __restrict__ dW, int accumulate) {
  const int co = blockIdx.x, n = cin_g * k;
  const float sg = sigma[0], coef = dot[0] / (sg * sg) * u[co], inv = 1.f / sg;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ci = i / k, j = i - ci * k;
    const float val = dw[(int64_t)co * dw_ld + span_goff(co, cin_g, cout_g, span) + j * span + ci] * inv - coef * v[i];
    float* o = dW + (int64_t)co * n + i;
    if (accumulate) atomicAdd(o, val);   // the fake and the real pass may un-fold this layer concurrently (two streams)
    else *o = val;
  }
}

// ---------------------------------------------------------------------------------------------- multi-tensor
// One launch for ALL weight-normed convs of a network: the per-layer kernels above are 5-10 us each and a
// network has 45 (G) / 31 (D) of them, i.e. more launch gaps than work.  `items` is a device-resident table.
__device__ __forceinline__ int find_item(const StgFoldItem* __restrict__ items, int n, int idx, bool by_tile) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int first = by_tile ? items[mid].tile0 : items[mid].row0;
    if (first <= idx) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) wn_scale_multi_kernel(const StgFoldItem* __restrict__ items, int n) {
  __shared__ float red[32];
  const int it = find_item(items, n, blockIdx.x, false);
  const StgFoldItem d = items[it];
  const int co = blockIdx.x - d.row0, len = d.cin_g * d.k;
  const float* row = d.v + (int64_t)co * len;
  float s = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) { const float x = row[i]; s = fmaf(x, x, s); }
  s = block_sum(s, red);
  if (threadIdx.x == 0) d.scale[co] = d.g[co] / sqrtf(s);
}

// one 32 x 32 (co' x ci') tile of one tap of one pack group: writes the wf tile and the transposed wd tile
template <typename T>
__global__ void __launch_bounds__(256) pack_multi_kernel(const StgFoldItem* __restrict__ items, int n) {
  __shared__ float tile[32][33];
  const int it = find_item(items, n, blockIdx.x, true);
  const StgFoldItem d = items[it];
  int t = blockIdx.x - d.tile0;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  T* wf = static_cast<T*>(d.wf);
  T* wd = static_cast<T*>(d.wd);
  if (d.flags & STG_PACK_UNFOLD) {
    // tiles over (co, q) with q = j*c_in + c in [0, Kp)
    const int Kp = (d.k * d.cin_g + 7) / 8 * 8, tq = (Kp + 31) / 32;
    const int co0 = (t / tq) * 32, q0 = (t % tq) * 32;
    for (int r = ty; r < 32; r += 8) {  // r: co, tx: q
      const int co = co0 + r, q = q0 + tx;
      float w = 0.f;
      if (co < d.c_out && q < d.k * d.cin_g) {
        const int j = q / d.cin_g, c = q - j * d.cin_g;
        w = d.v[((int64_t)co * d.cin_g + c) * d.k + j] * d.scale[co];
      }
      tile[r][tx] = w;
      if (wf && co < d.c_out && q < Kp) wf[(int64_t)co * Kp + q] = from_f<T>(w);
    }
    __syncthreads();
    if (wd) {
      for (int r = ty; r < 32; r += 8) {  // r: q, tx: co
        const int q = q0 + r, co = co0 + tx;
        if (q < Kp && co < d.c_out) wd[(int64_t)q * d.c_out + co] = from_f<T>(tile[tx][r]);
      }
    }
    return;
  }
  const int pg = d.pg, c_in = d.cin_g * d.groups, cin_gp = c_in / pg, cout_gp = d.c_out / pg, cout_g = d.c_out / d.groups;
  const int tco = (cout_gp + 31) / 32, tci = (cin_gp + 31) / 32;
  const int ci0 = (t % tci) * 32; t /= tci;
  const int co0 = (t % tco) * 32; t /= tco;
  const int j = t % d.k, gp = t / d.k;
  for (int r = ty; r < 32; r += 8) {  // r: co', tx: ci'
    const int cop = co0 + r, cip = ci0 + tx;
    float w = 0.f;
    if (cop < cout_gp && cip < cin_gp) {
      const int co = gp * cout_gp + cop, c = gp * cin_gp + cip;
      const int gc = c / d.cin_g;
      if (gc == co / cout_g) w = d.v[((int64_t)co * d.cin_g + (c - gc * d.cin_g)) * d.k + j] * d.scale[co];
      if (wf) wf[((int64_t)j * d.c_out + co) * cin_gp + cip] = from_f<T>(w);
    }
    tile[r][tx] = w;
  }
  __syncthreads();
  if (wd) {
    for (int r = ty; r < 32; r += 8) {  // r: ci', tx: co'
      const int cip = ci0 + r, cop = co0 + tx;
      if (cip < cin_gp && cop < cout_gp)
        wd[((int64_t)j * c_in + gp * cin_gp + cip) * cout_gp + cop] = from_f<T>(tile[tx][r]);
    }
  }
}

// Row form of the fold (forward pack only - the tcgen05 data-gradient reads it too) and of its backward: one WARP per
// output channel (weight row), see fold_row_warp below.  Replaced the scale + 32x32-tile pack pair, whose v reads were
// strided by k, and a block-per-row version that transposed (ci, j) -> (j, ci) through shared memory.
#ifndef STG_FOLD_RPW
#define STG_FOLD_RPW 1
#endif
constexpr int FOLD_RPW = STG_FOLD_RPW;     // rows per warp: 1 = one table search per row, but 4x the warps in flight (4 rows per warp
                                           // left the discriminator's 7.8 K rows with 13 warps per SM, each a serial chain of rows)
constexpr int FOLD_RPB = 8 * FOLD_RPW;   // rows per block: one table search per warp / block, the item is then walked forward
// Both row kernels are pure HBM streams (v 4 B + pack 2 B per weight; dw + v + dv 12 B per weight).  The first versions
// moved 4 bytes per thread per instruction with one block per row: ~1 KB in flight per block, 1.2-1.4 TB/s.
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool al16_host(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- warp-per-row fast path (plain convs: groups == pack groups == 1, k in {1, 3, 5}, c_in % 4 == 0, 16-byte aligned).
// 4 input channels x k taps of the torch layout v[co][ci][j] are exactly k consecutive float4: a lane reads them, has the
// (ci, j) -> (j, ci) transposition in REGISTERS and writes 4 consecutive channels per tap (8 B bf16 / 16 B fp32, the
// warp's stores contiguous).  No shared memory, no block barrier: a warp streams its row twice (norm pass, then scale +
// store pass - the second read hits L1/L2) with two channel groups per lane in flight.  ncu before: 1.7 TB/s,
// sm__throughput 55 % (shared-memory transposition + block reductions).
template <typename T> __device__ __forceinline__ void store4(T* dst, float a, float b, float c, float d);
template <> __device__ __forceinline__ void store4<float>(float* dst, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(dst) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<bf16>(bf16* dst, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  *reinterpret_cast<uint2*>(dst) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
__device__ __forceinline__ bool fold_fast(const StgFoldItem& d) {
  return d.flags == 0 && d.groups == 1 && d.pg == 1 && (d.k == 1 || d.k == 3 || d.k == 5) && (d.cin_g & 3) == 0 &&
         al16(d.v) && al16(d.wf);
}
__device__ __forceinline__ bool fold_bwd_fast(const StgFoldItem& d) {
  const int span = d.dw_span > 0 ? d.dw_span : d.cin_g, ld = d.dw_ld > 0 ? d.dw_ld : span * d.k;
  return d.flags == 0 && d.groups == 1 && (d.k == 1 || d.k == 3 || d.k == 5) && (d.cin_g & 3) == 0 && (span & 3) == 0 &&
         (ld & 3) == 0 && al16(d.v) && al16(d.dv) && al16(d.dw);
}

template <typename T, int K>
__device__ __forceinline__ void fold_row_warp(const StgFoldItem& d, int co, int lane) {
  const int cin = d.cin_g, ng = cin >> 2;
  const float4* v4 = reinterpret_cast<const float4*>(d.v + (int64_t)co * cin * K);
  float ss = 0.f;
  for (int g = lane; g < ng; g += 64) {
    float4 a[K], b[K];
    const bool two = g + 32 < ng;
#pragma unroll
    for (int q = 0; q < K; ++q) a[q] = v4[g * K + q];
#pragma unroll
    for (int q = 0; q < K; ++q) b[q] = two ? v4[(g + 32) * K + q] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < K; ++q)
      ss += a[q].x * a[q].x + a[q].y * a[q].y + a[q].z * a[q].z + a[q].w * a[q].w +
            b[q].x * b[q].x + b[q].y * b[q].y + b[q].z * b[q].z + b[q].w * b[q].w;
  }
  ss = warp_sum(ss);
  const float sc = d.g[co] / sqrtf(ss);
  if (lane == 0) d.scale[co] = sc;
  T* wf = static_cast<T*>(d.wf);
  for (int g = lane; g < ng; g += 32) {
    float f[4 * K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      const float4 x = v4[g * K + q];
      f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < K; ++j)   // f[c * K + j] = v[co][4g + c][j]
      store4<T>(wf + ((int64_t)j * d.c_out + co) * cin + 4 * g, f[j] * sc, f[K + j] * sc, f[2 * K + j] * sc, f[3 * K + j] * sc);
  }
}

template <int K>
__device__ __forceinline__ void fold_bwd_row_warp(const StgFoldItem& d, int co, int lane, int accumulate) {
  const int cin = d.cin_g, ng = cin >> 2;
  const int span = d.dw_span > 0 ? d.dw_span : cin, ld = d.dw_ld > 0 ? d.dw_ld : span * K;
  const float4* v4 = reinterpret_cast<const float4*>(d.v + (int64_t)co * cin * K);
  const float* dr = d.dw + (int64_t)co * ld;
  float4* o4 = reinterpret_cast<float4*>(d.dv + (int64_t)co * cin * K);
  float ss = 0.f, dot = 0.f;
  for (int g = lane; g < ng; g += 32) {
    float f[4 * K], w[K][4];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      const float4 x = v4[g * K + q];
      f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float4 x = *reinterpret_cast<const float4*>(dr + j * span + 4 * g);
      w[j][0] = x.x; w[j][1] = x.y; w[j][2] = x.z; w[j][3] = x.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < K; ++j) { ss = fmaf(f[c * K + j], f[c * K + j], ss); dot = fmaf(f[c * K + j], w[j][c], dot); }
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float norm = sqrtf(ss), gg = d.g[co];
  const float a = gg / norm, bcoef = gg * dot / (norm * ss);
  for (int g = lane; g < ng; g += 32) {
    float f[4 * K], w[K][4], o[4 * K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      const float4 x = v4[g * K + q];
      f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const float4 x = *reinterpret_cast<const float4*>(dr + j * span + 4 * g);
      w[j][0] = x.x; w[j][1] = x.y; w[j][2] = x.z; w[j][3] = x.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int j = 0; j < K; ++j) o[c * K + j] = a * w[j][c] - bcoef * f[c * K + j];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      float4 r = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
      if (accumulate) { const float4 p = o4[g * K + q]; r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w; }
      o4[g * K + q] = r;
    }
  }
  if (lane == 0) d.dg[co] = accumulate ? (d.dg[co] + dot / norm) : dot / norm;
}

// generic warp-per-row path (grouped convs with block-diagonal pack groups, im2col packs, odd layouts): these rows are
// short (k37 grouped: 16-32 channels x 37 taps; first layers: 8 channels), so the strided re-reads of v stay in L1
template <typename T>
__device__ __forceinline__ void fold_row_warp_generic(const StgFoldItem& d, int co, int lane) {
  const int n = d.cin_g * d.k, k = d.k, cin_g = d.cin_g;
  const float* vr = d.v + (int64_t)co * n;
  float ss = 0.f;
  for (int i = lane; i < n; i += 32) { const float x = vr[i]; ss = fmaf(x, x, ss); }
  ss = warp_sum(ss);
  const float sc = d.g[co] / sqrtf(ss);
  if (lane == 0) d.scale[co] = sc;
  T* wf = static_cast<T*>(d.wf);
  if (d.flags & STG_PACK_UNFOLD) {           // wf[co][q], q = j*c_in + c  (groups == 1)
    const int Kp = (k * cin_g + 7) / 8 * 8;
    for (int q = lane; q < Kp; q += 32) {
      float w = 0.f;
      if (q < k * cin_g) { const int j = q / cin_g, c = q - j * cin_g; w = vr[c * k + j] * sc; }
      wf[(int64_t)co * Kp + q] = from_f<T>(w);
    }
    return;
  }
  const int pg = d.pg, c_in = cin_g * d.groups, cin_gp = c_in / pg, cout_gp = d.c_out / pg, cout_g = d.c_out / d.groups;
  // this row's own group occupies columns [off, off + cin_g) of its pack group; the rest of the row is zero
  const int off = ((co / cout_g) - (co / cout_gp) * (d.groups / pg)) * cin_g;
  for (int j = 0; j < k; ++j) {
    T* dst = wf + ((int64_t)j * d.c_out + co) * cin_gp;
    for (int cip = lane; cip < cin_gp; cip += 32) {
      const int c = cip - off;
      dst[cip] = from_f<T>((c >= 0 && c < cin_g) ? vr[c * k + j] * sc : 0.f);
    }
  }
}

// Data-gradient packs of the grouped convs from their FORWARD packs (row form of the multi-tensor fold): one block per
// (item, tap, pack group) transposes the contiguous [cout_gp x cin_gp] slab wf[j][u*cout_gp ..][:] into the contiguous
// [cin_gp x cout_gp] slab wd[j][u*cin_gp ..][:] through shared memory - both sides fully coalesced (a per-row scatter
// of 2-byte elements cost 40 us per discriminator fold).  Items without wd have no blocks (tile0 = running block count).
template <typename T>
__global__ void __launch_bounds__(256) wd_from_wf_kernel(const StgFoldItem* __restrict__ items, int n_items) {
  __shared__ T slab[8192];
  const int it = find_item(items, n_items, blockIdx.x, true);
  const StgFoldItem d = items[it];
  const int local = blockIdx.x - d.tile0, j = local / d.pg, u = local - j * d.pg;
  const int c_in = d.cin_g * d.groups, cin_gp = c_in / d.pg, cout_gp = d.c_out / d.pg, n = cin_gp * cout_gp;
  const T* src = static_cast<const T*>(d.wf) + ((int64_t)j * d.c_out + (int64_t)u * cout_gp) * cin_gp;
  T* dst = static_cast<T*>(d.wd) + ((int64_t)j * c_in + (int64_t)u * cin_gp) * cout_gp;
  if (n <= 8192) {
    for (int i = threadIdx.x; i < n; i += 256) slab[i] = src[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) { const int c = i / cout_gp, r = i - c * cout_gp; dst[i] = slab[r * cin_gp + c]; }
  } else {   // (wide block-diagonal fallback packs: plain strided reads)
    for (int i = threadIdx.x; i < n; i += 256) { const int c = i / cout_gp, r = i - c * cout_gp; dst[i] = src[(int64_t)r * cin_gp + c]; }
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) wn_fold_rows_kernel(const StgFoldItem* __restrict__ items, int n_items, int total_rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w_begin = blockIdx.x * FOLD_RPB + warp * FOLD_RPW, w_end = min(total_rows, w_begin + FOLD_RPW);
  if (w_begin >= w_end) return;
  int it = find_item(items, n_items, w_begin, false);
  StgFoldItem d = items[it];
  for (int r = w_begin; r < w_end; ++r) {
    while (r >= d.row0 + d.c_out) d = items[++it];
    const int co = r - d.row0;
    if (!fold_fast(d)) fold_row_warp_generic<T>(d, co, lane);
    else if (d.k == 3) fold_row_warp<T, 3>(d, co, lane);
    else if (d.k == 1) fold_row_warp<T, 1>(d, co, lane);
    else fold_row_warp<T, 5>(d, co, lane);
  }
}

__device__ __forceinline__ void fold_bwd_row_warp_generic(const StgFoldItem& d, int co, int lane, int accumulate) {
  const int n = d.cin_g * d.k, k = d.k, cin_g = d.cin_g;
  const int span = d.dw_span > 0 ? d.dw_span : cin_g, ld = d.dw_ld > 0 ? d.dw_ld : span * k;
  const float* vr = d.v + (int64_t)co * n;
  const float* dr = d.dw + (int64_t)co * ld + span_goff(co, cin_g, d.c_out / d.groups, span);
  float* orow = d.dv + (int64_t)co * n;
  // First pass: lanes walk the GRADIENT in its own order (tap-major: dr[j * span + ci], consecutive ci) so that its first
  // touch is coalesced; the matching v elements (ci * k + j) are k floats apart - a row is at most a few KB and stays in
  // L1.  (Walking v's order made every warp load of the k = 37 grouped rows touch 32 different lines of dw.)
  float ss = 0.f, dot = 0.f;
  const int total = k * cin_g;
  for (int i = lane; i < total; i += 32) {
    const int j = i / cin_g, ci = i - j * cin_g;
    const float x = vr[ci * k + j];
    ss = fmaf(x, x, ss);
    dot = fmaf(x, dr[j * span + ci], dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float norm = sqrtf(ss), gg = d.g[co];
  const float a = gg / norm, bcoef = gg * dot / (norm * ss);
  for (int i = lane; i < n; i += 32) {   // second pass in v / dv order (coalesced stores); the row's dw lines are in L1 now
    const int ci = i / k, j = i - ci * k;
    const float val = a * dr[j * span + ci] - bcoef * vr[i];
    orow[i] = accumulate ? (orow[i] + val) : val;
  }
  if (lane == 0) d.dg[co] = accumulate ? (d.dg[co] + dot / norm) : dot / norm;
}

__global__ void __launch_bounds__(256, 4) wn_bwd_multi_kernel(const StgFoldItem* __restrict__ items, int n_items, int row_base,
                                                           int total_rows, int accumulate) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows [row_base, row_base + total_rows)
  const int w_begin = row_base + blockIdx.x * FOLD_RPB + warp * FOLD_RPW, w_end = min(row_base + total_rows, w_begin + FOLD_RPW);
  if (w_begin >= w_end) return;
  int it = find_item(items, n_items, w_begin, false);
  StgFoldItem d = items[it];
  for (int r = w_begin; r < w_end; ++r) {
    while (r >= d.row0 + d.c_out) d = items[++it];
    const int co = r - d.row0;
    if (!fold_bwd_fast(d)) fold_bwd_row_warp_generic(d, co, lane, accumulate);
    else if (d.k == 3) fold_bwd_row_warp<3>(d, co, lane, accumulate);
    else if (d.k == 1) fold_bwd_row_warp<1>(d, co, lane, accumulate);
    else fold_bwd_row_warp<5>(d, co, lane, accumulate);
  }
}

// Forward pack of a plain conv (groups == 1, k in {1, 3, 5}, c_in % 4 == 0) with a given per-row scale: the register
// transposition of fold_row_warp without the norm pass.  Used for the spectral-norm layers (scale = 1 / sigma), whose
// largest (512 -> 1024, k 5: 10.5 MB) took 23 us per power iteration in the generic strided pack.
template <typename T, int K>
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ v, const float* __restrict__ scale, int c_out,
                                                        int cin, T* __restrict__ wf) {
  const int co = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (co >= c_out) return;
  const float sc = scale[co];
  const float4* v4 = reinterpret_cast<const float4*>(v + (int64_t)co * cin * K);
  for (int g = lane; g < (cin >> 2); g += 32) {
    float f[4 * K];
#pragma unroll
    for (int q = 0; q < K; ++q) {
      const float4 x = v4[g * K + q];
      f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
      store4<T>(wf + ((int64_t)j * c_out + co) * cin + 4 * g, f[j] * sc, f[K + j] * sc, f[2 * K + j] * sc, f[3 * K + j] * sc);
  }
}

// the same transposition for ONE conv (per-layer folds: the spectral-norm layers re-pack on every forward)
template <typename T>
__global__ void __launch_bounds__(256) wd_from_wf_direct_kernel(const T* __restrict__ wf, T* __restrict__ wd, int c_out, int c_in,
                                                                int pg) {
  __shared__ T slab[8192];
  const int j = blockIdx.x / pg, u = blockIdx.x - j * pg;
  const int cin_gp = c_in / pg, cout_gp = c_out / pg, n = cin_gp * cout_gp;
  const T* src = wf + ((int64_t)j * c_out + (int64_t)u * cout_gp) * cin_gp;
  T* dst = wd + ((int64_t)j * c_in + (int64_t)u * cin_gp) * cout_gp;
  if (n <= 8192) {
    for (int i = threadIdx.x; i < n; i += 256) slab[i] = src[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += 256) { const int c = i / cout_gp, r = i - c * cout_gp; dst[i] = slab[r * cin_gp + c]; }
  } else {
    for (int i = threadIdx.x; i < n; i += 256) { const int c = i / cout_gp, r = i - c * cout_gp; dst[i] = src[(int64_t)r * cin_gp + c]; }
  }
}

template <typename T>
int launch_packs(const float* v, const float* scale, int c_out, int cin_g, int k, int groups, int pg, int flags, T* wf,
                 T* wd, cudaStream_t s) {
  if (flags & STG_PACK_UNFOLD) {
    if (groups != 1) return STG_EINVAL;
    const int Kp = (k * cin_g + 7) / 8 * 8;
    pack_unfold_kernel<T><<<ceil_div(c_out * Kp, 256), 256, 0, s>>>(v, scale, c_out, cin_g, k, Kp, wf, wd);
    STG_LAUNCH_CHECK();
    return STG_OK;
  }
  if (pg <= 0) pg = groups;
  if (groups % pg) return STG_EINVAL;
  const int c_in = cin_g * groups, cin_gp = c_in / pg, cout_gp = c_out / pg;
  if (wf && groups == 1 && (k == 1 || k == 3 || k == 5) && (cin_g & 3) == 0 && al16_host(v) && al16_host(wf)) {
    const int nb = ceil_div(c_out, 8);
    if (k == 1) pack_rows_kernel<T, 1><<<nb, 256, 0, s>>>(v, scale, c_out, cin_g, wf);
    else if (k == 3) pack_rows_kernel<T, 3><<<nb, 256, 0, s>>>(v, scale, c_out, cin_g, wf);
    else pack_rows_kernel<T, 5><<<nb, 256, 0, s>>>(v, scale, c_out, cin_g, wf);
    STG_LAUNCH_CHECK();
  } else if (wf) {
    dim3 g1(ceil_div(cin_gp * k, 256), c_out);
    pack_fwd_kernel<T><<<g1, 256, 0, s>>>(v, scale, c_out, cin_g, k, groups, pg, wf);
    STG_LAUNCH_CHECK();
  }
  if (wd && wf && sizeof(T) == 2 && groups > 1) {
    // grouped convs on the tensor engine (bf16): the K-major data-gradient pack is the per-group transposition of the forward
    // pack just written - contiguous slabs in, contiguous slabs out (the spectral-norm layers re-pack on every forward)
    wd_from_wf_direct_kernel<T><<<k * pg, 256, 0, s>>>(wf, wd, c_out, c_in, pg);
    STG_LAUNCH_CHECK();
  } else if (wd) {
    dim3 g2(ceil_div(cout_gp, 32), ceil_div(cin_gp, 32), k * pg);
    pack_dgrad_kernel<T><<<g2, 256, 0, s>>>(v, scale, c_out, cin_g, k, groups, pg, wd);
    STG_LAUNCH_CHECK();
  }
  return STG_OK;
}

}  // namespace
}  // namespace stg

using namespace stg;

extern "C" int stg_weightnorm_fold(const float* v, const float* g, int c_out, int cin_g, int k, int groups,
                                   int pack_groups, int flags, int dtype, void* wf, void* wd, float* scale,
                                   stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!v || !g || !scale || c_out < 1 || cin_g < 1 || k < 1 || groups < 1 || c_out % groups) return STG_EINVAL;
  wn_scale_kernel<<<c_out, 256, 0, s>>>(v, g, cin_g * k, scale);
  STG_LAUNCH_CHECK();
  if (dtype == STG_F32) return launch_packs<float>(v, scale, c_out, cin_g, k, groups, pack_groups, flags, (float*)wf, (float*)wd, s);
  if (dtype == STG_BF16) return launch_packs<bf16>(v, scale, c_out, cin_g, k, groups, pack_groups, flags, (bf16*)wf, (bf16*)wd, s);
  return STG_EINVAL;
}

extern "C" int stg_weightnorm_fold_bwd(const float* dw, int dw_ld, int dw_span, const float* v, const float* g, int c_out,
                                       int cin_g, int k, int groups, float* dv, float* dg, int accumulate,
                                       stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!dw || !v || !g || !dv || !dg || groups < 1) return STG_EINVAL;
  if (dw_span <= 0) dw_span = cin_g;
  if (dw_ld <= 0) dw_ld = dw_span * k;
  wn_bwd_kernel<<<c_out, 256, 0, s>>>(dw, v, g, cin_g, k, dw_ld, dw_span, c_out / groups, dv, dg, accumulate);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_spectralnorm_fold(const float* w_orig, float* u, float* v, int c_out, int cin_g, int k, int groups,
                                     int pack_groups, int flags, int training, int dtype, void* wf, void* wd,
                                     float* sigma_out, float* scratch, float* u_used, float* v_used, stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!w_orig || !u || !v || !sigma_out || !scratch) return STG_EINVAL;
  const int n = cin_g * k;
  float* raw_u = scratch;          // [c_out]
  float* raw_v = scratch + c_out;  // [n]
  float* scale = raw_u;            // reused after u is final
  const float eps = 1e-12f;
  if (training) {
    STG_CUDA_CHECK(cudaMemsetAsync(raw_v, 0, sizeof(float) * n, s));
    dim3 g1(ceil_div(n, 256), ceil_div(c_out, 32));
    sn_wt_u_kernel<<<g1, 256, 0, s>>>(w_orig, u, c_out, n, raw_v);
    STG_LAUNCH_CHECK();
    // v = normalize(W^T u) folded into the W v product; u = raw/max(||raw||,eps), sigma = <u, W v> = ||raw||^2 / max(||raw||, eps),
    // the per-channel scale 1/sigma and the (u, v) copies that the backward of THIS forward needs: four launches in all
    sn_w_v_norm_kernel<<<c_out, 256, 0, s>>>(w_orig, raw_v, n, eps, raw_u, v, v_used);
    STG_LAUNCH_CHECK();
    sn_finish_u_kernel<<<1, 1024, 0, s>>>(raw_u, c_out, eps, u, u_used, sigma_out, scale);
    STG_LAUNCH_CHECK();
  } else {
    sn_w_v_kernel<<<c_out, 256, 0, s>>>(w_orig, v, n, raw_u);
    STG_LAUNCH_CHECK();
    sn_dot_kernel<<<1, 1024, 0, s>>>(u, raw_u, c_out, sigma_out);
    STG_LAUNCH_CHECK();
    sn_fill_scale_kernel<<<ceil_div(c_out, 256), 256, 0, s>>>(sigma_out, c_out, scale);
    STG_LAUNCH_CHECK();
    if (u_used) STG_CUDA_CHECK(cudaMemcpyAsync(u_used, u, sizeof(float) * c_out, cudaMemcpyDeviceToDevice, s));
    if (v_used) STG_CUDA_CHECK(cudaMemcpyAsync(v_used, v, sizeof(float) * n, cudaMemcpyDeviceToDevice, s));
  }
  if (dtype == STG_F32) return launch_packs<float>(w_orig, scale, c_out, cin_g, k, groups, pack_groups, flags, (float*)wf, (float*)wd, s);
  if (dtype == STG_BF16) return launch_packs<bf16>(w_orig, scale, c_out, cin_g, k, groups, pack_groups, flags, (bf16*)wf, (bf16*)wd, s);
  return STG_EINVAL;
}

extern "C" int stg_spectralnorm_fold_bwd(const float* dw, int dw_ld, int dw_span, const float* w_orig, const float* u,
                                         const float* v, const float* sigma, int c_out, int cin_g, int k, int groups,
                                         float* dw_orig, int accumulate, float* scratch, stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!dw || !w_orig || !u || !v || !sigma || !dw_orig || !scratch || groups < 1) return STG_EINVAL;
  if (dw_span <= 0) dw_span = cin_g;
  if (dw_ld <= 0) dw_ld = dw_span * k;
  STG_CUDA_CHECK(cudaMemsetAsync(scratch, 0, sizeof(float), s));
  sn_bwd_dot_kernel<<<c_out, 256, 0, s>>>(dw, w_orig, cin_g, k, dw_ld, dw_span, c_out / groups, scratch);
  STG_LAUNCH_CHECK();
  sn_bwd_kernel<<<c_out, 256, 0, s>>>(dw, u, v, sigma, scratch, cin_g, k, dw_ld, dw_span, c_out / groups, dw_orig, accumulate);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_weightnorm_fold_multi(const StgFoldItem* items, int n_items, int total_rows, int total_tiles, int dtype,
                                         stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!items || n_items < 1 || total_rows < 1) return STG_EINVAL;
  if (total_tiles <= 0) {  // row form: forward packs; -total_tiles = transposition blocks of the grouped convs' wd packs
    if (dtype == STG_F32) wn_fold_rows_kernel<float><<<ceil_div(total_rows, FOLD_RPB), 256, 0, s>>>(items, n_items, total_rows);
    else if (dtype == STG_BF16) wn_fold_rows_kernel<bf16><<<ceil_div(total_rows, FOLD_RPB), 256, 0, s>>>(items, n_items, total_rows);
    else return STG_EINVAL;
    STG_LAUNCH_CHECK();
    if (total_tiles < 0) {
      if (dtype == STG_F32) return STG_EINVAL;           // (fp32 packs come from the tile form)
      wd_from_wf_kernel<bf16><<<-total_tiles, 256, 0, s>>>(items, n_items);
      STG_LAUNCH_CHECK();
    }
    return STG_OK;
  }
  wn_scale_multi_kernel<<<total_rows, 256, 0, s>>>(items, n_items);
  STG_LAUNCH_CHECK();
  if (dtype == STG_F32) pack_multi_kernel<float><<<total_tiles, 256, 0, s>>>(items, n_items);
  else if (dtype == STG_BF16) pack_multi_kernel<bf16><<<total_tiles, 256, 0, s>>>(items, n_items);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_weightnorm_fold_bwd_multi(const StgFoldItem* items, int n_items, int total_rows, int accumulate,
                                             stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!items || n_items < 1 || total_rows < 1) return STG_EINVAL;
  wn_bwd_multi_kernel<<<ceil_div(total_rows, FOLD_RPB), 256, 0, s>>>(items, n_items, 0, total_rows, accumulate);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_weightnorm_fold_bwd_range(const StgFoldItem* items, int n_items, int row_base, int n_rows, int accumulate,
                                             stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!items || n_items < 1 || row_base < 0 || n_rows < 1) return STG_EINVAL;
  wn_bwd_multi_kernel<<<ceil_div(n_rows, FOLD_RPB), 256, 0, s>>>(items, n_items, row_base, n_rows, accumulate);
  STG_LAUNCH_CHECK();
  return STG_OK;
}
