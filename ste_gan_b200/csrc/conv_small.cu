// 1-channel logits convolutions (the `output` layers of every sub-discriminator, models/discriminator.py:31,59,
// 81,111: Conv(C -> 1, k, stride 1)).  With one output channel these are matrix-vector products: HBM/L2-bound,
// no tensor-core shape fits them, and the generic engines spent 15-45 us of fixed cost on each.  bf16 storage,
// fp32 accumulation, period views supported (taps step `phases` rows).
//
//   forward     y[b][r]      = bias + sum_j sum_c x[b][r + (j - pad)*P][c] * w[j][c]                 (fp32 logits)
//   data grad   dx[b][r][c]  = (sum_j dy[b][r + (pad - j)*P] * w[j][c] + add_pre[b][r][c]) * act'(mask[b][r][c])
//   weight grad dw[j][c]    += sum_{b,r} dy[b][r] * x[b][r + (j - pad)*P][c] ;  dbias += sum dy
// (rows r = h*P + phase; a tap leaves the sample when h + j - pad is outside [0, H)).
#include "common.cuh"

namespace stg {
namespace {

constexpr int MAXK = 8;

__device__ __forceinline__ void unpack8(const uint4& q, float* o) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    o[2 * i] = __low2float(h);
    o[2 * i + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

struct C1P {
  int rows, P, H, C, k, pad;       // rows = H * P per sample
  const bf16 *x, *w;               // x [B][rows][C], w [k][C]
  const float* bias;
  float* y;                        // [B][rows]
};

// One warp per input row: the k dot products of that row with the k taps, scattered (atomically) to the k output
// rows it feeds (y is zeroed first; the centre tap adds the bias).  grid: (ceil(rows / 16), B), 256 threads = 8 warps x
// 2 rows.  CI = C / 256 (0: any C): with CI known every 16-byte load of a row (and, once per warp, the lane's slice of
// the taps, kept packed in registers) is issued before the first use - the first version walked 4 rows x C/256 chunks
// with one load in flight per lane and staged fp32 taps in shared memory behind a block barrier (12-16 us per launch
// for 13 MB that sit in L2).
__device__ __forceinline__ float dot8(const uint4& a, const uint4& b) {
  float x[8], w[8];
  unpack8(a, x); unpack8(b, w);
  return x[0] * w[0] + x[1] * w[1] + x[2] * w[2] + x[3] * w[3] + x[4] * w[4] + x[5] * w[5] + x[6] * w[6] + x[7] * w[7];
}
template <int K, int CI>
__global__ void __launch_bounds__(256) c1_fwd_kernel(const C1P p) {
  constexpr int RW = 2;                       // rows per warp
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const float bias0 = p.bias ? p.bias[0] : 0.f;
  const int r0 = (blockIdx.x * 8 + warp) * RW;
  if (r0 >= p.rows) return;
  float acc[RW][K];
#pragma unroll
  for (int rr = 0; rr < RW; ++rr)
#pragma unroll
    for (int j = 0; j < K; ++j) acc[rr][j] = 0.f;
  if constexpr (CI > 0) {
    uint4 wq[K][CI], xq[RW][CI];
#pragma unroll
    for (int rr = 0; rr < RW; ++rr) {
      const bf16* xr = p.x + ((int64_t)b * p.rows + min(r0 + rr, p.rows - 1)) * p.C;
#pragma unroll
      for (int i = 0; i < CI; ++i) xq[rr][i] = *reinterpret_cast<const uint4*>(xr + i * 256 + lane * 8);
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
      for (int i = 0; i < CI; ++i) wq[j][i] = __ldg(reinterpret_cast<const uint4*>(p.w + j * p.C + i * 256 + lane * 8));
#pragma unroll
    for (int rr = 0; rr < RW; ++rr)
#pragma unroll
      for (int i = 0; i < CI; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) acc[rr][j] += dot8(xq[rr][i], wq[j][i]);
  } else {
    for (int rr = 0; rr < RW; ++rr) {
      const bf16* xr = p.x + ((int64_t)b * p.rows + min(r0 + rr, p.rows - 1)) * p.C;
      for (int c = lane * 8; c < p.C; c += 256) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xr + c);
#pragma unroll
        for (int j = 0; j < K; ++j) acc[rr][j] += dot8(xv, __ldg(reinterpret_cast<const uint4*>(p.w + j * p.C + c)));
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < RW; ++rr) {
#pragma unroll
    for (int j = 0; j < K; ++j) acc[rr][j] = warp_sum(acc[rr][j]);
    const int r = r0 + rr;
    if (lane == 0 && r < p.rows) {
      const int h = r / p.P;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int ho = h - (j - p.pad);          // output row whose tap j reads input row h
        // the centre tap reaches every output row exactly once: it carries the bias (y starts at 0)
        if (ho >= 0 && ho < p.H) atomicAdd(p.y + (int64_t)b * p.rows + r - (j - p.pad) * p.P, acc[rr][j] + (j == p.pad ? bias0 : 0.f));
      }
    }
  }
}

struct C1D {
  int rows, P, H, C, k, pad, mask_mode;
  const bf16 *dy, *w, *mask, *add_pre;   // dy [B][rows], w [k][C]
  bf16* dx;                              // [B][rows][C]
};
// thread per 8 channels of one row; grid: (ceil(rows*C/8 / 256), B)
template <int K>
__global__ void __launch_bounds__(256) c1_dgrad_kernel(const C1D p) {
  const int cg = p.C / 8;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)p.rows * cg) return;
  const int r = (int)(idx / cg), c = (int)(idx - (int64_t)r * cg) * 8, b = blockIdx.y;
  const int h = r / p.P;
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int hs = h + p.pad - j;              // dy row whose tap j reads this input row
    if (hs < 0 || hs >= p.H) continue;
    const float g = to_f(p.dy[(int64_t)b * p.rows + r + (p.pad - j) * p.P]);
    float wv[8];
    unpack8(*reinterpret_cast<const uint4*>(p.w + j * p.C + c), wv);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(g, wv[i], v[i]);
  }
  const int64_t off = ((int64_t)b * p.rows + r) * p.C + c;
  if (p.add_pre) {
    float t[8];
    unpack8(*reinterpret_cast<const uint4*>(p.add_pre + off), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += t[i];
  }
  if (p.mask) {
    float t[8];
    unpack8(*reinterpret_cast<const uint4*>(p.mask + off), t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= act_grad_from_output(p.mask_mode, t[i]);
  }
  *reinterpret_cast<uint4*>(p.dx + off) = pack8(v);
}

struct C1W {
  int rows, P, H, C, k, pad, rows_per_block;
  const bf16 *x, *dy;
  float *dw, *dbias;   // dw [k][C]
};
// block: a chunk of rows of one sample, thread per 8 channels (C/8 <= 256 threads active per row lane)
// grid: (ceil(rows / rows_per_block), B)
template <int K>
__global__ void __launch_bounds__(256) c1_wgrad_kernel(const C1W p) {
  __shared__ float red[32];
  const int cg = p.C / 8, lanes = 256 / cg;     // row lanes
  const int rl = threadIdx.x / cg, c = (threadIdx.x - rl * cg) * 8, b = blockIdx.y;
  const int r_begin = blockIdx.x * p.rows_per_block, r_end = min(p.rows, r_begin + p.rows_per_block);
  float acc[K][8];
#pragma unroll
  for (int j = 0; j < K; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
  float bsum = 0.f;
  if (rl < lanes) {
#pragma unroll 8
    for (int r = r_begin + rl; r < r_end; r += lanes) {
      const int h = r / p.P;
      float xv[8];
      unpack8(*reinterpret_cast<const uint4*>(p.x + ((int64_t)b * p.rows + r) * p.C + c), xv);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int ho = h - (j - p.pad);        // output row whose tap j reads input row h
        if (ho < 0 || ho >= p.H) continue;
        const float g = to_f(p.dy[(int64_t)b * p.rows + r - (j - p.pad) * p.P]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(g, xv[i], acc[j][i]);
      }
    }
  }
  if (p.dbias) {
    for (int r = r_begin + threadIdx.x; r < r_end; r += 256) bsum += to_f(p.dy[(int64_t)b * p.rows + r]);
    bsum = block_sum(bsum, red);
    if (threadIdx.x == 0) atomicAdd(p.dbias, bsum);
  }
  // reduce the row lanes through shared memory, then one atomic per (tap, channel) per block
  __shared__ float part[256 * 8];
  for (int j = 0; j < K; ++j) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) part[threadIdx.x * 8 + i] = acc[j][i];
    __syncthreads();
    if (rl == 0) {
      for (int l = 1; l < lanes; ++l)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][i] += part[(threadIdx.x + l * cg) * 8 + i];
      float* dst = p.dw + j * p.C + c;
      if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {   // 16-byte vector reductions: a quarter of the L2 atomic operations
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(acc[j][0]), "f"(acc[j][1]), "f"(acc[j][2]), "f"(acc[j][3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(acc[j][4]), "f"(acc[j][5]), "f"(acc[j][6]), "f"(acc[j][7]) : "memory");
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(dst + i, acc[j][i]);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// First layer of a period stack, fused: reflect-pad right (discriminator.py:36,86) + period view + Conv2d(C -> c_out,
// (k,1), stride (s,1), padding (pad,0)) + bias + LeakyReLU, from the fp32 discriminator input straight to the bf16
// channels-last feature map.  K = k*C is 24-40: as im2col rows + a one-stage tensor-core GEMM this was three launches
// (pad, unfold, conv: 4 + 6 + 14 us) for ~3 MB of output; here one thread produces 8 output channels of one output row.
// Inputs and weights are rounded to bf16 exactly as that path did (fp32 accumulation).
struct PFirst {
  int B, T, C, P, H, Ho, c_out, k, stride, pad, Kp;
  float slope;
  const float* x;      // [B][T][C] fp32
  const bf16* w;       // unfold pack [c_out][Kp], q = j*C + c
  const float* bias;   // [c_out]
  bf16* y;             // [B][Ho*P][c_out]
};
__global__ void __launch_bounds__(256) period_first_kernel(const PFirst p) {
  extern __shared__ float wsm[];                 // [c_out][k*C + 1] (+1: the 4 channel groups of a row hit 4 banks) | bias
  const int kc = p.k * p.C, ld = kc + 1;
  float* bsm = wsm + p.c_out * ld;
  for (int i = threadIdx.x; i < p.c_out * kc; i += 256) { const int co = i / kc, q = i - co * kc; wsm[co * ld + q] = to_f(p.w[co * p.Kp + q]); }
  for (int i = threadIdx.x; i < p.c_out; i += 256) bsm[i] = p.bias ? p.bias[i] : 0.f;
  __syncthreads();
  const int cg = p.c_out >> 3, rows = p.Ho * p.P;
  const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (int64_t)p.B * rows * cg) return;
  const int g = (int)(idx % cg);
  const int64_t rr = idx / cg;
  const int row = (int)(rr % rows), b = (int)(rr / rows);
  const int ho = row / p.P, ph = row - ho * p.P;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = bsm[g * 8 + o];
  for (int j = 0; j < p.k; ++j) {
    const int hi = ho * p.stride + j - p.pad;
    if (hi < 0 || hi >= p.H) continue;
    int t = hi * p.P + ph;
    if (t >= p.T) t = 2 * (p.T - 1) - t;          // reflected tail
    const float* xr = p.x + ((int64_t)b * p.T + t) * p.C;
    for (int c = 0; c < p.C; ++c) {
      const float xv = __bfloat162float(__float2bfloat16_rn(xr[c]));
      const float* wr = wsm + (g * 8) * ld + j * p.C + c;
#pragma unroll
      for (int o = 0; o < 8; ++o) acc[o] = fmaf(xv, wr[o * ld], acc[o]);
    }
  }
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = fmaxf(acc[o], p.slope * acc[o]);
  *reinterpret_cast<uint4*>(p.y + ((int64_t)b * rows + row) * p.c_out + g * 8) = pack8(acc);
}

}  // namespace

bool conv_c1_supported(const StgConv* d) {
  if (d->dtype != STG_BF16 || d->groups != 1 || d->stride != 1 || d->dilation != 1 || d->k > MAXK) return false;
  if (d->pair_sum || d->dup_rows || d->add_post || d->post_shift) return false;
  if (!d->transposed) {
    return d->c_dst == 1 && (d->c_src % 8) == 0 && d->y_raw && d->out_f32 && !d->y_act && !d->add_pre && !d->mask &&
           d->t_src == d->t_dst;
  }
  return d->c_src == 1 && (d->c_dst % 8) == 0 && d->c_dst <= 2048 && d->y_raw && !d->out_f32 && !d->y_act && !d->bias &&
         d->t_src == d->t_dst;
}

template <int K>
static int launch_c1(const StgConv* d, cudaStream_t s) {
  const int rows = d->t_dst * d->phases;
  if (!d->transposed) {
    C1P p{rows, d->phases, d->t_dst, d->c_src, d->k, d->pad, static_cast<const bf16*>(d->src), static_cast<const bf16*>(d->w),
          d->bias, static_cast<float*>(d->y_raw)};
    if (d->pad < 0 || d->pad >= K) return STG_EUNSUPPORTED;   // the centre tap carries the bias
    STG_CUDA_CHECK(cudaMemsetAsync(p.y, 0, sizeof(float) * (size_t)d->n_samples * rows, s));
    dim3 grid(ceil_div(rows, 16), d->n_samples);
    switch (d->c_src) {
      case 256: c1_fwd_kernel<K, 1><<<grid, 256, 0, s>>>(p); break;
      case 512: c1_fwd_kernel<K, 2><<<grid, 256, 0, s>>>(p); break;
      case 1024: c1_fwd_kernel<K, 4><<<grid, 256, 0, s>>>(p); break;
      default: c1_fwd_kernel<K, 0><<<grid, 256, 0, s>>>(p); break;
    }
    STG_LAUNCH_CHECK();
    return STG_OK;
  }
  C1D p{rows, d->phases, d->t_dst, d->c_dst, d->k, d->pad, d->mask_mode, static_cast<const bf16*>(d->src),
        static_cast<const bf16*>(d->w), static_cast<const bf16*>(d->mask), static_cast<const bf16*>(d->add_pre),
        static_cast<bf16*>(d->y_raw)};
  dim3 grid((unsigned)ceil_div64((int64_t)rows * (d->c_dst / 8), 256), d->n_samples);
  c1_dgrad_kernel<K><<<grid, 256, 0, s>>>(p);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

int conv_c1(const StgConv* d, cudaStream_t s) {
  switch (d->k) {
    case 1: return launch_c1<1>(d, s);
    case 3: return launch_c1<3>(d, s);
    case 5: return launch_c1<5>(d, s);
    default: return STG_EUNSUPPORTED;
  }
}

bool wgrad_c1_supported(const StgWgrad* d) {
  return d->dtype == STG_BF16 && d->groups == 1 && d->stride == 1 && d->dilation == 1 && d->c_out == 1 &&
         (d->c_in % 8) == 0 && d->c_in >= 8 && d->c_in <= 2048 && (d->k == 1 || d->k == 3 || d->k == 5) && d->t_in == d->t_out;
}

template <int K>
static int launch_c1w(const StgWgrad* d, cudaStream_t s) {
  const int rows = d->t_out * d->phases;
  C1W p{rows, d->phases, d->t_out, d->c_in, d->k, d->pad, 0, static_cast<const bf16*>(d->x), static_cast<const bf16*>(d->dy),
        d->dw, d->dbias};
  // every block ends with k*C reductions onto the same k*C addresses, so the block count is the contention per
  // address (with scalar atomics 304 blocks made this kernel 28 us for 13 MB of input)
  const int lanes = 256 / (d->c_in / 8);
  int blocks_per_sample = ceil_div(296, d->n_samples);   // (with 16-byte vector reductions; scalar atomics were limited to ~96 blocks)
  int rpb = ceil_div(rows, blocks_per_sample);
  if (rpb < 4 * lanes) rpb = 4 * lanes;
  p.rows_per_block = rpb;
  dim3 grid(ceil_div(rows, rpb), d->n_samples);
  c1_wgrad_kernel<K><<<grid, 256, 0, s>>>(p);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

int wgrad_c1(const StgWgrad* d, cudaStream_t s) {
  if (!d->dw) return STG_EINVAL;
  switch (d->k) {
    case 1: return launch_c1w<1>(d, s);
    case 3: return launch_c1w<3>(d, s);
    case 5: return launch_c1w<5>(d, s);
    default: return STG_EUNSUPPORTED;
  }
}

}  // namespace stg

extern "C" int stg_period_first_layer(const float* x, const void* wf, const float* bias, int B, int T, int C, int period,
                                      int c_out, int k, int stride, int pad, float slope, void* y, stg_stream_t stream) {
  using namespace stg;
  if (!x || !wf || !y || B < 1 || T < 2 || C < 1 || period < 1 || k < 1 || stride < 1 || pad < 0) return STG_EINVAL;
  if ((c_out & 7) != 0 || c_out > 256 || k * C > 256) return STG_EUNSUPPORTED;
  PFirst p;
  p.B = B; p.T = T; p.C = C; p.P = period; p.c_out = c_out; p.k = k; p.stride = stride; p.pad = pad; p.slope = slope;
  const int t_pad = T + (period - T % period);      // reflect pad is always >= 1 (discriminator.py:36,86)
  if (t_pad - T >= T) return STG_EUNSUPPORTED;
  p.H = t_pad / period;
  p.Ho = (p.H + 2 * pad - (k - 1) - 1) / stride + 1;
  p.Kp = (k * C + 7) / 8 * 8;
  p.x = x; p.w = static_cast<const bf16*>(wf); p.bias = bias; p.y = static_cast<bf16*>(y);
  const int64_t total = (int64_t)B * p.Ho * period * (c_out / 8);
  const size_t smem = sizeof(float) * ((size_t)c_out * (k * C + 1) + c_out);
  period_first_kernel<<<(unsigned)((total + 255) / 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

