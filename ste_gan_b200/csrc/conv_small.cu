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
// rows it feeds (y is zeroed first; the centre tap adds the bias).  grid: (ceil(rows/32), B), 256 threads.
template <int K>
__global__ void __launch_bounds__(256) c1_fwd_kernel(const C1P p) {
  extern __shared__ float wsm[];  // [K][C] fp32 taps
  for (int i = threadIdx.x; i < K * p.C; i += 256) wsm[i] = to_f(p.w[i]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const float bias0 = p.bias ? p.bias[0] : 0.f;
  for (int rr = 0; rr < 4; ++rr) {             // 32 rows per block: the taps are staged once per 32 rows
    const int r = blockIdx.x * 32 + warp * 4 + rr;
    if (r >= p.rows) return;
    const bf16* xr = p.x + ((int64_t)b * p.rows + r) * p.C;
    float acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = 0.f;
    for (int c = lane * 8; c < p.C; c += 256) {
      float xv[8];
      unpack8(*reinterpret_cast<const uint4*>(xr + c), xv);
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(&wsm[j * p.C + c]);
        const float4 w1 = *reinterpret_cast<const float4*>(&wsm[j * p.C + c + 4]);
        acc[j] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y + xv[6] * w1.z + xv[7] * w1.w;
      }
    }
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
      const int h = r / p.P;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int ho = h - (j - p.pad);          // output row whose tap j reads input row h
        // the centre tap reaches every output row exactly once: it carries the bias (y starts at 0)
        if (ho >= 0 && ho < p.H) atomicAdd(p.y + (int64_t)b * p.rows + r - (j - p.pad) * p.P, acc[j] + (j == p.pad ? bias0 : 0.f));
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
#pragma unroll 4
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
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(p.dw + j * p.C + c + i, acc[j][i]);
    }
  }
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
    dim3 grid(ceil_div(rows, 32), d->n_samples);
    c1_fwd_kernel<K><<<grid, 256, K * d->c_src * sizeof(float), s>>>(p);
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
  // ~96 blocks in total: every block ends with k*C atomics onto the same k*C addresses, so the block count is the
  // contention per address (304 blocks made this kernel 28 us for 13 MB of input)
  const int lanes = 256 / (d->c_in / 8);
  int blocks_per_sample = ceil_div(96, d->n_samples);
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
