// Small HBM-bound data-movement kernels around the convolutions: generator input
// assembly (units + session embedding), discriminator input preparation (reflect pad,
// AvgPool1d(4,2,1)), casts, pair-sums, flat AdamW.  All coalesced along channels.
#include "common.cuh"

namespace stg {
namespace {

template <typename T>
__global__ void embed_concat_kernel(const float* __restrict__ units, const float* __restrict__ emb,
                                    const int64_t* __restrict__ ids, int T_, int du, int de, int64_t total,
                                    T* __restrict__ x0) {
  const int C = du + de;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bt = i / C;
    const int b = (int)(bt / T_);
    const float v = c < du ? units[bt * du + c] : emb[ids[b] * de + (c - du)];
    x0[i] = from_f<T>(v);
  }
}

// demb[ids[b]][e] += sum_t dx0[b][t][du + e]     grid: B blocks, de threads (<= 1024)
template <typename T>
__global__ void embed_bwd_kernel(const T* __restrict__ dx0, const int64_t* __restrict__ ids, int T_, int du, int de,
                                 float* __restrict__ demb) {
  const int b = blockIdx.x, e = threadIdx.x;
  if (e >= de) return;
  const int C = du + de;
  float s = 0.f;
  for (int t = 0; t < T_; ++t) s += to_f(dx0[((int64_t)b * T_ + t) * C + du + e]);
  atomicAdd(demb + ids[b] * de + e, s);
}

__device__ __forceinline__ int reflect_right(int t, int T_) { return t < T_ ? t : 2 * (T_ - 1) - t; }

template <typename T>
__global__ void reflect_pad_kernel(const float* __restrict__ x, int T_, int C, int Tp, int64_t total, T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bt = i / C;
    const int t = (int)(bt % Tp), b = (int)(bt / Tp);
    out[i] = from_f<T>(x[((int64_t)b * T_ + reflect_right(t, T_)) * C + c]);
  }
}
template <typename T>
__global__ void reflect_pad_bwd_kernel(const T* __restrict__ dout, int T_, int C, int Tp, int64_t total_in,
                                       float* __restrict__ dx) {
  // gather form: each input position t receives dout[t] (+ dout[2(T-1)-t] if that lies in the padded tail)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_in; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bt = i / C;
    const int t = (int)(bt % T_), b = (int)(bt / T_);
    float g = to_f(dout[((int64_t)b * Tp + t) * C + c]);
    const int m = 2 * (T_ - 1) - t;
    if (m >= T_ && m < Tp) g += to_f(dout[((int64_t)b * Tp + m) * C + c]);
    atomicAdd(dx + i, g);   // the period stacks' backward branches run on different streams and all add into dx
  }
}

// AvgPool1d(kernel 4, stride 2, padding 1, count_include_pad): out[o] = (x[2o-1]+x[2o]+x[2o+1]+x[2o+2])/4
__global__ void avgpool4_kernel(const float* __restrict__ x, int T_, int C, int To, int64_t total, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bo = i / C;
    const int o = (int)(bo % To), b = (int)(bo / To);
    float s = 0.f;
#pragma unroll
    for (int j = -1; j <= 2; ++j) {
      const int t = 2 * o + j;
      if (t >= 0 && t < T_) s += x[((int64_t)b * T_ + t) * C + c];
    }
    out[i] = 0.25f * s;
  }
}
__global__ void avgpool4_bwd_kernel(const float* __restrict__ dout, int T_, int C, int To, int64_t total_in,
                                    float* __restrict__ dx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_in; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t bt = i / C;
    const int t = (int)(bt % T_), b = (int)(bt / T_);
    // outputs o with 2o-1 <= t <= 2o+2  <=>  (t-2)/2 <= o <= (t+1)/2
    float s = 0.f;
    const int lo = (t - 2 + 1) >> 1;  // ceil((t-2)/2), valid for t-2 >= -1 via arithmetic shift
    for (int o = lo; o <= (t + 1) / 2; ++o)
      if (o >= 0 && o < To && 2 * o - 1 <= t && t <= 2 * o + 2) s += dout[((int64_t)b * To + o) * C + c];
    dx[i] += 0.25f * s;
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f<D>(to_f(src[i]));
}

// out rows (dup: rows 2r and 2r+1) = relu(src row r): the ReLU -> nn.Upsample(2) prologue of a stand-alone GBlock
template <typename T>
__global__ void relu_rows_kernel(const T* __restrict__ src, int64_t rows, int C, int dup, T* __restrict__ out) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    const T v = from_f<T>(fmaxf(to_f(src[i]), 0.f));
    if (dup) { out[(2 * r) * C + c] = v; out[(2 * r + 1) * C + c] = v; }
    else out[i] = v;
  }
}

template <typename T>
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int mode, int64_t n,
                               T* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = from_f<T>(dy[i] * act_grad_from_output(mode, y[i]));
}

template <typename T>
__global__ void pair_sum_kernel(const T* __restrict__ in, int64_t rows_out, int C, T* __restrict__ out) {
  const int64_t total = rows_out * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    out[i] = from_f<T>(to_f(in[(2 * r) * C + c]) + to_f(in[(2 * r + 1) * C + c]));
  }
}

template <typename T>
__global__ void axpy_kernel(float* __restrict__ y, const T* __restrict__ x, float alpha, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, to_f(x[i]), y[i]);
}

__global__ void step_inc_kernel(int64_t* step, const int32_t* __restrict__ enable) {
  if (enable == nullptr || *enable != 0) *step += 1;
}
__global__ void flag_clear_kernel(int32_t* flag) { *flag = 0; }

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr_host,
                                                    const float* __restrict__ lr_dev, float b1, float b2, float eps,
                                                    float wd, const int64_t* __restrict__ step, float gscale,
                                                    const int32_t* __restrict__ enable) {
  // `enable` (device flag, or NULL = always): a CUDA graph that holds this update replays it unconditionally; the flag
  // turns the replay into a no-op when there is no pending gradient (first replay, or after an explicit flush)
  if (enable != nullptr && *enable == 0) return;
  const float lr = lr_dev ? *lr_dev : lr_host;   // device-resident learning rate: a schedule changes it without re-capturing graphs
  const float t = (float)(*step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2), decay = 1.f - lr * wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = p[i] * decay - step_size * (mi / denom);
  }
}

inline int grid_for(int64_t n) {
  int64_t b = ceil_div64(n, 256);
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// im2col rows for tiny-channel first layers: one thread per (row, tap) copies C channels (C = 8 -> one 16-byte move)
template <typename T>
__global__ void unfold_kernel(const T* __restrict__ src, int phases, int t_src, int t_dst, int C, int k, int dil, int stride,
                              int pad, int Kp, int64_t total, T* __restrict__ out) {
  const int slots = Kp / C + ((Kp % C) ? 1 : 0);  // k real taps (+ one zero slot when Kp > k*C)
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % slots);
    const int64_t row = i / slots;                      // b*t_dst*phases + t*phases + ph
    const int ph = (int)(row % phases);
    const int64_t bt = row / phases;
    const int t = (int)(bt % t_dst);
    const int64_t b = bt / t_dst;
    const int ts = t * stride + j * dil - pad;
    const bool ok = j < k && ts >= 0 && ts < t_src;
    const T* s = src + ((b * t_src + ts) * phases + ph) * C;
    T* o = out + row * Kp + (int64_t)j * C;
    const int n = min(C, Kp - j * C);
    for (int c = 0; c < n; ++c) o[c] = ok ? s[c] : from_f<T>(0.f);
  }
}
template <typename T>
__global__ void unfold_bwd_kernel(const T* __restrict__ dout, int phases, int t_src, int t_dst, int C, int k, int dil,
                                  int stride, int pad, int Kp, int64_t total, float* __restrict__ dsrc) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t row = i / C;                          // b*t_src*phases + ts*phases + ph
    const int ph = (int)(row % phases);
    const int64_t bt = row / phases;
    const int ts = (int)(bt % t_src);
    const int64_t b = bt / t_src;
    float g = 0.f;
    for (int j = 0; j < k; ++j) {
      const int num = ts + pad - j * dil;
      if (num < 0) break;
      const int t = num / stride;
      if (t * stride == num && t < t_dst) g += to_f(dout[((b * t_dst + t) * phases + ph) * Kp + j * C + c]);
    }
    dsrc[i] += g;
  }
}

}  // namespace
}  // namespace stg

using namespace stg;
#define S_ static_cast<cudaStream_t>(stream)

extern "C" int stg_embed_concat(const float* units, const float* emb, const int64_t* ids, int B, int T, int d_units,
                                int d_emb, int dtype, void* x0, stg_stream_t stream) {
  if (!units || !x0 || (d_emb > 0 && (!emb || !ids))) return STG_EINVAL;
  const int64_t total = (int64_t)B * T * (d_units + d_emb);
  if (dtype == STG_F32) embed_concat_kernel<float><<<grid_for(total), 256, 0, S_>>>(units, emb, ids, T, d_units, d_emb, total, (float*)x0);
  else embed_concat_kernel<bf16><<<grid_for(total), 256, 0, S_>>>(units, emb, ids, T, d_units, d_emb, total, (bf16*)x0);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_embed_concat_bwd(const void* dx0, const int64_t* ids, int B, int T, int d_units, int d_emb, int dtype,
                                    float* demb, stg_stream_t stream) {
  if (!dx0 || !ids || !demb || d_emb > 1024) return STG_EINVAL;
  if (dtype == STG_F32) embed_bwd_kernel<float><<<B, d_emb, 0, S_>>>((const float*)dx0, ids, T, d_units, d_emb, demb);
  else embed_bwd_kernel<bf16><<<B, d_emb, 0, S_>>>((const bf16*)dx0, ids, T, d_units, d_emb, demb);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_reflect_pad_right(const float* x, int B, int T, int C, int T_pad, int dtype, void* out, stg_stream_t stream) {
  if (!x || !out || T_pad < T || T_pad - T > T - 1) return STG_EINVAL;
  const int64_t total = (int64_t)B * T_pad * C;
  if (dtype == STG_F32) reflect_pad_kernel<float><<<grid_for(total), 256, 0, S_>>>(x, T, C, T_pad, total, (float*)out);
  else reflect_pad_kernel<bf16><<<grid_for(total), 256, 0, S_>>>(x, T, C, T_pad, total, (bf16*)out);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_reflect_pad_right_bwd(const void* dout, int B, int T, int C, int T_pad, int dtype, float* dx, stg_stream_t stream) {
  if (!dout || !dx) return STG_EINVAL;
  const int64_t total = (int64_t)B * T * C;
  if (dtype == STG_F32) reflect_pad_bwd_kernel<float><<<grid_for(total), 256, 0, S_>>>((const float*)dout, T, C, T_pad, total, dx);
  else reflect_pad_bwd_kernel<bf16><<<grid_for(total), 256, 0, S_>>>((const bf16*)dout, T, C, T_pad, total, dx);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_avgpool4(const float* x, int B, int T, int C, float* out, stg_stream_t stream) {
  if (!x || !out) return STG_EINVAL;
  const int To = (T + 2 - 4) / 2 + 1;
  const int64_t total = (int64_t)B * To * C;
  avgpool4_kernel<<<grid_for(total), 256, 0, S_>>>(x, T, C, To, total, out);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_avgpool4_bwd(const float* dout, int B, int T, int C, float* dx, stg_stream_t stream) {
  if (!dout || !dx) return STG_EINVAL;
  const int To = (T + 2 - 4) / 2 + 1;
  const int64_t total = (int64_t)B * T * C;
  avgpool4_bwd_kernel<<<grid_for(total), 256, 0, S_>>>(dout, T, C, To, total, dx);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_cast(const void* src, int sd, void* dst, int dd, int64_t n, stg_stream_t stream) {
  if (!src || !dst) return STG_EINVAL;
  const int g = grid_for(n);
  if (sd == STG_F32 && dd == STG_BF16) cast_kernel<float, bf16><<<g, 256, 0, S_>>>((const float*)src, (bf16*)dst, n);
  else if (sd == STG_BF16 && dd == STG_F32) cast_kernel<bf16, float><<<g, 256, 0, S_>>>((const bf16*)src, (float*)dst, n);
  else if (sd == STG_F32 && dd == STG_F32) cast_kernel<float, float><<<g, 256, 0, S_>>>((const float*)src, (float*)dst, n);
  else if (sd == STG_BF16 && dd == STG_BF16) cast_kernel<bf16, bf16><<<g, 256, 0, S_>>>((const bf16*)src, (bf16*)dst, n);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_relu_rows(const void* src, int dtype, int64_t rows, int C, int dup, void* out, stg_stream_t stream) {
  if (!src || !out || rows < 1 || C < 1) return STG_EINVAL;
  if (dtype == STG_F32) relu_rows_kernel<float><<<grid_for(rows * C), 256, 0, S_>>>((const float*)src, rows, C, dup, (float*)out);
  else if (dtype == STG_BF16) relu_rows_kernel<bf16><<<grid_for(rows * C), 256, 0, S_>>>((const bf16*)src, rows, C, dup, (bf16*)out);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_act_bwd(const float* dy, const float* y, int mode, int64_t n, int dtype, void* out, stg_stream_t stream) {
  if (!dy || !y || !out) return STG_EINVAL;
  if (dtype == STG_F32) act_bwd_kernel<float><<<grid_for(n), 256, 0, S_>>>(dy, y, mode, n, (float*)out);
  else act_bwd_kernel<bf16><<<grid_for(n), 256, 0, S_>>>(dy, y, mode, n, (bf16*)out);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_pair_sum_rows(const void* in, int64_t rows_out, int C, int dtype, void* out, stg_stream_t stream) {
  if (!in || !out) return STG_EINVAL;
  const int g = grid_for(rows_out * C);
  if (dtype == STG_F32) pair_sum_kernel<float><<<g, 256, 0, S_>>>((const float*)in, rows_out, C, (float*)out);
  else pair_sum_kernel<bf16><<<g, 256, 0, S_>>>((const bf16*)in, rows_out, C, (bf16*)out);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_axpy_f32(float* y, const void* x, int x_dtype, float alpha, int64_t n, stg_stream_t stream) {
  if (!y || !x) return STG_EINVAL;
  if (x_dtype == STG_F32) axpy_kernel<float><<<grid_for(n), 256, 0, S_>>>(y, (const float*)x, alpha, n);
  else axpy_kernel<bf16><<<grid_for(n), 256, 0, S_>>>(y, (const bf16*)x, alpha, n);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, const float* lr_dev, float beta1,
                         float beta2, float eps, float weight_decay, int64_t* step_count, float grad_scale, int32_t* enable,
                         int flags, stg_stream_t stream) {
  if (!p || !g || !m || !v || !step_count) return STG_EINVAL;
  if (!(flags & STG_ADAMW_KEEP_STEP)) {
    step_inc_kernel<<<1, 1, 0, S_>>>(step_count, enable);
    STG_LAUNCH_CHECK();
  }
  adamw_kernel<<<grid_for(n), 256, 0, S_>>>(p, g, m, v, n, lr, lr_dev, beta1, beta2, eps, weight_decay, step_count, grad_scale, enable);
  STG_LAUNCH_CHECK();
  if (enable) {                    // consumed: the next replay is a no-op until somebody sets the flag again
    flag_clear_kernel<<<1, 1, 0, S_>>>(enable);
    STG_LAUNCH_CHECK();
  }
  return STG_OK;
}

extern "C" int stg_unfold(const void* src, int dtype, int B, int phases, int t_src, int t_dst, int C, int k, int dilation,
                          int stride, int pad, void* out, stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!src || !out || B < 1 || phases < 1 || C < 1 || k < 1 || stride < 1 || dilation < 1) return STG_EINVAL;
  const int Kp = (k * C + 7) / 8 * 8;
  const int slots = Kp / C + ((Kp % C) ? 1 : 0);
  const int64_t total = (int64_t)B * t_dst * phases * slots;
  const int64_t nb = ceil_div64(total, 256);
  const int blocks = (int)(nb < 148 * 16 ? nb : 148 * 16);
  if (dtype == STG_F32) unfold_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(src), phases, t_src, t_dst, C, k, dilation, stride, pad, Kp, total, static_cast<float*>(out));
  else if (dtype == STG_BF16) unfold_kernel<bf16><<<blocks, 256, 0, s>>>(static_cast<const bf16*>(src), phases, t_src, t_dst, C, k, dilation, stride, pad, Kp, total, static_cast<bf16*>(out));
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}

extern "C" int stg_unfold_bwd(const void* dout, int dtype, int B, int phases, int t_src, int t_dst, int C, int k,
                              int dilation, int stride, int pad, float* dsrc, stg_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!dout || !dsrc || B < 1 || phases < 1 || C < 1 || k < 1 || stride < 1 || dilation < 1) return STG_EINVAL;
  const int Kp = (k * C + 7) / 8 * 8;
  const int64_t total = (int64_t)B * t_src * phases * C;
  const int64_t nb = ceil_div64(total, 256);
  const int blocks = (int)(nb < 148 * 16 ? nb : 148 * 16);
  if (dtype == STG_F32) unfold_bwd_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(dout), phases, t_src, t_dst, C, k, dilation, stride, pad, Kp, total, dsrc);
  else if (dtype == STG_BF16) unfold_bwd_kernel<bf16><<<blocks, 256, 0, s>>>(static_cast<const bf16*>(dout), phases, t_src, t_dst, C, k, dilation, stride, pad, Kp, total, dsrc);
  else return STG_EINVAL;
  STG_LAUNCH_CHECK();
  return STG_OK;
}
