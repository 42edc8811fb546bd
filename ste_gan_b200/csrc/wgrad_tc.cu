// tcgen05 weight-gradient kernel, bf16 in / fp32 accumulate.
//
//   dw[co][j][ci] += sum_{sample n, row r} dy[n][r][co] * x[n][r*stride + j*dilation - pad][ci]
//
// GEMM view: M = 128 output channels (TMEM lanes), N = BNW input channels per tap (TMEM
// columns, one accumulator block per tap), K = time rows in chunks of 64.  The
// contraction index (time) is the slow axis of the channels-last tensors, so both
// operands are MN-major: a TMA box of (64 channels x 64 rows) lands in shared memory as
// 64 rows of 128 B with 128-byte swizzle, which is exactly the canonical MN-major SW128
// UMMA layout ((8,n),(8,k)):((1,LBO),(8,SBO)) with LBO = bytes between 64-channel boxes
// and SBO = 1024.  Zero padding / dilation / sample boundaries again come from TMA
// out-of-bounds fill.  Split-K over (sample, row-chunk) pairs; partial tiles are reduced
// into the fp32 gradient with red.global.add.f32.
#include "tc_common.cuh"

namespace stg {
namespace {

using namespace tc;

constexpr int RK = 64;                 // time rows per K chunk
constexpr int BOX_BYTES = RK * 128;    // one (64 ch x 64 rows) box
constexpr int MAX_STAGES = 6;

struct WgTcP {
  int phases, t_out, c_in, c_out, k, stride;
  int cin_g, cout_g, ci_span, x_tiles;  // groups: input channels per group, ..., input-channel span of one 128-row co tile
  int n_taps, tap_groups, bnw, nbox_b, stages, tmem_cols;
  int chunks_per_sample, total_chunks, chunks_per_split;
  int tap_off[STG_MAX_TAPS];
  float* dw;
};

__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const WgTcP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int a_bytes = 2 * BOX_BYTES;
  const int b_bytes = p.nbox_b * BOX_BYTES;  // per tap
  const int stage_bytes = a_bytes + p.n_taps * b_bytes;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * MAX_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co_tile = blockIdx.y / p.tap_groups, tg = blockIdx.y - co_tile * p.tap_groups;
  const int co0 = co_tile * 128;
  // grouped convolution: the 128 output channels of this tile only meet the input channels of their own groups,
  // a span of ci_span channels starting at the first group's base; the off-diagonal part of the tile is computed
  // (the MMA is dense) and dropped in the epilogue
  const int ci0 = (co0 / p.cout_g) * p.cin_g + blockIdx.x * p.bnw;
  const int tap0 = tg * p.n_taps;
  const int ntap = min(p.n_taps, p.k - tap0);
  const int q0 = blockIdx.z * p.chunks_per_split, q1 = min(p.total_chunks, q0 + p.chunks_per_split);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmY);
    prefetch_tmap(&tmX);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)(a_bytes + ntap * b_bytes);
      for (int q = q0; q < q1; ++q) {
        const int it = q - q0, s = it % p.stages, phs = (it / p.stages) & 1;
        const int n = q / p.chunks_per_sample, rc = q - n * p.chunks_per_sample;
        const int b = n / p.phases, ph = n - b * p.phases;
        const int r0 = rc * RK;
        mbar_wait(empty_bar(s), phs ^ 1);
        mbar_expect_tx(full_bar(s), tx_bytes);
        const uint32_t a_dst = smem_base + s * stage_bytes;
        tma_load_4d(a_dst, &tmY, full_bar(s), co0, r0, ph, b);
        tma_load_4d(a_dst + BOX_BYTES, &tmY, full_bar(s), co0 + 64, r0, ph, b);
        for (int tl = 0; tl < ntap; ++tl)
          for (int bx = 0; bx < p.nbox_b; ++bx)
            tma_load_4d(a_dst + a_bytes + tl * b_bytes + bx * BOX_BYTES, &tmX, full_bar(s), ci0 + bx * 64,
                        r0 * p.stride + p.tap_off[tap0 + tl], ph, b);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_bf16_f32(128, p.bnw, 1, 1);
    for (int q = q0; q < q1; ++q) {
      const int it = q - q0, s = it % p.stages, phs = (it / p.stages) & 1;
      mbar_wait(full_bar(s), phs);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_base + s * stage_bytes;
        const uint64_t adesc = smem_desc_mnmajor_sw128(a_addr, BOX_BYTES, 1024);
        for (int tl = 0; tl < ntap; ++tl) {
          const uint64_t bdesc = smem_desc_mnmajor_sw128(a_addr + a_bytes + tl * b_bytes, BOX_BYTES, 1024);
#pragma unroll
          for (int ks = 0; ks < RK / 16; ++ks)  // 16 rows = 2048 B further along K
            umma_bf16(tmem_base + (uint32_t)(tl * p.bnw), adesc + (uint64_t)(ks * 128), bdesc + (uint64_t)(ks * 128), idesc,
                      (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        if (q == q1 - 1) umma_commit(tmem_full_bar);
      }
      __syncwarp();
    }
  } else {
    const int sub = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int co = co0 + sub * 32 + lane;
    const int KK = p.k * p.cin_g;
    const int cg_lo = (co < p.c_out ? co / p.cout_g : 0) * p.cin_g;  // input channels [cg_lo, cg_lo + cin_g) belong to co's group
    for (int tl = 0; tl < ntap; ++tl) {
      for (int c = 0; c < p.bnw; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(tl * p.bnw + c), v);
        if (co < p.c_out) {
          float* dst = p.dw + (int64_t)co * KK + (int64_t)(tap0 + tl) * p.cin_g + (ci0 + c - cg_lo);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = ci0 + c + i;
            if (ci >= cg_lo && ci < cg_lo + p.cin_g) atomicAdd(dst + i, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace

bool wgrad_tc_supported(const StgWgrad* d) {
  if (d->dtype != STG_BF16) return false;
  if (d->k < 1 || d->k > STG_MAX_TAPS) return false;
  if ((d->c_in % 8) || (d->c_out % 8)) return false;
  if (d->stride > 4) return false;
  if (d->c_out < 32) return false;  // 1- and 8-channel outputs: CUDA-core engine
  if (d->groups > 1) {
    const int cout_g = d->c_out / d->groups;
    if (cout_g < 128 ? (128 % cout_g) : (cout_g % 128)) return false;  // co tiles aligned to group boundaries
  }
  return true;
}

int wgrad_tc(const StgWgrad* d, cudaStream_t s) {
  if (!wgrad_tc_supported(d)) return STG_EUNSUPPORTED;
  WgTcP p;
  p.phases = d->phases; p.t_out = d->t_out; p.c_in = d->c_in; p.c_out = d->c_out; p.k = d->k; p.stride = d->stride;
  p.cin_g = d->c_in / d->groups; p.cout_g = d->c_out / d->groups;
  p.ci_span = d->groups == 1 ? d->c_in : (p.cout_g >= 128 ? p.cin_g : (128 / p.cout_g) * p.cin_g);
  p.bnw = p.ci_span > 64 ? 128 : 64;
  p.x_tiles = ceil_div(p.ci_span, p.bnw);
  p.nbox_b = p.bnw / 64;
  int taps = 512 / p.bnw;                       // TMEM columns
  const int smem_budget = 200 * 1024;
  while (taps > 1 && (2 * BOX_BYTES + taps * p.nbox_b * BOX_BYTES) * 2 > smem_budget) --taps;  // >= 2 stages
  if (taps > d->k) taps = d->k;
  p.n_taps = taps;
  p.tap_groups = ceil_div(d->k, taps);
  p.tmem_cols = 32;
  while (p.tmem_cols < taps * p.bnw) p.tmem_cols *= 2;
  for (int j = 0; j < d->k; ++j) p.tap_off[j] = j * d->dilation - d->pad;
  const int stage_bytes = 2 * BOX_BYTES + taps * p.nbox_b * BOX_BYTES;
  int stages = smem_budget / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.chunks_per_sample = ceil_div(d->t_out, RK);
  p.total_chunks = d->n_samples * d->phases * p.chunks_per_sample;
  const int gx = p.x_tiles, gy = ceil_div(d->c_out, 128) * p.tap_groups;
  int want = ceil_div(148, gx * gy);
  if (want < 1) want = 1;
  if (want > p.total_chunks) want = p.total_chunks;
  p.chunks_per_split = ceil_div(p.total_chunks, want);
  const int nsplit = ceil_div(p.total_chunks, p.chunks_per_split);
  if (stages > p.chunks_per_split) stages = p.chunks_per_split;
  if (stages < 1) stages = 1;
  p.stages = stages;
  p.dw = d->dw;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 8 * (2 * MAX_STAGES + 2);

  CUtensorMap tmY, tmX;
  {
    const uint64_t C = d->c_out, P = d->phases, T = d->t_out, B = d->n_samples;
    const uint64_t dims[4] = {C, T, P, B};
    const uint64_t strides[3] = {P * C * 2, C * 2, T * P * C * 2};
    const uint32_t box[4] = {64, RK, 1, 1};
    int r = make_tmap_bf16(&tmY, d->dy, 4, dims, strides, box, nullptr);
    if (r) return r;
  }
  {
    const uint64_t C = d->c_in, P = d->phases, T = d->t_in, B = d->n_samples;
    const uint64_t dims[4] = {C, T, P, B};
    const uint64_t strides[3] = {P * C * 2, C * 2, T * P * C * 2};
    const uint32_t box[4] = {64, (uint32_t)(RK * d->stride), 1, 1};
    const uint32_t es[4] = {1, (uint32_t)d->stride, 1, 1};
    int r = make_tmap_bf16(&tmX, d->x, 4, dims, strides, box, es);
    if (r) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    STG_CUDA_CHECK(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  dim3 grid(gx, gy, nsplit);
  wgrad_tc_kernel<<<grid, 192, smem, s>>>(tmY, tmX, p);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

}  // namespace stg
