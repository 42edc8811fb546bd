// Hardware probe (debug entry point, not on the product path): does a K-major SWIZZLE_128B UMMA operand accept a
// start address that is NOT a multiple of the 8-row swizzle pattern (i.e. a row-shifted view of a larger tile)?
// One CTA: TMA-loads a [rows_a x 64] bf16 tile and a [64 x 64] weight tile, issues one 128x64x64 MMA whose A
// descriptor starts `shift` rows into the tile, with descriptor bits [49,52) ("base offset") set to `base_off`,
// and writes the fp32 accumulator [128][64] out.  Host compares with X[shift:shift+128] @ W^T.
#include "tc_common.cuh"

namespace stg {
namespace {
using namespace tc;

__global__ void __launch_bounds__(128, 1)
rowshift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int rows_a, int shift,
                      int base_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, w_addr = base + 32768, bar = base + 32768 + 8192, done = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)(rows_a * 128 + 64 * 128));
    tma_load_3d(a_addr, &tmA, bar, 0, 0, 0);
    tma_load_3d(w_addr, &tmW, bar, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16_f32(128, 64, 0, 0);
    uint64_t adesc = smem_desc_kmajor_sw128(a_addr + (uint32_t)shift * 128u) | ((uint64_t)(base_off & 7) << 49);
    const uint64_t bdesc = smem_desc_kmajor_sw128(w_addr);
    for (int ks = 0; ks < 4; ++ks) umma_bf16(tmem, adesc + 2 * ks, bdesc + 2 * ks, idesc, ks > 0 ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  for (int c = 0; c < 64; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}
// Second probe: the per-group MMAs of a grouped convolution.  A = [128][64] bf16 (SWIZZLE_128B, 64 / cin_g groups side by
// side on K), W = compact [n_g * cout_g][cin_g] bf16 loaded with the swizzle that matches its row length (32 / 64 / 128 B).
// Group q: D[:, q*cout_g .. +cout_g] = A[:, q*cin_g .. +cin_g] * W[q*cout_g .. +cout_g][:]^T as cin_g/16 MMAs of N = cout_g
// whose A descriptor starts q*cin_g*2 bytes into the swizzle atom, whose B descriptor starts q*cout_g rows into the narrow tile
// and whose accumulator starts q*cout_g TMEM columns in.
__global__ void __launch_bounds__(128, 1)
group_mma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int cin_g, int cout_g,
                       float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_addr = base, w_addr = base + 16384, bar = base + 16384 + 32768, done = bar + 8, slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_g = 64 / cin_g, row_bytes = cin_g * 2, n_cols = n_g * cout_g;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 256); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)(128 * 128 + n_cols * row_bytes));
    tma_load_3d(a_addr, &tmA, bar, 0, 0, 0);
    tma_load_3d(w_addr, &tmW, bar, 0, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = idesc_bf16_f32(128, cout_g, 0, 0);
    const uint64_t adesc = smem_desc_kmajor_sw128(a_addr);
    for (int q = 0; q < n_g; ++q) {
      const uint64_t bdesc = smem_desc_kmajor_narrow(w_addr + (uint32_t)(q * cout_g * row_bytes), row_bytes);
      for (int ks = 0; ks < cin_g / 16; ++ks)
        umma_bf16(tmem + (uint32_t)(q * cout_g), adesc + 2 * (q * (cin_g / 16) + ks), bdesc + 2 * ks, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  for (int c = 0; c < n_cols; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * n_cols + c + i] = v[i];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}
// Third probe: what one SM can ingest through TMA when `gridDim.x` SMs pull [box_rows x 64] bf16 boxes (128-byte rows,
// SWIZZLE_128B) at once - from disjoint rows (mode 0: activation-like) or all from the same rows (mode 1: weight-like).
// One producer thread and one consumer thread per CTA around a `stages`-deep ring; nothing else runs.  clk[cta] = clocks
// for n_iters boxes.
__global__ void __launch_bounds__(64, 1)
tma_bw_probe_kernel(const __grid_constant__ CUtensorMap tm, long long n_rows, int box_rows, int n_iters, int mode, int stages,
                    long long* clk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t box_bytes = (uint32_t)box_rows * 128u;
  const uint32_t bars = base + (uint32_t)stages * box_bytes;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (16 + s); };
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  const long long n_boxes = n_rows / box_rows;
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int s = 0; uint32_t ph = 0;
    const int wrap = (int)(n_boxes * box_rows);
    int row = mode == 1 ? 0 : (int)((((long long)blockIdx.x * n_iters) % n_boxes) * box_rows);   // no division inside the loop
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(empty(s), ph ^ 1);
      mbar_expect_tx(full(s), box_bytes);
      tma_load_3d(base + (uint32_t)s * box_bytes, &tm, full(s), 0, row, 0);
      row += box_rows; if (row >= wrap) row = 0;
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
    // all boxes landed: wait until the consumer has released the last one
    const int ls = (n_iters - 1) % stages; const uint32_t lph = (uint32_t)(((n_iters - 1) / stages) & 1);
    mbar_wait(empty(ls), lph);
    clk[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 32) {
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < n_iters; ++it) {
      mbar_wait(full(s), ph);
      mbar_arrive(empty(s));
      if (++s == stages) { s = 0; ph ^= 1u; }
    }
  }
}
}  // namespace
}  // namespace stg

using namespace stg;
/* buf: bf16 [n_rows][64]; clk: int64 [grid] */
extern "C" int stg_debug_tma_bw(const void* buf, long long n_rows, int box_rows, int n_iters, int mode, int stages, int grid,
                                long long* clk, stg_stream_t stream) {
  if (box_rows < 8 || box_rows > 256 || stages < 1 || stages > 16 || stages * box_rows * 128 > 200 * 1024 || grid < 1) return STG_EINVAL;
  CUtensorMap tm;
  { const uint64_t dims[3] = {64, (uint64_t)n_rows, 1}; const uint64_t st[2] = {128, (uint64_t)n_rows * 128};
    const uint32_t box[3] = {64, (uint32_t)box_rows, 1}; int r = make_tmap_bf16(&tm, buf, 3, dims, st, box, nullptr); if (r) return r; }
  STG_CUDA_CHECK(cudaFuncSetAttribute(tma_bw_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const size_t smem = (size_t)stages * box_rows * 128 + 1024 + 512;
  tma_bw_probe_kernel<<<grid, 64, smem, static_cast<cudaStream_t>(stream)>>>(tm, n_rows, box_rows, n_iters, mode, stages, clk);
  STG_LAUNCH_CHECK();
  return STG_OK;
}
/* x: bf16 [128][64], w: bf16 [(64/cin_g)*cout_g][cin_g], out: float [128][(64/cin_g)*cout_g]; cin_g in {16,32,64}, cout_g % 16 == 0 */
extern "C" int stg_debug_group_mma(const void* x, const void* w, int cin_g, int cout_g, float* out, stg_stream_t stream) {
  if ((cin_g != 16 && cin_g != 32 && cin_g != 64) || cout_g % 16 || cout_g < 16 || (64 / cin_g) * cout_g > 256) return STG_EINVAL;
  const int n_cols = (64 / cin_g) * cout_g;
  CUtensorMap tmA, tmW;
  { const uint64_t dims[3] = {64, 128, 1}; const uint64_t st[2] = {128, 128 * 128};
    const uint32_t box[3] = {64, 128, 1}; int r = make_tmap_bf16(&tmA, x, 3, dims, st, box, nullptr); if (r) return r; }
  { const uint64_t dims[3] = {(uint64_t)cin_g, (uint64_t)n_cols, 1}; const uint64_t st[2] = {(uint64_t)cin_g * 2, (uint64_t)n_cols * cin_g * 2};
    const uint32_t box[3] = {(uint32_t)cin_g, (uint32_t)n_cols, 1};
    int r = make_tmap_bf16(&tmW, w, 3, dims, st, box, nullptr, cin_g * 2); if (r) return r; }
  STG_CUDA_CHECK(cudaFuncSetAttribute(group_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  group_mma_probe_kernel<<<1, 128, 56 * 1024, static_cast<cudaStream_t>(stream)>>>(tmA, tmW, cin_g, cout_g, out);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

/* x: bf16 [rows_a][64] (rows_a <= 256), w: bf16 [64][64], out: float [128][64] */
extern "C" int stg_debug_rowshift(const void* x, const void* w, int rows_a, int shift, int base_off, float* out,
                                  stg_stream_t stream) {
  if (rows_a > 256 || shift + 128 > rows_a) return STG_EINVAL;
  CUtensorMap tmA, tmW;
  { const uint64_t dims[3] = {64, (uint64_t)rows_a, 1}; const uint64_t st[2] = {128, (uint64_t)rows_a * 128};
    const uint32_t box[3] = {64, (uint32_t)rows_a, 1}; int r = make_tmap_bf16(&tmA, x, 3, dims, st, box, nullptr); if (r) return r; }
  { const uint64_t dims[3] = {64, 64, 1}; const uint64_t st[2] = {128, 64 * 128};
    const uint32_t box[3] = {64, 64, 1}; int r = make_tmap_bf16(&tmW, w, 3, dims, st, box, nullptr); if (r) return r; }
  STG_CUDA_CHECK(cudaFuncSetAttribute(rowshift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  rowshift_probe_kernel<<<1, 128, 48 * 1024, static_cast<cudaStream_t>(stream)>>>(tmA, tmW, rows_a, shift, base_off, out);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

// Host-only: the row classes the tcgen05 convolution uses for a sample of t_dst rows (conv_tc.cu: row_classes).
namespace stg { int debug_row_classes(int t_dst, int* out); }
extern "C" int stg_debug_row_classes(int t_dst, int* out) {
  if (t_dst < 1 || !out) return STG_EINVAL;
  return stg::debug_row_classes(t_dst, out);
}
