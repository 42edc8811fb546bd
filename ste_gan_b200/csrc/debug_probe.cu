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
}  // namespace
}  // namespace stg

using namespace stg;
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
