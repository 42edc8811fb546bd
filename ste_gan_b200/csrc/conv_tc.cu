// tcgen05 implicit-GEMM convolution (forward and data-gradient), bf16 in / fp32 accumulate.
//
// GEMM view (channels-last activations):   D[t][c_dst] = sum_taps sum_{c_src} A_tap[t][c_src] * W_tap[c_dst][c_src]
//   M = 128 time rows of one (virtual) sample   -> TMEM lanes
//   N = BN <= 256 output channels               -> TMEM columns
//   K = 64-channel chunks, one per (tap, chunk) -> one pipeline stage each
// A tiles come straight from the activation tensor by TMA: the row coordinate is
// r0*stride + tap_offset, may be negative or run past the sample and is zero-filled by
// the TMA unit, which implements zero padding, dilation and per-sample boundaries with
// no im2col buffer.  Period views (DiscriminatorP) are a 4-D tensor map (C, rows, phase, B).
// W tiles come from the packed [k][c_dst][c_src/g] weights.  Both operands are K-major
// with 128-byte swizzle.
//
// PERSISTENT kernel: one CTA per SM walks the tile list (column tile fastest, so CTAs that run
// together share A rows in L2).  Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM
// alloc + single-thread tcgen05.mma issuer, warps 2..9 = epilogue.  The shared-memory ring runs
// continuously across tiles; the accumulator is DOUBLE-BUFFERED in TMEM (2 x BN columns), so
// the epilogue of tile i (tcgen05.ld -> bias / pair-sum / add_pre / activation mask / residual /
// activation / row duplication -> global stores) overlaps the MMAs of tile i+1.  The eight
// epilogue warps split the tile by TMEM sub-partition (warp % 4) and column half, and prefetch the
// epilogue operands of the next 16-column chunk while the current one is processed.
#include "tc_common.cuh"

namespace stg {

namespace {

using namespace tc;

constexpr int TM = 128;        // rows per CTA tile
constexpr int KC = 64;         // channels per K chunk (128 B of bf16)
constexpr int A_BYTES = TM * KC * 2;
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 32 * (2 + EPI_WARPS);
constexpr int ACC_COLS = 256;  // TMEM column distance between the two accumulator buffers

struct TcEpi {
  int phases, t_out, c_dst, post_shift, mask_mode, act, dup_rows, out_f32, pair_sum;
  const float* bias;
  const bf16 *add_pre, *mask, *add_post;
  void* y_raw;
  void* y_act;
};

constexpr int MAX_RES = 8;  // residues of a strided data-gradient (= stride)

struct TcP {
  int phases, t_dst, stride, k_chunks, bn, stages, a_boxes, tmem_cols;
  int cs_g, cd_g;          // source / destination channels per (packed) group
  int n_res, tiles_m;      // output-row residues (1 unless transposed && stride > 1), row tiles per residue
  int tiles_n, n_tiles;    // column tiles, total tiles = n_vs * n_res * tiles_m * tiles_n
  int res_first[MAX_RES + 1];  // taps of residue r: [res_first[r], res_first[r+1])
  int tap_off[STG_MAX_TAPS];   // source-row offset of the tap (rows of the A tile start at r0*stride + tap_off)
  int tap_w[STG_MAX_TAPS];     // tap index into the packed weights
  TcEpi e;
};

__device__ __forceinline__ void ld16(const bf16* p, float (&o)[16]) {
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *reinterpret_cast<const uint4*>(p + 8);
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    o[2 * i] = __low2float(h);
    o[2 * i + 1] = __high2float(h);
  }
}
__device__ __forceinline__ void st16(bf16* p, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  *reinterpret_cast<uint4*>(p + 8) = make_uint4(w[4], w[5], w[6], w[7]);
}
__device__ __forceinline__ void st16f(float* p, const float (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// epilogue operands of one output row, 16 consecutive channels starting at `col` (prefetched one chunk ahead)
struct EpiIn {
  float pre[16], mk[16], post[16];
};

__device__ __forceinline__ int64_t epi_off(const TcEpi& e, int b, int ph, int row, int col) {
  return ((int64_t)b * e.t_out + row) * ((int64_t)e.phases * e.c_dst) + (int64_t)ph * e.c_dst + col;
}

__device__ __forceinline__ void epi_load(const TcEpi& e, int b, int ph, int row, int col, bool row_ok, EpiIn& in) {
  if (!row_ok || col >= e.c_dst) return;
  const int ncols = min(16, e.c_dst - col);
  const bool vec = (ncols == 16) && ((e.c_dst & 7) == 0);
  const int64_t off = epi_off(e, b, ph, row, col);
  if (e.add_pre) {
    if (vec) ld16(e.add_pre + off, in.pre); else for (int i = 0; i < 16; ++i) in.pre[i] = i < ncols ? to_f(e.add_pre[off + i]) : 0.f;
  }
  if (e.mask) {
    if (vec) ld16(e.mask + off, in.mk); else for (int i = 0; i < 16; ++i) in.mk[i] = i < ncols ? to_f(e.mask[off + i]) : 0.f;
  }
  if (e.add_post) {
    const int t_post = e.t_out >> e.post_shift;
    const int64_t o2 = ((int64_t)b * t_post + (row >> e.post_shift)) * ((int64_t)e.phases * e.c_dst) + (int64_t)ph * e.c_dst + col;
    if (vec) ld16(e.add_post + o2, in.post); else for (int i = 0; i < 16; ++i) in.post[i] = i < ncols ? to_f(e.add_post[o2 + i]) : 0.f;
  }
}

// v: accumulator (+bias, pair-summed) of one output row, 16 channels
__device__ __forceinline__ void epi_store(const TcEpi& e, int b, int ph, int row, int col, float (&v)[16], const EpiIn& in) {
  const int64_t pitch = (int64_t)e.phases * e.c_dst;
  const int64_t off = epi_off(e, b, ph, row, col);
  const int ncols = min(16, e.c_dst - col);
  const bool vec = (ncols == 16) && ((e.c_dst & 7) == 0);
  if (e.add_pre) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += in.pre[i];
  }
  if (e.mask) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] *= act_grad_from_output(e.mask_mode, in.mk[i]);
  }
  if (e.add_post) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += in.post[i];
  }
  if (e.y_raw) {
    if (e.out_f32) {
      float* p = static_cast<float*>(e.y_raw) + off;
      if (vec) st16f(p, v); else for (int i = 0; i < ncols; ++i) p[i] = v[i];
    } else {
      bf16* p = static_cast<bf16*>(e.y_raw) + off;
      if (vec) st16(p, v); else for (int i = 0; i < ncols; ++i) p[i] = __float2bfloat16_rn(v[i]);
    }
  }
  if (e.y_act) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = act_apply(e.act, v[i]);
    const int64_t o0 = e.dup_rows ? ((int64_t)b * 2 * e.t_out + 2 * row) * pitch + (int64_t)ph * e.c_dst + col : off;
    if (e.out_f32) {
      float* p0 = static_cast<float*>(e.y_act) + o0;
      if (vec) st16f(p0, a); else for (int i = 0; i < ncols; ++i) p0[i] = a[i];
      if (e.dup_rows) { float* p1 = p0 + pitch; if (vec) st16f(p1, a); else for (int i = 0; i < ncols; ++i) p1[i] = a[i]; }
    } else {
      bf16* p0 = static_cast<bf16*>(e.y_act) + o0;
      if (vec) st16(p0, a); else for (int i = 0; i < ncols; ++i) p0[i] = __float2bfloat16_rn(a[i]);
      if (e.dup_rows) { bf16* p1 = p0 + pitch; if (vec) st16(p1, a); else for (int i = 0; i < ncols; ++i) p1[i] = __float2bfloat16_rn(a[i]); }
    }
  }
}

struct Tile {
  int b, ph, res, r0, col0, ch0, tap0, n_iters;
};
__device__ __forceinline__ Tile decode_tile(const TcP& p, int t) {
  Tile x;
  const int tn = t % p.tiles_n; int u = t / p.tiles_n;
  const int tm = u % p.tiles_m; u /= p.tiles_m;
  x.res = u % p.n_res; const int n = u / p.n_res;
  x.b = n / p.phases; x.ph = n - x.b * p.phases;
  // strided data-gradient: output rows r = r' * n_res + res are produced per residue class `res` from the
  // taps j with (res + pad - j*dilation) % stride == 0, as a stride-1 correlation over r'
  x.r0 = tm * TM;
  x.col0 = tn * p.bn;
  x.ch0 = (x.col0 / p.cd_g) * p.cs_g;  // first source channel of this column tile's group
  x.tap0 = p.res_first[x.res];
  x.n_iters = (p.res_first[x.res + 1] - x.tap0) * p.k_chunks;
  return x;
}

__global__ void __launch_bounds__(NTHREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const TcP p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for SWIZZLE_128B tiles
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int b_bytes = p.bn * KC * 2;
  const int stage_bytes = A_BYTES + b_bytes;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;  // 8-byte aligned (multiple of 1024)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * MAX_STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmW);
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tmem_full_bar(a), 1); mbar_init(tmem_empty_bar(a), EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int rows_per_box = TM / p.a_boxes;
      int itg = 0;  // stage counter, continuous across tiles
      for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const Tile x = decode_tile(p, t);
        for (int it = 0; it < x.n_iters; ++it, ++itg) {
          const int s = itg % p.stages, phs = (itg / p.stages) & 1;
          const int tl = it / p.k_chunks, chunk = it - tl * p.k_chunks, tap = x.tap0 + tl;
          mbar_wait(empty_bar(s), phs ^ 1);
          mbar_expect_tx(full_bar(s), (uint32_t)stage_bytes);
          const uint32_t a_dst = smem_base + s * stage_bytes;
          for (int bx = 0; bx < p.a_boxes; ++bx)
            tma_load_4d(a_dst + bx * rows_per_box * KC * 2, &tmA, full_bar(s), x.ch0 + chunk * KC,
                        (x.r0 + bx * rows_per_box) * p.stride + p.tap_off[tap], x.ph, x.b);
          tma_load_3d(a_dst + A_BYTES, &tmW, full_bar(s), chunk * KC, x.col0, p.tap_w[tap]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = idesc_bf16_f32(TM, p.bn, 0, 0);
    int itg = 0, acc_i = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const Tile x = decode_tile(p, t);
      if (x.n_iters == 0) continue;
      const int as = acc_i & 1, aph = (acc_i >> 1) & 1;
      mbar_wait(tmem_empty_bar(as), aph ^ 1);  // epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_COLS);
      for (int it = 0; it < x.n_iters; ++it, ++itg) {
        const int s = itg % p.stages, phs = (itg / p.stages) & 1;
        mbar_wait(full_bar(s), phs);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_base + s * stage_bytes;
          const uint64_t adesc = smem_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = smem_desc_kmajor_sw128(a_addr + A_BYTES);
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks)  // +32 bytes (16 bf16) along K inside the swizzle atom
            umma_bf16(d_tmem, adesc + 2 * ks, bdesc + 2 * ks, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          umma_commit(empty_bar(s));
          if (it == x.n_iters - 1) umma_commit(tmem_full_bar(as));
        }
        __syncwarp();
      }
      ++acc_i;
    }
  } else {
    // ===== epilogue =====
    const int ew = warp - 2;
    const int sub = warp & 3;        // TMEM sub-partition this warp may read
    const int half = ew >> 2;        // column half of the tile
    const TcEpi& e = p.e;
    // columns of this warp: chunks of 16, split between the two warps of a sub-partition
    const int n_chunks = p.bn / 16;
    const int c_lo = (half == 0) ? 0 : (n_chunks + 1) / 2, c_hi = (half == 0) ? (n_chunks + 1) / 2 : n_chunks;
    int acc_i = 0;
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      const Tile x = decode_tile(p, t);
      const int as = acc_i & 1, aph = (acc_i >> 1) & 1;
      const int arow = (x.r0 + sub * 32 + lane) * p.n_res + x.res;  // output row of this thread (before pair_sum)
      const bool row_ok = arow < p.t_dst && (!e.pair_sum || (lane & 1) == 0);
      const int orow = e.pair_sum ? (arow >> 1) : arow;
      EpiIn inA, inB;
      if (c_lo < c_hi) epi_load(e, x.b, x.ph, orow, x.col0 + c_lo * 16, row_ok, inA);
      if (x.n_iters > 0) {
        mbar_wait(tmem_full_bar(as), aph);
        tc_fence_after();
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(as * ACC_COLS);
      auto process = [&](int c, const EpiIn& in) {
        float v[16];
        if (x.n_iters > 0) {
          tmem_ld16(t_addr + (uint32_t)(c * 16), v);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
        }
        const int col = x.col0 + c * 16;
        if (col >= e.c_dst) return;  // warp-uniform
        if (e.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += (col + i < e.c_dst) ? e.bias[col + i] : 0.f;
        }
        if (e.pair_sum) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 1);
        }
        if (row_ok) epi_store(e, x.b, x.ph, orow, col, v, in);
      };
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 2) {
        if (c + 1 < c_hi) epi_load(e, x.b, x.ph, orow, x.col0 + (c + 1) * 16, row_ok, inB);
        process(c, inA);
        if (c + 1 < c_hi) {
          if (c + 2 < c_hi) epi_load(e, x.b, x.ph, orow, x.col0 + (c + 2) * 16, row_ok, inA);
          process(c + 1, inB);
        }
      }
      if (x.n_iters > 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(as));
        ++acc_i;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_cuda_error(cudaErrorNotSupported, "cuTensorMapEncodeTiled entry point"); return STG_ECUDA; }
  cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_cuda_error(cudaErrorInvalidValue, "cuTensorMapEncodeTiled failed");
    return STG_ECUDA;
  }
  return STG_OK;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1)
      n = 148;
  }
  return n;
}

static int pick_bn(int cd_g, int groups) {
  if (groups > 1) {  // a column tile must not straddle groups
    if (cd_g % 256 == 0) return 256;
    if (cd_g % 128 == 0) return 128;
    if (cd_g <= 256 && cd_g % 16 == 0) return cd_g;
    return 0;
  }
  if (cd_g <= 16) return 16;
  if (cd_g <= 256 && (cd_g % 16) == 0) return cd_g;
  if (cd_g % 256 == 0) return 256;
  if (cd_g % 192 == 0) return 192;
  if (cd_g % 128 == 0) return 128;
  if (cd_g % 64 == 0) return 64;
  return 128;
}

bool conv_tc_supported(const StgConv* d) {
  if (d->dtype != STG_BF16) return false;
  if (d->k > STG_MAX_TAPS || d->k < 1) return false;
  if ((d->c_src % 8) != 0) return false;                 // 16-byte global strides for TMA
  if (d->groups > 1) {
    if ((d->c_src / d->groups) % KC) return false;       // whole K chunks per group (see stg_tc_pack_groups)
    if (pick_bn(d->c_dst / d->groups, d->groups) == 0) return false;
  }
  if (d->transposed && d->stride > MAX_RES) return false;
  if (d->transposed && d->stride > 1 && d->pair_sum) return false;
  if (!d->transposed && d->stride > 4) return false;     // A box rows = 64*stride <= 256
  if (d->pair_sum && (d->t_dst & 1)) return false;
  if (d->add_pre == nullptr && d->mask == nullptr && d->add_post == nullptr && d->y_raw == nullptr && d->y_act == nullptr)
    return false;
  return true;
}

int tc_pack_groups(int c_in, int c_out, int groups) {
  if (groups <= 1) return 1;
  const int cin_g = c_in / groups, cout_g = c_out / groups;
  // smallest merge factor f | groups with (cin_g * f) % 64 == 0 and (cout_g * f) % 64 == 0 (dgrad K chunks)
  for (int f = 1; f <= groups; ++f) {
    if (groups % f) continue;
    if ((cin_g * f) % KC == 0 && (cout_g * f) % KC == 0) return groups / f;
  }
  return groups;
}

int conv_tc(const StgConv* d, cudaStream_t s) {
  if (!conv_tc_supported(d)) return STG_EUNSUPPORTED;
  TcP p;
  p.phases = d->phases; p.t_dst = d->t_dst; p.stride = d->transposed ? 1 : d->stride;
  p.cs_g = d->c_src / d->groups; p.cd_g = d->c_dst / d->groups;
  p.k_chunks = ceil_div(p.cs_g, KC);
  p.bn = pick_bn(p.cd_g, d->groups);
  p.a_boxes = (p.stride * TM <= 256) ? 1 : 2;
  p.tmem_cols = p.bn <= 128 ? 256 : 512;  // two accumulator buffers ACC_COLS apart (bn <= 32: second one still at +256)
  p.tmem_cols = 512;
  int max_taps = 0;
  if (d->transposed && d->stride > 1) {
    p.n_res = d->stride;
    int n = 0;
    for (int r = 0; r < p.n_res; ++r) {
      p.res_first[r] = n;
      for (int j = 0; j < d->k; ++j) {
        const int num = r + d->pad - j * d->dilation;
        if (((num % d->stride) + d->stride) % d->stride != 0) continue;
        p.tap_off[n] = (num >= 0) ? num / d->stride : -((-num) / d->stride);  // exact division
        p.tap_w[n] = j;
        ++n;
      }
      if (n - p.res_first[r] > max_taps) max_taps = n - p.res_first[r];
    }
    p.res_first[p.n_res] = n;
    p.tiles_m = ceil_div(ceil_div(d->t_dst, p.n_res), TM);
  } else {
    p.n_res = 1;
    for (int j = 0; j < d->k; ++j) {
      p.tap_off[j] = d->transposed ? (d->pad - j * d->dilation) : (j * d->dilation - d->pad);
      p.tap_w[j] = j;
    }
    p.res_first[0] = 0; p.res_first[1] = d->k;
    max_taps = d->k;
    p.tiles_m = ceil_div(d->t_dst, TM);
  }
  const int stage_bytes = A_BYTES + p.bn * KC * 2;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 1) stages = 1;
  p.stages = stages;
  (void)max_taps;
  const size_t smem = (size_t)stages * stage_bytes + 1024 /*align slack*/ + 8 * (2 * MAX_STAGES + 6);

  TcEpi& e = p.e;
  e.phases = d->phases; e.t_out = d->pair_sum ? d->t_dst / 2 : d->t_dst; e.c_dst = d->c_dst;
  e.post_shift = d->post_shift; e.mask_mode = d->mask_mode; e.act = d->act; e.dup_rows = d->dup_rows;
  e.out_f32 = d->out_f32; e.pair_sum = d->pair_sum; e.bias = d->bias;
  e.add_pre = static_cast<const bf16*>(d->add_pre); e.mask = static_cast<const bf16*>(d->mask);
  e.add_post = static_cast<const bf16*>(d->add_post); e.y_raw = d->y_raw; e.y_act = d->y_act;

  CUtensorMap tmA, tmW;
  {
    const uint64_t C = d->c_src, P = d->phases, T = d->t_src, B = d->n_samples;
    const uint64_t dims[4] = {C, T, P, B};
    const uint64_t strides[3] = {P * C * 2, C * 2, T * P * C * 2};
    const uint32_t box[4] = {(uint32_t)KC, (uint32_t)((TM / p.a_boxes) * p.stride), 1, 1};
    const uint32_t es[4] = {1, (uint32_t)p.stride, 1, 1};
    int r = make_tmap_bf16(&tmA, d->src, 4, dims, strides, box, es);
    if (r) return r;
  }
  {
    const uint64_t C = p.cs_g, N = d->c_dst, K = d->k;
    const uint64_t dims[3] = {C, N, K};
    const uint64_t strides[2] = {C * 2, N * C * 2};
    const uint32_t box[3] = {(uint32_t)KC, (uint32_t)p.bn, 1};
    int r = make_tmap_bf16(&tmW, d->w, 3, dims, strides, box, nullptr);
    if (r) return r;
  }
  static bool attr_set = false;
  if (!attr_set) {
    STG_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  p.tiles_n = ceil_div(d->c_dst, p.bn);
  const int64_t n_tiles = (int64_t)d->n_samples * d->phases * p.n_res * p.tiles_m * p.tiles_n;
  if (n_tiles > 0x7fffffff) return STG_EINVAL;
  p.n_tiles = (int)n_tiles;
  const int grid = p.n_tiles < sm_count() ? p.n_tiles : sm_count();
  conv_tc_kernel<<<grid, NTHREADS, smem, s>>>(tmA, tmW, p);
  STG_LAUNCH_CHECK();
  return STG_OK;
}

}  // namespace stg
