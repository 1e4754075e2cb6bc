// Sparse convolution filter gradient on the tensor cores (BF16x3-class precision, no atomics).
//
//   dW[k][ci][co] = sum over output rows o of  features[nbr[k][o]][ci] * out_grad[o][co]      (spconv_ops.h:363-456,
//   the `torch::mm_out(filterGradSub, inputBuffer.t(), outputBuffer)` per kernel offset of indiceConvBackward)
//
// The reduction runs over ROWS, so both operands are "MN-major" for the tensor core: a row of the operand tile is one
// reduction index and holds the M (or N) elements contiguously -- which is exactly how the split images of
// spconv_v6.cu store a row (per 16 channels: 16 bf16 hi | 16 bf16 lo).  The kernel therefore
//   * gathers the 128-byte channel slabs of 64 feature rows (through nbr, missing rows as zeros) with cp.async into
//     SWIZZLE_128B tiles, exactly like the forward kernel, and loads the matching 64 out-grad rows with TMA tile
//     copies (they are contiguous);
//   * treats the hi and lo halves of a row as SEPARATE M (N) elements: one M = 128 x N = 2 Cout x K = 16 MMA with
//     MN-major descriptors produces hi.hi, hi.lo, lo.hi and lo.lo partial products of 64 input channels at once, in
//     separate accumulator rows / columns; the epilogue adds the four (two columns in registers, two TMEM lanes
//     through one shuffle);
//   * gives every CTA work items (kernel offset k, chunk of rows): the accumulator (2 Cin x 2 Cout fp32 = up to all
//     512 TMEM columns) lives in TMEM for the whole item and is written once as a partial dW; a second kernel adds the
//     partials of the chunks in a fixed order (bit-reproducible, like the SIMT path it replaces).
// Warp roles (10 warps, one CTA per SM): 0-7 feature-row gather + epilogue, 8 MMA issue + TMEM allocation,
// 9 out-grad loader (TMA).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace bevb200 {

int spconv_v6_split_rows(const float *features, int n_cap, const int32_t *n_dev, int c_in, void *split,
                         cudaStream_t st);

constexpr int kWgtThreads = 10 * 32;
constexpr int kWgtRows = 64;                       // reduction rows per stage
constexpr int kWgtSlab = kWgtRows * 128;           // 64 rows x 128 B (32 channels, hi | lo)
constexpr int kWgtMaxStages = 8;

struct WgtParams {
  const uint8_t *fsplit;       // [n_in][c_in * 4 B]
  const int32_t *nbr;          // [kvol][n_out]
  float *partial;              // [n_chunks][kvol][c_in][c_out]
  int n_in, n_out, c_in, c_out, kvol;
  int n_chunks, tiles_per_chunk, n_tiles;
  int n_a, n_b;                // 128-byte slabs per row: c_in / 32, c_out / 32
  int stages, stage_bytes;
  int m_blocks;                // accumulator row blocks of 64 input channels
  int tmem_cols;
};

// UMMA shared-memory descriptor, MN-major, SWIZZLE_128B (cute/atom/mma_traits_sm100.hpp, canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): a reduction row is 128 B, eight rows form a 1024-byte swizzle
// atom (SBO = distance between 8-row groups), LBO = distance between 64-element atoms along M / N.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void wgt_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void wgt_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tma_tile_2d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(kWgtThreads, 1)
spconv_wgrad_tc_kernel(const WgtParams p, const __grid_constant__ CUtensorMap gmap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[2 * kWgtMaxStages + 2];
  __shared__ uint32_t tmem_base_s;
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[kWgtMaxStages]);
  const uint32_t acc_full = smem_u32(&bars[2 * kWgtMaxStages]), acc_empty = smem_u32(&bars[2 * kWgtMaxStages + 1]);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, n_a = p.n_a, n_b = p.n_b;
  const uint32_t stage_bytes = (uint32_t)p.stage_bytes;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full + 8 * s, 8 + 1);      // eight gather warps + the out-grad loader's expect_tx arrive
      mbar_init(empty + 8 * s, 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  const int n_items = p.n_chunks * p.kvol;
  const int n2 = 2 * p.c_out;            // accumulator columns per row block

  if (warp < 8) {
    // =============================== feature-row gather + epilogue ==========================
    const int m = lane >> 3, c = lane & 7;
    const uint32_t row_bytes = (uint32_t)p.c_in * 4u;
    const unsigned long long fbase = reinterpret_cast<unsigned long long>(p.fsplit);
    // this lane's two copies per slab: rows 8 w + m and 8 w + 4 + m; it writes physical chunk c and fetches the
    // logical chunk c ^ (row & 7) (lane-order stores, as in spconv_v6.cu)
    const int r0 = 8 * warp + m, r1 = r0 + 4;
    const uint32_t dst0 = (uint32_t)(r0 * 128 + (c << 4)), dst1 = (uint32_t)(r1 * 128 + (c << 4));
    const uint32_t sc0 = (uint32_t)((c ^ (r0 & 7)) << 4), sc1 = (uint32_t)((c ^ (r1 & 7)) << 4);
    int s = 0, pend = -1;
    uint32_t ph = 0, acc_ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / p.kvol, k = item - chunk * p.kvol;
      const int t_begin = chunk * p.tiles_per_chunk, t_end = min(p.n_tiles, t_begin + p.tiles_per_chunk);
      const int32_t *nk = p.nbr + (long long)k * p.n_out;
      auto load_idx = [&](int t, int &i0, int &i1) {
        const int o0 = t * kWgtRows + r0, o1 = o0 + 4;
        i0 = o0 < p.n_out ? __ldg(nk + o0) : -1;
        i1 = o1 < p.n_out ? __ldg(nk + o1) : -1;
      };
      int i0 = -1, i1 = -1;
      if (t_begin < t_end) load_idx(t_begin, i0, i1);
      for (int t = t_begin; t < t_end; ++t) {
        int j0 = i0, j1 = i1;
        if (j0 >= p.n_in) j0 = -1;
        if (j1 >= p.n_in) j1 = -1;
        if (t + 1 < t_end) load_idx(t + 1, i0, i1);
        mbar_wait(empty + 8 * s, ph ^ 1u);
        const uint32_t a_stage = ring + (uint32_t)s * stage_bytes;
        const unsigned long long src0 = fbase + (unsigned long long)(uint32_t)max(j0, 0) * row_bytes;
        const unsigned long long src1 = fbase + (unsigned long long)(uint32_t)max(j1, 0) * row_bytes;
        for (int sl = 0; sl < n_a; ++sl) {
          cp_async16_row(a_stage + (uint32_t)sl * kWgtSlab + dst0, src0 + (unsigned)(sl * 128) + sc0, j0);
          cp_async16_row(a_stage + (uint32_t)sl * kWgtSlab + dst1, src1 + (unsigned)(sl * 128) + sc1, j1);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (pend >= 0) {                 // the previous stage's copies have had a whole stage's time
          asm volatile("cp.async.wait_group 1;" ::: "memory");
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(full + 8 * pend);
        }
        pend = s;
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
      if (pend >= 0) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * pend);
        pend = -1;
      }
      // ------------------------------- epilogue of the item ---------------------------------
      mbar_wait(acc_full, acc_ph);
      acc_ph ^= 1u;
      tc_fence_after();
      const int q = warp & 3, half = warp >> 2;
      float *dst_item = p.partial + ((long long)chunk * p.kvol + k) * p.c_in * p.c_out;
      for (int mb = 0; mb < p.m_blocks; ++mb) {
        const int ci = mb * 64 + q * 16 + (lane & 15);
        for (int cg = half; cg < (p.c_out >> 4); cg += 2) {
          float v[32];
          wgt_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mb * n2 + cg * 32), v);
          float sum[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            sum[j] = v[j] + v[j + 16];                                   // hi.co + lo.co columns
            sum[j] += __shfl_down_sync(0xffffffffu, sum[j], 16);          // + the lo.ci accumulator row
          }
          if (lane < 16 && ci < p.c_in) {
            float4 *dst = reinterpret_cast<float4 *>(dst_item + (long long)ci * p.c_out + cg * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ==========================================
    const uint32_t idesc = umma_idesc_bf16(128, n2) | (1u << 15) | (1u << 16);      // A and B MN-major
    int s = 0;
    uint32_t ph = 0, pe = 0;
    bool first_item = true;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item / p.kvol;
      const int t_begin = chunk * p.tiles_per_chunk, t_end = min(p.n_tiles, t_begin + p.tiles_per_chunk);
      if (!first_item) {
        mbar_wait(acc_empty, pe);
        pe ^= 1u;
        tc_fence_after();
      }
      first_item = false;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(full + 8 * s, ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_stage = ring + (uint32_t)s * stage_bytes;
          const uint32_t b_stage = a_stage + (uint32_t)n_a * kWgtSlab;
#pragma unroll
          for (int ks = 0; ks < kWgtRows / 16; ++ks) {
            const uint64_t bdesc = umma_desc_mn_sw128(b_stage + (uint32_t)(ks * 2048), kWgtSlab);
            for (int mb = 0; mb < p.m_blocks; ++mb) {
              const uint64_t adesc = umma_desc_mn_sw128(a_stage + (uint32_t)(2 * mb) * kWgtSlab + (uint32_t)(ks * 2048), kWgtSlab);
              wgt_mma(tmem_base + (uint32_t)(mb * n2), adesc, bdesc, idesc, (t == t_begin && ks == 0) ? 0u : 1u);
            }
          }
          tc_commit(empty + 8 * s);
          if (t == t_end - 1) tc_commit(acc_full);
        }
        __syncwarp();
        if (++s == stages) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // =============================== out-grad loader (TMA) ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / p.kvol;
        const int t_begin = chunk * p.tiles_per_chunk, t_end = min(p.n_tiles, t_begin + p.tiles_per_chunk);
        for (int t = t_begin; t < t_end; ++t) {
          mbar_wait(empty + 8 * s, ph ^ 1u);
          mbar_arrive_expect_tx(full + 8 * s, (uint32_t)n_b * kWgtSlab);
          const uint32_t b_stage = ring + (uint32_t)s * stage_bytes + (uint32_t)n_a * kWgtSlab;
          for (int sl = 0; sl < n_b; ++sl)
            tma_tile_2d(b_stage + (uint32_t)sl * kWgtSlab, &gmap, sl * 64, t * kWgtRows, full + 8 * s);
          if (++s == stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// dW[e] = sum over chunks (ascending) of partial[chunk][e]
__global__ void spconv_wgrad_tc_reduce_kernel(const float *__restrict__ partial, long long elems, int n_chunks,
                                              float *__restrict__ w_grad) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < elems; e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int ch = 0; ch < n_chunks; ++ch) s += partial[(long long)ch * elems + e];
    w_grad[e] = s;
  }
}

typedef CUresult (*WgtEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WgtEncodeTiled wgt_encode_fn() {
  static WgtEncodeTiled fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (WgtEncodeTiled)f;
  }();
  return fn;
}

bool spconv_wgrad_tc_ok(int c_in, int c_out, int kvol) {
  static const bool enabled = [] {
    const char *e = getenv("BEVB200_WGRAD_TC");
    return e == nullptr || atoi(e) != 0;
  }();
  auto ok = [](int c) { return c == 32 || c == 64 || c == 128; };
  return enabled && ok(c_in) && ok(c_out) && kvol >= 1 && kvol <= 27;
}

// row chunks per kernel offset: ~3 work items per SM, every chunk non-empty
static int wgt_chunks(int n_out, int kvol, int *tiles_per_chunk = nullptr) {
  const int n_tiles = (n_out + kWgtRows - 1) / kWgtRows;
  int n_chunks = (3 * kNumSMs) / kvol;
  if (n_chunks < 1) n_chunks = 1;
  if (n_chunks > n_tiles) n_chunks = n_tiles > 0 ? n_tiles : 1;
  const int tpc = n_tiles > 0 ? (n_tiles + n_chunks - 1) / n_chunks : 1;
  n_chunks = n_tiles > 0 ? (n_tiles + tpc - 1) / tpc : 1;
  if (tiles_per_chunk) *tiles_per_chunk = tpc;
  return n_chunks;
}

size_t spconv_wgrad_tc_workspace_bytes(int n_in, int n_out, int c_in, int c_out, int kvol) {
  if (!spconv_wgrad_tc_ok(c_in, c_out, kvol) || n_in <= 0 || n_out <= 0) return 0;
  return align_up((size_t)n_in * c_in * 4) + align_up((size_t)n_out * c_out * 4) +
         align_up((size_t)wgt_chunks(n_out, kvol) * kvol * c_in * c_out * sizeof(float));
}

int spconv_wgrad_tc(const float *features, const float *out_grad, const int32_t *nbr, int n_in, int n_out,
                    int c_in, int c_out, int kvol, float *weight_grad, void *workspace, cudaStream_t st) {
  BEVB200_REQUIRE(spconv_wgrad_tc_ok(c_in, c_out, kvol), "shape has no tensor-core filter gradient");
  uint8_t *fsplit = (uint8_t *)workspace;
  uint8_t *gsplit = fsplit + align_up((size_t)n_in * c_in * 4);
  float *partial = (float *)(gsplit + align_up((size_t)n_out * c_out * 4));
  int rc = spconv_v6_split_rows(features, n_in, nullptr, c_in, fsplit, st);
  if (!rc) rc = spconv_v6_split_rows(out_grad, n_out, nullptr, c_out, gsplit, st);
  if (rc) return rc;
  WgtParams p;
  memset(&p, 0, sizeof(p));
  p.fsplit = fsplit;
  p.nbr = nbr;
  p.partial = partial;
  p.n_in = n_in; p.n_out = n_out; p.c_in = c_in; p.c_out = c_out; p.kvol = kvol;
  p.n_tiles = (n_out + kWgtRows - 1) / kWgtRows;
  p.n_chunks = wgt_chunks(n_out, kvol, &p.tiles_per_chunk);
  p.n_a = c_in / 32;
  p.n_b = c_out / 32;
  p.stage_bytes = (p.n_a + p.n_b) * kWgtSlab;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > kWgtMaxStages) p.stages = kWgtMaxStages;
  p.m_blocks = c_in >= 64 ? c_in / 64 : 1;
  int cols = p.m_blocks * 2 * c_out;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  BEVB200_REQUIRE(p.tmem_cols <= 512 && p.stages >= 2, "filter gradient tile does not fit");
  // out-grad split image as a 2-D bf16 tensor [n_out][c_out * 2]; one box = 64 rows x 128 B
  WgtEncodeTiled enc = wgt_encode_fn();
  BEVB200_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available");
  CUtensorMap gmap;
  const cuuint64_t dims[2] = {(cuuint64_t)c_out * 2, (cuuint64_t)n_out};
  const cuuint64_t strides[1] = {(cuuint64_t)c_out * 4};
  const cuuint32_t box[2] = {64, (cuuint32_t)kWgtRows};
  const cuuint32_t estr[2] = {1, 1};
  BEVB200_REQUIRE(enc(&gmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gsplit, dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS,
                  "cuTensorMapEncodeTiled failed for the out-grad image");
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const int n_items = p.n_chunks * kvol;
  const int grid = n_items < kNumSMs ? n_items : kNumSMs;
  BEVB200_CUDA(cudaFuncSetAttribute(spconv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  BEVB200_LAUNCH(spconv_wgrad_tc_kernel, grid, kWgtThreads, smem, st, p, gmap);
  const long long elems = (long long)kvol * c_in * c_out;
  BEVB200_LAUNCH(spconv_wgrad_tc_reduce_kernel, grid_for(elems, 256), 256, 0, st, partial, elems, p.n_chunks, weight_grad);
  return BEVB200_OK;
}

}  // namespace bevb200
