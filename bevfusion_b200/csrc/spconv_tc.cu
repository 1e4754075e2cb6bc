// Sparse convolution forward on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   out[o, :] = epilogue( sum_k features[nbr[k, o], :] @ W[k] )          (spconv_ops.h:260-361)
//
// One CTA owns 128 output rows (UMMA M = 128, cta_group::1) and all Cout <= 128 columns; the
// fp32 accumulator lives in TMEM (128 lanes x Cout columns) across all kernel offsets and
// Cin blocks, so every output row is written exactly once, after the fused BN / residual /
// ReLU epilogue.  Pipeline (NS stages, mbarrier full/empty pairs):
//
//   warps 0-7  producers: thread (row r, half h) gathers 64 B of features[nbr[k, row0+r]]
//              (one 32-float K block = 128 B per row per stage) with 16-byte loads, splits every
//              value into tf32 hi + lo parts (3xTF32: a = hi + lo, hi = a & ~0x1fff) and stores both
//              into shared memory in the UMMA canonical K-major SWIZZLE_128B layout
//              (16-byte chunk index XOR (row & 7)).  Missing neighbours become zero rows.
//              Thread 0 also issues ONE cp.async.bulk (TMA 1-D) per stage for the weights of
//              (offset k, K block): they are pre-packed in global memory as the exact swizzled
//              shared-memory image [hi | lo][Cout rows][128 B].
//   warp 8     lane 0 issues tcgen05.mma.kind::tf32 (M128 x N=Cout x K8): per K step
//              hi*hi + hi*lo + lo*hi (fp32 accumulate in TMEM) and tcgen05.commit's the stage
//              back to the producers; after the last stage it commits to the epilogue barrier.
//   epilogue   warps 0-7 read the accumulator with tcgen05.ld (32 lanes x 16 columns per
//              instruction; warps w and w+4 share TMEM lane quarter w and split the columns),
//              apply scale/shift (folded BatchNorm1d), residual and ReLU, and store the row.
//
// BEVB200_PREC_TF32X3 keeps fp32-class accuracy (error ~2^-21 per product) at 3 MMAs per K
// step; BEVB200_PREC_TF32 issues only hi*hi (single-pass TF32, ~1e-3 relative).
#include "common.cuh"

namespace bevb200 {

constexpr int kTcProducerThreads = 256;
constexpr int kTcThreads = kTcProducerThreads + 32;
constexpr int kTileM = 128;
constexpr int kKBlock = 32;                       // floats per K block = one 128-byte swizzle row
constexpr int kABlockBytes = kTileM * 128;        // one split part of the A stage: 16 KB

// ---- PTX wrappers ----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void *src, uint32_t bytes,
                                              uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 (= 1024 B between 8-row groups), [46,48) version = 1,
//   [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = (uint64_t)((smem_addr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor for kind::tf32: c=F32 (bit 4), a=b=TF32 (2 at bits 7, 10), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ inline uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcParams {
  const float *features;
  const float *wpacked;   // [K][nkb][nsplit][Cout][32] floats, swizzled smem image
  const int32_t *nbr;
  const float *scale, *shift, *residual;
  float *out;
  int n_in, n_out, c_in, c_out, kvol, relu;
  int nkb;        // K blocks of 32 floats over the concatenated (offset, channel) axis
  int cin_shift;  // log2(c_in): c_in is a power of two >= 16 on this path
  int nstages;
  int tmem_cols;  // power of two >= max(32, c_out)
};

template <int NSPLIT>
__global__ void __launch_bounds__(kTcThreads, 2) spconv_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // align the stage area to 1024 B (SWIZZLE_128B atoms)
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_part_bytes = p.c_out * 128;                       // one split part of B: Cout x 128 B
  const int stage_bytes = NSPLIT * (kABlockBytes + b_part_bytes);
  __shared__ uint64_t bars[2 * 8 + 1];
  __shared__ uint32_t tmem_base_s;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]), accbar = smem_u32(&bars[16]);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * kTileM;
  const int NS = p.nstages;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8 * s, kTcProducerThreads + 1);
      mbar_init(empty0 + 8 * s, 1);
    }
    mbar_init(accbar, 1);
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
  const int n_iters = p.nkb;

  if (warp < 8) {
    // =============================== producers ===========================================
    const int r = tid & 127, half = tid >> 7;
    const uint32_t sw = (uint32_t)(r & 7);
    const uint32_t row_off = (uint32_t)r * 128u;
    // neighbour rows of this tile for every offset, staged once in shared memory
    int32_t *nbr_s = reinterpret_cast<int32_t *>(smem + (size_t)NS * stage_bytes);
    for (int i = tid; i < p.kvol * kTileM; i += kTcProducerThreads) {
      const int k = i >> 7, rr = i & 127, o = row0 + rr;
      int v = o < p.n_out ? __ldg(p.nbr + (long long)k * p.n_out + o) : -1;
      if (v >= p.n_in) v = -1;
      nbr_s[i] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kTcProducerThreads) : "memory");

    // K block `it` covers concatenated-K indices [32*it, 32*it+32); this thread owns 16 of them
    // (4 float4 chunks), which lie inside ONE kernel offset because c_in % 16 == 0.
    auto issue = [&](int it, float4 (&v)[4]) {
      const int kk = it * kKBlock + half * 16;
      const int k = kk >> p.cin_shift;
      const int ch = kk & (p.c_in - 1);
      const int src = k < p.kvol ? nbr_s[k * kTileM + r] : -1;
      if (src >= 0) {
        const float4 *q = reinterpret_cast<const float4 *>(p.features + (long long)src * p.c_in + ch);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldg(q + j);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    constexpr int PD = 4;  // register prefetch distance (K blocks in flight per thread)
    float4 v[PD][4];
#pragma unroll
    for (int j = 0; j < PD; ++j)
      if (j < n_iters) issue(j, v[j]);
    for (int it0 = 0; it0 < n_iters; it0 += PD) {
#pragma unroll
      for (int jj = 0; jj < PD; ++jj) {
        const int it = it0 + jj;
        if (it < n_iters) {
          const int s = it % NS;
          const uint32_t ph = (uint32_t)(it / NS) & 1u;
          mbar_wait(empty0 + 8 * s, ph ^ 1u);
          const uint32_t stage = smem_base + (uint32_t)s * (uint32_t)stage_bytes;
          if (tid == 0) {
            const uint32_t bbytes = (uint32_t)(NSPLIT * b_part_bytes);
            mbar_arrive_expect_tx(full0 + 8 * s, bbytes);
            const float *wsrc = p.wpacked + (long long)it * (long long)(NSPLIT * p.c_out * 32);
            bulk_copy_g2s(stage + NSPLIT * kABlockBytes, wsrc, bbytes, full0 + 8 * s);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t chunk = (uint32_t)(half * 4 + j);
            const uint32_t off = row_off + ((chunk ^ sw) << 4);
            uint4 hi;
            hi.x = __float_as_uint(v[jj][j].x) & 0xffffe000u;
            hi.y = __float_as_uint(v[jj][j].y) & 0xffffe000u;
            hi.z = __float_as_uint(v[jj][j].z) & 0xffffe000u;
            hi.w = __float_as_uint(v[jj][j].w) & 0xffffe000u;
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(stage + off), "r"(hi.x), "r"(hi.y),
                         "r"(hi.z), "r"(hi.w) : "memory");
            if (NSPLIT == 2) {
              float4 lo;
              lo.x = v[jj][j].x - __uint_as_float(hi.x);
              lo.y = v[jj][j].y - __uint_as_float(hi.y);
              lo.z = v[jj][j].z - __uint_as_float(hi.z);
              lo.w = v[jj][j].w - __uint_as_float(hi.w);
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(stage + kABlockBytes + off),
                           "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w) : "memory");
            }
          }
          fence_proxy_async();           // generic-proxy stores -> visible to the tensor core (async proxy)
          mbar_arrive(full0 + 8 * s);
          if (it + PD < n_iters) issue(it + PD, v[jj]);
        }
      }
    }
    // =============================== epilogue ============================================
    mbar_wait(accbar, 0);
    tc_fence_after();
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int orow = row0 + q * 32 + lane;  // accumulator lane == output row of the tile
    const int ncol_half = p.c_out / 2;      // warps q and q+4 split the columns
    const int col_begin = (warp >> 2) * ncol_half;
    for (int c0 = col_begin; c0 < col_begin + ncol_half; c0 += 16) {
      float acc[16];
      tc_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, acc);
      if (orow < p.n_out) {
        float *dst = p.out + (long long)orow * p.c_out + c0;
        const float *res = p.residual ? p.residual + (long long)orow * p.c_out + c0 : nullptr;
        const int ncols = min(16, col_begin + ncol_half - c0);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          if (j < ncols) {
            float y[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float t = acc[j + e];
              if (p.scale) t *= __ldg(p.scale + c0 + j + e);
              if (p.shift) t += __ldg(p.shift + c0 + j + e);
              y[e] = t;
            }
            if (res) {
              float4 rv = __ldg(reinterpret_cast<const float4 *>(res + j));
              y[0] += rv.x; y[1] += rv.y; y[2] += rv.z; y[3] += rv.w;
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) y[e] = fmaxf(y[e], 0.f);
            }
            *reinterpret_cast<float4 *>(dst + j) = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
      }
    }
  } else {
    // =============================== MMA issuer (warp 8; lane 0 issues) ===================
    const uint32_t idesc = umma_idesc_tf32(kTileM, p.c_out);
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % NS;
      const uint32_t ph = (uint32_t)(it / NS) & 1u;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t stage = smem_base + (uint32_t)s * (uint32_t)stage_bytes;
        const uint64_t a_hi = umma_desc_sw128(stage);
        const uint64_t a_lo = umma_desc_sw128(stage + kABlockBytes);
        const uint64_t b_hi = umma_desc_sw128(stage + NSPLIT * kABlockBytes);
        const uint64_t b_lo = umma_desc_sw128(stage + NSPLIT * kABlockBytes + b_part_bytes);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 2);  // +32 B along K inside the 128-byte swizzle row
          if (NSPLIT == 2) {
            // small terms first, then the dominant product
            tc_mma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            tc_mma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
            tc_mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
          } else {
            tc_mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          }
        }
        tc_commit(empty0 + 8 * s);   // frees the smem slot once these MMAs have read it
        if (it == n_iters - 1) tc_commit(accbar);   // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// weight [K][Cin][Cout] fp32 -> packed [nkb][nsplit][Cout][32] in the swizzled smem image.  The K
// axis is the concatenation over kernel offsets of the Cin channels (kk = k*Cin + ci), cut into
// blocks of 32; element (n, c) of block kb sits at float index n*32 + (((c/4) ^ (n&7)) * 4) + (c%4).
__global__ void spconv_pack_weights_kernel(const float *__restrict__ w, int kvol, int c_in, int c_out,
                                           int nkb, int nsplit, float *__restrict__ packed) {
  const long long total = (long long)nkb * c_out * 32;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % 32);
    const int n = (int)((t / 32) % c_out);
    const int kb = (int)(t / (32ll * c_out));
    const int kk = kb * 32 + c;
    const int k = kk / c_in, ci = kk % c_in;
    const float v = k < kvol ? w[((long long)k * c_in + ci) * c_out + n] : 0.f;
    const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    const long long blk = (long long)kb * nsplit;
    const int pos = n * 32 + ((((c >> 2) ^ (n & 7)) << 2) | (c & 3));
    packed[(blk + 0) * c_out * 32 + pos] = hi;
    if (nsplit == 2) packed[(blk + 1) * c_out * 32 + pos] = v - hi;
  }
}

int spconv_forward_simt(const float *features, const float *weight, const int32_t *nbr, int n_in,
                        int n_out, int c_in, int c_out, int kvol, const float *scale,
                        const float *shift, const float *residual, int relu, float *out,
                        cudaStream_t st);

static bool tc_shape_ok(int c_in, int c_out, int kvol) {
  return (c_out == 16 || c_out == 32 || c_out == 64 || c_out == 128) &&
         (c_in == 16 || c_in == 32 || c_in == 64 || c_in == 128) && kvol >= 1 && kvol <= 27;
}

static int tc_nkb(int c_in, int kvol) { return (kvol * c_in + kKBlock - 1) / kKBlock; }

size_t spconv_packed_bytes(int c_in, int c_out, int kvol, int precision) {
  if (!tc_shape_ok(c_in, c_out, kvol)) return 0;
  const int nsplit = precision == BEVB200_PREC_TF32X3 ? 2 : 1;
  return (size_t)tc_nkb(c_in, kvol) * nsplit * c_out * 32 * sizeof(float);
}

int spconv_pack_weights(const float *weight, int c_in, int c_out, int kvol, int precision,
                        float *packed, cudaStream_t st) {
  const int nsplit = precision == BEVB200_PREC_TF32X3 ? 2 : 1;
  const int nkb = tc_nkb(c_in, kvol);
  BEVB200_LAUNCH(spconv_pack_weights_kernel, grid_for((long long)nkb * c_out * 32, 256), 256, 0, st,
                 weight, kvol, c_in, c_out, nkb, nsplit, packed);
  return BEVB200_OK;
}

// `packed` may be null: the weights are then packed into a stream-ordered temporary.
int spconv_forward_tc(const float *features, const float *weight, const float *packed_in,
                      const int32_t *nbr, int n_in, int n_out, int c_in, int c_out, int kvol,
                      const float *scale, const float *shift, const float *residual, int relu,
                      int precision, float *out, cudaStream_t st) {
  const bool ok = tc_shape_ok(c_in, c_out, kvol) && ((uintptr_t)features % 16 == 0) &&
                  ((uintptr_t)out % 16 == 0) && (residual == nullptr || (uintptr_t)residual % 16 == 0);
  if (!ok) {
    // shapes the UMMA tile cannot take (e.g. conv_input, Cin = 5): exact-fp32 SIMT kernel
    if (weight == nullptr) {
      snprintf(g_last_error, sizeof(g_last_error), "spconv_forward: shape needs the unpacked weights");
      return BEVB200_EINVAL;
    }
    return spconv_forward_simt(features, weight, nbr, n_in, n_out, c_in, c_out, kvol, scale, shift,
                               residual, relu, out, st);
  }
  const int nsplit = precision == BEVB200_PREC_TF32X3 ? 2 : 1;
  TcParams p;
  p.features = features; p.nbr = nbr; p.scale = scale; p.shift = shift; p.residual = residual;
  p.out = out; p.n_in = n_in; p.n_out = n_out; p.c_in = c_in; p.c_out = c_out; p.kvol = kvol;
  p.relu = relu;
  p.nkb = tc_nkb(c_in, kvol);
  p.cin_shift = 0;
  while ((1 << p.cin_shift) < c_in) ++p.cin_shift;
  p.tmem_cols = c_out < 32 ? 32 : c_out;
  const int stage_bytes = nsplit * (kABlockBytes + c_out * 128);
  // two CTAs per SM when the stages are small enough (overlaps one tile's epilogue with the
  // other's main loop); otherwise one CTA with a deeper ring
  const int nbr_bytes = kvol * kTileM * 4;
  int ns = 2;
  if (2 * stage_bytes + nbr_bytes + 1024 > 111 * 1024) {
    ns = (215 * 1024 - nbr_bytes) / stage_bytes;
    if (ns > 4) ns = 4;
  }
  p.nstages = ns;
  const size_t smem = (size_t)ns * stage_bytes + nbr_bytes + 1024;
  float *packed = nullptr;
  if (packed_in == nullptr) {
    BEVB200_CUDA(cudaMallocAsync((void **)&packed, spconv_packed_bytes(c_in, c_out, kvol, precision), st));
    int rc = spconv_pack_weights(weight, c_in, c_out, kvol, precision, packed, st);
    if (rc) return rc;
    p.wpacked = packed;
  } else {
    p.wpacked = packed_in;
  }
  const int grid = (n_out + kTileM - 1) / kTileM;
  if (nsplit == 2) {
    BEVB200_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    BEVB200_LAUNCH(spconv_tc_kernel<2>, grid, kTcThreads, smem, st, p);
  } else {
    BEVB200_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    BEVB200_LAUNCH(spconv_tc_kernel<1>, grid, kTcThreads, smem, st, p);
  }
  if (packed) BEVB200_CUDA(cudaFreeAsync(packed, st));
  return BEVB200_OK;
}

}  // namespace bevb200
