// Dev probe: throughput and layout of TMA tile::gather4 row gathers into SWIZZLE_128B tiles, against the
// LDGSTS (cp.async 16 B) gather the generation-6 sparse-conv kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/gather4_probe tools/gather4_probe.cu
//   ./tools/bin/gather4_probe
// One "item" is a 128-row x (slabs x 128 B) operand tile: what one (row tile, kernel offset) pair of the
// sparse convolution stages.  Indices are random rows with a given valid fraction (missing = -1 -> zeros).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static __device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n));
}
static __device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
      ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
static __device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                               uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)) : "memory");
}

constexpr int kStageBytes = 16384;   // 128 rows x 128 B

// One producer warp per CTA group of `warps`; each warp owns its own stage ring and its own items.
// stages: ring depth per warp.  Items are (tile, offset) pairs taken round-robin over all warps of the grid.
__global__ void __launch_bounds__(256) gather4_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ nbr,
                                                       int n_out, int kvol, int slabs, int stages, int items_total,
                                                       unsigned char* dump, int dump_item) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  unsigned char* ring = smem + (size_t)warp * stages * slabs * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)warps * stages * slabs * kStageBytes) + warp * stages;
  if (lane == 0)
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int tiles = (n_out + 127) / 128;
  const int gw = blockIdx.x * warps + warp, nw = gridDim.x * warps;
  int it = 0;
  for (int item = gw; item < items_total; item += nw, ++it) {
    const int s = it % stages;
    if (it >= stages) mbar_wait(&bars[s], ((it / stages) - 1) & 1);     // the copy issued `stages` items ago has landed
    const int tile = item % tiles, k = item / tiles;
    const int row0 = tile * 128 + lane * 4;
    int4 r = make_int4(-1, -1, -1, -1);
    if (row0 + 3 < n_out) r = *reinterpret_cast<const int4*>(nbr + (size_t)k * n_out + row0);
    if (lane == 0) mbar_expect_tx(&bars[s], slabs * kStageBytes);
    __syncwarp();
    unsigned char* dst = ring + (size_t)s * slabs * kStageBytes + lane * 512;
    for (int sl = 0; sl < slabs; ++sl) gather4(dst + sl * kStageBytes, &map, sl * 64, r.x, r.y, r.z, r.w, &bars[s]);
    if (item == dump_item) {
      mbar_wait(&bars[s], (it / stages) & 1);
      for (int i = lane; i < slabs * kStageBytes / 16; i += 32)
        reinterpret_cast<int4*>(dump)[i] = reinterpret_cast<const int4*>(ring + (size_t)s * slabs * kStageBytes)[i];
      // keep the phase bookkeeping of the ring consistent: this stage's phase was consumed here
    }
  }
  // drain
  for (int j = max(0, it - stages); j < it; ++j) {
    const int s = j % stages;
    mbar_wait(&bars[s], (j / stages) & 1);
  }
}

static __device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.b32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}

// Same ring, but ONE elected thread issues the 32 gather4 of an item (indices staged through shared memory), so the
// compiler emits straight UTMALDG instead of a 32-trip ELECT / R2UR / BRA.U.ANY loop around each one.
// zero_row >= 0: missing rows (-1) are redirected to that (all-zero) row instead of relying on the out-of-bounds fill.
__global__ void __launch_bounds__(256) gather4_elect_kernel(const __grid_constant__ CUtensorMap map, const int* __restrict__ nbr,
                                                             int n_out, int kvol, int slabs, int stages, int items_total,
                                                             int zero_row) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  unsigned char* ring = smem + (size_t)warp * stages * slabs * kStageBytes;
  unsigned char* tail = smem + (size_t)warps * stages * slabs * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail) + warp * stages;
  int4* idx_s = reinterpret_cast<int4*>(tail + 512) + warp * 64;          // two index buffers of 32 int4 per warp
  if (lane == 0)
    for (int s = 0; s < stages; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const int tiles = (n_out + 127) / 128;
  const int gw = blockIdx.x * warps + warp, nw = gridDim.x * warps;
  int it = 0;
  for (int item = gw; item < items_total; item += nw, ++it) {
    const int s = it % stages;
    const int tile = item % tiles, k = item / tiles;
    const int row0 = tile * 128 + lane * 4;
    int4 r = make_int4(-1, -1, -1, -1);
    if (row0 + 3 < n_out) r = *reinterpret_cast<const int4*>(nbr + (size_t)k * n_out + row0);
    if (zero_row >= 0) {
      r.x = r.x < 0 ? zero_row : r.x; r.y = r.y < 0 ? zero_row : r.y;
      r.z = r.z < 0 ? zero_row : r.z; r.w = r.w < 0 ? zero_row : r.w;
    }
    int4* buf = idx_s + (it & 1) * 32;
    buf[lane] = r;
    __syncwarp();
    if (it >= stages) mbar_wait(&bars[s], ((it / stages) - 1) & 1);
    if (elect_one()) {
      mbar_expect_tx(&bars[s], slabs * kStageBytes);
      unsigned char* dst = ring + (size_t)s * slabs * kStageBytes;
#pragma unroll 8
      for (int l = 0; l < 32; ++l) {
        const int4 v = buf[l];
        for (int sl = 0; sl < slabs; ++sl) gather4(dst + sl * kStageBytes + l * 512, &map, sl * 64, v.x, v.y, v.z, v.w, &bars[s]);
      }
    }
    __syncwarp();
  }
  for (int j = max(0, it - stages); j < it; ++j) mbar_wait(&bars[j % stages], (j / stages) & 1);
}

// LDGSTS reference: 8 lanes copy one 128 B row chunk, a warp covers 4 rows per instruction, 32 instructions per
// stage; rows are zero-filled through src-size 0.  Same ring discipline through commit groups.
__global__ void __launch_bounds__(256) ldgsts_kernel(const unsigned char* __restrict__ image, int row_bytes,
                                                      const int* __restrict__ nbr, int n_out, int kvol, int slabs,
                                                      int stages, int items_total) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  unsigned char* ring = smem + (size_t)warp * stages * slabs * kStageBytes;
  const int tiles = (n_out + 127) / 128;
  const int gw = blockIdx.x * warps + warp, nw = gridDim.x * warps;
  const int m = lane >> 3, c = lane & 7;
  int it = 0;
  for (int item = gw; item < items_total; item += nw, ++it) {
    const int s = it % stages;
    const int tile = item % tiles, k = item / tiles;
    const int row0 = tile * 128 + lane * 4;
    int4 r = make_int4(-1, -1, -1, -1);
    if (row0 + 3 < n_out) r = *reinterpret_cast<const int4*>(nbr + (size_t)k * n_out + row0);
    unsigned char* st = ring + (size_t)s * slabs * kStageBytes;
#pragma unroll 4
    for (int g = 0; g < 32; ++g) {                       // rows 4g .. 4g+3 : held by lane g
      const int rx = __shfl_sync(0xffffffffu, r.x, g), ry = __shfl_sync(0xffffffffu, r.y, g);
      const int rz = __shfl_sync(0xffffffffu, r.z, g), rw = __shfl_sync(0xffffffffu, r.w, g);
      const int row = m == 0 ? rx : m == 1 ? ry : m == 2 ? rz : rw;
      const int rr = 4 * g + m;
      const unsigned char* src = image + (size_t)(row < 0 ? 0 : row) * row_bytes;
      const uint32_t nbytes = row < 0 ? 0u : 16u;
      for (int sl = 0; sl < slabs; ++sl) {
        const uint32_t d = smem_u32(st + sl * kStageBytes + rr * 128 + ((c ^ (rr & 7)) << 4));
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src + sl * 128 + c * 16), "r"(nbytes) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // at most `stages - 1` groups stay in flight  (stages is 2, 4 or 8 here)
    if (stages == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else if (stages == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
    else asm volatile("cp.async.wait_group 7;" ::: "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int n_in = 308113, n_out = 308112, kvol = 27;           // n_out multiple of 4 keeps the int4 index loads aligned
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d\n", sms);
  for (int slabs : {1, 2, 4}) {
    const int row_bytes = slabs * 128;
    std::vector<uint16_t> h_img((size_t)(n_in + 1) * row_bytes / 2, 0);      // row n_in: all zero
    for (size_t i = 0; i < (size_t)n_in * row_bytes / 2; ++i) h_img[i] = (uint16_t)((i * 2654435761u) >> 13);
    unsigned char* d_img;
    CK(cudaMalloc(&d_img, h_img.size() * 2));
    CK(cudaMemcpy(d_img, h_img.data(), h_img.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)row_bytes / 2, (cuuint64_t)n_in + 1};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_img, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
    for (float valid : {0.45f, 1.0f}) {
      std::vector<int> h_nbr((size_t)kvol * n_out);
      uint64_t st = 88172645463325252ull;
      auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
      for (auto& v : h_nbr) {
        const bool ok = (rnd() % 1000) < (uint64_t)(valid * 1000);
        v = ok ? (int)(rnd() % n_in) : -1;
      }
      int* d_nbr;
      CK(cudaMalloc(&d_nbr, h_nbr.size() * 4));
      CK(cudaMemcpy(d_nbr, h_nbr.data(), h_nbr.size() * 4, cudaMemcpyHostToDevice));
      const int tiles = (n_out + 127) / 128, items = tiles * kvol;
      unsigned char* d_dump;
      CK(cudaMalloc(&d_dump, slabs * kStageBytes));
      // ---- layout check (one item) ----
      {
        const int dump_item = 12345 % items;
        const int warps = 1, stages = 2;
        const size_t sm = (size_t)warps * stages * slabs * kStageBytes + 1024;
        CK(cudaFuncSetAttribute(gather4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        CK(cudaMemset(d_dump, 0xee, slabs * kStageBytes));
        gather4_kernel<<<1, 32, sm>>>(map, d_nbr, n_out, kvol, slabs, stages, dump_item + 1, d_dump, dump_item);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> h_dump(slabs * kStageBytes);
        CK(cudaMemcpy(h_dump.data(), d_dump, h_dump.size(), cudaMemcpyDeviceToHost));
        const int tile = dump_item % tiles, k = dump_item / tiles;
        long bad = 0, zero_rows = 0;
        for (int r = 0; r < 128; ++r) {
          const int idx = h_nbr[(size_t)k * n_out + tile * 128 + r];
          if (idx < 0) ++zero_rows;
          for (int sl = 0; sl < slabs; ++sl)
            for (int c = 0; c < 8; ++c) {
              const unsigned char* got = &h_dump[sl * kStageBytes + r * 128 + ((c ^ (r & 7)) << 4)];
              unsigned char want[16];
              if (idx < 0) memset(want, 0, 16);
              else memcpy(want, (const unsigned char*)h_img.data() + (size_t)idx * row_bytes + sl * 128 + c * 16, 16);
              if (memcmp(got, want, 16) != 0) ++bad;
            }
        }
        printf("slabs %d valid %.2f  layout check: %ld bad 16-byte chunks of %d (%ld missing rows)\n", slabs, valid, bad,
               128 * slabs * 8, zero_rows);
      }
      // ---- throughput ----
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      const double bytes = (double)items * slabs * kStageBytes;
      for (int ctas_per_sm : {1, 2})
        for (int warps : {1, 2, 4, 8})
          for (int stages : {2, 4}) {
            const size_t sm = (size_t)warps * stages * slabs * kStageBytes + 1024;
            if ((sm + 8192) * ctas_per_sm > 220 * 1024) continue;
            CK(cudaFuncSetAttribute(gather4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            CK(cudaFuncSetAttribute(ldgsts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            float ms_t = 0, ms_e = 0, ms_z = 0;
            CK(cudaFuncSetAttribute(gather4_elect_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm + 8192));
            for (int rep = 0; rep < 3; ++rep) {
              CK(cudaEventRecord(e0));
              gather4_kernel<<<sms * ctas_per_sm, warps * 32, sm>>>(map, d_nbr, n_out, kvol, slabs, stages, items, d_dump, -1);
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
              CK(cudaEventElapsedTime(&ms_t, e0, e1));
            }
            for (int rep = 0; rep < 3; ++rep) {
              CK(cudaEventRecord(e0));
              gather4_elect_kernel<<<sms * ctas_per_sm, warps * 32, sm + 8192>>>(map, d_nbr, n_out, kvol, slabs, stages, items, -1);
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
              CK(cudaEventElapsedTime(&ms_e, e0, e1));
            }
            for (int rep = 0; rep < 3; ++rep) {
              CK(cudaEventRecord(e0));
              gather4_elect_kernel<<<sms * ctas_per_sm, warps * 32, sm + 8192>>>(map, d_nbr, n_out, kvol, slabs, stages, items, n_in);
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
              CK(cudaEventElapsedTime(&ms_z, e0, e1));
            }
            const double bc = bytes / 1e-3 / sms / 1.965e9;
            printf("slabs %d valid %.2f ctas/SM %d warps %d stages %d : all-lanes %7.1f us %5.1f B/clk/SM | elected %7.1f us %5.1f | elected, zero row %7.1f us %5.1f\n",
                   slabs, valid, ctas_per_sm, warps, stages, ms_t * 1e3, bc / ms_t, ms_e * 1e3, bc / ms_e, ms_z * 1e3, bc / ms_z);
            fflush(stdout);
          }
      CK(cudaFree(d_nbr));
      CK(cudaFree(d_dump));
    }
    CK(cudaFree(d_img));
  }
  return 0;
}
