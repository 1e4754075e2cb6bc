// libbevfusion_b200: error state, launch counter, device-wide exclusive scan.
#include "common.cuh"

namespace bevb200 {

thread_local char g_last_error[512] = "";
thread_local long long g_launch_count = 0;

// ---------------------------------------------------------------------------------------
// Exclusive scan of uint32 (optionally of popc(word)): reduce tiles -> scan tile sums ->
// apply.  Tiles are 4096 elements (256 threads x 16 consecutive elements).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

// Block-wide exclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024).
// Returns the exclusive prefix; *block_total gets the sum over the block.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *smem_warp /*[32]*/,
                                                    uint32_t *block_total) {
  const int lane = lane_id(), warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  uint32_t incl = warp_incl_scan(v);
  if (lane == 31) smem_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < nwarp ? smem_warp[lane] : 0u;
    uint32_t wi = warp_incl_scan(w);
    smem_warp[lane] = wi - w;  // exclusive warp offsets
    if (lane == 31) smem_warp[32] = wi;
  }
  __syncthreads();
  uint32_t res = incl - v + smem_warp[warp];
  *block_total = smem_warp[32];
  __syncthreads();
  return res;
}

template <bool POPC>
__global__ void __launch_bounds__(256) scan_tile_reduce(const uint32_t *__restrict__ in,
                                                        size_t count,
                                                        uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t sw[33];
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 16;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    size_t i = base + j;
    if (i < count) s += POPC ? (uint32_t)__popc(in[i]) : in[i];
  }
  uint32_t total;
  block_excl_scan(s, sw, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_tile_sums(uint32_t *__restrict__ tile_sums,
                                                       size_t ntiles,
                                                       uint32_t *__restrict__ total_out) {
  __shared__ uint32_t sw[33];
  uint32_t carry = 0;
  for (size_t base = 0; base < ntiles; base += blockDim.x) {
    size_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0u;
    uint32_t tot;
    uint32_t ex = block_excl_scan(v, sw, &tot);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <bool POPC>
__global__ void __launch_bounds__(256) scan_tile_apply(const uint32_t *__restrict__ in,
                                                       uint32_t *__restrict__ out, size_t count,
                                                       const uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t sw[33];
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 16;
  uint32_t v[16];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    size_t i = base + j;
    v[j] = i < count ? (POPC ? (uint32_t)__popc(in[i]) : in[i]) : 0u;
    s += v[j];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan(s, sw, &total) + tile_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    size_t i = base + j;
    if (i < count) out[i] = ex;
    ex += v[j];
  }
}

int exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t count, uint32_t *tile_sums,
                       uint32_t *total, bool popc_input, cudaStream_t stream) {
  size_t ntiles = (count + kScanTile - 1) / kScanTile;
  if (ntiles == 0) {
    if (total) BEVB200_CUDA(cudaMemsetAsync(total, 0, sizeof(uint32_t), stream));
    return BEVB200_OK;
  }
  if (popc_input) {
    BEVB200_LAUNCH(scan_tile_reduce<true>, (unsigned)ntiles, 256, 0, stream, in, count, tile_sums);
  } else {
    BEVB200_LAUNCH(scan_tile_reduce<false>, (unsigned)ntiles, 256, 0, stream, in, count, tile_sums);
  }
  BEVB200_LAUNCH(scan_tile_sums, 1, 1024, 0, stream, tile_sums, ntiles, total);
  if (popc_input) {
    BEVB200_LAUNCH(scan_tile_apply<true>, (unsigned)ntiles, 256, 0, stream, in, out, count, tile_sums);
  } else {
    BEVB200_LAUNCH(scan_tile_apply<false>, (unsigned)ntiles, 256, 0, stream, in, out, count, tile_sums);
  }
  return BEVB200_OK;
}

}  // namespace bevb200

extern "C" {
int bevb200_version(void) { return 100; }
const char *bevb200_last_error(void) { return bevb200::g_last_error; }
long long bevb200_launch_count(void) { return bevb200::g_launch_count; }
void bevb200_reset_launch_count(void) { bevb200::g_launch_count = 0; }
}
