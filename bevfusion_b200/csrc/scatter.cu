// DynamicScatter for sm_100a: reduce point features into the voxels named by per-point integer
// coordinates (mmdet3d/ops/voxel/src/scatter_points_cuda.cu:187-241 forward, :243-315 backward).
//
// The reference runs at::unique_dim (a generic lexicographic row sort) and then reduces with
// float atomics, so its sums depend on the arrival order.  Here the points are ordered by one
// stable radix sort of a packed coordinate key; a voxel is a run of equal keys and its points sit
// in ascending point-index order, so each (voxel, channel) is reduced by one thread in a fixed
// order: bit-reproducible, no atomics, and the arg-max for the backward pass falls out for free.
//
//   K1 ds_keys      coors row -> 64-bit key (most significant column first = unique_dim's
//                   lexicographic order); rows with a negative entry get the invalid key
//   (cub)           stable SortPairs(key, point index)
//   K2 ds_heads     head[i] = key[i] valid and != key[i-1]
//   (scan)          exclusive scan of head = voxel id per sorted position; total = M
//   K3 ds_segments  coors_map, out_coors, segment starts
//   K4 ds_reduce    one thread per (voxel, channel): sum / mean / max over the segment
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace bevb200 {

enum { kReduceSum = 0, kReduceMean = 1, kReduceMax = 2 };  // scatter_points_cuda.cu:7

struct KeyLayout {
  int ndim, bits;
  __host__ __device__ unsigned long long invalid() const { return 1ull << (ndim * bits); }
  __host__ __device__ int sort_bits() const { return ndim * bits + 1; }
};

static KeyLayout key_layout(int ndim) { return KeyLayout{ndim, ndim <= 3 ? 20 : 15}; }

__global__ void __launch_bounds__(256)
    ds_keys_kernel(const int32_t *__restrict__ coors, int n, KeyLayout kl,
                   unsigned long long *__restrict__ keys, uint32_t *__restrict__ idx,
                   int32_t *__restrict__ meta) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned long long key = 0;
    bool neg = false, big = false;
    for (int k = 0; k < kl.ndim; ++k) {
      int v = coors[(long long)i * kl.ndim + k];
      neg |= v < 0;
      big |= v >= (1 << kl.bits);
      key = (key << kl.bits) | (unsigned long long)(v & ((1 << kl.bits) - 1));
    }
    if (big && !neg) atomicAdd(meta + 1, 1);  // reported to the host as an error
    keys[i] = (neg || big) ? kl.invalid() : key;
    idx[i] = (uint32_t)i;
  }
}

__global__ void __launch_bounds__(256)
    ds_heads_kernel(const unsigned long long *__restrict__ keys, int n, KeyLayout kl,
                    uint32_t *__restrict__ head) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned long long k = keys[i];
    head[i] = (k < kl.invalid() && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
  }
}

__global__ void __launch_bounds__(256)
    ds_segments_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ idx,
                       const uint32_t *__restrict__ head, const uint32_t *__restrict__ vid_ex,
                       const uint32_t *__restrict__ total, int n, KeyLayout kl,
                       int32_t *__restrict__ coors_map, int32_t *__restrict__ out_coors,
                       uint32_t *__restrict__ seg_start, int32_t *__restrict__ meta) {
  if (blockIdx.x == 0 && threadIdx.x == 0) meta[0] = (int32_t)*total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    unsigned long long k = keys[i];
    const bool valid = k < kl.invalid();
    const int v = (int)(vid_ex[i] + head[i]) - 1;
    coors_map[idx[i]] = valid ? v : -1;  // scatter_points_cuda.cu:205-215 (the -1 row removed)
    if (!valid) continue;
    if (head[i]) {
      seg_start[v] = (uint32_t)i;
      for (int d = kl.ndim - 1; d >= 0; --d) {
        out_coors[(long long)v * kl.ndim + d] = (int32_t)(k & ((1ull << kl.bits) - 1));
        k >>= kl.bits;
      }
    }
    if (i == n - 1 || keys[i + 1] >= kl.invalid()) seg_start[v + 1] = (uint32_t)(i + 1);
  }
}

__global__ void __launch_bounds__(256)
    ds_reduce_kernel(const float *__restrict__ feats, const uint32_t *__restrict__ idx,
                     const uint32_t *__restrict__ seg_start, const int32_t *__restrict__ meta, int c,
                     int reduce_type, float *__restrict__ reduced, int32_t *__restrict__ count,
                     int32_t *__restrict__ reduce_from) {
  const long long total = (long long)meta[0] * c;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(t / c), ch = (int)(t % c);
    const uint32_t s = seg_start[v], e = seg_start[v + 1];
    float acc;
    int from = -1;
    if (reduce_type == kReduceMax) {
      acc = -INFINITY;  // scatter_points_cuda.cu:223; first (= smallest index) maximum wins (:156)
      for (uint32_t j = s; j < e; ++j) {
        const uint32_t p = idx[j];
        const float x = feats[(long long)p * c + ch];
        if (x > acc) { acc = x; from = (int)p; }            // NaN never wins, like fmaxf (:22-30)
        else if (from < 0 && x == acc) from = (int)p;      // a voxel whose maximum is -inf
      }
    } else {
      acc = 0.f;
      for (uint32_t j = s; j < e; ++j) acc = __fadd_rn(acc, feats[(long long)idx[j] * c + ch]);
      if (reduce_type == kReduceMean) acc = __fdiv_rn(acc, (float)(e - s));  // :234-235
    }
    reduced[t] = acc;
    if (reduce_from) reduce_from[t] = from;
    if (ch == 0) count[v] = (int32_t)(e - s);
  }
}

// ---- backward ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    ds_bwd_add_kernel(const float *__restrict__ grad_reduced, const int32_t *__restrict__ coors_map,
                      const int32_t *__restrict__ count, long long total, int c, int mean,
                      float *__restrict__ grad_feats) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / c), ch = (int)(t % c);
    const int v = coors_map[i];
    float g = 0.f;
    if (v >= 0) {
      g = grad_reduced[(long long)v * c + ch];
      if (mean) g = __fdiv_rn(g, (float)count[v]);  // scatter_points_cuda.cu:133-137
    }
    grad_feats[t] = g;
  }
}

__global__ void __launch_bounds__(256)
    ds_bwd_traceback_kernel(const float *__restrict__ feats, const float *__restrict__ reduced,
                            const int32_t *__restrict__ coors_map, long long total, int c,
                            int32_t *__restrict__ reduce_from) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(t / c), ch = (int)(t % c);
    const int v = coors_map[i];
    if (v < 0) continue;
    if (feats[t] == reduced[(long long)v * c + ch]) atomicMin(reduce_from + (long long)v * c + ch, i);
  }
}

__global__ void __launch_bounds__(256)
    ds_bwd_max_kernel(const float *__restrict__ grad_reduced, const int32_t *__restrict__ reduce_from,
                      long long total, int c, int n, float *__restrict__ grad_feats) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int from = reduce_from[t];
    if (from >= 0 && from < n) grad_feats[(long long)from * c + (int)(t % c)] = grad_reduced[t];
  }
}

struct ScatterWs {
  unsigned long long *keys_a, *keys_b;
  uint32_t *idx_a, *idx_b, *head, *vid_ex, *seg_start, *tiles, *total;
  void *cub_tmp;
  size_t cub_bytes;
};

static size_t scatter_layout(Arena &a, int n, ScatterWs &w) {
  w.keys_a = a.take<unsigned long long>(n);
  w.keys_b = a.take<unsigned long long>(n);
  w.idx_a = a.take<uint32_t>(n);
  w.idx_b = a.take<uint32_t>(n);
  w.head = a.take<uint32_t>(n);
  w.vid_ex = a.take<uint32_t>(n);
  w.seg_start = a.take<uint32_t>((size_t)n + 1);
  w.tiles = a.take<uint32_t>(scan_scratch_elems(n));
  w.total = a.take<uint32_t>(1);
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned long long *)nullptr,
                                  (unsigned long long *)nullptr, (const uint32_t *)nullptr,
                                  (uint32_t *)nullptr, n, 0, 64);
  w.cub_bytes = cub_bytes;
  w.cub_tmp = a.take<char>(cub_bytes);
  return a.off;
}

}  // namespace bevb200

using namespace bevb200;

extern "C" {

size_t bevb200_dynamic_scatter_workspace_bytes(int num_points) {
  Arena a(nullptr, 0);
  ScatterWs w;
  return scatter_layout(a, num_points > 0 ? num_points : 1, w);
}

int bevb200_dynamic_scatter(const float *feats, const int32_t *coors, int num_points,
                            int num_features, int ndim, int reduce_type, float *reduced_feats,
                            int32_t *out_coors, int32_t *coors_map, int32_t *reduce_count,
                            int32_t *reduce_from, int32_t *meta, void *workspace,
                            size_t workspace_bytes, void *stream) {
  BEVB200_REQUIRE(num_points >= 0 && num_features > 0, "bad sizes");
  BEVB200_REQUIRE(ndim >= 1 && ndim <= 4, "coors must have 1..4 columns");
  BEVB200_REQUIRE(reduce_type >= kReduceSum && reduce_type <= kReduceMax, "bad reduce type");
  BEVB200_REQUIRE(meta, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  BEVB200_CUDA(cudaMemsetAsync(meta, 0, 2 * sizeof(int32_t), st));
  if (num_points == 0) return BEVB200_OK;
  BEVB200_REQUIRE(feats && coors && reduced_feats && out_coors && coors_map && reduce_count && workspace,
                  "null argument");
  const int n = num_points;
  Arena a(workspace, workspace_bytes);
  ScatterWs w;
  scatter_layout(a, n, w);
  if (!a.ok()) {
    snprintf(g_last_error, sizeof(g_last_error), "dynamic_scatter: workspace too small");
    return BEVB200_EWORKSPACE;
  }
  const KeyLayout kl = key_layout(ndim);
  const int grid = grid_for(n, 256);
  BEVB200_LAUNCH(ds_keys_kernel, grid, 256, 0, st, coors, n, kl, w.keys_a, w.idx_a, meta);
  size_t cub_bytes = w.cub_bytes;
  BEVB200_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_tmp, cub_bytes, (const unsigned long long *)w.keys_a,
                                               w.keys_b, (const uint32_t *)w.idx_a, w.idx_b, n, 0,
                                               kl.sort_bits(), st));
  g_launch_count += (kl.sort_bits() + 7) / 8 + 2;  // cub: histogram + one onesweep pass per 8 bits
  BEVB200_LAUNCH(ds_heads_kernel, grid, 256, 0, st, w.keys_b, n, kl, w.head);
  int rc = exclusive_scan_u32(w.head, w.vid_ex, n, w.tiles, w.total, false, st);
  if (rc != BEVB200_OK) return rc;
  BEVB200_LAUNCH(ds_segments_kernel, grid, 256, 0, st, w.keys_b, w.idx_b, w.head, w.vid_ex, w.total,
                 n, kl, coors_map, out_coors, w.seg_start, meta);
  BEVB200_LAUNCH(ds_reduce_kernel, grid_for((long long)n * num_features, 256), 256, 0, st, feats,
                 w.idx_b, w.seg_start, meta, num_features, reduce_type, reduced_feats, reduce_count,
                 reduce_from);
  return BEVB200_OK;
}

int bevb200_dynamic_scatter_backward(const float *grad_reduced_feats, const float *feats,
                                     const float *reduced_feats, const int32_t *coors_map,
                                     const int32_t *reduce_count, int32_t *reduce_from,
                                     int reduce_from_valid, int num_points, int num_reduced,
                                     int num_features, int reduce_type, float *grad_feats,
                                     void *stream) {
  BEVB200_REQUIRE(num_points >= 0 && num_reduced >= 0 && num_features > 0, "bad sizes");
  BEVB200_REQUIRE(reduce_type >= kReduceSum && reduce_type <= kReduceMax, "bad reduce type");
  if (num_points == 0) return BEVB200_OK;
  BEVB200_REQUIRE(grad_feats, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int c = num_features;
  const long long np = (long long)num_points * c, nr = (long long)num_reduced * c;
  if (reduce_type != kReduceMax) {
    BEVB200_REQUIRE(num_reduced == 0 || (grad_reduced_feats && reduce_count), "null argument");
    BEVB200_REQUIRE(coors_map, "null argument");
    BEVB200_LAUNCH(ds_bwd_add_kernel, grid_for(np, 256), 256, 0, st, grad_reduced_feats, coors_map,
                   reduce_count, np, c, reduce_type == kReduceMean, grad_feats);
    return BEVB200_OK;
  }
  BEVB200_CUDA(cudaMemsetAsync(grad_feats, 0, (size_t)np * sizeof(float), st));  // :257
  if (num_reduced == 0) return BEVB200_OK;
  BEVB200_REQUIRE(grad_reduced_feats && reduce_from, "null argument");
  if (!reduce_from_valid) {  // rebuild the arg-max by traceback (:287-299)
    BEVB200_REQUIRE(feats && reduced_feats && coors_map, "null argument");
    BEVB200_CUDA(cudaMemsetAsync(reduce_from, 0x7f, (size_t)nr * sizeof(int32_t), st));
    BEVB200_LAUNCH(ds_bwd_traceback_kernel, grid_for(np, 256), 256, 0, st, feats, reduced_feats,
                   coors_map, np, c, reduce_from);
  }
  BEVB200_LAUNCH(ds_bwd_max_kernel, grid_for(nr, 256), 256, 0, st, grad_reduced_feats, reduce_from, nr,
                 c, num_points, grad_feats);
  return BEVB200_OK;
}

}  // extern "C"
