// Sparse-conv rulebook for sm_100a (replaces spconv::getIndicePair<3>, spconv_ops.h:27-141,
// kernels indice.cu.h:22-203, geometry.h:24-85 `getValidOutPos`).
//
// The reference fills a dense int32 grid of the whole output volume (85 M ints = 340 MB at
// 1440x1440x41, re-allocated for every conv) and, for strided convs, sorts candidate outputs
// with torch::_unique.  Here the active OUTPUT sites live in a BITMAP of the dense output grid
// (1 bit per site: 10.6 MB at 1440x1440x41, L2 resident) with a popcount prefix per word:
//     row(site) = prefix[word] + popc(bits[word] & below(bit))
// which IS the ascending-flat-index order the reference's GPU path produces for strided convs
// (no sort), and a membership test + rank lookup for SubM (through rank2row, because SubM
// keeps the input row order).
//
// Output: offset-major neighbour table nbr[k, o] (input row or -1); converters to / from the
// reference's indicePairs[K,2,N] + indiceNum[K] layout are provided for the drop-in API.
#include "common.cuh"

namespace bevb200 {

struct ConvGeom {
  int in_shape[3], out_shape[3], ksize[3], stride[3], pad[3], dil[3];
  int batch, kvol;
};

__device__ __forceinline__ long long flat_site(int b, int x, int y, int z, const int shape[3]) {
  // tensorview.h:453-464 rowArrayIdx + batch * volume (indice.cu.h:59-60)
  return (((long long)b * shape[0] + x) * shape[1] + y) * shape[2] + z;
}

__device__ __forceinline__ int site_rank(const uint32_t *__restrict__ bits,
                                         const uint32_t *__restrict__ prefix, long long site) {
  uint32_t w = __ldg(bits + (site >> 5));
  uint32_t bit = 1u << (site & 31);
  if (!(w & bit)) return -1;
  return (int)(__ldg(prefix + (site >> 5)) + __popc(w & (bit - 1)));
}

// ---- SubM -----------------------------------------------------------------------------
__global__ void rb_mark_inputs_kernel(const int32_t *__restrict__ indices, int n, ConvGeom g,
                                      uint32_t *__restrict__ bits) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = *reinterpret_cast<const int4 *>(indices + 4ll * i);  // (b, x, y, z)
    if ((unsigned)c.x >= (unsigned)g.batch || (unsigned)c.y >= (unsigned)g.out_shape[0] ||
        (unsigned)c.z >= (unsigned)g.out_shape[1] || (unsigned)c.w >= (unsigned)g.out_shape[2])
      continue;
    long long s = flat_site(c.x, c.y, c.z, c.w, g.out_shape);
    atomicOr(bits + (s >> 5), 1u << (s & 31));
  }
}

__global__ void rb_rank2row_kernel(const int32_t *__restrict__ indices, int n, ConvGeom g,
                                   const uint32_t *__restrict__ bits,
                                   const uint32_t *__restrict__ prefix,
                                   int32_t *__restrict__ rank2row) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 c = *reinterpret_cast<const int4 *>(indices + 4ll * i);
    if ((unsigned)c.x >= (unsigned)g.batch || (unsigned)c.y >= (unsigned)g.out_shape[0] ||
        (unsigned)c.z >= (unsigned)g.out_shape[1] || (unsigned)c.w >= (unsigned)g.out_shape[2])
      continue;
    int r = site_rank(bits, prefix, flat_site(c.x, c.y, c.z, c.w, g.out_shape));
    if (r >= 0) rank2row[r] = i;  // duplicate coordinates: any one row wins (reference: last)
  }
}

// nbr[k, o] for SubM: input site = out site - pad + off*dil (stride 1, pad = k/2)
__global__ void __launch_bounds__(256)
    rb_subm_nbr_kernel(const int32_t *__restrict__ indices, int n, ConvGeom g,
                       const uint32_t *__restrict__ bits, const uint32_t *__restrict__ prefix,
                       const int32_t *__restrict__ rank2row, int32_t *__restrict__ nbr) {
  // grid (row tiles, kernel offsets): the offset is uniform per block, no per-element division
  const int k = blockIdx.y;
  const int kz = k % g.ksize[2], ky = (k / g.ksize[2]) % g.ksize[1], kx = k / (g.ksize[2] * g.ksize[1]);
  const int dx = kx * g.dil[0] - g.pad[0], dy = ky * g.dil[1] - g.pad[1], dz = kz * g.dil[2] - g.pad[2];
  int32_t *dst = nbr + (long long)k * n;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
    const int4 c = *reinterpret_cast<const int4 *>(indices + 4ll * o);
    const int x = c.y + dx, y = c.z + dy, z = c.w + dz;
    int row = -1;
    if ((unsigned)c.x < (unsigned)g.batch && (unsigned)x < (unsigned)g.out_shape[0] &&
        (unsigned)y < (unsigned)g.out_shape[1] && (unsigned)z < (unsigned)g.out_shape[2]) {
      int r = site_rank(bits, prefix, flat_site(c.x, x, y, z, g.out_shape));
      if (r >= 0) row = rank2row ? rank2row[r] : r;
    }
    dst[o] = row;
  }
}

// ---- strided (regular) sparse conv ----------------------------------------------------
// output site reached from input site q through kernel offset k: p = (q + pad - k*dil) / stride
// (exists iff divisible and inside the output grid) -- the set getValidOutPos enumerates.
// One thread per input voxel.  Per dimension only the kernel offsets k with
// (q + pad - k*dil) % stride == 0 reach an output site, so the thread walks the <= prod(ceil(K/s))
// valid (kx, ky, kz) combinations (8 of 27 for k3 s2) instead of testing all of them.
template <bool FILL>
__global__ void rb_conv_sites_kernel(const int32_t *__restrict__ indices, int n, ConvGeom g,
                                     uint32_t *__restrict__ bits,
                                     const uint32_t *__restrict__ prefix, int n_out,
                                     int32_t *__restrict__ nbr) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const int4 c = *reinterpret_cast<const int4 *>(indices + 4ll * j);
    if ((unsigned)c.x >= (unsigned)g.batch) continue;
    const int q[3] = {c.y, c.z, c.w};
    for (int kx = 0; kx < g.ksize[0]; ++kx) {
      int vx = q[0] + g.pad[0] - kx * g.dil[0];
      if (vx < 0 || vx % g.stride[0]) continue;
      vx /= g.stride[0];
      if (vx >= g.out_shape[0]) continue;
      for (int ky = 0; ky < g.ksize[1]; ++ky) {
        int vy = q[1] + g.pad[1] - ky * g.dil[1];
        if (vy < 0 || vy % g.stride[1]) continue;
        vy /= g.stride[1];
        if (vy >= g.out_shape[1]) continue;
        for (int kz = 0; kz < g.ksize[2]; ++kz) {
          int vz = q[2] + g.pad[2] - kz * g.dil[2];
          if (vz < 0 || vz % g.stride[2]) continue;
          vz /= g.stride[2];
          if (vz >= g.out_shape[2]) continue;
          const long long s = flat_site(c.x, vx, vy, vz, g.out_shape);
          if (FILL) {
            const int k = (kx * g.ksize[1] + ky) * g.ksize[2] + kz;
            const int o = site_rank(bits, prefix, s);
            if (o >= 0 && o < n_out) nbr[(long long)k * n_out + o] = j;
          } else {
            const uint32_t bit = 1u << (s & 31);
            if (!(bits[s >> 5] & bit)) atomicOr(bits + (s >> 5), bit);
          }
        }
      }
    }
  }
}

__global__ void rb_out_indices_kernel(const uint32_t *__restrict__ bits,
                                      const uint32_t *__restrict__ prefix, size_t nwords,
                                      ConvGeom g, int n_out, int32_t *__restrict__ out_indices) {
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < nwords;
       w += (size_t)gridDim.x * blockDim.x) {
    uint32_t m = bits[w];
    if (!m) continue;
    int r = (int)prefix[w];
    while (m) {
      int bit = __ffs(m) - 1;
      m &= m - 1;
      long long s = ((long long)w << 5) + bit;
      int z = (int)(s % g.out_shape[2]); s /= g.out_shape[2];
      int y = (int)(s % g.out_shape[1]); s /= g.out_shape[1];
      int x = (int)(s % g.out_shape[0]); s /= g.out_shape[0];
      if (r < n_out)
        *reinterpret_cast<int4 *>(out_indices + 4ll * r) = make_int4((int)s, x, y, z);
      ++r;
    }
  }
}

// ---- layout converters ----------------------------------------------------------------
// one CTA per kernel offset: ordered compaction of the valid (in, out) pairs
__global__ void __launch_bounds__(1024)
    rb_to_pairs_kernel(const int32_t *__restrict__ nbr, int n_out, int n_in,
                       int32_t *__restrict__ pairs, int32_t *__restrict__ num) {
  __shared__ int warp_cnt[32];
  __shared__ int carry_s;
  const int k = blockIdx.x, lane = lane_id(), warp = threadIdx.x >> 5;
  int32_t *pin = pairs + (2ll * k) * n_in, *pout = pairs + (2ll * k + 1) * n_in;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_out; base += blockDim.x) {
    int o = base + threadIdx.x;
    int j = o < n_out ? nbr[(long long)k * n_out + o] : -1;
    bool v = j >= 0;
    unsigned m = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int off = carry_s;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    int pos = off + __popc(m & ((1u << lane) - 1));
    if (v && pos < n_in) { pin[pos] = j; pout[pos] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += warp_cnt[w];
      carry_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) num[k] = min(carry_s, n_in);
}

__global__ void rb_pairs_to_nbr_kernel(const int32_t *__restrict__ pairs,
                                       const int32_t *__restrict__ num, int kvol, int pairs_dim,
                                       int n_out, int inverse, int32_t *__restrict__ nbr) {
  const long long total = (long long)kvol * pairs_dim;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    int k = (int)(t / pairs_dim), p = (int)(t % pairs_dim);
    if (p >= num[k]) continue;
    int a = pairs[(2ll * k) * pairs_dim + p], b = pairs[(2ll * k + 1) * pairs_dim + p];
    int in = inverse ? b : a, out = inverse ? a : b;
    if (in >= 0 && out >= 0 && out < n_out) nbr[(long long)k * n_out + out] = in;
  }
}

// ---- host side ------------------------------------------------------------------------
struct RbWs {
  uint32_t *bits, *prefix, *tiles, *total;
  int32_t *rank2row;
  size_t nwords;
};

static size_t rb_layout(int n_in, int batch, const int32_t *out_shape, void *ws, size_t ws_bytes,
                        RbWs *out) {
  Arena a(ws, ws_bytes);
  RbWs w;
  size_t sites = (size_t)batch * out_shape[0] * out_shape[1] * out_shape[2];
  w.nwords = (sites + 31) / 32;
  w.bits = a.take<uint32_t>(w.nwords);
  w.prefix = a.take<uint32_t>(w.nwords);
  w.tiles = a.take<uint32_t>(scan_scratch_elems(w.nwords));
  w.total = a.take<uint32_t>(64);
  w.rank2row = a.take<int32_t>(n_in > 0 ? n_in : 1);
  if (out) *out = w;
  return a.off;
}

static int make_geom(int batch, const int32_t *in_shape, const int32_t *out_shape,
                     const int32_t *ksize, const int32_t *stride, const int32_t *pad,
                     const int32_t *dil, int subm, ConvGeom *g) {
  g->batch = batch;
  g->kvol = 1;
  for (int d = 0; d < 3; ++d) {
    g->in_shape[d] = in_shape[d];
    g->out_shape[d] = out_shape[d];
    g->ksize[d] = ksize[d];
    g->dil[d] = dil[d];
    // spconv_ops.h:74-83: SubM forces stride 1 and padding k/2
    g->stride[d] = subm ? 1 : stride[d];
    g->pad[d] = subm ? ksize[d] / 2 : pad[d];
    if (ksize[d] <= 0 || g->stride[d] <= 0 || dil[d] <= 0 || out_shape[d] <= 0 || in_shape[d] <= 0 ||
        g->pad[d] < 0)
      return -1;
    g->kvol *= ksize[d];
  }
  return 0;
}

}  // namespace bevb200

using namespace bevb200;

#define RB_COMMON_CHECKS()                                                                     \
  BEVB200_REQUIRE(n_in >= 0 && batch_size > 0, "bad sizes");                                   \
  BEVB200_REQUIRE(spatial_shape_host && out_shape_host && ksize_host && stride_host &&         \
                      padding_host && dilation_host, "null argument");                         \
  ConvGeom g;                                                                                  \
  BEVB200_REQUIRE(make_geom(batch_size, spatial_shape_host, out_shape_host, ksize_host,        \
                            stride_host, padding_host, dilation_host, subm, &g) == 0,          \
                  "bad convolution geometry");                                                 \
  BEVB200_REQUIRE(g.kvol <= 4096, "kernel volume > 4096 (spconv_ops.h:50)");                   \
  if (subm)                                                                                    \
    for (int d = 0; d < 3; ++d)                                                                \
      BEVB200_REQUIRE(out_shape_host[d] == spatial_shape_host[d], "SubM keeps the spatial shape"); \
  RbWs w;                                                                                      \
  size_t need = rb_layout(n_in, batch_size, out_shape_host, workspace, workspace_bytes, &w);   \
  if (workspace == nullptr || workspace_bytes < need) {                                        \
    snprintf(g_last_error, sizeof(g_last_error), "rulebook: workspace too small (%zu < %zu)",  \
             workspace_bytes, need);                                                           \
    return BEVB200_EWORKSPACE;                                                                 \
  }                                                                                            \
  cudaStream_t st = (cudaStream_t)stream

extern "C" {

size_t bevb200_rulebook_workspace_bytes(int n_in, int batch_size, const int32_t *out_shape_host) {
  if (n_in < 0 || batch_size <= 0 || !out_shape_host) return 0;
  return rb_layout(n_in, batch_size, out_shape_host, nullptr, 0, nullptr);
}

int bevb200_rulebook_prepare(const int32_t *indices, int n_in, int batch_size,
                             const int32_t *spatial_shape_host, const int32_t *out_shape_host,
                             const int32_t *ksize_host, const int32_t *stride_host,
                             const int32_t *padding_host, const int32_t *dilation_host, int subm,
                             int32_t *n_out, void *workspace, size_t workspace_bytes,
                             void *stream) {
  RB_COMMON_CHECKS();
  BEVB200_REQUIRE(n_out != nullptr, "null n_out");
  BEVB200_CUDA(cudaMemsetAsync(w.bits, 0, w.nwords * sizeof(uint32_t), st));
  if (n_in == 0) {
    BEVB200_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st));
    BEVB200_CUDA(cudaMemsetAsync(w.prefix, 0, w.nwords * sizeof(uint32_t), st));
    return BEVB200_OK;
  }
  BEVB200_REQUIRE(indices != nullptr, "null indices");
  if (subm) {
    BEVB200_LAUNCH(rb_mark_inputs_kernel, grid_for(n_in, 256), 256, 0, st, indices, n_in, g, w.bits);
  } else {
    BEVB200_LAUNCH(rb_conv_sites_kernel<false>, grid_for(n_in, 128), 128, 0, st, indices, n_in, g, w.bits,
                   (const uint32_t *)nullptr, 0, (int32_t *)nullptr);
  }
  int rc = exclusive_scan_u32(w.bits, w.prefix, w.nwords, w.tiles, w.total, true, st);
  if (rc) return rc;
  if (subm) {
    BEVB200_CUDA(cudaMemsetAsync(w.rank2row, 0xff, (size_t)n_in * sizeof(int32_t), st));
    BEVB200_LAUNCH(rb_rank2row_kernel, grid_for(n_in, 256), 256, 0, st, indices, n_in, g, w.bits,
                   w.prefix, w.rank2row);
    // SubM: outputs are the inputs, in input order (spconv_ops.h:101)
    int32_t n32 = n_in;
    BEVB200_CUDA(cudaMemcpyAsync(n_out, &n32, sizeof(int32_t), cudaMemcpyHostToDevice, st));
  } else {
    BEVB200_CUDA(cudaMemcpyAsync(n_out, w.total, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return BEVB200_OK;
}

int bevb200_rulebook_fill(const int32_t *indices, int n_in, int batch_size,
                          const int32_t *spatial_shape_host, const int32_t *out_shape_host,
                          const int32_t *ksize_host, const int32_t *stride_host,
                          const int32_t *padding_host, const int32_t *dilation_host, int subm,
                          int n_out, int32_t *out_indices, int32_t *nbr, void *workspace,
                          size_t workspace_bytes, void *stream) {
  RB_COMMON_CHECKS();
  BEVB200_REQUIRE(n_out >= 0, "negative n_out");
  if (n_out == 0 || n_in == 0) return BEVB200_OK;
  BEVB200_REQUIRE(indices && nbr, "null argument");
  if (subm) {
    BEVB200_REQUIRE(n_out == n_in, "SubM: n_out must equal n_in");
    if (out_indices && out_indices != indices)
      BEVB200_CUDA(cudaMemcpyAsync(out_indices, indices, (size_t)n_in * 4 * sizeof(int32_t),
                                   cudaMemcpyDeviceToDevice, st));
    BEVB200_LAUNCH(rb_subm_nbr_kernel, dim3(grid_for(n_in, 256, kNumSMs * 2), g.kvol), 256, 0, st,
                   indices, n_in, g, w.bits, w.prefix, w.rank2row, nbr);
  } else {
    BEVB200_REQUIRE(out_indices != nullptr, "null out_indices");
    BEVB200_CUDA(cudaMemsetAsync(nbr, 0xff, (size_t)g.kvol * n_out * sizeof(int32_t), st));
    BEVB200_LAUNCH(rb_out_indices_kernel, grid_for((long long)w.nwords, 256), 256, 0, st, w.bits,
                   w.prefix, w.nwords, g, n_out, out_indices);
    BEVB200_LAUNCH(rb_conv_sites_kernel<true>, grid_for(n_in, 128), 128, 0, st, indices, n_in, g, w.bits,
                   (const uint32_t *)w.prefix, n_out, nbr);
  }
  return BEVB200_OK;
}

int bevb200_rulebook_fill_subm_sorted(const int32_t *indices, int n, int batch_size,
                                      const int32_t *spatial_shape_host, const int32_t *ksize_host,
                                      const int32_t *dilation_host, int32_t *nbr, void *workspace,
                                      size_t workspace_bytes, void *stream) {
  BEVB200_REQUIRE(n >= 0 && batch_size > 0, "bad sizes");
  BEVB200_REQUIRE(spatial_shape_host && ksize_host && dilation_host, "null argument");
  ConvGeom g;
  const int32_t ones[3] = {1, 1, 1}, zeros[3] = {0, 0, 0};
  BEVB200_REQUIRE(make_geom(batch_size, spatial_shape_host, spatial_shape_host, ksize_host, ones, zeros,
                            dilation_host, 1, &g) == 0, "bad convolution geometry");
  RbWs w;
  size_t need = rb_layout(0, batch_size, spatial_shape_host, workspace, workspace_bytes, &w);
  if (workspace == nullptr || workspace_bytes < need) {
    snprintf(g_last_error, sizeof(g_last_error), "rulebook: workspace too small (%zu < %zu)", workspace_bytes, need);
    return BEVB200_EWORKSPACE;
  }
  if (n == 0) return BEVB200_OK;
  BEVB200_REQUIRE(indices && nbr, "null argument");
  BEVB200_LAUNCH(rb_subm_nbr_kernel, dim3(grid_for(n, 256, kNumSMs * 2), g.kvol), 256, 0, (cudaStream_t)stream,
                 indices, n, g, w.bits, w.prefix, (const int32_t *)nullptr, nbr);
  return BEVB200_OK;
}

int bevb200_rulebook_to_pairs(const int32_t *nbr, int kernel_volume, int n_out, int n_in,
                              int32_t *indice_pairs, int32_t *indice_num, void *stream) {
  BEVB200_REQUIRE(kernel_volume > 0 && n_out >= 0 && n_in >= 0, "bad sizes");
  BEVB200_REQUIRE(indice_num != nullptr, "null indice_num");
  cudaStream_t st = (cudaStream_t)stream;
  BEVB200_CUDA(cudaMemsetAsync(indice_num, 0, (size_t)kernel_volume * sizeof(int32_t), st));
  if (n_in == 0) return BEVB200_OK;
  BEVB200_REQUIRE(indice_pairs != nullptr, "null indice_pairs");
  BEVB200_CUDA(cudaMemsetAsync(indice_pairs, 0xff, (size_t)kernel_volume * 2 * n_in * sizeof(int32_t), st));
  if (n_out == 0) return BEVB200_OK;
  BEVB200_REQUIRE(nbr != nullptr, "null nbr");
  BEVB200_LAUNCH(rb_to_pairs_kernel, kernel_volume, 1024, 0, st, nbr, n_out, n_in, indice_pairs,
                 indice_num);
  return BEVB200_OK;
}

int bevb200_pairs_to_nbr(const int32_t *indice_pairs, const int32_t *indice_num,
                         int kernel_volume, int pairs_dim, int n_out, int inverse, int32_t *nbr,
                         void *stream) {
  BEVB200_REQUIRE(kernel_volume > 0 && pairs_dim >= 0 && n_out >= 0, "bad sizes");
  if (n_out == 0) return BEVB200_OK;
  BEVB200_REQUIRE(nbr != nullptr, "null nbr");
  cudaStream_t st = (cudaStream_t)stream;
  BEVB200_CUDA(cudaMemsetAsync(nbr, 0xff, (size_t)kernel_volume * n_out * sizeof(int32_t), st));
  if (pairs_dim == 0) return BEVB200_OK;
  BEVB200_REQUIRE(indice_pairs && indice_num, "null argument");
  BEVB200_LAUNCH(rb_pairs_to_nbr_kernel, grid_for((long long)kernel_volume * pairs_dim, 256), 256, 0,
                 st, indice_pairs, indice_num, kernel_volume, pairs_dim, n_out, inverse, nbr);
  return BEVB200_OK;
}

}  // extern "C"
