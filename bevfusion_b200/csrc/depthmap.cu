// LiDAR points -> per-camera sparse depth images for sm_100a: the per-frame python loop of
// BaseDepthTransform.forward (mmdet3d/models/vtransforms/base.py:279-329) as two kernels.
//
// The reference runs, per sample, ~25 small torch ops plus a python loop over cameras whose
// `depth[b, c, 0, rows, cols] = dist` index_put leaves colliding pixels to whichever thread writes
// last.  Here:
//   K1 depth_project   one thread per (point, camera): undo the lidar augmentation, project with
//                      lidar2image, clamp z to [1e-5, 1e5], divide, apply the image augmentation
//                      (the exact op order of :289-304, fp32, no FMA contraction so the numpy oracle
//                      reproduces it bit for bit); on-image points take the pixel with
//                      atomicMax(point index) -- "the last point wins", the sequential meaning of the
//                      index_put -- or, for one-hot depth, set their (bin, row, col) cell.
//   K2 depth_resolve   one thread per (camera, pixel): writes the winner's distance (and, with
//                      add_depth_features, its feature row) or zero: the output is written once,
//                      coalesced, with no separate zero fill.
#include "common.cuh"

namespace bevb200 {

struct DepthCams {
  // per camera: lidar2image R (9) t (3), img_aug R (9) t (3)
  const float *lidar2image, *img_aug;  // [ncam, 4, 4] row-major, device
  const float *lidar_aug;              // [4, 4] row-major, device
};

struct Mat34 {
  float r[9], t[3];
};

__device__ __forceinline__ float dot3(const float *m, float x, float y, float z) {
  // explicit fp32 multiply / add, left to right: no FMA contraction
  return __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z));
}

// inverse of the 3x3 lidar augmentation by the adjugate (base.py:291 torch.inverse)
__device__ void inverse3(const float *a, float *inv) {
  const float c00 = __fsub_rn(__fmul_rn(a[4], a[8]), __fmul_rn(a[5], a[7]));
  const float c01 = __fsub_rn(__fmul_rn(a[3], a[8]), __fmul_rn(a[5], a[6]));
  const float c02 = __fsub_rn(__fmul_rn(a[3], a[7]), __fmul_rn(a[4], a[6]));
  const float det = __fadd_rn(__fsub_rn(__fmul_rn(a[0], c00), __fmul_rn(a[1], c01)), __fmul_rn(a[2], c02));
  inv[0] = __fdiv_rn(c00, det);
  inv[1] = __fdiv_rn(__fsub_rn(__fmul_rn(a[2], a[7]), __fmul_rn(a[1], a[8])), det);
  inv[2] = __fdiv_rn(__fsub_rn(__fmul_rn(a[1], a[5]), __fmul_rn(a[2], a[4])), det);
  inv[3] = __fdiv_rn(-c01, det);
  inv[4] = __fdiv_rn(__fsub_rn(__fmul_rn(a[0], a[8]), __fmul_rn(a[2], a[6])), det);
  inv[5] = __fdiv_rn(__fsub_rn(__fmul_rn(a[2], a[3]), __fmul_rn(a[0], a[5])), det);
  inv[6] = __fdiv_rn(c02, det);
  inv[7] = __fdiv_rn(__fsub_rn(__fmul_rn(a[1], a[6]), __fmul_rn(a[0], a[7])), det);
  inv[8] = __fdiv_rn(__fsub_rn(__fmul_rn(a[0], a[4]), __fmul_rn(a[1], a[3])), det);
}

struct DepthSmem {
  float aug_inv[9], aug_t[3];
  Mat34 l2i[16], ia[16];
};

__device__ void load_cams(DepthSmem &sm, const DepthCams &cams, int ncam) {
  if (threadIdx.x == 0) {
    float a[9];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) a[3 * r + c] = cams.lidar_aug[4 * r + c];
      sm.aug_t[r] = cams.lidar_aug[4 * r + 3];
    }
    inverse3(a, sm.aug_inv);
  }
  for (int t = threadIdx.x; t < ncam * 12; t += blockDim.x) {
    const int cam = t / 12, e = t % 12, r = e / 4, c = e % 4;
    const float v0 = cams.lidar2image[16 * cam + 4 * r + c], v1 = cams.img_aug[16 * cam + 4 * r + c];
    if (c < 3) { sm.l2i[cam].r[3 * r + c] = v0; sm.ia[cam].r[3 * r + c] = v1; }
    else       { sm.l2i[cam].t[r] = v0;         sm.ia[cam].t[r] = v1; }
  }
  __syncthreads();
}

// base.py:289-313 for one (point, camera).  Returns true when the point lands on the image.
__device__ __forceinline__ bool project(const DepthSmem &sm, int cam, float px, float py, float pz, int H,
                                        int W, int &row, int &col, float &dist) {
  const float x1 = __fsub_rn(px, sm.aug_t[0]), y1 = __fsub_rn(py, sm.aug_t[1]), z1 = __fsub_rn(pz, sm.aug_t[2]);
  const float x2 = dot3(sm.aug_inv + 0, x1, y1, z1), y2 = dot3(sm.aug_inv + 3, x1, y1, z1),
              z2 = dot3(sm.aug_inv + 6, x1, y1, z1);
  const Mat34 &L = sm.l2i[cam];
  float x3 = __fadd_rn(dot3(L.r + 0, x2, y2, z2), L.t[0]);
  float y3 = __fadd_rn(dot3(L.r + 3, x2, y2, z2), L.t[1]);
  float z3 = __fadd_rn(dot3(L.r + 6, x2, y2, z2), L.t[2]);
  if (!(z3 == z3)) return false;       // torch.clamp keeps NaN, which then fails the on-image test
  z3 = fminf(fmaxf(z3, 1e-5f), 1e5f);  // torch.clamp; `dist` aliases the clamped row (:298-299)
  x3 = __fdiv_rn(x3, z3);
  y3 = __fdiv_rn(y3, z3);
  const Mat34 &A = sm.ia[cam];
  const float u = __fadd_rn(dot3(A.r + 0, x3, y3, z3), A.t[0]);  // image x -> column
  const float v = __fadd_rn(dot3(A.r + 3, x3, y3, z3), A.t[1]);  // image y -> row
  if (!(v < (float)H && v >= 0.f && u < (float)W && u >= 0.f)) return false;  // :309-314
  row = (int)v;  // .long(): truncation of a non-negative float
  col = (int)u;
  dist = z3;
  return true;
}

__global__ void __launch_bounds__(256)
    depth_project_kernel(const float *__restrict__ points, int n, int nf, DepthCams cams, int ncam, int H,
                         int W, int one_hot, int D, int channels, int32_t *__restrict__ winner,
                         float *__restrict__ depth) {
  __shared__ DepthSmem sm;
  load_cams(sm, cams, ncam);
  const long long total = (long long)n * ncam;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cam = (int)(t / n), i = (int)(t % n);  // point fastest: coalesced point reads
    const float *p = points + (long long)i * nf;
    int row, col;
    float dist;
    if (!project(sm, cam, p[0], p[1], p[2], H, W, row, col, dist)) continue;
    const long long pix = (long long)row * W + col;
    if (winner) atomicMax(winner + (long long)cam * H * W + pix, i);
    if (one_hot) {  // :321-325 clamp(max=D-1).long() picks the bin
      const int bin = (int)fminf(dist, (float)(D - 1));
      depth[((long long)cam * channels + bin) * H * W + pix] = 1.0f;
    }
  }
}

__global__ void __launch_bounds__(256)
    depth_resolve_kernel(const float *__restrict__ points, int nf, DepthCams cams, int ncam, int H, int W,
                         int scalar, int feat_channels, int channels,
                         const int32_t *__restrict__ winner, float *__restrict__ depth) {
  __shared__ DepthSmem sm;
  load_cams(sm, cams, ncam);
  const long long hw = (long long)H * W, total = hw * ncam;
  const int first_feat = channels - feat_channels;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cam = (int)(t / hw);
    const long long pix = t % hw;
    const int w = winner[t];
    float *out = depth + (long long)cam * channels * hw + pix;
    const float *p = points + (long long)(w < 0 ? 0 : w) * nf;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (w >= 0) { px = p[0]; py = p[1]; pz = p[2]; }
    if (scalar) {
      float dist = 0.f;
      int row, col;
      if (w >= 0) project(sm, cam, px, py, pz, H, W, row, col, dist);
      out[0] = w >= 0 ? dist : 0.f;
    }
    for (int k = 0; k < feat_channels; ++k) {
      // :329 writes points[b][idx].T; the reference's `cur_coords -= trans` (:290) has already
      // shifted the xyz columns of points[b] in place, so those carry the un-translated xyz
      float v = 0.f;
      if (w >= 0) v = k < 3 ? __fsub_rn(p[k], sm.aug_t[k]) : p[k];
      out[(long long)(first_feat + k) * hw] = v;
    }
  }
}

}  // namespace bevb200

using namespace bevb200;

extern "C" {

size_t bevb200_depth_rasterize_workspace_bytes(int ncam, int height, int width) {
  if (ncam <= 0 || height <= 0 || width <= 0) return 0;
  return align_up((size_t)ncam * height * width * sizeof(int32_t));
}

int bevb200_depth_rasterize(const float *points, int num_points, int num_features,
                            const float *lidar_aug_matrix, const float *lidar2image,
                            const float *img_aug_matrix, int ncam, int height, int width, int one_hot,
                            int depth_bins, int add_features, float *depth, void *workspace,
                            size_t workspace_bytes, void *stream) {
  BEVB200_REQUIRE(num_points >= 0 && num_features >= 3, "bad point tensor shape");
  BEVB200_REQUIRE(ncam > 0 && ncam <= 16 && height > 0 && width > 0, "bad camera / image size");
  BEVB200_REQUIRE(!one_hot || depth_bins > 0, "one-hot depth needs the number of bins");
  BEVB200_REQUIRE(lidar_aug_matrix && lidar2image && img_aug_matrix && depth, "null argument");
  BEVB200_REQUIRE(num_points == 0 || points, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int feat = add_features ? num_features : 0;
  const int channels = (one_hot ? depth_bins : 1) + feat;
  const long long hw = (long long)height * width;
  const bool need_winner = !one_hot || feat > 0;
  int32_t *winner = nullptr;
  if (need_winner) {
    if (workspace == nullptr || workspace_bytes < bevb200_depth_rasterize_workspace_bytes(ncam, height, width)) {
      snprintf(g_last_error, sizeof(g_last_error), "depth_rasterize: workspace too small");
      return BEVB200_EWORKSPACE;
    }
    winner = (int32_t *)workspace;
    BEVB200_CUDA(cudaMemsetAsync(winner, 0xff, (size_t)ncam * hw * sizeof(int32_t), st));  // -1
  }
  if (one_hot)  // the bin planes are written sparsely
    BEVB200_CUDA(cudaMemsetAsync(depth, 0, (size_t)ncam * channels * hw * sizeof(float), st));
  DepthCams cams{lidar2image, img_aug_matrix, lidar_aug_matrix};
  if (num_points > 0)
    BEVB200_LAUNCH(depth_project_kernel, grid_for((long long)num_points * ncam, 256), 256, 0, st, points,
                   num_points, num_features, cams, ncam, height, width, one_hot, depth_bins, channels,
                   winner, depth);
  if (need_winner)
    BEVB200_LAUNCH(depth_resolve_kernel, grid_for(hw * ncam, 256), 256, 0, st, points, num_features, cams,
                   ncam, height, width, one_hot ? 0 : 1, feat, channels, winner, depth);
  return BEVB200_OK;
}

}  // extern "C"
