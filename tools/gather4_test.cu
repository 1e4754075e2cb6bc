// Dev probe: does cp.async.bulk.tensor.2d.tile::gather4 work with a plain tiled tensor map, and
// with which box shape?  Builds a [rows, cols] fp32 matrix, gathers 4 rows by index, checks them.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int cols, int r0, int r1, int r2, int r3,
                      float *out, int *status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(4 * cols * 4) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(&tm), "r"(bar_a), "r"(0), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  // bounded wait (so a wrong descriptor cannot hang the GPU)
  int ok = 0;
  for (int spin = 0; spin < 2000000 && !ok; ++spin) {
    uint32_t p;
    asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0; selp.b32 %0, 1, 0, P; }"
                 : "=r"(p) : "r"(bar_a) : "memory");
    ok = p;
  }
  if (threadIdx.x == 0) *status = ok;
  __syncthreads();
  if (ok)
    for (int i = threadIdx.x; i < 4 * cols; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main(int argc, char **argv) {
  int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  int cols = argc > 2 ? atoi(argv[2]) : 80;
  const int rows = 1000;
  float *h = (float *)malloc(sizeof(float) * rows * cols);
  for (int i = 0; i < rows * cols; ++i) h[i] = (float)i;
  float *d, *out; int *status;
  cudaMalloc(&d, sizeof(float) * rows * cols); cudaMalloc(&out, sizeof(float) * 4 * cols); cudaMalloc(&status, 4);
  cudaMemcpy(d, h, sizeof(float) * rows * cols, cudaMemcpyHostToDevice);
  cudaMemset(out, 0, sizeof(float) * 4 * cols);
  void *fn = nullptr; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
  CUtensorMap tm;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box_rows=%d cols=%d -> %d\n", box_rows, cols, (int)r);
  if (r != CUDA_SUCCESS) return 3;
  int idx[4] = {7, 123, 999, 42};
  probe<<<1, 128, 4 * cols * 4 + 1024>>>(tm, cols, idx[0], idx[1], idx[2], idx[3], out, status);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel -> %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 4;
  int st; cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
  float *ho = (float *)malloc(sizeof(float) * 4 * cols);
  cudaMemcpy(ho, out, sizeof(float) * 4 * cols, cudaMemcpyDeviceToHost);
  int good = st;
  for (int j = 0; j < 4 && good; ++j)
    for (int c = 0; c < cols; ++c)
      if (ho[j * cols + c] != (float)(idx[j] * cols + c)) { good = 0; printf("mismatch row %d col %d: %f\n", j, c, ho[j * cols + c]); break; }
  printf("barrier_completed=%d data_ok=%d\n", st, good);
  // OOB row index -> zero fill?
  probe<<<1, 128, 4 * cols * 4 + 1024>>>(tm, cols, 3, 5000, -1, 9, out, status);
  e = cudaDeviceSynchronize();
  cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ho, out, sizeof(float) * 4 * cols, cudaMemcpyDeviceToHost);
  printf("oob probe: %s completed=%d row1[0]=%f row2[0]=%f row3[0]=%f\n", cudaGetErrorString(e), st, ho[cols], ho[2 * cols], ho[3 * cols]);
  return 0;
}
