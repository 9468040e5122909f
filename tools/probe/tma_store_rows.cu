// Is the TMA *tensor* store engine limited by rows per second?  Persistent CTAs, 8 issuing warps each, store the same
// shared-memory tile to successive positions of a [rows x pitch] uint16 matrix through a tiled tensor map, for several
// box shapes of equal or doubled size.  Reports GB/s and SM cycles per box row.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_store_rows tools/probe/tma_store_rows.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256, 1) store_kernel(const __grid_constant__ CUtensorMap tm, int box_cols, int box_rows,
                                                       int n_cols, int n_rows, int in_flight) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = box_cols * 2 * box_rows;
  uint8_t* mine = sm + warp * tile_bytes;
  for (int i = lane; i < tile_bytes / 4; i += 32) ((uint32_t*)mine)[i] = i * 2654435761u + warp;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long tiles_x = n_cols / box_cols, tiles_y = n_rows / box_rows, ntiles = tiles_x * tiles_y;
  const long long stride = (long long)gridDim.x * 8;
  if (lane == 0) {
    for (long long t = (long long)blockIdx.x * 8 + warp; t < ntiles; t += stride) {
      const int ty = (int)(t / tiles_x), tx = (int)(t % tiles_x);
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm),
                   "r"(smem_u32(mine)), "r"(tx * box_cols), "r"(ty * box_rows) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (in_flight == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (in_flight == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int n_cols = 4096, n_rows = 86 * 4096;  // the config-2 rank tensor
  const size_t bytes = (size_t)n_cols * n_rows * 2;
  void* buf;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  struct Cfg { int cols, rows; CUtensorMapSwizzle swz; const char* name; };
  Cfg cfgs[] = {{32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "64B x 32 rows (2 KB)"},
                {64, 32, CU_TENSOR_MAP_SWIZZLE_128B, "128B x 32 rows (4 KB)"},
                {64, 16, CU_TENSOR_MAP_SWIZZLE_128B, "128B x 16 rows (2 KB)"},
                {128, 16, CU_TENSOR_MAP_SWIZZLE_NONE, "256B x 16 rows (4 KB)"},
                {128, 8, CU_TENSOR_MAP_SWIZZLE_NONE, "256B x 8 rows (2 KB)"},
                {16, 64, CU_TENSOR_MAP_SWIZZLE_32B, "32B x 64 rows (2 KB)"}};
  for (const Cfg& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)n_cols, (cuuint64_t)n_rows};
    cuuint64_t strides[1] = {(cuuint64_t)n_cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     c.swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    for (int inflight : {1, 2, 4}) {
      const int smem = 8 * c.cols * 2 * c.rows;
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      for (int w = 0; w < 2; ++w) store_kernel<<<148, 256, smem>>>(tm, c.cols, c.rows, n_cols, n_rows, inflight);
      cudaDeviceSynchronize();
      cudaEventRecord(a);
      const int iters = 5;
      for (int i = 0; i < iters; ++i) store_kernel<<<148, 256, smem>>>(tm, c.cols, c.rows, n_cols, n_rows, inflight);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= iters;
      const double rows_per_sm = (double)n_rows * (n_cols / c.cols) / 148.0;
      printf("%-24s in-flight/warp %d : %.3f ms  %5.0f GB/s  %.2f SM-cycles per box row (at %.2f GHz nominal)\n", c.name,
             inflight, ms, bytes / ms / 1e6, ms * 1e-3 * clk_khz * 1e3 / rows_per_sm, clk_khz / 1e6);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
