// How many bytes per SM-cycle does one SM's TMA unit move?  L2-resident traffic only (no HBM in the way):
//   loads : warp 0 keeps `depth` 16 KB tensor loads ([128 rows x 64 bf16], 128B swizzle) in flight from a 2 MB matrix
//   stores: warps 1..8 each store a 2 KB tile ([32 rows x 32 uint16], 64B swizzle) to a per-CTA 512 KB region, 2 in flight
//   both  : at the same time
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_unit_rate tools/probe/tma_unit_rate.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.b32 %0, 1, 0, P; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
}

constexpr int kDepth = 4;
__global__ void __launch_bounds__(288, 1) rate_kernel(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmS,
                                                      int do_loads, int do_stores, int iters, int load_rows, int store_rows_per_cta) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = sm;                                  // kDepth x 16 KB
  uint8_t* tiles = sm + kDepth * 16384;                // 8 x 2 x 2 KB
  __shared__ __align__(8) unsigned long long bars[kDepth];
  if (threadIdx.x == 0) for (int i = 0; i < kDepth; ++i) mbar_init(smem_u32(&bars[i]), 1);
  for (int i = threadIdx.x; i < 8 * 4096 / 4; i += blockDim.x) ((uint32_t*)tiles)[i] = i;
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0 && do_loads) {
    if (lane == 0) {
      const int nblk = load_rows / 128;
      for (int i = 0; i < iters + kDepth; ++i) {
        const int s = i % kDepth;
        if (i >= kDepth) mbar_wait(smem_u32(&bars[s]), ((i / kDepth) - 1) & 1);
        if (i < iters) {
          mbar_expect(smem_u32(&bars[s]), 16384);
          const int r = ((i * 37 + blockIdx.x * 11) % nblk) * 128, k = ((i + blockIdx.x) & 3) * 64;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(ring + s * 16384)), "l"((uint64_t)&tmL), "r"(smem_u32(&bars[s])), "r"(k), "r"(r) : "memory");
        }
      }
    }
  } else if (warp >= 1 && do_stores) {
    if (lane == 0) {
      const int w = warp - 1;
      const int ntile = store_rows_per_cta / 32 * 4;   // region: store_rows_per_cta rows x 128 uint16 (4 tiles wide)
      for (int i = 0; i < iters * 8 / 8; ++i) {        // each store warp issues `iters` stores of 2 KB: 8 warps -> iters * 16 KB
        const int t = (i * 8 + w) % ntile;
        const int row = blockIdx.x * store_rows_per_cta + (t / 4) * 32, col = (t % 4) * 32;
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tmS),
                     "r"(smem_u32(tiles + w * 4096 + (i & 1) * 2048)), "r"(col), "r"(row) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const int load_rows = 4096, load_cols = 256;                       // 2 MB bf16 matrix (like z_cols): L2-resident
  const int store_rows_per_cta = 2048, store_cols = 128;             // 512 KB per CTA, 74 MB in total: L2-resident
  void *lbuf, *sbuf;
  cudaMalloc(&lbuf, (size_t)load_rows * load_cols * 2);
  cudaMalloc(&sbuf, (size_t)148 * store_rows_per_cta * store_cols * 2);
  cudaMemset(lbuf, 0, (size_t)load_rows * load_cols * 2);
  CUtensorMap tmL, tmS;
  {
    cuuint64_t dims[2] = {(cuuint64_t)load_cols, (cuuint64_t)load_rows}; cuuint64_t st[1] = {(cuuint64_t)load_cols * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    enc(&tmL, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lbuf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)store_cols, (cuuint64_t)148 * store_rows_per_cta}; cuuint64_t st[1] = {(cuuint64_t)store_cols * 2};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    enc(&tmS, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, sbuf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  const int smem = kDepth * 16384 + 8 * 4096;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  const char* names[3] = {"loads only ", "stores only", "loads+stores"};
  for (int mode = 0; mode < 3; ++mode) {
    const int dl = mode != 1, ds = mode != 0;
    rate_kernel<<<148, 288, smem>>>(tmL, tmS, dl, ds, 2000, load_rows, store_rows_per_cta);
    cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    rate_kernel<<<148, 288, smem>>>(tmL, tmS, dl, ds, iters, load_rows, store_rows_per_cta);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lb = dl ? (double)iters * 16384 : 0, sb = ds ? (double)iters * 8 * 2048 : 0;  // per SM
    const double cyc = ms * 1e-3 * clk_khz * 1e3;
    printf("%s: %.3f ms  loads %.1f B/clk/SM  stores %.1f B/clk/SM  total %.1f B/clk/SM (%.2f TB/s chip-wide) at %.2f GHz nominal\n",
           names[mode], ms, lb / cyc, sb / cyc, (lb + sb) / cyc, (lb + sb) * 148 / ms / 1e9, clk_khz / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
