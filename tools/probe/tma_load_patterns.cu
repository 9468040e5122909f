// How fast does one SM's TMA unit stream [128 rows x 64 k] bf16 panels (16 KB, 128B swizzle) of an L2-resident weight
// matrix [rows x K] when (a) every CTA picks its own pseudo-random panels, (b) ALL CTAs walk the same panels in the same
// order (what the encoder's nn.Linear GEMMs do with their weights), (c) the same walk rotated per CTA?  `depth` loads in
// flight per CTA.  Question behind it: in_proj (K = 512, N = 1536) receives its weights at 26 B/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_load_patterns tools/probe/tma_load_patterns.cu -lcuda
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

constexpr int kMaxDepth = 8;
__global__ void __launch_bounds__(64, 1) load_kernel(const __grid_constant__ CUtensorMap tm, int pattern, int depth, int iters,
                                                     int n_rowblk, int n_kpanel) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bars[kMaxDepth];
  if (threadIdx.x == 0) for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    // no divisions in the issue loop (a single thread issues every load: index arithmetic must not be the bound)
    const int total = n_rowblk * n_kpanel;
    int idx = pattern == 0 ? (blockIdx.x * 101) % total : (pattern == 1 ? 0 : (blockIdx.x * n_kpanel) % total);
    int rb = idx / n_kpanel, kp = idx - rb * n_kpanel;
    const int step = pattern == 0 ? 37 % total : 1;
    const int step_rb = step / n_kpanel, step_kp = step - step_rb * n_kpanel;
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < iters + depth; ++i) {
      if (i >= depth) mbar_wait(smem_u32(&bars[s]), ph ^ 1);
      if (i < iters) {
        mbar_expect(smem_u32(&bars[s]), 16384);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sm + s * 16384)), "l"((uint64_t)&tm), "r"(smem_u32(&bars[s])), "r"(kp * 64), "r"(rb * 128) : "memory");
        kp += step_kp; rb += step_rb;
        if (kp >= n_kpanel) { kp -= n_kpanel; ++rb; }
        if (rb >= n_rowblk) rb -= n_rowblk;
      }
      if (++s == depth) { s = 0; ph ^= 1; }
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
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDepth * 16384);
  struct Shape { int rows, K; const char* name; };
  const Shape shapes[] = {{1536, 512, "in_proj weights [1536 x 512]"}, {4096, 256, "z_cols [4096 x 256]"}, {512, 512, "out_proj weights [512 x 512]"}};
  const char* pnames[3] = {"own random walk", "all CTAs in lockstep", "lockstep, rotated per CTA"};
  for (const Shape& sh : shapes) {
    void* buf; cudaMalloc(&buf, (size_t)sh.rows * sh.K * 2); cudaMemset(buf, 0, (size_t)sh.rows * sh.K * 2);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)sh.K, (cuuint64_t)sh.rows}; cuuint64_t st[1] = {(cuuint64_t)sh.K * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int pattern = 0; pattern < 3; ++pattern)
      for (int depth : {3, 4, 8}) {
        const int iters = 20000;
        load_kernel<<<148, 64, depth * 16384>>>(tm, pattern, depth, 2000, sh.rows / 128, sh.K / 64);
        cudaDeviceSynchronize();
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        load_kernel<<<148, 64, depth * 16384>>>(tm, pattern, depth, iters, sh.rows / 128, sh.K / 64);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double cyc = ms * 1e-3 * clk_khz * 1e3;
        printf("%-30s %-26s depth %d : %.1f B/clk/SM (%.2f TB/s chip-wide)\n", sh.name, pnames[pattern], depth,
               (double)iters * 16384 / cyc, (double)iters * 16384 * 148 / ms / 1e9);
      }
    cudaFree(buf);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
