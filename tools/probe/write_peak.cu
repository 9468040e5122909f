// Write-only HBM bandwidth probes (context for the rank kernel's roofline: its traffic is ~all writes, while the
// contract's HBM denominator is a copy).  Three persistent kernels over the same buffer:
//   bulk   : cp.async.bulk shared -> global, contiguous 16 KB pieces, up to 8 bulk groups in flight per CTA
//   tile   : cp.async.bulk of 32 separate 64-byte rows at a row pitch (the access shape of a 32x32 uint16 TMA tile store)
//   plain  : st.global.v4 streaming stores, 16 B per thread, fully coalesced
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/write_peak tools/probe/write_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) bulk_kernel(uint8_t* out, size_t bytes, int piece) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < piece / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t npieces = bytes / piece;
    for (size_t p = blockIdx.x; p < npieces; p += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + p * piece), "r"(smem_u32(sm)),
                   "r"(piece) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// 32 rows x 64 B at pitch `pitch` bytes per tile; tiles walk along the row first (like adjacent 32-column chunks)
__global__ void __launch_bounds__(128, 1) tile_kernel(uint8_t* out, size_t rows, size_t pitch) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const size_t tiles_per_row = pitch / 64, row_blocks = rows / 32, ntiles = tiles_per_row * row_blocks;
  if (threadIdx.x < 32) {
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const size_t rb = t / tiles_per_row, cb = t % tiles_per_row;
      uint8_t* dst = out + (rb * 32 + threadIdx.x) * pitch + cb * 64;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 64;" ::"l"(dst),
                   "r"(smem_u32(sm) + threadIdx.x * 64) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void __launch_bounds__(512) plain_kernel(uint4* out, size_t n16) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

template <typename F>
float time_ms(F f, int iters) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / iters;
}

int main() {
  const size_t rows = 86ull * 4096, pitch = 8192, bytes = rows * pitch;  // the config-2 rank tensor: 2.886 GB
  uint8_t* buf;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int piece : {2048, 4096, 16384, 65536}) {
    float ms = time_ms([&] { bulk_kernel<<<sms, 128, piece>>>(buf, bytes, piece); }, 10);
    printf("bulk  piece=%6d B : %.3f ms  %.0f GB/s\n", piece, ms, bytes / ms / 1e6);
  }
  for (int ctas : {sms, 2 * sms, 4 * sms}) {
    float ms = time_ms([&] { bulk_kernel<<<ctas, 128, 16384>>>(buf, bytes, 16384); }, 10);
    printf("bulk  16 KB, %d CTAs : %.3f ms  %.0f GB/s\n", ctas, ms, bytes / ms / 1e6);
  }
  {
    float ms = time_ms([&] { tile_kernel<<<sms, 128, 2048>>>(buf, rows, pitch); }, 10);
    printf("tile  32 rows x 64 B (pitch 8 KB), %d CTAs : %.3f ms  %.0f GB/s\n", sms, ms, bytes / ms / 1e6);
    ms = time_ms([&] { tile_kernel<<<4 * sms, 128, 2048>>>(buf, rows, pitch); }, 10);
    printf("tile  32 rows x 64 B (pitch 8 KB), %d CTAs : %.3f ms  %.0f GB/s\n", 4 * sms, ms, bytes / ms / 1e6);
  }
  {
    float ms = time_ms([&] { plain_kernel<<<sms * 4, 512>>>((uint4*)buf, bytes / 16); }, 10);
    printf("plain st.global.v4 : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    ms = time_ms([&] { cudaMemsetAsync(buf, 0, bytes); }, 10);
    printf("cudaMemset : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
