// Do tcgen05.mma operand reads from shared memory (SS mode) take bandwidth away from the LSU's shared-memory loads?
// The rank epilogue is bound by its exact-LUT gathers (ncu: LSU wavefront pipe 77 % busy) while the tensor pipe is 27 %
// busy reading A and B panels from shared memory at 128 B per busy cycle.  One CTA per SM:
//   warp 0        : back-to-back 128x128x16 bf16 MMAs (A, B from shared memory; or A from TMEM with mode bit 4)
//   warps 4..11   : the epilogue's look-up pattern: random 4-byte loads from an 8192-entry table, ~9 ALU/FMA
//                   instructions per load
// modes: 1 = MMA only, 2 = gathers only, 3 = both, 5 / 7 = the same with the A operand in TMEM.
// `duty` throttles the MMA warp (percent of time it keeps the tensor pipe fed) to mimic the kernel's 27 %.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/smem_contention tools/probe/smem_contention.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../../madrigal_b200/csrc/mdg_ptx.cuh"

using namespace mdg;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, 1, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc) : "memory");
}

constexpr int kGatherWarps = 8;
__global__ void __launch_bounds__(384, 1) probe(int mode, int mma_groups, int gather_iters, int idle_cycles,
                                                unsigned long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 8 * 16384, sLut = sB + 2 * 16384, sBar = sLut + 32768;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (10 * 16384 + 32768) / 4; i += blockDim.x)
    st_shared_u32(base + 4 * i, i < 10 * 4096 ? 0x3F803F80u : (i * 2654435761u) >> 8);
  if (threadIdx.x == 0) {
    mbar_init(sBar, 1);
    mbar_init(sBar + 8, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  if (warp == 2) tmem_alloc(sBar + 32, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sBar + 32));
  const long long t0 = clock64();
  if (warp == 0 && (mode & 1)) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, 128);
    uint32_t ph[2] = {0, 0};
    for (int g = 0; g < mma_groups; ++g) {
      // one "B panel" of the real kernel: 4 K steps x 2 row sub-tiles = 8 MMAs (8 x 64 tensor-pipe cycles)
      if (g >= 2) {
        mbar_wait(sBar + 8 * (g & 1), ph[g & 1], 1);
        ph[g & 1] ^= 1;
      }
      if (elect_one()) {
        const uint64_t bdesc = umma_desc_kmajor_sw128(sB + (g & 1) * 16384);
        const uint64_t a0 = umma_desc_kmajor_sw128(sA + (g & 3) * 16384), a1 = umma_desc_kmajor_sw128(sA + (4 + (g & 3)) * 16384);
        for (int k = 0; k < 4; ++k) {
          if (mode & 4) umma_bf16_ts(tmem_base, tmem_base + 256 + 8 * k, bdesc + 2 * k, idesc);
          else umma_bf16(tmem_base, a0 + 2 * k, bdesc + 2 * k, idesc, 1);
        }
        for (int k = 0; k < 4; ++k) {
          if (mode & 4) umma_bf16_ts(tmem_base + 128, tmem_base + 384 + 8 * k, bdesc + 2 * k, idesc);
          else umma_bf16(tmem_base + 128, a1 + 2 * k, bdesc + 2 * k, idesc, 1);
        }
        umma_commit(sBar + 8 * (g & 1));
      }
      __syncwarp();
      if (idle_cycles > 0) {
        const long long t = clock64();
        while (clock64() - t < idle_cycles) {}
      }
    }
    for (int g = mma_groups; g < mma_groups + 2; ++g)
      if (g >= 2) {
        mbar_wait(sBar + 8 * (g & 1), ph[g & 1], 2);
        ph[g & 1] ^= 1;
      }
    if (lane == 0) out[blockIdx.x * 16 + 0] = clock64() - t0;
  } else if (warp >= 4 && warp < 4 + kGatherWarps && (mode & 2)) {
    uint32_t x = (threadIdx.x + 1) * 2654435761u + blockIdx.x * 40503u;
    uint32_t acc = 0;
    const float scale = 1.0f / 4294967296.0f;
    for (int it = 0; it < gather_iters; ++it) {
      uint32_t a[8], v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // same instruction mix as the epilogue: FFMA.SAT, FMUL, IMAD.HI, LDS, LOP3, SHF, POPC, IADD
        x = x * 1664525u + 1013904223u;
        const float y = __saturatef(fmaf(static_cast<float>(x >> 8), scale * 256.f, 0.0f));
        const uint32_t kb = __float_as_uint(__fmul_rn(y, __uint_as_float(131071u)));
        a[j] = kb;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t bucket;
        asm("mul.hi.u32 %0, %1, %2;" : "=r"(bucket) : "r"(a[j]), "r"(1u << 28));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[j]) : "r"(sLut + bucket * 4));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += v[j] + __popc(v[j] >> ((~a[j] & 15u) | 16u));
    }
    if (lane == 0) {
      out[blockIdx.x * 16 + 1 + (warp - 4)] = clock64() - t0;
      out[blockIdx.x * 16 + 9] = acc;
    }
  }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

int main(int argc, char** argv) {
  const int groups = 20000, iters = 20000;
  unsigned long long* d;
  cudaMalloc(&d, 148 * 16 * 8);
  const int smem = 10 * 16384 + 32768 + 1024 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<unsigned long long> h(148 * 16);
  struct Case { int mode, idle; const char* name; };
  const Case cases[] = {
      {1, 0, "MMA only, SS (A,B in smem), full rate"},
      {2, 0, "gathers only"},
      {3, 0, "both, MMA full rate"},
      {3, 1000, "both, MMA ~1/3 duty (idle 1000 cyc per 512-cycle group)"},
      {3, 1400, "both, MMA ~27 % duty (idle 1400)"},
      {5, 0, "MMA only, TS (A in TMEM)"},
      {7, 0, "both, TS, full rate"},
      {7, 1400, "both, TS, ~27 % duty"},
      {1, 1400, "MMA only, SS, ~27 % duty"},
  };
  for (const Case& c : cases) {
    cudaMemset(d, 0, 148 * 16 * 8);
    const int g = c.idle ? groups / 4 : groups;
    probe<<<148, 384, smem>>>(c.mode, g, iters, c.idle, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, 148 * 16 * 8, cudaMemcpyDeviceToHost);
    double mma = 0, gat = 0;
    for (int b = 0; b < 148; ++b) {
      mma += h[b * 16];
      double w = 0;
      for (int k = 0; k < kGatherWarps; ++k) w = w > h[b * 16 + 1 + k] ? w : (double)h[b * 16 + 1 + k];
      gat += w;
    }
    mma /= 148; gat /= 148;
    printf("%-58s:", c.name);
    if (c.mode & 1) printf("  MMA %.1f cyc per 128x128x16 (incl. idle)", mma / (g * 8.0));
    if (c.mode & 2) printf("  gathers: %.2f cyc per warp-wide LDS (8 warps; %.0f cycles total)", gat / (iters * 8.0 * kGatherWarps), gat);
    printf("\n");
  }
  return 0;
}
