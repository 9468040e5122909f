// Per-outcome reference-quantile rank lookup: the device function shared by the fused decoder epilogue, the
// stand-alone lookup kernel and the table builder, plus the builder itself.
//
// Semantics to reproduce (oracle/oracle.py: quantile_rank): rank = np.searchsorted(thresholds[l], x, side='right')
// = #{ i : thresholds[l, i] <= x }.   In the reference the distribution each score is ranked against is the
// strict-lower-triangle score set of the same outcome (notebooks/normalize_scores.py:36-60, 67); a Q-point quantile
// table of that set bounds |rank/Q - reference normalised rank| by 1/Q ("rank semantics").
//
// How the lookup is made O(1) and still exact:
//   cell(x) = round(sat(x*scale + bias) * (2^17 - 1)): a 17-bit fixed-point image of x (8192 buckets x 16 sub-cells)
//             under a per-outcome affine map, computed with two FFMAs; monotone non-decreasing in x because each
//             fp32 rounding is monotone;
//   every threshold is SNAPPED by the builder to   t_i = min{ fp32 f : cell(f) >= c_i }   for a strictly increasing
//   cell sequence c_i, so that   x >= t_i  <=>  cell(x) >= c_i   (=> by monotonicity, <= by minimality);
//   hence #{t_i <= x} = #{c_i <= cell(x)} = lut[bucket].base + popc(occupied sub-cells <= sub(x)).
// One 32-bit shared-memory load per score instead of a ~14-step dependent binary search.
#pragma once
#include <stdint.h>

#include "../../include/madrigal_b200.h"

namespace mdg {

constexpr int kRankBucketBits = MDG_RANK_BUCKET_BITS;
constexpr int kRankSubBits = MDG_RANK_SUB_BITS;
constexpr int kRankCellBits = kRankBucketBits + kRankSubBits;  // 17
constexpr int kRankCells = 1 << kRankCellBits;
constexpr int kRankLutEntries = MDG_RANK_LUT_ENTRIES;
static_assert(kRankSubBits == 4, "LUT entry packs a 16-bit occupancy bitmap");

// cell(x) = round(sat(x*scale + bias) * (2^17 - 1)) in [0, 2^17), produced directly as an INTEGER bit pattern:
// multiplying by the subnormal constant (2^17 - 1) * 2^-149 yields a subnormal result whose bit pattern is the
// rounded integer (subnormals are exact multiples of 2^-149; FMUL rounds to nearest-even at full speed, no FTZ).
// One FFMA + one FMUL, monotone non-decreasing in x for scale > 0.
__device__ __forceinline__ uint32_t rank_key_bits(float x, float scale, float bias) {
  const float y = __saturatef(fmaf(x, scale, bias));  // [0, 1]
  const float c = __uint_as_float(static_cast<uint32_t>(kRankCells - 1));  // subnormal: (2^17 - 1) * 2^-149
  return __float_as_uint(__fmul_rn(y, c));
}
__device__ __forceinline__ uint32_t rank_cell(float x, float scale, float bias) {
  return rank_key_bits(x, scale, bias);
}

// LUT entry: base (thresholds in earlier buckets) in the low half-word, occupancy bitmap in the high half-word with
// sub-cell j at bit 31 - j, so   #{occupied sub-cells <= sub} = popc(entry >> (31 - sub)).
// Returns base + count with garbage in the high half-word: callers take the low 16 bits (the fused epilogue does
// it for free when packing two ranks with PRMT).
__device__ __forceinline__ uint32_t rank_finish(uint32_t entry, uint32_t key_bits) {
  uint32_t sh;  // 31 - sub = (~key & 15) | 16 in ONE LOP3
  asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(sh) : "r"(key_bits), "r"(15u), "r"(16u));
  return entry + __popc(entry >> sh);
}
__device__ __forceinline__ uint32_t rank_lookup_raw(const uint32_t* lut, float x, float scale, float bias) {
  uint32_t kb = rank_key_bits(x, scale, bias);
  return rank_finish(lut[kb >> kRankSubBits], kb);
}

// ------------------------------------------------------------------------------------------------ builder
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  return __uint_as_float(u);
}

// One block per outcome. cells_ws: [L, Q] uint32 scratch = thresholds_out reinterpretation is NOT used; we keep
// the cell codes in the LUT-independent scratch carried in the upper part of thresholds_out until overwritten.
__global__ void __launch_bounds__(256) rank_table_build_kernel(const float* __restrict__ quantiles, int Q,
                                                               float* __restrict__ thresholds_out,
                                                               uint32_t* __restrict__ lut_out,
                                                               float* __restrict__ affine_out) {
  const int l = blockIdx.x;
  const float* q = quantiles + static_cast<size_t>(l) * Q;
  float* thr = thresholds_out + static_cast<size_t>(l) * Q;
  uint32_t* cells = reinterpret_cast<uint32_t*>(thr);  // codes first, overwritten by the snapped floats at the end
  uint32_t* lut = lut_out + static_cast<size_t>(l) * kRankLutEntries;
  __shared__ float s_scale, s_bias;
  __shared__ uint32_t s_cnt[kRankLutEntries];

  if (threadIdx.x == 0) {
    float lo = q[0], hi = q[Q - 1];
    float r = hi - lo;
    if (!(r > 0.f)) r = fmaxf(fabsf(lo), 1.0f) * 1e-3f;
    float lo2 = lo - 0.01f * r, hi2 = hi + 0.01f * r;
    float scale = 1.0f / (hi2 - lo2);
    s_scale = scale;
    s_bias = -lo2 * scale;
    affine_out[2 * l + 0] = s_scale;
    affine_out[2 * l + 1] = s_bias;
  }
  for (int b = threadIdx.x; b < kRankLutEntries; b += blockDim.x) s_cnt[b] = 0;
  __syncthreads();
  const float scale = s_scale, bias = s_bias;

  // 1. raw cell codes
  for (int i = threadIdx.x; i < Q; i += blockDim.x) cells[i] = rank_cell(q[i], scale, bias);
  __syncthreads();
  // 2. make them strictly increasing and keep them inside the grid (sequential; the builder is not a hot path)
  if (threadIdx.x == 0) {
    uint32_t prev = cells[0];
    for (int i = 1; i < Q; ++i) {
      uint32_t c = cells[i];
      c = (c > prev) ? c : prev + 1;
      cells[i] = c;
      prev = c;
    }
    // backward clamp: c_i <= kRankCells-1 - (Q-1-i)   (needs Q <= kRankCells; checked on the host)
    for (int i = Q - 1; i >= 0; --i) {
      uint32_t cap = static_cast<uint32_t>(kRankCells - 1 - (Q - 1 - i));
      if (cells[i] > cap) cells[i] = cap; else break;
    }
  }
  __syncthreads();
  // 3. occupancy bitmap + per-bucket counts
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    uint32_t c = cells[i];
    atomicAdd(&s_cnt[c >> kRankSubBits], 1u);
  }
  __syncthreads();
  // exclusive scan of bucket counts -> base (sequential over 8192 buckets, one thread)
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int b = 0; b < kRankLutEntries; ++b) {
      uint32_t c = s_cnt[b];
      s_cnt[b] = run;  // base = #thresholds in earlier buckets  (<= 65535)
      run += c;
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kRankLutEntries; b += blockDim.x) lut[b] = s_cnt[b];
  __syncthreads();
  __threadfence_block();
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    uint32_t c = cells[i];
    atomicOr(&lut[c >> kRankSubBits], 1u << (31 - (c & 15u)));
  }
  __syncthreads();
  // 4. snapped thresholds: t_i = min fp32 f with cell(f) >= c_i  (bisection over the ordered-uint image of fp32)
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    uint32_t c = cells[i];
    float t;
    uint32_t lo_o = float_to_ordered(-3.0e38f), hi_o = float_to_ordered(3.0e38f);
    if (rank_cell(ordered_to_float(lo_o), scale, bias) >= c) {
      t = ordered_to_float(lo_o);  // c == 0: every finite score is >= the threshold
    } else {
      // invariant: cell(lo) < c <= cell(hi)   (cell(+3e38) = kRankCells-1 >= c)
      while (hi_o - lo_o > 1u) {
        uint32_t mid = lo_o + ((hi_o - lo_o) >> 1);
        if (rank_cell(ordered_to_float(mid), scale, bias) >= c) hi_o = mid; else lo_o = mid;
      }
      t = ordered_to_float(hi_o);
    }
    thr[i] = t;  // overwrites cells[i] (same index, same thread)
  }
}

__global__ void __launch_bounds__(256) rank_lookup_kernel(const float* __restrict__ logits, int64_t n,
                                                          const uint32_t* __restrict__ lut_all,
                                                          const float* __restrict__ affine,
                                                          uint16_t* __restrict__ ranks) {
  const int l = blockIdx.y;
  const uint32_t* lut = lut_all + static_cast<size_t>(l) * kRankLutEntries;
  const float scale = affine[2 * l], bias = affine[2 * l + 1];
  const float* x = logits + static_cast<size_t>(l) * n;
  uint16_t* r = ranks + static_cast<size_t>(l) * n;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    r[i] = static_cast<uint16_t>(rank_lookup_raw(lut, x[i], scale, bias) & 0xFFFFu);
}

}  // namespace mdg
