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

// ---- MDG_RANK_PWL: the same 17-bit cell grid, but the table is a 256-bin histogram CDF with linear interpolation:
//   piece = cell >> 9, sub = cell & 511,   rank = base[piece] + (((sub + 1) * slope[piece]) >> 16)
// with base[piece] = #{thresholds whose cell < 512*piece} (exact at the 257 knots) and slope = floor(128 * (base[p+1]
// - base[p])) capped at 65535, so the function is monotone non-decreasing in the cell and hence in x.  One entry is
// base << 16 | slope; the 256 entries are REPLICATED 32x (entry p of copy c at word 32*p + c) so that lane c always
// reads bank c: the shared-memory lookup is conflict-free (1 wavefront per warp instead of ~4.9 for the exact LUT).
// The table the ranks are bit-exact against (np.searchsorted) is thresholds_out[i] = min{x : rank(x) >= i + 1}.
constexpr int kRankPwlPieceBits = 8;
constexpr int kRankPwlPieces = 1 << kRankPwlPieceBits;                 // 256
constexpr int kRankPwlSubBits = kRankCellBits - kRankPwlPieceBits;      // 9
static_assert(kRankPwlPieces * 32 == kRankLutEntries, "the replicated PWL table fills the LUT storage exactly");
__device__ __forceinline__ uint32_t rank_pwl_finish(uint32_t entry, uint32_t key_bits) {
  const uint32_t sub = key_bits & ((1u << kRankPwlSubBits) - 1u);
  const uint32_t m = entry & 0xFFFFu;
  return (entry >> 16) + ((sub * m + m) >> 16);  // (sub + 1) * slope: reaches base[p + 1] in the last cell of a full bin
}
__device__ __forceinline__ uint32_t rank_pwl_raw(const uint32_t* lut, float x, float scale, float bias) {
  const uint32_t kb = rank_key_bits(x, scale, bias);
  return rank_pwl_finish(lut[(kb >> kRankPwlSubBits) * 32], kb);
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

// One block per outcome: histogram-CDF (MDG_RANK_PWL) table from ascending quantiles.  max_dev_out[l] (optional) =
// max over the quantiles of |table rank at q_i - exact rank (i + 1)|: how far the interpolated CDF is from the
// order statistics it was built from, in rank units.
__global__ void __launch_bounds__(256) rank_table_build_pwl_kernel(const float* __restrict__ quantiles, int Q,
                                                                   float* __restrict__ thresholds_out,
                                                                   uint32_t* __restrict__ lut_out,
                                                                   float* __restrict__ affine_out,
                                                                   float* __restrict__ max_dev_out) {
  const int l = blockIdx.x;
  const float* q = quantiles + static_cast<size_t>(l) * Q;
  float* thr = thresholds_out + static_cast<size_t>(l) * Q;
  uint32_t* lut = lut_out + static_cast<size_t>(l) * kRankLutEntries;
  __shared__ float s_scale, s_bias;
  __shared__ uint32_t s_base[kRankPwlPieces + 1];
  __shared__ uint32_t s_entry[kRankPwlPieces];
  __shared__ int s_dev;
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
    s_dev = 0;
  }
  __syncthreads();
  const float scale = s_scale, bias = s_bias;
  // base[p] = #{i : cell(q_i) < 512 p}: the quantiles ascend, so do their cells -> binary search per knot
  for (int pce = threadIdx.x; pce <= kRankPwlPieces; pce += blockDim.x) {
    const uint32_t edge = static_cast<uint32_t>(pce) << kRankPwlSubBits;
    int lo = 0, hi = Q;  // first index with cell >= edge
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (rank_cell(q[mid], scale, bias) >= edge) hi = mid; else lo = mid + 1;
    }
    s_base[pce] = static_cast<uint32_t>(lo);
  }
  __syncthreads();
  for (int pce = threadIdx.x; pce < kRankPwlPieces; pce += blockDim.x) {
    const uint32_t d = s_base[pce + 1] - s_base[pce];
    uint32_t m = d << (16 - kRankPwlSubBits);  // floor(d * 65536 / 512)
    if (m > 0xFFFFu) m = 0xFFFFu;
    s_entry[pce] = (s_base[pce] << 16) | m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRankLutEntries; i += blockDim.x) lut[i] = s_entry[i >> 5];
  auto f = [&](float x) -> uint32_t {
    const uint32_t kb = rank_key_bits(x, scale, bias);
    return rank_pwl_finish(s_entry[kb >> kRankPwlSubBits], kb);
  };
  // snapped thresholds: t_i = min fp32 x with f(x) >= i + 1 (bisection over the ordered-uint image of fp32)
  for (int i = threadIdx.x; i < Q; i += blockDim.x) {
    const uint32_t want = static_cast<uint32_t>(i) + 1u;
    uint32_t lo_o = float_to_ordered(-3.0e38f), hi_o = float_to_ordered(3.0e38f);
    float t;
    if (f(ordered_to_float(lo_o)) >= want) {
      t = ordered_to_float(lo_o);
    } else if (f(ordered_to_float(hi_o)) < want) {
      t = __uint_as_float(0x7f800000u);  // +inf: no finite score reaches this rank
    } else {
      while (hi_o - lo_o > 1u) {
        const uint32_t mid = lo_o + ((hi_o - lo_o) >> 1);
        if (f(ordered_to_float(mid)) >= want) hi_o = mid; else lo_o = mid;
      }
      t = ordered_to_float(hi_o);
    }
    // deviation from the order statistic the table was built from (last of a run of equal quantiles only)
    if (i + 1 == Q || q[i + 1] != q[i]) {
      const int d = static_cast<int>(f(q[i])) - (i + 1);
      atomicMax(&s_dev, d < 0 ? -d : d);
    }
    thr[i] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0 && max_dev_out != nullptr) max_dev_out[l] = static_cast<float>(s_dev);
}

template <bool PWL>
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
    r[i] = static_cast<uint16_t>((PWL ? rank_pwl_raw(lut, x[i], scale, bias) : rank_lookup_raw(lut, x[i], scale, bias)) &
                                 0xFFFFu);
}

}  // namespace mdg
