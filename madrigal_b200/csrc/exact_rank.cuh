// Exact in-sample normalised rank (reference: classwise_normalized_rank_3d_numpy + run_slice,
// notebooks/normalize_scores.py:36-74): per outcome, rank the M = N(N-1)/2 strict-lower-triangle scores
// (argsort(argsort) + 1), divide by M in float64, store float32 at [i,j] and [j,i], zero diagonal.
//
// The sort itself is CUB's device radix sort (library code, like the reference's use of numpy's sort); the gather,
// the key transform and the rank scatter are ours.  Keys are the order-preserving uint32 image of fp32; the radix
// sort is stable, so equal scores are ranked in (i, j) row-major order == np.argsort(kind='stable').
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_segmented_radix_sort.cuh>
#include <stdint.h>

namespace mdg {

__device__ __forceinline__ void tri_unflatten(unsigned long long p, unsigned int* i_out, unsigned int* j_out) {
  // p = i(i-1)/2 + j with 0 <= j < i
  unsigned long long i = static_cast<unsigned long long>((1.0 + sqrt(1.0 + 8.0 * static_cast<double>(p))) * 0.5);
  while (i * (i - 1) / 2 > p) --i;
  while ((i + 1) * i / 2 <= p) ++i;
  *i_out = static_cast<unsigned int>(i);
  *j_out = static_cast<unsigned int>(p - i * (i - 1) / 2);
}

__global__ void __launch_bounds__(256) tri_gather_keys_kernel(const float* __restrict__ scores, int N,
                                                              unsigned long long M, uint32_t* __restrict__ keys,
                                                              uint32_t* __restrict__ idx) {
  for (unsigned long long p = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; p < M;
       p += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    unsigned int i, j;
    tri_unflatten(p, &i, &j);
    float x = scores[static_cast<size_t>(i) * N + j];
    if (x == 0.f) x = 0.f;  // -0.0 and +0.0 are one value for the reference's float comparison
    const uint32_t u = __float_as_uint(x);
    keys[p] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    idx[p] = static_cast<uint32_t>(p);
  }
}

__global__ void __launch_bounds__(256) tri_scatter_rank_kernel(const uint32_t* __restrict__ sorted_idx, int N,
                                                               unsigned long long M, float* __restrict__ out) {
  for (unsigned long long r = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; r < M;
       r += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    unsigned int i, j;
    tri_unflatten(sorted_idx[r], &i, &j);
    // normalize_scores.py:57: rank / (N*(N-1)/2) in float64, stored into a float32 memmap (:72)
    const float v = static_cast<float>(static_cast<double>(r + 1) / static_cast<double>(M));
    out[static_cast<size_t>(i) * N + j] = v;
    out[static_cast<size_t>(j) * N + i] = v;
  }
  for (unsigned long long d = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
       d < static_cast<unsigned long long>(N); d += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
    out[d * N + d] = 0.f;  // normalize_scores.py:69
}

// The scatter above writes 4 bytes to two random places per pair: at the reference's size (67 M pairs per outcome)
// that is 4.2 GB of partial-sector read-modify-write traffic and 70 % of the exact-rank time.  The placement below
// replaces it.  Two more radix passes order the (pair index, rank) pairs by the pair index's bits >= 13; since every
// pair index 0..M-1 occurs exactly once, slice b of 8192 consecutive array elements then holds exactly the pair
// indices [8192 b, 8192 (b+1)).  One block per slice drops the ranks into shared memory at (index - 8192 b) and writes
// the slice out in index order = row-major order of the strict lower triangle: coalesced, every sector written once.
// A tiled transpose then mirrors the triangle (coalesced both ways).
constexpr int kRankSliceBits = 13;
constexpr int kRankSlice = 1 << kRankSliceBits;
__global__ void __launch_bounds__(256) tri_place_rank_kernel(const uint32_t* __restrict__ part_idx,
                                                             const uint32_t* __restrict__ part_rank, int N,
                                                             unsigned long long M, float* __restrict__ out) {
  __shared__ uint32_t ranks[kRankSlice];
  const unsigned long long base = static_cast<unsigned long long>(blockIdx.x) * kRankSlice;
  const int n = static_cast<int>(min(static_cast<unsigned long long>(kRankSlice), M - base));
  for (int k = threadIdx.x; k < n; k += 256) ranks[part_idx[base + k] - static_cast<uint32_t>(base)] = part_rank[base + k];
  __syncthreads();
  const double inv_den = static_cast<double>(M);
  for (int k = threadIdx.x; k < n; k += 256) {
    unsigned int i, j;
    tri_unflatten(base + k, &i, &j);
    // normalize_scores.py:57: rank / (N*(N-1)/2) in float64, stored into a float32 memmap (:72)
    out[static_cast<size_t>(i) * N + j] = static_cast<float>(static_cast<double>(ranks[k] + 1u) / inv_den);
  }
}

// out[j, i] = out[i, j] for i > j, diagonal 0 (normalize_scores.py:69-70); one 32x32 tile pair per block of (32, 8)
__global__ void __launch_bounds__(256) tri_mirror_kernel(float* __restrict__ out, int N) {
  __shared__ float tile[32][33];
  unsigned int bi, bj;  // blockIdx.x = bi (bi + 1) / 2 + bj, bj <= bi
  {
    const unsigned long long t = blockIdx.x;
    unsigned long long b = static_cast<unsigned long long>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
    while (b * (b + 1) / 2 > t) --b;
    while ((b + 1) * (b + 2) / 2 <= t) ++b;
    bi = static_cast<unsigned int>(b);
    bj = static_cast<unsigned int>(t - b * (b + 1) / 2);
  }
  const int x = threadIdx.x;
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = bi * 32 + y, c = bj * 32 + x;
    tile[y][x] = (r < N && c < N && c < r) ? out[static_cast<size_t>(r) * N + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = bj * 32 + y, c = bi * 32 + x;  // transposed position: value of (row c, col r)
    if (r < N && c < N && c >= r) out[static_cast<size_t>(r) * N + c] = tile[x][y];  // c == r: tile value is 0
  }
}

// Q order statistics (ranks ceil(i*M/Q), i = 1..Q) of the sorted keys -> ascending fp32 quantiles
__global__ void __launch_bounds__(256) pick_quantiles_kernel(const uint32_t* __restrict__ sorted_keys,
                                                             unsigned long long M, int Q, float* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const unsigned long long rank = (static_cast<unsigned long long>(q + 1) * M + Q - 1) / Q;  // 1-based
  const uint32_t o = sorted_keys[rank - 1];
  const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  out[q] = __uint_as_float(u);
}

// ---- top-k finalisation: candidates appended by the EPI_TOPK epilogue -> per-outcome top k
// sort key = (~ordered(score) << 32) | pair index: ascending order = descending score, ties by ascending pair index;
// unused slots are all-ones and sort last.
__global__ void __launch_bounds__(256) topk_make_keys_kernel(const unsigned long long* __restrict__ cand,
                                                             const unsigned int* __restrict__ counts, int cap,
                                                             unsigned long long* __restrict__ keys,
                                                             int* __restrict__ offsets, int L) {
  const int l = blockIdx.y;
  const unsigned int n = min(counts[l], static_cast<unsigned int>(cap));
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
    unsigned long long key = ~0ull;
    if (static_cast<unsigned int>(i) < n) {
      const unsigned long long c = cand[static_cast<size_t>(l) * cap + i];
      const uint32_t u = static_cast<uint32_t>(c >> 32);
      const uint32_t o = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      key = (static_cast<unsigned long long>(~o) << 32) | (c & 0xFFFFFFFFull);
    }
    keys[static_cast<size_t>(l) * cap + i] = key;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    offsets[l] = l * cap;
    if (l == L - 1) offsets[L] = L * cap;
  }
}

__global__ void __launch_bounds__(256) topk_emit_kernel(const unsigned long long* __restrict__ sorted,
                                                        const unsigned int* __restrict__ counts, int cap, int k,
                                                        int ncols, float* __restrict__ scores, int* __restrict__ rows,
                                                        int* __restrict__ cols, int* __restrict__ status) {
  const int l = blockIdx.y;
  const unsigned int n_raw = counts[l];
  const int n = n_raw < static_cast<unsigned int>(cap) ? static_cast<int>(n_raw) : cap;
  if (blockIdx.x == 0 && threadIdx.x == 0)
    status[l] = (n_raw > static_cast<unsigned int>(cap)) ? 2 : (n < k ? 1 : 0);  // 2: overflow, 1: fewer than k
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    float sc = -INFINITY;
    int r = -1, c = -1;
    if (i < n) {
      const unsigned long long key = sorted[static_cast<size_t>(l) * cap + i];
      const uint32_t o = ~static_cast<uint32_t>(key >> 32);
      const uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
      sc = __uint_as_float(u);
      const uint32_t idx = static_cast<uint32_t>(key);
      r = static_cast<int>(idx / static_cast<uint32_t>(ncols));
      c = static_cast<int>(idx % static_cast<uint32_t>(ncols));
    }
    scores[static_cast<size_t>(l) * k + i] = sc;
    rows[static_cast<size_t>(l) * k + i] = r;
    cols[static_cast<size_t>(l) * k + i] = c;
  }
}

struct ExactRankWs {
  uint32_t *keys_in, *keys_out, *idx_in, *idx_out;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

inline ExactRankWs plan_exact_rank(void* ws, long long N) {
  ExactRankWs w;
  const unsigned long long M = static_cast<unsigned long long>(N) * (N - 1) / 2;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<const uint32_t*>(nullptr),
                                  static_cast<uint32_t*>(nullptr), static_cast<long long>(M));
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t arr = static_cast<size_t>(M ? M : 1) * 4;
  size_t o0 = take(arr), o1 = take(arr), o2 = take(arr), o3 = take(arr), o4 = take(cub_bytes ? cub_bytes : 1);
  uint8_t* b = static_cast<uint8_t*>(ws);
  w.keys_in = reinterpret_cast<uint32_t*>(b + o0);
  w.keys_out = reinterpret_cast<uint32_t*>(b + o1);
  w.idx_in = reinterpret_cast<uint32_t*>(b + o2);
  w.idx_out = reinterpret_cast<uint32_t*>(b + o3);
  w.cub_temp = b + o4;
  w.cub_bytes = cub_bytes;
  w.total = off;
  return w;
}

}  // namespace mdg
