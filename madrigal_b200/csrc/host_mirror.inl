// Host half of the packed device-to-host transfer (MDG_PAIRS_PACKED_TILES): pure data movement, no arithmetic of the
// scoring path.  The GPU computes every rank; to halve the PCIe volume it ships each unordered pair once, as 32 x 32
// lower-triangular tiles, and this routine scatters the tiles into the normaliser's [L, N, N] layout on the host —
// rank at [i, j] and [j, i], zero diagonal (notebooks/normalize_scores.py:67-70) — with a pool of worker threads.
//
// Work unit = (outcome l, block row bi): the 32 output rows 32 bi .. 32 bi + 31, written left to right in full
// 64-byte lines: tiles (bi, bj < bi) as they are, the diagonal tile OR-ed with its transpose, and tiles (bj > bi, bi)
// transposed (8 x 8 uint16 SSE2 transposes into an L1-resident block).  Destination lines are written with
// non-temporal stores when rows are 64-byte aligned: the output is 2 x the input and is not read again here.
// Included inside the extern "C" block of madrigal_b200.cu.
}  // extern "C"  (C++ helpers below)

#include <emmintrin.h>

#include <thread>
#include <vector>

namespace {

// 8 x 8 uint16 transpose: rows r[0..7] (each 8 lanes) -> columns
inline void transpose8x8_u16(__m128i (&r)[8]) {
  const __m128i a0 = _mm_unpacklo_epi16(r[0], r[1]), a1 = _mm_unpackhi_epi16(r[0], r[1]);
  const __m128i a2 = _mm_unpacklo_epi16(r[2], r[3]), a3 = _mm_unpackhi_epi16(r[2], r[3]);
  const __m128i a4 = _mm_unpacklo_epi16(r[4], r[5]), a5 = _mm_unpackhi_epi16(r[4], r[5]);
  const __m128i a6 = _mm_unpacklo_epi16(r[6], r[7]), a7 = _mm_unpackhi_epi16(r[6], r[7]);
  const __m128i b0 = _mm_unpacklo_epi32(a0, a2), b1 = _mm_unpackhi_epi32(a0, a2);
  const __m128i b2 = _mm_unpacklo_epi32(a1, a3), b3 = _mm_unpackhi_epi32(a1, a3);
  const __m128i b4 = _mm_unpacklo_epi32(a4, a6), b5 = _mm_unpackhi_epi32(a4, a6);
  const __m128i b6 = _mm_unpacklo_epi32(a5, a7), b7 = _mm_unpackhi_epi32(a5, a7);
  r[0] = _mm_unpacklo_epi64(b0, b4);
  r[1] = _mm_unpackhi_epi64(b0, b4);
  r[2] = _mm_unpacklo_epi64(b1, b5);
  r[3] = _mm_unpackhi_epi64(b1, b5);
  r[4] = _mm_unpacklo_epi64(b2, b6);
  r[5] = _mm_unpackhi_epi64(b2, b6);
  r[6] = _mm_unpacklo_epi64(b3, b7);
  r[7] = _mm_unpackhi_epi64(b3, b7);
}

// dst[c][r] = src[r][c] for a 32 x 32 uint16 tile (both contiguous, 64-byte rows)
inline void transpose_tile32(const uint16_t* src, uint16_t* dst) {
  for (int rb = 0; rb < 4; ++rb)
    for (int cb = 0; cb < 4; ++cb) {
      __m128i v[8];
      for (int i = 0; i < 8; ++i)
        v[i] = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + (8 * rb + i) * 32 + 8 * cb));
      transpose8x8_u16(v);
      for (int i = 0; i < 8; ++i)
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + (8 * cb + i) * 32 + 8 * rb), v[i]);
    }
}

template <bool kStream>
inline void put_line(uint16_t* dst, const uint16_t* src) {  // one 64-byte piece of an output row
  for (int q = 0; q < 4; ++q) {
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 8 * q));
    if (kStream) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + 8 * q), v);
    else _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + 8 * q), v);
  }
}

// block row bi of outcome l.  `tiles` = this outcome's [T, 32, 32] array, `out` = this outcome's [N, N] array.
template <bool kStream>
void mirror_block_row(const uint16_t* tiles, int64_t N, int64_t nb, int64_t bi, uint16_t* out) {
  alignas(64) uint16_t blk[32 * 32];
  const int64_t r0 = 32 * bi;
  const int rows = static_cast<int>(N - r0 < 32 ? N - r0 : 32);
  for (int64_t bj = 0; bj < nb; ++bj) {
    const uint16_t* src;
    if (bj < bi) {
      src = tiles + (bi * (bi + 1) / 2 + bj) * 1024;
    } else if (bj > bi) {
      // tiles (bj, bi) of one block column lie (bj + 1) * 2 KB apart: no hardware prefetcher follows that, so the tiles
      // two and three steps ahead are requested here (measured: 2.4x on one thread)
      for (int64_t pj = bj + 2; pj <= bj + 3 && pj < nb; ++pj) {
        const char* nx = reinterpret_cast<const char*>(tiles + (pj * (pj + 1) / 2 + bi) * 1024);
        for (int ln = 0; ln < 2048; ln += 64) _mm_prefetch(nx + ln, _MM_HINT_T0);
      }
      transpose_tile32(tiles + (bj * (bj + 1) / 2 + bi) * 1024, blk);
      src = blk;
    } else {  // diagonal tile: the device left col >= row at zero
      const uint16_t* d = tiles + (bi * (bi + 1) / 2 + bi) * 1024;
      transpose_tile32(d, blk);
      for (int i = 0; i < 1024; i += 8) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(d + i));
        const __m128i b = _mm_load_si128(reinterpret_cast<const __m128i*>(blk + i));
        _mm_store_si128(reinterpret_cast<__m128i*>(blk + i), _mm_or_si128(a, b));
      }
      for (int i = 0; i < 32; ++i) blk[i * 32 + i] = 0;  // normalize_scores.py:69
      src = blk;
    }
    const int64_t c0 = 32 * bj;
    if (c0 + 32 <= N) {
      for (int r = 0; r < rows; ++r) put_line<kStream>(out + (r0 + r) * N + c0, src + r * 32);
    } else {  // ragged right edge
      const int cols = static_cast<int>(N - c0);
      for (int r = 0; r < rows; ++r) memcpy(out + (r0 + r) * N + c0, src + r * 32, static_cast<size_t>(cols) * 2);
    }
  }
}

}  // namespace

extern "C" {

int mdg_host_mirror_tiles(const uint16_t* tiles_host, int64_t L, int64_t N, uint16_t* out_host, int32_t threads) {
  if (!tiles_host || !out_host) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_host_mirror_tiles: NULL pointer");
  if (L < 0 || N < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_host_mirror_tiles: negative size");
  if (L == 0 || N == 0) return MDG_OK;
  const int64_t nb = (N + 31) / 32;
  const int64_t T = nb * (nb + 1) / 2;
  const int64_t units = L * nb;
  int nt = threads > 0 ? threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (nt > 256) nt = 256;
  if (nt > units) nt = static_cast<int>(units);
  // full 64-byte destination lines <=> rows are 64-byte aligned
  const bool stream = (N % 32 == 0) && (reinterpret_cast<uintptr_t>(out_host) % 64 == 0);
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      // a few block rows at a time, largest-offset-free order: unit u = (l, bi)
      const int64_t u = next.fetch_add(1, std::memory_order_relaxed);
      if (u >= units) break;
      const int64_t l = u / nb, bi = u - l * nb;
      const uint16_t* tl = tiles_host + l * T * 1024;
      uint16_t* ol = out_host + l * N * N;
      if (stream) mirror_block_row<true>(tl, N, nb, bi, ol);
      else mirror_block_row<false>(tl, N, nb, bi, ol);
    }
    if (stream) _mm_sfence();
  };
  if (nt == 1) {
    worker();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(static_cast<size_t>(nt - 1));
    try {
      for (int i = 0; i < nt - 1; ++i) pool.emplace_back(worker);
    } catch (...) {  // could not spawn: the calling thread finishes the work
    }
    worker();
    for (auto& t : pool) t.join();
  }
  return MDG_OK;
}
