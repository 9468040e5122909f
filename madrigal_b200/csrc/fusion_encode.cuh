// Fusion-encoder glue kernels (reference: TransformerFusion.forward, madrigal/models/models.py:401-455, with the
// eval-mode arithmetic of torch 1.13 nn.TransformerEncoderLayer / nn.MultiheadAttention, restated in
// oracle/oracle.py:fusion_forward).  All nn.Linear layers run on the tcgen05 GEMM kernel in pair_score.cuh
// (EPI_LINEAR); the kernels here are the non-GEMM parts: operand conversion, LayerNorm, the tiny per-(drug, head)
// masked attention (T <= 32 tokens: one warp, keys on lanes, warp-shuffle softmax), pooling and token assembly.
#pragma once
#include <cuda_bf16.h>
#include <math_constants.h>
#include <stdint.h>

namespace mdg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// attention inputs: the packed q|k|v rows are fp32 in the fp32-parity mode and bf16 in the bf16 mode (half the traffic
// of the in_proj GEMM's output and of the attention kernels' input)
__device__ __forceinline__ float ld_in(const float* p) { return *p; }
__device__ __forceinline__ float ld_in(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ void store_bf16_split(__nv_bfloat16* row, int k, int k_pad, int split, float x) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(x);
  row[k] = hi;
  if (split) row[k_pad + k] = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// x [rows, K] fp32 (row pitch ld_in) -> bf16 GEMM operand rows [hi(k_pad) | lo(k_pad)], zero for k >= K.
// Optional gather: row r reads input row r * row_stride (used to pick the CLS token of every drug).
__global__ void __launch_bounds__(256) convert_rows_kernel(const float* __restrict__ x, long long rows, int K,
                                                           long long ld_in, long long row_stride, int k_pad,
                                                           int split, __nv_bfloat16* __restrict__ out) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * row_stride * ld_in;
  __nv_bfloat16* o = out + row * static_cast<long long>(split ? 2 * k_pad : k_pad);
  for (int k = lane; k < k_pad; k += 32) store_bf16_split(o, k, k_pad, split, k < K ? xr[k] : 0.f);
}

// LayerNorm (eps 1e-5, biased variance) of each row of h [rows, D] (+ optional per-column vector `addvec` first),
// written as a bf16 GEMM operand and/or back to fp32.  do_ln == 0: conversion only.  One warp per row.
__global__ void __launch_bounds__(256) ln_convert_kernel(const float* __restrict__ h, long long rows, int D,
                                                         const float* __restrict__ addvec,
                                                         const float* __restrict__ w, const float* __restrict__ b,
                                                         int do_ln, float* __restrict__ out_f32,
                                                         __nv_bfloat16* __restrict__ out_bf16, int k_pad, int split) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* hr = h + row * D;
  float mean = 0.f, rstd = 1.f;
  if (do_ln) {
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s += hr[k] + (addvec ? addvec[k] : 0.f);
    mean = warp_sum(s) / D;
    float v = 0.f;
    for (int k = lane; k < D; k += 32) {
      const float d = hr[k] + (addvec ? addvec[k] : 0.f) - mean;
      v += d * d;
    }
    rstd = 1.0f / sqrtf(warp_sum(v) / D + 1e-5f);
  }
  __nv_bfloat16* ob = out_bf16 ? out_bf16 + row * static_cast<long long>(split ? 2 * k_pad : k_pad) : nullptr;
  const int kmax = ob ? k_pad : D;
  for (int k = lane; k < kmax; k += 32) {
    float y = 0.f;
    if (k < D) {
      y = hr[k] + (addvec ? addvec[k] : 0.f);
      if (do_ln) y = (y - mean) * rstd * w[k] + b[k];
      if (out_f32) out_f32[row * D + k] = y;
    }
    if (ob) store_bf16_split(ob, k, k_pad, split, y);
  }
}

// Masked multi-head self-attention core.
//   qkv [B*T, 3*Dl] fp32 (packed in_proj output: q | k | v), key_mask [B, T] (non-zero = masked key),
//   src_mask [T, T] or NULL (non-zero = query row i may not attend key j)
//   -> out: bf16 operand rows [B*T, (hi | lo)(k_pad)], head h at columns h*hd .. h*hd+hd-1.
// A warp processes 32/TP (drug, head) items at once, TP = T rounded up to a power of two: lane = (item slot, key).
// K/V tiles are staged in shared memory (row pitch hd+1), the query row is broadcast from shared memory, softmax is
// a segmented warp-shuffle max/sum over the TP lanes of an item.  Scores use q * (1/sqrt(hd)) as
// nn.MultiheadAttention does; masked scores are -inf.
template <typename TIn>
__global__ void attention_kernel(const TIn* __restrict__ qkv, long long ld, const uint8_t* __restrict__ key_mask,
                                 const uint8_t* __restrict__ src_mask, long long B, int T, int TP, int H, int hd,
                                 __nv_bfloat16* __restrict__ out, int k_pad, int split) {
  extern __shared__ float att_smem[];
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = 32 / TP;  // items per warp
  const int pitch = hd + 1;
  const int item_floats = 2 * T * pitch + hd;
  float* wbase = att_smem + static_cast<size_t>(wid) * G * item_floats;
  const int slot = lane / TP, j = lane - slot * TP;
  float* Ks = wbase + slot * item_floats;
  float* Vs = Ks + T * pitch;
  float* Qs = Vs + T * pitch;
  const int Dl = H * hd;
  const float qscale = 1.0f / sqrtf(static_cast<float>(hd));
  const long long total = B * H;
  const long long groups = (total + G - 1) / G;
  for (long long grp = static_cast<long long>(blockIdx.x) * warps + wid; grp < groups;
       grp += static_cast<long long>(gridDim.x) * warps) {
    const long long item = grp * G + slot;
    const bool live = item < total;
    const long long b = live ? item / H : 0;
    const int h = live ? static_cast<int>(item - b * H) : 0;
    const TIn* base = qkv + (b * T) * ld + h * hd;
    __syncwarp();
    if (live) {
      for (int idx = j; idx < T * hd; idx += TP) {
        const int r = idx / hd, d = idx - r * hd;
        Ks[r * pitch + d] = ld_in(base + static_cast<long long>(r) * ld + Dl + d);
        Vs[r * pitch + d] = ld_in(base + static_cast<long long>(r) * ld + 2 * Dl + d);
      }
    }
    const bool key_ok = live && j < T && key_mask[b * T + j] == 0;
    for (int i = 0; i < T; ++i) {
      __syncwarp();
      if (live)
        for (int d = j; d < hd; d += TP) Qs[d] = ld_in(base + static_cast<long long>(i) * ld + d) * qscale;
      __syncwarp();
      float sc = -CUDART_INF_F;
      if (key_ok && !(src_mask != nullptr && src_mask[i * T + j] != 0)) {
        sc = 0.f;
        const float* kr = Ks + j * pitch;
        for (int d = 0; d < hd; ++d) sc = fmaf(Qs[d], kr[d], sc);
      }
      float m = sc;
      for (int o = TP >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float pexp = (sc == -CUDART_INF_F) ? 0.f : expf(sc - m);
      float denom = pexp;
      for (int o = TP >> 1; o > 0; o >>= 1) denom += __shfl_xor_sync(0xffffffffu, denom, o);
      const float pj = pexp / denom;  // NaN if every key is masked, like torch.softmax over all -inf
      __nv_bfloat16* orow = out + (b * T + i) * static_cast<long long>(split ? 2 * k_pad : k_pad) + h * hd;
      for (int d0 = 0; d0 < hd; d0 += TP) {
        const int d = d0 + j;
        float acc = 0.f;
        for (int jj = 0; jj < T; ++jj) {
          const float pb = __shfl_sync(0xffffffffu, pj, slot * TP + jj);
          if (d < hd) acc = fmaf(pb, Vs[jj * pitch + d], acc);
        }
        if (live && d < hd) store_bf16_split(orow, d, k_pad, split, acc);
      }
    }
  }
}

// Register-resident variant for 8 < T <= 32 and head_dim 32 / 64 (the shipped production shapes: T = 21 / 23, 8 heads of
// 64).  One (drug, head) per warp.  Scores use lane = key with that key's K ROW in registers (HD floats); the
// P.V product uses lane = head dimension with the V COLUMNS in registers; the query rows are broadcast from shared
// memory with 16-byte loads.  Per query: HD/4 LDS.128 + HD FMA + two 5-step shuffle reductions + T shuffle-FMA pairs,
// against 2*HD + 2*T scalar shared-memory loads per lane in the generic kernel above.
template <int HD, typename TIn>
__global__ void __launch_bounds__(128) attention_rows_kernel(const TIn* __restrict__ qkv, long long ld,
                                                             const uint8_t* __restrict__ key_mask,
                                                             const uint8_t* __restrict__ src_mask, long long B, int T,
                                                             int H, __nv_bfloat16* __restrict__ out, int k_pad,
                                                             int split) {
  constexpr int NV = HD / 32;        // head dimensions per lane in the P.V stage
  constexpr int KP = HD + 4;         // K staging pitch (floats): 16-byte loads of 8 consecutive rows hit distinct banks
  extern __shared__ float att_smem[];
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = att_smem + static_cast<size_t>(wid) * T * (KP + HD);
  float* Qs = Ks + T * KP;
  const int Dl = H * HD;
  const float qscale = 1.0f / sqrtf(static_cast<float>(HD));
  // src_mask is the same for every item: lane j keeps, as a bit set over queries i, "query i may not attend key j"
  uint32_t blocked_q = 0;
  if (src_mask != nullptr && lane < T)
    for (int i = 0; i < T; ++i) blocked_q |= (src_mask[i * T + lane] != 0 ? 1u : 0u) << i;
  const long long total = B * H;
  for (long long item = static_cast<long long>(blockIdx.x) * warps + wid; item < total;
       item += static_cast<long long>(gridDim.x) * warps) {
    const long long b = item / H;
    const int h = static_cast<int>(item - b * H);
    const TIn* base = qkv + (b * T) * ld + h * HD;
    __syncwarp();
    // coalesced loads: row r of q / k (lane = dimension) -> shared memory; V columns straight into registers
    float v[NV][32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      if (r < T) {
        const TIn* row = base + static_cast<long long>(r) * ld;
#pragma unroll
        for (int e = 0; e < NV; ++e) {
          Qs[r * HD + lane + 32 * e] = ld_in(row + lane + 32 * e) * qscale;
          Ks[r * KP + lane + 32 * e] = ld_in(row + Dl + lane + 32 * e);
          v[e][r] = ld_in(row + 2 * Dl + lane + 32 * e);
        }
      } else {
#pragma unroll
        for (int e = 0; e < NV; ++e) v[e][r] = 0.f;
      }
    }
    __syncwarp();
    float kreg[HD];  // this lane's key row
    const int jr = lane < T ? lane : 0;
#pragma unroll
    for (int c = 0; c < HD; c += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(Ks + jr * KP + c);
      kreg[c] = k4.x; kreg[c + 1] = k4.y; kreg[c + 2] = k4.z; kreg[c + 3] = k4.w;
    }
    const bool key_ok = lane < T && key_mask[b * T + lane] == 0;
    const long long out_pitch = split ? 2 * k_pad : k_pad;
    for (int i = 0; i < T; ++i) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;  // four independent FMA chains
#pragma unroll
      for (int c = 0; c < HD; c += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(Qs + i * HD + c);  // broadcast
        s0 = fmaf(q4.x, kreg[c], s0);
        s1 = fmaf(q4.y, kreg[c + 1], s1);
        s2 = fmaf(q4.z, kreg[c + 2], s2);
        s3 = fmaf(q4.w, kreg[c + 3], s3);
      }
      float sc = (s0 + s1) + (s2 + s3);
      if (!key_ok || ((blocked_q >> i) & 1u)) sc = -CUDART_INF_F;
      const float m = warp_max(sc);
      const float pexp = (sc == -CUDART_INF_F) ? 0.f : expf(sc - m);
      const float pj = pexp / warp_sum(pexp);  // NaN if every key is masked, like torch.softmax over all -inf
      float acc[NV][2];  // even / odd keys: independent chains
#pragma unroll
      for (int e = 0; e < NV; ++e) acc[e][0] = acc[e][1] = 0.f;
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        if (jj < T) {
          const float pb = __shfl_sync(0xffffffffu, pj, jj);
#pragma unroll
          for (int e = 0; e < NV; ++e) acc[e][jj & 1] = fmaf(pb, v[e][jj], acc[e][jj & 1]);
        }
      }
      __nv_bfloat16* orow = out + (b * T + i) * out_pitch + h * HD;
#pragma unroll
      for (int e = 0; e < NV; ++e) store_bf16_split(orow, lane + 32 * e, k_pad, split, acc[e][0] + acc[e][1]);
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2_rn(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Tensor-core variant for bf16 q|k|v rows, head_dim 64 / 128 / 256, T <= 32 (the shipped production shapes: T = 23 with
// 8 heads of 64, T = 21 with 2 heads of 256).  One (drug, head) per warp, FlashAttention-style on warp-level MMAs: the
// item's q, k, v rows (T x 2 HD bytes each) are copied with 16-byte cp.async into XOR-swizzled shared memory,
// S = Q.K^T is m16n8k16 MMAs on ldmatrix fragments, the softmax runs on the accumulator fragments (a row lives in 4
// lanes: two shuffles per reduction), P is re-used in place as the A operand of O = P.V (V fragments from
// ldmatrix.trans, 64 head dimensions at a time), and O goes back through the Q tile so that global stores are 16-byte
// coalesced rows.  Against the FMA kernels (~180 instructions per query at HD = 64, far more at 256) the kernel becomes
// a stream over the q|k|v rows.  P is rounded to bf16 for the second product (bf16 mode only; the fp32-parity mode
// keeps the FMA kernels).  Rows >= T of the staged tiles stay zero; their scores are masked (-inf) and their outputs
// never stored.
template <int HD>
__global__ void __launch_bounds__(128) attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld,
                                                            const uint8_t* __restrict__ key_mask,
                                                            const uint8_t* __restrict__ src_mask, long long B, int T,
                                                            int H, __nv_bfloat16* __restrict__ out, int k_pad) {
  constexpr int PITCH = HD * 2;            // bytes per staged row
  constexpr int PIECES = HD / 8;           // 16-byte pieces per row
  constexpr uint32_t TILE = 32u * PITCH;   // one 32-row tile
  extern __shared__ __align__(128) uint8_t att_mma_smem[];
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sQ = smem_u32(att_mma_smem) + static_cast<uint32_t>(wid) * 3u * TILE, sK = sQ + TILE, sV = sK + TILE;
  for (uint32_t i = lane; i < 3u * TILE / 16u; i += 32) st_shared_v4(sQ + i * 16, 0u, 0u, 0u, 0u);
  const int Dl = H * HD;
  const float qscale = 1.0f / sqrtf(static_cast<float>(HD));
  const int g = lane >> 2, t = lane & 3;
  const int n_mt = T > 16 ? 2 : 1;        // 16-row query tiles
  const int n_nt = (T + 7) >> 3;          // 8-key score tiles
  const int n_kk = T > 16 ? 2 : 1;        // 16-key steps of P.V
  // src_mask (same for every item): for each query row this lane owns, the set of blocked keys
  uint32_t blocked[2][2] = {{0u, 0u}, {0u, 0u}};
  if (src_mask != nullptr)
    for (int mt = 0; mt < 2; ++mt)
      for (int hh = 0; hh < 2; ++hh) {
        const int row = 16 * mt + g + 8 * hh;
        if (row < T)
          for (int j = 0; j < T; ++j) blocked[mt][hh] |= (src_mask[row * T + j] != 0 ? 1u : 0u) << j;
      }
  const uint32_t beyond = T >= 32 ? 0u : ~((1u << T) - 1u);  // keys >= T
  auto swz = [](int row, int piece) { return static_cast<uint32_t>(row * PITCH + ((piece ^ (row & 7)) << 4)); };
  const long long total = B * H;
  for (long long item = static_cast<long long>(blockIdx.x) * warps + wid; item < total;
       item += static_cast<long long>(gridDim.x) * warps) {
    const long long b = item / H;
    const int h = static_cast<int>(item - b * H);
    const __nv_bfloat16* base = qkv + (b * T) * ld + h * HD;
    __syncwarp();  // the previous item's output rows have left the Q tile
    for (int e = lane; e < T * PIECES; e += 32) {
      const int r = e / PIECES, pc = e - r * PIECES;
      const __nv_bfloat16* src = base + static_cast<long long>(r) * ld + pc * 8;
      const uint32_t off = swz(r, pc);
      cp_async_16(sQ + off, src);
      cp_async_16(sK + off, src + Dl);
      cp_async_16(sV + off, src + 2 * Dl);
    }
    cp_async_commit();
    const uint32_t km = __ballot_sync(0xffffffffu, lane < T && key_mask[b * T + lane] != 0) | beyond;
    cp_async_wait_all();
    __syncwarp();
    // ---- S = Q.K^T
    float S[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) S[mt][nt][e] = 0.f;
#pragma unroll 4
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        if (mt < n_mt) {
          const int row = 16 * mt + (lane & 7) + 8 * ((lane >> 3) & 1), piece = 2 * ks + (lane >> 4);
          ldmatrix_x4(sQ + swz(row, piece), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
        }
#pragma unroll
      for (int np = 0; np < 2; ++np)  // pairs of key tiles: one ldmatrix.x4 = (nt, nt+1) x (pieces 2ks, 2ks+1)
        if (2 * np < n_nt) {
          const int row = 16 * np + (lane & 7) + 8 * (lane >> 4), piece = 2 * ks + ((lane >> 3) & 1);
          uint32_t b0, b1, b2, b3;
          ldmatrix_x4(sK + swz(row, piece), b0, b1, b2, b3);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
            if (mt < n_mt) {
              mma_bf16_16816(S[mt][2 * np], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b0, b1);
              if (2 * np + 1 < n_nt) mma_bf16_16816(S[mt][2 * np + 1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b2, b3);
            }
        }
    }
    // ---- masked softmax on the fragments; P packed as the A operand of the second product
    uint32_t P[2][4][2];
    float inv[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const uint32_t dead = km | blocked[mt][hh];
        float m = -CUDART_INF_F;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int j = 8 * nt + 2 * t + e;
            float& sv = S[mt][nt][2 * hh + e];
            sv = ((dead >> j) & 1u) ? -CUDART_INF_F : sv * qscale;
            m = fmaxf(m, sv);
          }
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            float& sv = S[mt][nt][2 * hh + e];
            sv = (sv == -CUDART_INF_F) ? 0.f : __expf(sv - m);
            sum += sv;
          }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        inv[mt][hh] = 1.0f / sum;  // inf if every key is masked: 0 * inf = NaN, like torch.softmax over all -inf
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) P[mt][nt][hh] = pack_bf16x2_rn(S[mt][nt][2 * hh], S[mt][nt][2 * hh + 1]);
      }
    __syncwarp();  // every lane has read its Q fragments: the Q tile becomes the output staging tile
    // ---- O = P.V, 64 head dimensions at a time; normalised rows are staged in the Q tile
    for (int dc = 0; dc < HD / 64; ++dc) {
      float O[2][8][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int dn = 0; dn < 8; ++dn)
#pragma unroll
          for (int e = 0; e < 4; ++e) O[mt][dn][e] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk)
        if (kk < n_kk) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-dimension tiles
            const int row = 16 * kk + (lane & 7) + 8 * ((lane >> 3) & 1), piece = 8 * dc + 2 * dp + (lane >> 4);
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4_trans(sV + swz(row, piece), b0, b1, b2, b3);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
              if (mt < n_mt) {
                mma_bf16_16816(O[mt][2 * dp], P[mt][2 * kk][0], P[mt][2 * kk][1], P[mt][2 * kk + 1][0], P[mt][2 * kk + 1][1], b0, b1);
                mma_bf16_16816(O[mt][2 * dp + 1], P[mt][2 * kk][0], P[mt][2 * kk][1], P[mt][2 * kk + 1][0], P[mt][2 * kk + 1][1], b2, b3);
              }
          }
        }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
        if (mt < n_mt)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int row = 16 * mt + g + 8 * hh;
#pragma unroll
            for (int dn = 0; dn < 8; ++dn)
              st_shared_u32(sQ + swz(row, 8 * dc + dn) + t * 4,
                            pack_bf16x2_rn(O[mt][dn][2 * hh] * inv[mt][hh], O[mt][dn][2 * hh + 1] * inv[mt][hh]));
          }
    }
    __syncwarp();
    __nv_bfloat16* obase = out + (b * T) * static_cast<long long>(k_pad) + h * HD;
    for (int e = lane; e < T * PIECES; e += 32) {
      const int r = e / PIECES, pc = e - r * PIECES;
      uint4 v4;
      ld_shared_v4(sQ + swz(r, pc), v4.x, v4.y, v4.z, v4.w);
      *reinterpret_cast<uint4*>(obase + static_cast<long long>(r) * k_pad + pc * 8) = v4;
    }
  }
}

// Few-token variant (T <= TT <= 8, hd <= 32*DPT): one (drug, head) per warp with the HEAD DIMENSION on lanes.  q/k/v
// rows are read straight from global memory as coalesced 128-byte rows into registers (no shared memory), the T*T
// scores are warp-shuffle reductions, softmax and the P.V product are lane-local.  This is the shape of BASELINE
// config 5 (4 modality tokens) where the keys-on-lanes kernel above would leave 28 of 32 lanes idle.
template <int TT, int DPT, typename TIn>
__global__ void __launch_bounds__(256) attention_small_kernel(const TIn* __restrict__ qkv, long long ld,
                                                              const uint8_t* __restrict__ key_mask,
                                                              const uint8_t* __restrict__ src_mask, long long B,
                                                              int T, int H, int hd, __nv_bfloat16* __restrict__ out,
                                                              int k_pad, int split) {
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dl = H * hd;
  const float qscale = 1.0f / sqrtf(static_cast<float>(hd));
  const long long total = B * H;
  for (long long item = static_cast<long long>(blockIdx.x) * warps + wid; item < total;
       item += static_cast<long long>(gridDim.x) * warps) {
    const long long b = item / H;
    const int h = static_cast<int>(item - b * H);
    const TIn* base = qkv + (b * T) * ld + h * hd;
    float q[TT][DPT], k[TT][DPT], v[TT][DPT];
    bool kok[TT];
#pragma unroll
    for (int t = 0; t < TT; ++t) {
      kok[t] = t < T && key_mask[b * T + t] == 0;
#pragma unroll
      for (int e = 0; e < DPT; ++e) {
        const int d = lane + 32 * e;
        const bool ok = t < T && d < hd;
        const TIn* r = base + static_cast<long long>(t) * ld + d;
        q[t][e] = ok ? ld_in(r) * qscale : 0.f;
        k[t][e] = ok ? ld_in(r + Dl) : 0.f;
        v[t][e] = ok ? ld_in(r + 2 * Dl) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < TT; ++i) {
      if (i < T) {
        float sc[TT];
        float m = -CUDART_INF_F;
#pragma unroll
        for (int j = 0; j < TT; ++j) {
          float part = 0.f;
#pragma unroll
          for (int e = 0; e < DPT; ++e) part = fmaf(q[i][e], k[j][e], part);
          part = warp_sum(part);
          const bool vis = kok[j] && !(src_mask != nullptr && src_mask[i * T + j] != 0);
          sc[j] = vis ? part : -CUDART_INF_F;
          m = fmaxf(m, sc[j]);
        }
        float denom = 0.f;
#pragma unroll
        for (int j = 0; j < TT; ++j) {
          sc[j] = (sc[j] == -CUDART_INF_F) ? 0.f : expf(sc[j] - m);
          denom += sc[j];
        }
        __nv_bfloat16* orow = out + (b * T + i) * static_cast<long long>(split ? 2 * k_pad : k_pad) + h * hd;
#pragma unroll
        for (int e = 0; e < DPT; ++e) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < TT; ++j) acc = fmaf(sc[j] / denom, v[j][e], acc);
          const int d = lane + 32 * e;
          if (d < hd) store_bf16_split(orow, d, k_pad, split, acc);
        }
      }
    }
  }
}

// x-attn pooling attention (models.py:422-440): one learned query per head (already projected and scaled, q_proj
// [Dl]), keys/values kv [B*T, 2*Dl] (k | v), constant key mask pool_mask [T].  out: fp32-exact bf16 operand rows
// [B, (hi | lo)(k_pad)] = concatenated heads, input of the MHA out_proj GEMM.
__global__ void pool_attention_kernel(const float* __restrict__ kv, const float* __restrict__ q_proj,
                                      const uint8_t* __restrict__ pool_mask, long long B, int T, int H, int hd,
                                      __nv_bfloat16* __restrict__ out, int k_pad, int split) {
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dl = H * hd;
  const long long total = B * H;
  for (long long item = static_cast<long long>(blockIdx.x) * warps + wid; item < total;
       item += static_cast<long long>(gridDim.x) * warps) {
    const long long b = item / H;
    const int h = static_cast<int>(item - b * H);
    const float* base = kv + (b * T) * 2LL * Dl + h * hd;
    float s = -CUDART_INF_F;
    if (lane < T && !(pool_mask != nullptr && pool_mask[lane] != 0)) {
      s = 0.f;
      const float* kr = base + static_cast<long long>(lane) * 2 * Dl;
      for (int d = 0; d < hd; ++d) s = fmaf(q_proj[h * hd + d], kr[d], s);
    }
    const float m = warp_max(s);
    const float pexp = (s == -CUDART_INF_F) ? 0.f : expf(s - m);
    const float pj = pexp / warp_sum(pexp);
    __nv_bfloat16* orow = out + b * static_cast<long long>(split ? 2 * k_pad : k_pad) + h * hd;
    for (int d0 = 0; d0 < hd; d0 += 32) {
      const int d = d0 + lane;
      float acc = 0.f;
      for (int j = 0; j < T; ++j) {
        const float pb = __shfl_sync(0xffffffffu, pj, j);
        if (d < hd) acc = fmaf(pb, base[static_cast<long long>(j) * 2 * Dl + Dl + d], acc);
      }
      if (d < hd) store_bf16_split(orow, d, k_pad, split, acc);
    }
  }
}

// The x-attn query path, once per call (models.py:423-429 + the q part of the packed in_proj):
//   q_res  = norm_first ? LN_q(x_attn_query) : x_attn_query          (added back after attention, models.py:440)
//   q_proj = (W_q . q_res + b_q) / sqrt(hd)
// Single block; W_q = first Dl rows of in_proj_weight.
__global__ void __launch_bounds__(256) xattn_query_kernel(const float* __restrict__ query,
                                                          const float* __restrict__ lnw,
                                                          const float* __restrict__ lnb, int norm_first,
                                                          const float* __restrict__ in_w,
                                                          const float* __restrict__ in_b, int Dl, int hd,
                                                          float* __restrict__ q_res, float* __restrict__ q_proj) {
  __shared__ float red[2];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, warps = blockDim.x >> 5;
  float mean = 0.f, rstd = 1.f;
  if (norm_first) {
    if (wid == 0) {
      float s = 0.f;
      for (int k = lane; k < Dl; k += 32) s += query[k];
      const float mu = warp_sum(s) / Dl;
      float v = 0.f;
      for (int k = lane; k < Dl; k += 32) v += (query[k] - mu) * (query[k] - mu);
      const float var = warp_sum(v) / Dl;
      if (lane == 0) {
        red[0] = mu;
        red[1] = 1.0f / sqrtf(var + 1e-5f);
      }
    }
    __syncthreads();
    mean = red[0];
    rstd = red[1];
  }
  for (int k = tid; k < Dl; k += blockDim.x)
    q_res[k] = norm_first ? (query[k] - mean) * rstd * lnw[k] + lnb[k] : query[k];
  __syncthreads();
  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  for (int n = wid; n < Dl; n += warps) {
    float s = 0.f;
    for (int k = lane; k < Dl; k += 32) s = fmaf(in_w[static_cast<long long>(n) * Dl + k], q_res[k], s);
    s = warp_sum(s);
    if (lane == 0) q_proj[n] = (s + in_b[n]) * scale;
  }
}

// 'mean' / 'max' aggregation (models.py:444-451): masked mean / max over the unmasked tokens of each drug.
//   e [B*T, E] fp32 (latent2embed output), key_mask [B, T] -> z [B, E].  Empty -> 0 (torch_scatter fill value).
__global__ void __launch_bounds__(256) masked_pool_kernel(const float* __restrict__ e,
                                                          const uint8_t* __restrict__ key_mask, long long B, int T,
                                                          int E, int is_max, float* __restrict__ z) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * E) return;
  const long long b = idx / E;
  const int d = static_cast<int>(idx - b * E);
  // is_max: 0 = mean, 1 = max, 2 = sum ('add' fusion, models.py:875-878); optional per-token L2 normalisation
  float acc = is_max == 1 ? -CUDART_INF_F : 0.f;
  int cnt = 0;
  for (int t = 0; t < T; ++t) {
    if (key_mask[b * T + t] == 0) {
      const float v = e[(b * T + t) * E + d];
      acc = is_max == 1 ? fmaxf(acc, v) : acc + v;
      ++cnt;
    }
  }
  z[idx] = cnt == 0 ? 0.f : (is_max == 0 ? acc / cnt : acc);
}

// chemCPA transcriptomic token (reference: TxAdaptingComPert.predict, chemcpa/chemCPA/model.py:678-697):
//   latent_treated = (latent_basal + dose_scale * drug_latent) + covariate_embedding[cell_line]
// dose_scale: doser 0 = the dosage itself (GeneralizedSigmoid nonlin=None, :272), 1 = 'sigm', 2 = 'logsigm'
// (:259-271: sigmoid(f(d) * beta[i] + bias[i]) - sigmoid(bias[i]), i = the row's drug index), 3 = precomputed per row
// (the 'amortized' doser MLP's output, :622-627).  drug_latent NULL = model built with use_drugs=False.
__device__ __forceinline__ float tx_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__global__ void __launch_bounds__(256) tx_latent_combine_kernel(
    const float* __restrict__ basal, const float* __restrict__ drug_latent, const float* __restrict__ dosage,
    const long long* __restrict__ drug_idx, const float* __restrict__ beta, const float* __restrict__ bias, int doser,
    const float* __restrict__ cov_table, const long long* __restrict__ cov_idx, long long B, int dim,
    float* __restrict__ out) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const long long b = idx / dim;
  const int d = static_cast<int>(idx - b * dim);
  float v = basal[idx];
  if (drug_latent != nullptr) {
    float sc = dosage[b];
    if (doser == 1 || doser == 2) {
      const long long i = drug_idx[b];
      const float bi = bias[i];
      const float x = doser == 2 ? log1pf(sc) : sc;
      sc = tx_sigmoid(x * beta[i] + bi) - tx_sigmoid(bi);
    }
    v = v + sc * drug_latent[idx];
  }
  if (cov_table != nullptr) v = v + cov_table[cov_idx[b] * dim + d];
  out[idx] = v;
}

// Per-drug 'mlp' dosers (reference: TxAdaptingComPert.__init__ / compute_drug_embeddings_, chemCPA/model.py:405-416,
// 609-621): scale[b] = sigmoid(MLP_i(dosage[b])), i = drug_idx[b], MLP_i = Linear(1, w) -> ReLU ->
// [Linear(w, w) -> ReLU] x (depth - 1) -> Linear(w, 1) with its own parameters per drug (batch_norm=False).
// One warp per sample: the hidden vector lives in the warp's shared-memory row, lane j owns units j, j + 32, ...
// Stacked parameters: w_in / b_in [nd, w], w_hid [nd, depth - 1, w(out), w(in)], b_hid [nd, depth - 1, w],
// w_out [nd, w], b_out [nd].  A drug index outside [0, nd) yields NaN (the reference raises IndexError).
constexpr int kDoserMaxWidth = 256;
__global__ void __launch_bounds__(128) doser_mlp_kernel(const float* __restrict__ dosage,
                                                        const long long* __restrict__ drug_idx, long long B, int nd,
                                                        int w, int depth, const float* __restrict__ w_in,
                                                        const float* __restrict__ b_in,
                                                        const float* __restrict__ w_hid,
                                                        const float* __restrict__ b_hid,
                                                        const float* __restrict__ w_out,
                                                        const float* __restrict__ b_out, float* __restrict__ scale) {
  __shared__ float hbuf[4][2][kDoserMaxWidth];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long b = static_cast<long long>(blockIdx.x) * 4 + wid;
  if (b >= B) return;
  const long long i = drug_idx[b];
  if (i < 0 || i >= nd) {
    if (lane == 0) scale[b] = CUDART_NAN_F;
    return;
  }
  const float x = dosage[b];
  float* h = hbuf[wid][0];
  float* h2 = hbuf[wid][1];
  for (int j = lane; j < w; j += 32) h[j] = fmaxf(fmaf(w_in[i * w + j], x, b_in[i * w + j]), 0.f);
  __syncwarp();
  for (int l = 0; l < depth - 1; ++l) {
    const float* W = w_hid + (static_cast<size_t>(i) * (depth - 1) + l) * w * w;
    const float* bb = b_hid + (static_cast<size_t>(i) * (depth - 1) + l) * w;
    for (int j = lane; j < w; j += 32) {
      const float* row = W + static_cast<size_t>(j) * w;
      float acc = 0.f;
      for (int k = 0; k < w; ++k) acc = fmaf(row[k], h[k], acc);
      h2[j] = fmaxf(acc + bb[j], 0.f);
    }
    __syncwarp();
    float* t = h;
    h = h2;
    h2 = t;
  }
  float acc = 0.f;
  for (int j = lane; j < w; j += 32) acc = fmaf(w_out[i * w + j], h[j], acc);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) scale[b] = tx_sigmoid(acc + b_out[i]);
}

// Token assembly (reference: NovelDDIEncoder.encode, models.py:772-852): builds the position-encoded fusion sequence
// and its key mask from the stacked modality embeddings.
//   embeds [B, M, E] (order [non-TX..., TX...], models.py:772), masks [B, M] (non-zero = missing)
//   sequence layout: [cls?] [non-TX (n_non_tx)] [bottleneck (nb)] [TX (M - n_non_tx)]
//   normalize: L2-normalise every token (eps 1e-12) BEFORE the positional encoding (models.py:849-852)
//   pe [pe_len, E]: added to the first pe_len tokens (sinusoidal buffers are zero beyond max_len; learnable: :602)
__global__ void __launch_bounds__(256) assemble_tokens_kernel(const float* __restrict__ embeds,
                                                              const uint8_t* __restrict__ masks, long long B, int M,
                                                              int E, int n_non_tx, int nb, int has_cls,
                                                              const float* __restrict__ bottleneck,
                                                              const float* __restrict__ cls,
                                                              const float* __restrict__ pe, int pe_len,
                                                              int normalize, float* __restrict__ seq,
                                                              uint8_t* __restrict__ seq_mask) {
  const int T = M + nb + has_cls;
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= B * T) return;
  const int lane = threadIdx.x & 31;
  const long long b = row / T;
  const int t = static_cast<int>(row - b * T);
  const int u = t - has_cls;  // position without CLS
  const float* src;
  uint8_t mk = 0;
  if (u < 0) {
    src = cls;
  } else if (u < n_non_tx) {
    src = embeds + (b * M + u) * E;
    mk = masks[b * M + u];
  } else if (u < n_non_tx + nb) {
    src = bottleneck + static_cast<long long>(u - n_non_tx) * E;
  } else {
    src = embeds + (b * M + (u - nb)) * E;
    mk = masks[b * M + (u - nb)];
  }
  float denom = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int k = lane; k < E; k += 32) ss += src[k] * src[k];
    denom = fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
  }
  for (int k = lane; k < E; k += 32) {
    float v = src[k] / denom;
    if (pe != nullptr && t < pe_len) v += pe[static_cast<long long>(t) * E + k];
    seq[row * E + k] = v;
  }
  if (lane == 0) seq_mask[row] = mk != 0;
}

}  // namespace mdg
