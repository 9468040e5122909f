// Fused fusion-encoder kernel (reference: TransformerFusion.forward, madrigal/models/models.py:401-455; pre-LN
// nn.TransformerEncoderLayer semantics as restated in oracle/oracle.py:fusion_forward).
//
// ONE persistent kernel runs the whole encoder for a tile of 128 token rows (= floor(128/T) drugs): embed2latent,
// every transformer layer (LN1 -> per-head QKV -> masked softmax attention -> out-proj -> LN2 -> FFN) and the
// pooling (cls / mean / max / x-attn) + latent2embed.  Nothing but the tokens, the masks and z touches HBM:
//
//   * the residual stream H [128 x Dl] fp32 lives in TMEM columns [0, 256): out-proj and FFN2 are tcgen05.mma's that
//     ACCUMULATE straight onto it; the biases they would add are carried as per-stage "pending bias" vectors that
//     are applied whenever H is read (LayerNorm / pooling);
//   * a second TMEM accumulator (columns [256, 512)) receives the per-head [q|k|v] projections, the FFN1 chunks and
//     the latent2embed output;
//   * the A operands (LN output, attention output, GELU output) are produced by the epilogue warps directly in the
//     128-byte-swizzled K-major shared-memory layout UMMA reads, never leaving the SM;
//   * weights (bf16, K-major rows, prepared once) stream from L2 through a 2-stage TMA ring.
//
// warp 0: TMA weight producer, warp 1: UMMA issuer, warp 2: TMEM allocator, warps 4-19: FOUR threads per token row
// (tcgen05.ld -> LayerNorm / attention / activation -> swizzled smem; the four split 32-column chunks in element-wise
// stages and (head, 16-dimension slice) in attention stages).  MMA phases and epilogue phases of a tile
// alternate strictly (two mbarriers, one arrival protocol), so hazards on the shared buffers are ordered by
// construction; the weight ring runs ahead independently.
//
// Supported: bf16 operands, norm_first, Dl <= 256 (multiple of 64), head_dim in {16, 32}, E <= 256 (multiple of
// 16), T <= 32, any FFN width, 16-byte aligned bias / LayerNorm vectors.  Everything else takes the generic multi-kernel path in capi_fusion.inl.
#pragma once
#include <cuda_bf16.h>
#include <math_constants.h>

#include "../../include/madrigal_b200.h"
#include "mdg_ptx.cuh"

namespace mdg {

constexpr int kFeRows = 128;
constexpr int kFeABufBytes = 128 * 256 * 2;      // 64 KB: 4 panels of [128 x 64] bf16
constexpr int kFeBStageBytes = 256 * 64 * 2;     // 32 KB: [<=256 rows x 64 k] bf16
constexpr int kFeBStages = 2;
constexpr int kFeKvBytes = 128 * 2 * 64 * 2;  // k|v rows of one head in bf16 (16-byte chunks XOR-swizzled by row)
constexpr int kFeSmemA = 0;
constexpr int kFeSmemO = kFeSmemA + kFeABufBytes;
constexpr int kFeSmemB = kFeSmemO + kFeABufBytes;
constexpr int kFeSmemKv = kFeSmemB + kFeBStages * kFeBStageBytes;
constexpr int kFeSmemBar = kFeSmemKv + kFeKvBytes;
constexpr int kFeSmemPar = kFeSmemBar + 128;          // 2 x 1 KB: the parameter vector of the current / next phase
constexpr int kFeParBytes = 1024;
constexpr int kFeSmemTotal = kFeSmemPar + 2 * kFeParBytes;
constexpr int kFeSmemBytes = kFeSmemTotal;  // the dynamic shared-memory base is 1024-byte aligned (checked in the kernel)
constexpr int kFeEpiWarps = 16;
constexpr int kFeThreads = (4 + kFeEpiWarps) * 32;
constexpr int kFeTmemH = 0;
constexpr int kFeTmemAcc = 256;
static_assert(kFeSmemBytes <= 232448, "fused encoder exceeds 227 KB of shared memory");

// Phase trace of CTA 0 (measurement hook, enabled per launch by FusedEncParams::trace): clock64 at every epilogue
// `wait_mma` return / `signal` and every MMA-warp `wait_epi` return / commit.  Read back with mdg_fusion_trace_read.
constexpr int kFeTraceLen = 512;
__device__ unsigned long long g_fe_trace[2 * kFeTraceLen];  // [0, 512): epilogue warp 4, [512, 1024): MMA warp

struct FusedEncParams {
  int trace;  // non-zero: CTA 0 records its phase timeline into g_fe_trace
  long long B;
  int T, E, Dl, F, H, hd, layers, act, agg;
  int kp_e, kp_d;  // 64-wide K panels of E and Dl
  int f_pad;       // F rounded up to 64
  int fc;          // FFN chunk width: 256, 128 or 64 (divides f_pad)
  const float* tokens;        // [B, T, E]
  const uint8_t* key_mask;    // [B, T]
  const uint8_t* src_mask;    // [T, T] or NULL
  const uint8_t* pool_mask;   // [T] or NULL (x-attn)
  float* z_out;               // [B, E]
  const float* pend;          // [(2*layers + 1), Dl] cumulative biases folded into reads of H
  const float* in_bias[MDG_MAX_LAYERS];
  const float* l1_bias[MDG_MAX_LAYERS];
  const float* l2e_bias;
  // x-attn pooling
  const float* xin_bias;   // x_attn in_proj bias [3*Dl] (k at Dl, v at 2*Dl)
  const float* xout_bias;  // [Dl]
  const float* xq_nw;      // x_attn_query_norm (applied after the residual when !norm_first; unused: norm_first only)
  const float* xq_nb;
  const float* q_res;      // [Dl]
  const float* xo_qres;    // [Dl] x_attn out_proj bias + q_res (what is added to the pooled out-proj output)
  const float* q_proj;     // [Dl] (already scaled)
  long long num_tiles;
  int drugs_per_tile;
};

// element (row, k) of an A buffer: panel k/64, 128-byte rows, 16-byte chunks XOR-swizzled with (row & 7)
__device__ __forceinline__ uint32_t fe_a_addr(uint32_t buf, int row, int k) {
  return buf + static_cast<uint32_t>((k >> 6) * 16384 + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2);
}

__device__ __forceinline__ uint32_t fe_pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// write 32 consecutive k (starting at k0, multiple of 32) of this thread's row as bf16 into a swizzled A buffer
__device__ __forceinline__ void fe_store_row32(uint32_t buf, int row, int k0, const float (&y)[32]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t addr = fe_a_addr(buf, row, k0 + 8 * c);
    st_shared_v4(addr, fe_pack2(y[8 * c], y[8 * c + 1]), fe_pack2(y[8 * c + 2], y[8 * c + 3]),
                 fe_pack2(y[8 * c + 4], y[8 * c + 5]), fe_pack2(y[8 * c + 6], y[8 * c + 7]));
  }
}

// Exact-erf GELU (F.gelu default, models.py:366 `activation='gelu'`) without the ~45-instruction erff():
//   gelu(x) = 0.5 x (1 + erf(x / sqrt2)) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt2),
//   erfc(u) = (a1 t + ... + a5 t^5) exp(-u^2),  t = 1 / (1 + 0.3275911 u)     (Abramowitz-Stegun 7.1.26,
//   |error| <= 1.5e-7 in erf) -- two MUFU ops + 9 FMA-pipe instructions; absolute error of gelu < 1e-6.
__device__ __forceinline__ float fe_gelu(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(ax, 0.3275911f * 0.70710678118654752440f, 1.0f));
  const float e = exp2f(x * x * (-0.5f * 1.4426950408889634f));
  float pl = fmaf(t, 1.061405429f, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  return fmaf(-0.5f * ax * pl, e, fmaxf(x, 0.f));
}
// The same GELU for two values at once on the packed fp32x2 pipe (fma.rn.f32x2 / mul.f32x2 / add.f32x2, sm_100):
// every polynomial step is ONE instruction for the pair; only |x|, max(x, 0) and the two MUFU ops stay scalar.
__device__ __forceinline__ float2 fe_gelu2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 den = __ffma2_rn(ax, make_float2(0.3275911f * 0.70710678118654752440f, 0.3275911f * 0.70710678118654752440f),
                                make_float2(1.0f, 1.0f));
  const float2 t = make_float2(__fdividef(1.0f, den.x), __fdividef(1.0f, den.y));
  const float2 xx = __fmul2_rn(x, x);
  const float2 ea = __fmul2_rn(xx, make_float2(-0.5f * 1.4426950408889634f, -0.5f * 1.4426950408889634f));
  const float2 e = make_float2(exp2f(ea.x), exp2f(ea.y));
  float2 pl = __ffma2_rn(t, make_float2(1.061405429f, 1.061405429f), make_float2(-1.453152027f, -1.453152027f));
  pl = __ffma2_rn(pl, t, make_float2(1.421413741f, 1.421413741f));
  pl = __ffma2_rn(pl, t, make_float2(-0.284496736f, -0.284496736f));
  pl = __ffma2_rn(pl, t, make_float2(0.254829592f, 0.254829592f));
  pl = __fmul2_rn(pl, t);
  const float2 w = __fmul2_rn(__fmul2_rn(ax, make_float2(-0.5f, -0.5f)), pl);
  return __ffma2_rn(w, e, make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f)));
}
__device__ __forceinline__ float fe_act(float a, int act) {
  if (act == 1) return fmaxf(a, 0.f);
  return fe_gelu(a);
}

// 16-byte vector load of 4 consecutive fp32 parameters (uniform address across the warp: one L1 wavefront)
__device__ __forceinline__ float4 fe_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 fe_lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <int HD>
__global__ void __launch_bounds__(kFeThreads, 1)
fused_encoder_kernel(const __grid_constant__ CUtensorMap tm_e2l, const __grid_constant__ CUtensorMap tm_in,
                     const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_l1,
                     const __grid_constant__ CUtensorMap tm_l2, const __grid_constant__ CUtensorMap tm_l2e,
                     const __grid_constant__ CUtensorMap tm_xin, const __grid_constant__ CUtensorMap tm_xout,
                     const __grid_constant__ FusedEncParams p) {
  constexpr int HP = 64 / HD;  // heads per attention phase (their q|k|v projections share one accumulator: 192 columns)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  if (base & 1023u) {  // the swizzled operand layouts need 1024-byte alignment; there is no slack to realign
    g_hang_code = 0x40000000u;
    asm volatile("trap;");
  }
  uint8_t* gbase = smem_raw;
  const uint32_t sA = base + kFeSmemA, sO = base + kFeSmemO, sB = base + kFeSmemB, sKv = base + kFeSmemKv,
                 sBar = base + kFeSmemBar, sPar = base + kFeSmemPar;
  const uint32_t bar_mma_done = sBar, bar_epi_done = sBar + 8, bar_acc_free = sBar + 48;
  auto bar_full = [&](int i) { return sBar + 16 + 8 * i; };
  auto bar_empty = [&](int i) { return sBar + 32 + 8 * i; };
  const uint32_t tmem_slot = sBar + 64;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = lane_id();
  if (threadIdx.x == 0) {
    mbar_init(bar_mma_done, 1);
    mbar_init(bar_epi_done, kFeEpiWarps);  // one arrival per epilogue warp
    mbar_init(bar_acc_free, kFeEpiWarps);  // "the q|k|v accumulator has been read out": lets the next head pair's
                                           // projection run under the current pair's softmax
    for (int i = 0; i < kFeBStages; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_e2l);
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_out);
    tma_prefetch_desc(&tm_l1);
    tma_prefetch_desc(&tm_l2);
    tma_prefetch_desc(&tm_l2e);
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + kFeSmemBar + 64);
  // PDL launch: the prologue above overlapped the previous kernel's tail; the tokens may be that kernel's output, so
  // everything else waits for it.  (Waiting only before the first z store instead was measured: -2 us per bench
  // step, not worth an aliasing rule on the inputs.)
  pdl_wait();
  pdl_launch_dependents();

  const int Dl = p.Dl, kp_d = p.kp_d, kp_e = p.kp_e;
  constexpr int hd = HD;
  const int FC = p.fc;  // FFN chunk width (divides f_pad)
  const int n_fchunks = p.f_pad / FC;
  const bool xattn = p.agg == MDG_AGG_XATTN;
  const int n_phases = (p.H + HP - 1) / HP;

  if (warp == 0) {
    // ======================================================================= weight producer (TMA)
    int stage = 0;
    uint32_t phase = 0;
    // one B k-panel = `nbox` boxes of `rows` rows each (row coordinate row_of(b)), stacked in one ring stage;
    // `layer` is the tensor map's batch coordinate (per-layer weights share one map)
    auto load_panel = [&](const CUtensorMap* tm, int layer, int nbox, int rows, int kc, auto row_of) {
      mbar_wait(bar_empty(stage), phase ^ 1, 21);
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_full(stage), static_cast<uint32_t>(nbox * rows * 128));
        const uint32_t dst = sB + stage * kFeBStageBytes;
        for (int b = 0; b < nbox; ++b) tma_load_3d(dst + b * rows * 128, tm, bar_full(stage), kc, row_of(b), layer);
      }
      __syncwarp();
      if (++stage == kFeBStages) {
        stage = 0;
        phase ^= 1;
      }
    };
    auto row0 = [](int) { return 0; };
    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int k = 0; k < kp_e; ++k) load_panel(&tm_e2l, 0, 1, Dl, k * 64, row0);  // embed2latent
      for (int l = 0; l < p.layers; ++l) {
        for (int ph = 0; ph < n_phases; ++ph) {  // [Wq_h; Wk_h; Wv_h] for each head of the phase
          const int h0 = ph * HP, nh = min(HP, p.H - h0);
          for (int k = 0; k < kp_d; ++k)
            load_panel(&tm_in, l, 3 * nh, hd, k * 64, [&](int b) { return (b % 3) * Dl + (h0 + b / 3) * hd; });
        }
        for (int k = 0; k < kp_d; ++k) load_panel(&tm_out, l, 1, Dl, k * 64, row0);  // out-proj
        for (int c = 0; c < n_fchunks; ++c) {
          if (c == 0)
            for (int k = 0; k < kp_d; ++k) load_panel(&tm_l1, l, 1, FC, k * 64, row0);
          // phase: ffn1(c+1) (runs under the tail of chunk c's activation) then ffn2(c)
          if (c + 1 < n_fchunks)
            for (int k = 0; k < kp_d; ++k) load_panel(&tm_l1, l, 1, FC, k * 64, [&](int) { return (c + 1) * FC; });
          for (int k = 0; k < FC / 64; ++k) load_panel(&tm_l2, l, 1, Dl, c * FC + k * 64, row0);
        }
      }
      if (xattn) {
        for (int ph = 0; ph < n_phases; ++ph) {  // [Wk_h; Wv_h] of the pooling MHA (tm_xin holds the k|v rows: 2*Dl)
          const int h0 = ph * HP, nh = min(HP, p.H - h0);
          for (int k = 0; k < kp_d; ++k)
            load_panel(&tm_xin, 0, 2 * nh, hd, k * 64, [&](int b) { return (b % 2) * Dl + (h0 + b / 2) * hd; });
        }
        for (int k = 0; k < kp_d; ++k) load_panel(&tm_xout, 0, 1, Dl, k * 64, row0);
      }
      for (int k = 0; k < kp_d; ++k) load_panel(&tm_l2e, 0, 1, p.E, k * 64, row0);  // latent2embed
    }
  } else if (warp == 1) {
    // ======================================================================= UMMA issuer
    const uint64_t desc_hi = umma_desc_kmajor_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_kmajor_sw128(0));
    auto desc_of = [&](uint32_t a) { return desc_hi | static_cast<uint64_t>(desc_lo0 | ((a & 0x3FFFFu) >> 4)); };
    int stage = 0;
    uint32_t phase = 0;
    uint32_t epi_waits = 0;
    // D[tmem_col .. +N) (+)= A[abuf panels 0 .. kp) . B^T, B panels from the ring
    auto gemm = [&](uint32_t abuf, int kp, int N, uint32_t tmem_col, bool accumulate) {
      const uint32_t idesc = umma_idesc_bf16_f32(128, N);
      const uint32_t d = tmem_base + tmem_col;
      for (int k = 0; k < kp; ++k) {
        mbar_wait(bar_full(stage), phase, 22);
        tc_fence_after_sync();
        const uint64_t adesc = desc_of(abuf + k * 16384);
        const uint64_t bdesc = desc_of(sB + stage * kFeBStageBytes);
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            umma_bf16(d, adesc + static_cast<uint64_t>(q * 2), bdesc + static_cast<uint64_t>(q * 2), idesc,
                      (accumulate || k > 0 || q > 0) ? 1u : 0u);
          umma_commit(bar_empty(stage));
        }
        __syncwarp();
        if (++stage == kFeBStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    };
    int tr_n = 0;
    const bool tracing = p.trace != 0 && blockIdx.x == 0 && lane == 0;
    auto wait_epi = [&]() {
      mbar_wait(bar_epi_done, epi_waits & 1, 23);
      ++epi_waits;
      tc_fence_after_sync();
      if (tracing && tr_n < kFeTraceLen) g_fe_trace[kFeTraceLen + tr_n++] = clock64();
    };
    auto signal = [&]() {
      if (elect_one()) umma_commit(bar_mma_done);
      __syncwarp();
      if (tracing && tr_n < kFeTraceLen) g_fe_trace[kFeTraceLen + tr_n++] = clock64();
    };
    uint32_t acc_waits = 0;
    auto wait_acc = [&]() {
      mbar_wait(bar_acc_free, acc_waits & 1, 25);
      ++acc_waits;
      tc_fence_after_sync();
    };
    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      wait_epi();  // tokens in A
      gemm(sA, kp_e, Dl, kFeTmemH, false);
      signal();
      for (int l = 0; l < p.layers; ++l) {
        for (int ph = 0; ph < n_phases; ++ph) {
          if (ph == 0) wait_epi();  // LN1 in A
          else wait_acc();          // the previous head pair's q|k|v have left the accumulator (their softmax still runs)
          gemm(sA, kp_d, 3 * hd * min(HP, p.H - ph * HP), kFeTmemAcc, false);
          signal();
        }
        wait_epi();  // attention output complete in O
        gemm(sO, kp_d, Dl, kFeTmemH, true);
        signal();
        wait_epi();  // LN2 in A
        gemm(sA, kp_d, FC, kFeTmemAcc, false);
        signal();
        for (int c = 0; c < n_fchunks; ++c) {
          if (c + 1 < n_fchunks) {
            wait_acc();  // every thread has its pieces of chunk c in registers: the accumulator can take chunk c + 1
            gemm(sA, kp_d, FC, kFeTmemAcc, false);
          }
          wait_epi();  // activation chunk c in O
          gemm(sO, FC / 64, Dl, kFeTmemH, true);
          signal();
        }
      }
      if (xattn) {
        for (int ph = 0; ph < n_phases; ++ph) {
          if (ph == 0) wait_epi();  // LN_kv in A
          else wait_acc();
          gemm(sA, kp_d, 2 * hd * min(HP, p.H - ph * HP), kFeTmemAcc, false);
          signal();
        }
        wait_epi();  // pooled attention output in O (rows = first token row of each drug)
        gemm(sO, kp_d, Dl, kFeTmemAcc, false);
        signal();
      }
      wait_epi();  // pooling input in A
      gemm(sA, kp_d, p.E, kFeTmemAcc, false);
      signal();
    }
  } else if (warp >= 4) {
    // ======================================================================= epilogue: FOUR threads per token row
    // warps 4-7 (group 0) .. 16-19 (group 3) all map lane -> TMEM lane (warp & 3) * 32 + lane.  The four threads of a
    // row split element-wise stages by 32-column chunk (chunk ci belongs to group ci % 4) and attention stages by
    // (head of the phase, 16-dimension slice of the head's output): a thread always accumulates 16 output dimensions.
    constexpr int TPR = kFeEpiWarps / 4;  // threads per row
    constexpr int TPH = TPR / HP;         // threads per head of a phase (hd 32: 2, hd 16: 1)
    constexpr int DPT = HD / TPH;         // output dimensions per thread (16)
    static_assert(TPR == 4 && DPT == 16, "epilogue mapping assumes 4 threads per row and 16 dimensions per thread");
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int g = ew >> 2;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int T = p.T;
    const int G = p.drugs_per_tile;
    // k|v exchange rows of one head: k (hd bf16) | v (hd bf16), pitch 4*hd bytes, 16-byte chunk c of row r stored at
    // chunk c ^ sw(r) so that the 8 lanes of a quarter-warp (consecutive rows, same logical chunk) hit distinct banks;
    // one 128-row region per head of the phase
    constexpr int kv_pitch = 4 * HD;
    constexpr int kKvShift = (HD == 16) ? 1 : 0;        // rows per 128-byte line = 128 / pitch (hd = 16: 2)
    constexpr uint32_t kKvMask = (HD == 16) ? 3u : 7u;  // chunks per row - 1, capped at 7
    auto kv_addr = [&](int hh, int r, int chunk) -> uint32_t {
      return sKv + static_cast<uint32_t>(hh * 128 * kv_pitch + r * kv_pitch) +
             ((static_cast<uint32_t>(chunk) ^ ((static_cast<uint32_t>(r) >> kKvShift) & kKvMask)) << 4);
    };
    uint32_t mma_waits = 0;
    int tr_n = 0;
    const bool tracing = p.trace != 0 && blockIdx.x == 0 && warp == 4 && lane == 0;
    auto wait_mma = [&]() {
      mbar_wait(bar_mma_done, mma_waits & 1, 24);
      ++mma_waits;
      tc_fence_after_sync();
      if (tracing && tr_n < kFeTraceLen) g_fe_trace[tr_n++] = (1ull << 56) | (clock64() & 0xFFFFFFFFFFFFull);
    };
    auto mark = [&](unsigned long long tag) {  // sub-phase marker (trace only)
      if (tracing && tr_n < kFeTraceLen) g_fe_trace[tr_n++] = (tag << 56) | (clock64() & 0xFFFFFFFFFFFFull);
    };
    auto signal = [&]() {  // smem writes -> async proxy, TMEM reads done
      mark(3);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (tracing && tr_n < kFeTraceLen) g_fe_trace[tr_n++] = (2ull << 56) | (clock64() & 0xFFFFFFFFFFFFull);
      if (lane == 0) mbar_arrive(bar_epi_done);
    };
    auto release_acc = [&]() {  // this warp's tcgen05.ld of the q|k|v accumulator are complete
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_free);
    };
    // Per-phase parameter vectors (pending bias, q|k|v biases of the phase's heads, FFN bias chunk, ...) are read by
    // every thread; with ~225 KB of shared memory the L1 cache is gone, so a global load is an L2 round trip.  They
    // are therefore STAGED: right after handing a phase to the MMA warp, the first threads copy the NEXT phase's
    // vector (<= 256 floats) into one of two 1 KB buffers (the copy overlaps the MMA phase); after the wait a
    // 512-thread barrier publishes it.  Buffer k&1 was last read two phases ago, which every warp has left.
    uint32_t par_n = 0;
    const int etid = ew * 32 + lane;
    auto stage = [&](int n4, auto src_of) {  // src_of(i) -> address of the i-th float4 of the vector
      if (etid < n4) {
        const float4 v = fe_ldg4(src_of(etid));
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(sPar + (par_n & 1u) * kFeParBytes + etid * 16),
                     "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
      ++par_n;
    };
    auto publish = [&]() -> uint32_t {  // after wait_mma: the vector staged last is now readable; returns its address
      named_bar_sync(6, kFeEpiWarps * 32);
      return sPar + ((par_n - 1u) & 1u) * kFeParBytes;
    };
    auto stage_vec = [&](const float* v, int n) { stage((n + 3) >> 2, [&](int i) { return v + 4 * i; }); };
    // q|k|v biases of the heads of attention phase ph, layout [head slot][q | k | v][HD]
    auto stage_qkv = [&](const float* ib, int ph) {
      const int h0 = ph * HP, nh = min(HP, p.H - h0);
      stage(nh * 3 * (HD / 4), [&](int i) {
        const int seg = i / (HD / 4), off = i - seg * (HD / 4);
        return ib + (seg % 3) * Dl + (h0 + seg / 3) * HD + 4 * off;
      });
    };
    // pooling phase: [head slot][k bias | v bias | projected query][HD]
    auto stage_pool = [&](int ph) {
      const int h0 = ph * HP, nh = min(HP, p.H - h0);
      stage(nh * 3 * (HD / 4), [&](int i) {
        const int seg = i / (HD / 4), off = i - seg * (HD / 4), part = seg % 3, h = h0 + seg / 3;
        return (part == 2 ? p.q_proj + h * HD : p.xin_bias + (part + 1) * Dl + h * HD) + 4 * off;
      });
    };
    // LayerNorm of (H + pend) for this row -> bf16 into `dst`.  Each of the row's threads owns every 4th 32-column
    // chunk; the partial sums meet in shared memory (the k|v exchange region is idle during LN stages).
    auto layer_norm_to = [&](uint32_t dst, uint32_t pend, bool do_ln) {
      float mean = 0.f, rstd = 1.f;
      if (do_ln) {
        float s = 0.f, ss = 0.f;
        for (int c = g * 32; c < Dl; c += 32 * TPR) {
          uint32_t v[32];
          tmem_ld_32x32(trow + kFeTmemH + c, v);
          float4 pd[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) pd[j] = fe_lds4(pend + (c + 4 * j) * 4);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x0 = __uint_as_float(v[4 * j]) + pd[j].x, x1 = __uint_as_float(v[4 * j + 1]) + pd[j].y,
                        x2 = __uint_as_float(v[4 * j + 2]) + pd[j].z, x3 = __uint_as_float(v[4 * j + 3]) + pd[j].w;
            s += (x0 + x1) + (x2 + x3);
            ss = fmaf(x0, x0, ss);
            ss = fmaf(x1, x1, ss);
            ss = fmaf(x2, x2, ss);
            ss = fmaf(x3, x3, ss);
          }
        }
        float2* part = reinterpret_cast<float2*>(gbase + kFeSmemKv);
        part[g * 128 + row] = make_float2(s, ss);
        mark(4);
        named_bar_sync(1, kFeEpiWarps * 32);
        mark(5);
        s = 0.f;
        ss = 0.f;
#pragma unroll
        for (int k = 0; k < TPR; ++k) {  // same order in all four threads: identical statistics
          const float2 o = part[k * 128 + row];
          s += o.x;
          ss += o.y;
        }
        mean = s / Dl;
        const float var = fmaxf(ss / Dl - mean * mean, 0.f);
        rstd = 1.0f / sqrtf(var + 1e-5f);
      }
      const float shift = -mean * rstd;
      for (int c = g * 32; c < Dl; c += 32 * TPR) {
        uint32_t v[32];
        tmem_ld_32x32(trow + kFeTmemH + c, v);
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 pd = fe_lds4(pend + (c + 4 * j) * 4);
          y[4 * j] = pd.x; y[4 * j + 1] = pd.y; y[4 * j + 2] = pd.z; y[4 * j + 3] = pd.w;
        }
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(v[j]);
        if (do_ln) {  // the LayerNorm weight / bias live in the next linear's weights (folded at prepare time)
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = fmaf(y[j], rstd, shift);
        }
        fe_store_row32(dst, row, c, y);
      }
    };
    // softmax(q . K^T) V over the T keys of this row's drug for head slot hh of the phase (k|v rows start at row r0),
    // output dimensions [d0, d0 + DPT) of the head; q is the full scaled query.  The (already normalised) result
    // goes to O columns [col0 + d0, col0 + d0 + DPT).
    auto attend = [&](int hh, int r0, uint32_t key_blocked, bool active, const float (&q)[HD], int d0, int col0) {
      float m = -CUDART_INF_F, lsum = 0.f;
      float acc[DPT];
#pragma unroll
      for (int d = 0; d < DPT; ++d) acc[d] = 0.f;
      if (active && T <= 4) {
        // few-token case (BASELINE's 4 modality tokens): all scores first, ONE max / normaliser, then P.V without the
        // running rescale of the online form below
        float sc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sc[j] = -CUDART_INF_F;
          if (j < T && !((key_blocked >> j) & 1u)) {
            float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);  // packed fp32x2 FMA chains
#pragma unroll
            for (int d8 = 0; d8 < HD / 8; ++d8) {
              uint32_t a, b, c2, e;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(e) : "r"(kv_addr(hh, r0 + j, d8)));
              const uint32_t ww[4] = {a, b, c2, e};
#pragma unroll
              for (int t2 = 0; t2 < 4; t2 += 2) {
                const float2 k0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[t2]));
                const float2 k1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[t2 + 1]));
                s01 = __ffma2_rn(make_float2(q[d8 * 8 + 2 * t2], q[d8 * 8 + 2 * t2 + 1]), k0, s01);
                s23 = __ffma2_rn(make_float2(q[d8 * 8 + 2 * t2 + 2], q[d8 * 8 + 2 * t2 + 3]), k1, s23);
              }
            }
            sc[j] = (s01.x + s01.y) + (s23.x + s23.y);
          }
        }
        m = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (sc[j] == -CUDART_INF_F) continue;  // masked (or beyond T): contributes exactly 0
          const float pj = __expf(sc[j] - m);
          lsum += pj;
#pragma unroll
          for (int d8 = 0; d8 < DPT / 8; ++d8) {
            uint32_t a, b, c2, e;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(e) : "r"(kv_addr(hh, r0 + j, HD / 8 + d0 / 8 + d8)));
            const uint32_t ww[4] = {a, b, c2, e};
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[t2]));
              const float2 r = __ffma2_rn(make_float2(pj, pj), vv, make_float2(acc[d8 * 8 + 2 * t2], acc[d8 * 8 + 2 * t2 + 1]));
              acc[d8 * 8 + 2 * t2] = r.x;
              acc[d8 * 8 + 2 * t2 + 1] = r.y;
            }
          }
        }
      } else if (active) {
        for (int j = 0; j < T; ++j) {
          if ((key_blocked >> j) & 1u) continue;
          const int rj = r0 + j;
          float s = 0.f;
#pragma unroll
          for (int d8 = 0; d8 < HD / 8; ++d8) {
            uint32_t a, b, c2, e;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(e) : "r"(kv_addr(hh, rj, d8)));
            const uint32_t ww[4] = {a, b, c2, e};
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              const float2 kk = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[t2]));
              s = fmaf(q[d8 * 8 + 2 * t2], kk.x, s);
              s = fmaf(q[d8 * 8 + 2 * t2 + 1], kk.y, s);
            }
          }
          const float m_new = fmaxf(m, s);
          const float corr = __expf(m - m_new);  // exp(-inf) = 0 on the first visible key
          const float pj = __expf(s - m_new);
          lsum = fmaf(lsum, corr, pj);
#pragma unroll
          for (int d8 = 0; d8 < DPT / 8; ++d8) {
            uint32_t a, b, c2, e;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c2), "=r"(e) : "r"(kv_addr(hh, rj, HD / 8 + d0 / 8 + d8)));
            const uint32_t ww[4] = {a, b, c2, e};
#pragma unroll
            for (int t2 = 0; t2 < 4; ++t2) {
              const float2 vv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[t2]));
              acc[d8 * 8 + 2 * t2] = fmaf(pj, vv.x, acc[d8 * 8 + 2 * t2] * corr);
              acc[d8 * 8 + 2 * t2 + 1] = fmaf(pj, vv.y, acc[d8 * 8 + 2 * t2 + 1] * corr);
            }
          }
          m = m_new;
        }
      }
      const float inv = active ? 1.0f / lsum : 0.f;  // all keys masked -> inf/NaN like torch.softmax
#pragma unroll
      for (int d8 = 0; d8 < DPT / 8; ++d8) {
        float o[8];
#pragma unroll
        for (int t2 = 0; t2 < 8; ++t2) o[t2] = active ? acc[d8 * 8 + t2] * inv : 0.f;
        st_shared_v4(fe_a_addr(sO, row, col0 + d0 + d8 * 8), fe_pack2(o[0], o[1]), fe_pack2(o[2], o[3]),
                     fe_pack2(o[4], o[5]), fe_pack2(o[6], o[7]));
      }
    };
    // 16 accumulator columns [acol, acol + 16) + bias -> bf16 into the k (part 0) / v (part 1) row of head slot hh at
    // dimension offset d0
    auto stash_kv16 = [&](int hh, int part, int d0, uint32_t acol, uint32_t bias) {
      uint32_t v[16];
      tmem_ld_32x16(trow + kFeTmemAcc + acol, v);
      float4 bb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = fe_lds4(bias + 16 * j);
      tmem_ld_wait();
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        w[2 * j] = fe_pack2(__uint_as_float(v[4 * j]) + bb[j].x, __uint_as_float(v[4 * j + 1]) + bb[j].y);
        w[2 * j + 1] = fe_pack2(__uint_as_float(v[4 * j + 2]) + bb[j].z, __uint_as_float(v[4 * j + 3]) + bb[j].w);
      }
      const int ch = (part * HD + d0) / 8;  // 16-byte chunk index within the row
      st_shared_v4(kv_addr(hh, row, ch), w[0], w[1], w[2], w[3]);
      st_shared_v4(kv_addr(hh, row, ch + 1), w[4], w[5], w[6], w[7]);
    };
    const int hh_mine = g % HP;          // head slot of a phase this thread works on
    const int d0_mine = (g / HP) * DPT;  // its slice of that head's dimensions

    for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int dloc = row / T, tok = row - dloc * T;
      const long long drug = tile * G + dloc;
      const bool valid = dloc < G && drug < p.B;
      // key visibility of this row's drug as a bitmask (bit j set = key j blocked for this query)
      uint32_t blocked = 0;
      if (valid) {
        for (int j = 0; j < T; ++j) {
          const bool m = p.key_mask[drug * T + j] != 0 || (p.src_mask != nullptr && p.src_mask[tok * T + j] != 0);
          blocked |= (m ? 1u : 0u) << j;
        }
      }

      // ---- tokens -> A (bf16, zero padded): each warp copies its 8 rows one at a time, 32 lanes x float4 per
      //      512-byte piece of the row (coalesced), converted to bf16 and written into the swizzled operand layout
      const float* pend_final = p.pend + static_cast<long long>(2 * p.layers) * Dl;
      {
        const int kw = kp_e * 64;
        for (int r = 0; r < 128 / kFeEpiWarps; ++r) {
          const int rr = ew * (128 / kFeEpiWarps) + r;
          const int dl2 = rr / T;
          const long long drug2 = tile * G + dl2;
          const bool ok = dl2 < G && drug2 < p.B;
          const float* src = p.tokens + (drug2 * T + (rr - dl2 * T)) * static_cast<long long>(p.E);
          for (int k = lane * 4; k < kw; k += 128) {
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok && k < p.E) x = __ldg(reinterpret_cast<const float4*>(src + k));  // E % 16 == 0
            const uint32_t addr = fe_a_addr(sA, rr, k);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(fe_pack2(x.x, x.y)), "r"(fe_pack2(x.z, x.w))
                         : "memory");
          }
        }
        signal();
        stage_vec(p.layers > 0 ? p.pend : pend_final, Dl);  // next phase: LN1 of layer 0 (or the pooling input)
      }
      for (int l = 0; l < p.layers; ++l) {
        const float* ib = p.in_bias[l];
        // ---- LN1
        wait_mma();
        layer_norm_to(sA, publish(), true);
        signal();
        stage_qkv(ib, 0);
        // ---- attention: HP heads per phase; a thread takes (head slot, 16-dimension slice)
        const float qscale = 1.0f / sqrtf(static_cast<float>(hd));
        for (int ph = 0; ph < n_phases; ++ph) {
          wait_mma();
          const uint32_t par = publish();
          const int h0 = ph * HP, nh = min(HP, p.H - h0);
          if (hh_mine < nh) {
            const int hh = hh_mine, h = h0 + hh;
            const uint32_t acol = static_cast<uint32_t>(hh * 3 * HD);
            const uint32_t bq = par + static_cast<uint32_t>(hh * 3 * HD) * 4;  // [q | k | v] biases of this head
            float q[HD];
#pragma unroll
            for (int c = 0; c < HD; c += 16) {
              uint32_t v[16];
              tmem_ld_32x16(trow + kFeTmemAcc + acol + c, v);
              float4 bb[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) bb[j] = fe_lds4(bq + (c + 4 * j) * 4);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                q[c + 4 * j] = (__uint_as_float(v[4 * j]) + bb[j].x) * qscale;
                q[c + 4 * j + 1] = (__uint_as_float(v[4 * j + 1]) + bb[j].y) * qscale;
                q[c + 4 * j + 2] = (__uint_as_float(v[4 * j + 2]) + bb[j].z) * qscale;
                q[c + 4 * j + 3] = (__uint_as_float(v[4 * j + 3]) + bb[j].w) * qscale;
              }
            }
            stash_kv16(hh, 0, d0_mine, acol + HD + d0_mine, bq + (HD + d0_mine) * 4);
            stash_kv16(hh, 1, d0_mine, acol + 2 * HD + d0_mine, bq + (2 * HD + d0_mine) * 4);
            if (ph + 1 < n_phases) release_acc();  // the next head pair's projection may overwrite the accumulator
            named_bar_sync(2 + hh, TPH * 128);  // k/v of head h for every row of the tile are in shared memory
            attend(hh, row - tok, blocked, valid, q, d0_mine, h * HD);
          } else if (ph + 1 < n_phases) {
            release_acc();
          }
          // the hand-over to the MMA warp happens once, after the last head pair (the attention output is complete);
          // between head pairs the MMA warp only waits for the accumulator (release_acc) and the publish() barrier of
          // the next phase keeps the k|v exchange region from being overwritten under a slower warp's softmax
          if (ph + 1 < n_phases) {
            stage_qkv(ib, ph + 1);
          } else {
            signal();
            stage_vec(p.pend + static_cast<long long>(2 * l + 1) * Dl, Dl);  // next: LN2
          }
        }
        // ---- LN2
        wait_mma();
        layer_norm_to(sA, publish(), true);
        signal();
        stage_vec(p.l1_bias[l], min(FC, p.F));
        // ---- FFN activation chunks (32-column pieces of the chunk alternate between the row's threads)
        for (int c = 0; c < n_fchunks; ++c) {
          wait_mma();
          const uint32_t b1 = publish();  // linear1 bias of columns [c * FC, c * FC + FC)
          const bool more_chunks = c + 1 < n_fchunks;
          if (more_chunks && g * 32 >= FC) release_acc();  // this thread reads nothing of the chunk
          for (int cc = g * 32; cc < FC; cc += 32 * TPR) {
            uint32_t v[32];
            tmem_ld_32x32(trow + kFeTmemAcc + cc, v);
            const bool last_piece = cc + 32 * TPR >= FC;
            float y[32];
            if (c * FC + cc + 32 <= p.F) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 bb = fe_lds4(b1 + (cc + 4 * j) * 4);
                y[4 * j] = bb.x; y[4 * j + 1] = bb.y; y[4 * j + 2] = bb.z; y[4 * j + 3] = bb.w;
              }
              tmem_ld_wait();
              if (more_chunks && last_piece) release_acc();  // FFN1 of the next chunk runs under this piece's GELU
              if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = fmaxf(__uint_as_float(v[j]) + y[j], 0.f);
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                  const float2 r = fe_gelu2(__fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])),
                                                       make_float2(y[j], y[j + 1])));
                  y[j] = r.x;
                  y[j + 1] = r.y;
                }
              }
            } else {
              tmem_ld_wait();
              if (more_chunks && last_piece) release_acc();
              const float* b1g = p.l1_bias[l] + c * FC;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const bool in = c * FC + cc + j < p.F;  // columns beyond F are K padding: exactly zero
                y[j] = in ? fe_act(__uint_as_float(v[j]) + __ldg(b1g + cc + j), p.act) : 0.f;
              }
            }
            fe_store_row32(sO, row, cc, y);
          }
          signal();
          if (c + 1 < n_fchunks) stage_vec(p.l1_bias[l] + (c + 1) * FC, min(FC, p.F - (c + 1) * FC));
          else stage_vec(l + 1 < p.layers ? p.pend + static_cast<long long>(2 * l + 2) * Dl : pend_final, Dl);
        }
      }
      if (xattn) {
        // ---- x-attn pooling (models.py:422-440): kv = LN_kv(h); one learned query per head; constant key mask
        wait_mma();
        layer_norm_to(sA, publish(), true);
        signal();
        stage_pool(0);
        uint32_t pblocked = 0;
        for (int j = 0; j < T; ++j)
          if (p.pool_mask != nullptr && p.pool_mask[j] != 0) pblocked |= 1u << j;
        const bool pool_row = valid && tok == 0;  // the first token row of each drug computes the pooled output
        for (int ph = 0; ph < n_phases; ++ph) {
          wait_mma();
          const uint32_t par = publish();
          const int h0 = ph * HP, nh = min(HP, p.H - h0);
          if (hh_mine < nh) {
            const int hh = hh_mine, h = h0 + hh;
            const uint32_t acol = static_cast<uint32_t>(hh * 2 * HD);
            const uint32_t bk = par + static_cast<uint32_t>(hh * 3 * HD) * 4;  // [k bias | v bias | query] of this head
            stash_kv16(hh, 0, d0_mine, acol + d0_mine, bk + d0_mine * 4);
            stash_kv16(hh, 1, d0_mine, acol + HD + d0_mine, bk + (HD + d0_mine) * 4);
            float q[HD];
#pragma unroll
            for (int c = 0; c < HD; c += 4) {
              const float4 qq = fe_lds4(bk + (2 * HD + c) * 4);
              q[c] = qq.x; q[c + 1] = qq.y; q[c + 2] = qq.z; q[c + 3] = qq.w;
            }
            if (ph + 1 < n_phases) release_acc();
            named_bar_sync(2 + hh, TPH * 128);
            attend(hh, row, pblocked, pool_row, q, d0_mine, h * HD);
          } else if (ph + 1 < n_phases) {
            release_acc();
          }
          if (ph + 1 < n_phases) {
            stage_pool(ph + 1);
          } else {
            signal();
            stage_vec(p.xo_qres, Dl);  // next: out-proj bias + residual query
          }
        }
        // ---- out-proj of the pooled query + residual query (norm_first: no LN here) -> A for latent2embed
        wait_mma();
        const uint32_t xo = publish();
        for (int c = g * 32; c < Dl; c += 32 * TPR) {
          uint32_t v[32];
          tmem_ld_32x32(trow + kFeTmemAcc + c, v);
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = fe_lds4(xo + (c + 4 * j) * 4);
            y[4 * j] = a.x; y[4 * j + 1] = a.y; y[4 * j + 2] = a.z; y[4 * j + 3] = a.w;
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(v[j]);
          fe_store_row32(sA, row, c, y);
        }
        signal();
      } else {
        // ---- pooling input: latent2embed is applied to every token row (models.py:415)
        wait_mma();
        layer_norm_to(sA, publish(), false);
        signal();
      }
      stage_vec(p.l2e_bias, p.E);
      // ---- latent2embed output -> pooled z (32-column chunks rotate over the row's threads)
      wait_mma();
      {
        const uint32_t l2b = publish();
        // per-group [128][33] fp32 exchange for mean / max pooling: groups 0-2 in the (idle) O buffer, group 3 in the
        // k|v exchange region
        float* xbuf = (g < 3) ? reinterpret_cast<float*>(gbase + kFeSmemO) + g * (128 * 33)
                              : reinterpret_cast<float*>(gbase + kFeSmemKv);
        for (int c = g * 32; c < p.E; c += 32 * TPR) {
          uint32_t v[32];
          tmem_ld_32x32(trow + kFeTmemAcc + c, v);
          tmem_ld_wait();
          if (p.agg == MDG_AGG_CLS || xattn) {  // the first token row of each drug IS the result
            if (valid && tok == 0) {
              float* zo = p.z_out + drug * p.E + c;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (c + j < p.E) {  // E % 16 == 0: whole float4s
                  const float4 bb = fe_lds4(l2b + (c + j) * 4);
                  *reinterpret_cast<float4*>(zo + j) =
                      make_float4(__uint_as_float(v[j]) + bb.x, __uint_as_float(v[j + 1]) + bb.y,
                                  __uint_as_float(v[j + 2]) + bb.z, __uint_as_float(v[j + 3]) + bb.w);
                }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) xbuf[row * 33 + j] = __uint_as_float(v[j]);
            named_bar_sync(2 + g, 128);
            if (valid && tok == 0) {
              float* zo = p.z_out + drug * p.E + c;
              for (int j = 0; j < 32; ++j) {
                if (c + j >= p.E) break;
                float accv = p.agg == MDG_AGG_MAX ? -CUDART_INF_F : 0.f;
                int cnt = 0;
                for (int t2 = 0; t2 < T; ++t2) {
                  if (p.key_mask[drug * T + t2] != 0) continue;
                  const float x = xbuf[(row + t2) * 33 + j];
                  accv = p.agg == MDG_AGG_MAX ? fmaxf(accv, x) : accv + x;
                  ++cnt;
                }
                const float bias = __ldg(p.l2e_bias + c + j);
                zo[j] = cnt == 0 ? 0.f : (p.agg == MDG_AGG_MAX ? accv + bias : accv / cnt + bias);
              }
            }
            named_bar_sync(2 + g, 128);
          }
        }
        tc_fence_before_sync();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// W' = W * diag(ln_w) as bf16 operand rows [N, k_pad] (zero padded) and b' = b + W . ln_b: the LayerNorm affine of the
// rows feeding a linear, folded into that linear.  One warp per output row.
__global__ void __launch_bounds__(256) fold_ln_linear_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                             const float* __restrict__ ln_w,
                                                             const float* __restrict__ ln_b, int N, int K, int k_pad,
                                                             __nv_bfloat16* __restrict__ w_out,
                                                             float* __restrict__ b_out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const float* wr = W + static_cast<size_t>(n) * K;
  __nv_bfloat16* o = w_out + static_cast<size_t>(n) * k_pad;
  float acc = 0.f;
  for (int k = lane; k < k_pad; k += 32) {
    float x = 0.f;
    if (k < K) {
      const float wv = wr[k];
      x = wv * ln_w[k];
      acc = fmaf(wv, ln_b[k], acc);
    }
    o[k] = __float2bfloat16_rn(x);
  }
#pragma unroll
  for (int s2 = 16; s2 > 0; s2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s2);
  if (lane == 0) b_out[n] = bias[n] + acc;
}

__global__ void vec_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// pend[s][d]: biases already "owed" to the TMEM-resident residual stream at stage s
//   s = 0: embed2latent bias; s = 2l+1: + out_proj bias of layer l; s = 2l+2: + linear2 bias of layer l
struct FusedPendArgs {
  const float* e2l_bias;
  const float* out_bias[MDG_MAX_LAYERS];
  const float* l2_bias[MDG_MAX_LAYERS];
  int layers, Dl;
};
__global__ void fused_pend_kernel(FusedPendArgs a, float* __restrict__ pend) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= a.Dl) return;
  float acc = a.e2l_bias[d];
  pend[d] = acc;
  for (int l = 0; l < a.layers; ++l) {
    acc += a.out_bias[l][d];
    pend[static_cast<size_t>(2 * l + 1) * a.Dl + d] = acc;
    acc += a.l2_bias[l][d];
    pend[static_cast<size_t>(2 * l + 2) * a.Dl + d] = acc;
  }
}

}  // namespace mdg
