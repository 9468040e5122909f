// All-pairs bilinear decoder on tcgen05 / TMEM / TMA  (reference: BilinearDDIScorer.bilinear,
// madrigal/models/models.py:537-539:  matmul(matmul(z1, W[L,D,D]), z2.T) -> [L, N1, N2]).
//
// The reference's association order is kept: GEMM 1  Y_l = z_rows . W_l   ([Nr,D] per outcome, small),
// then GEMM 2  S_l = Y_l . z_cols^T  ([Nr,Nc] per outcome, the N^2 part).  Both are "NT" GEMMs with K-major bf16
// operands, so ONE persistent warp-specialised kernel runs both, with a different epilogue:
//
//   warp 0      TMA producer   : A panels (resident for a whole task) + B panels (3-stage ring), 128B-swizzled
//   warp 1      UMMA issuer    : tcgen05.mma cta_group::1 kind::f16, 128x128x16, fp32 accumulators in TMEM,
//                                2 accumulator stages x up to 2 row sub-tiles = 512 TMEM columns
//   warp 2      TMEM allocator
//   warps 4..   epilogue       : tcgen05.ld -> registers -> {fp32 | sigmoid | u16 quantile rank | bf16 hi/lo}
//                                -> 64-byte-swizzled staging in smem -> TMA store (or guarded direct stores)
//
// A task = (outcome l, 128*msub-row block, chunk of 128-column blocks).  Within a task the A operand stays in
// shared memory and only B streams, so L2->SM traffic per output is K*2/(128*msub) bytes.
//
// Precision: MDG_PREC_BF16 -> one bf16 term (msub = 2).  MDG_PREC_FP32 -> bf16x3 split: operands are stored as
// [hi | lo] along K and the K loop runs hi*hi + lo*hi + hi*lo into one accumulator (msub = 1).
#pragma once
#include <cuda_bf16.h>
#include <math_constants.h>

#include "mdg_ptx.cuh"
#include "rank_table.cuh"

namespace mdg {

constexpr int kBM = 128;
constexpr int kBN = 128;
constexpr int kBK = 64;  // bf16 elements per K block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kPanelBytes = kBM * kBK * 2;  // 16 KB: one [128 x 64] bf16 panel
constexpr int kMaxAPanels = 8;              // 128 KB resident A
constexpr int kFirstEpiWarp = 4;
constexpr int kStagingBytesPerWarp = 2048;  // 32 rows x 64 B
constexpr int kTmemCols = 512;
constexpr int kTaskRing = 4;  // dynamic-scheduler ring depth

enum EpiMode : int {
  EPI_F32 = 0, EPI_SIGMOID = 1, EPI_RANK_U16 = 2, EPI_BF16_SPLIT = 3, EPI_LINEAR = 4, EPI_TOPK = 5,
  EPI_RANK_U16_MIRROR = 6,  // rank of row > col pairs, written at [row, col] and [col, row]
  EPI_RANK_U16_PWL = 7,         // the two rank modes with the histogram-CDF table (MDG_RANK_PWL): conflict-free lookups
  EPI_RANK_U16_MIRROR_PWL = 8
};
__host__ __device__ constexpr bool epi_is_rank(int e) {
  return e == EPI_RANK_U16 || e == EPI_RANK_U16_MIRROR || e == EPI_RANK_U16_PWL || e == EPI_RANK_U16_MIRROR_PWL;
}
__host__ __device__ constexpr bool epi_is_mirror(int e) { return e == EPI_RANK_U16_MIRROR || e == EPI_RANK_U16_MIRROR_PWL; }
__host__ __device__ constexpr bool epi_is_pwl(int e) { return e == EPI_RANK_U16_PWL || e == EPI_RANK_U16_MIRROR_PWL; }
// The normaliser-layout epilogue with 8 epilogue warps is the software-pipelined one (two staging tiles per warp).
__host__ __device__ constexpr bool epi_is_pipelined_mirror(int e, int ne) { return epi_is_mirror(e) && ne == 8; }
__host__ __device__ constexpr int epi_staging_bufs(int e, int ne) { return epi_is_pipelined_mirror(e, ne) ? 2 : 1; }
// does the instance use the 32 KB auxiliary shared-memory region (see PairSmem)?
__host__ __device__ constexpr bool epi_needs_aux32(int e) { return epi_is_rank(e) || e == EPI_BF16_SPLIT; }
// Accumulator chunks (32 columns each) fetched per tcgen05.wait::ld by the row-per-lane epilogues.  Measured (round
// 2, B200): fetching the warp's whole strip at once (4 chunks, stage released before the stores) does NOT pay outside
// the top-k epilogue — fp32 logits 1.24 vs 1.18 ms (the per-chunk loop overlaps the next TMEM load with the previous
// store), GEMM 1 unchanged, EPI_LINEAR with 2 chunks spills and slows the production encoder by 4 % — so it is 1.
__host__ __device__ constexpr int epi_burst(int e, int ne) {
  return (e == EPI_BF16_SPLIT && ne == 8) ? 4 : 1;
}

// Shared-memory plan for a kernel instance with NE epilogue warps (staging is per warp, so more epilogue warps
// trade one B stage for staging space).
// kAux32: the instance needs the 32 KB auxiliary region behind the staging tiles (rank table; GEMM 1's wide 4 KB tiles).
// The other epilogues only keep a second 2 KB staging tile per warp there, which leaves room for a FOURTH B stage with 8
// epilogue warps: the B ring is what bounds the tensor-bound instances (a stage comes back ~1,000 cycles after it was
// freed; 3 stages of 256 MMA cycles each keep the tensor pipe 61 % busy, 4 stages 81 %).
template <int NE, int NBUF = 1, bool kAux32 = true>
struct PairSmem {
  static constexpr int kBStages = (NE * NBUF > 8) ? 2 : ((kAux32 || NE > 8) ? 3 : 4);
  static constexpr int kA = 0;
  static constexpr int kB = kA + kMaxAPanels * kPanelBytes;
  static constexpr int kStaging = kB + kBStages * kPanelBytes;
  static constexpr int kLut = kStaging + NE * NBUF * kStagingBytesPerWarp;
  static constexpr int kBar = kLut + (kAux32 ? kRankLutEntries * 4 : NE * kStagingBytesPerWarp);
  static constexpr int kTotal = kBar + 640;  // mbarriers, TMEM slot, dynamic-scheduler task ring, per-warp residual barriers, B ring
  static constexpr int kBytes = kTotal + 1024;  // + slack for manual 1024-byte alignment
  static constexpr int kThreads = (kFirstEpiWarp + NE) * 32;
  static_assert(kBytes <= 232448, "exceeds 227 KB of shared memory per CTA");
};

struct PairScoreParams {
  int L;          // outcomes in this launch
  int rows;       // valid output rows
  int cols;       // valid output columns
  int kb;         // 64-wide K blocks per precision term (D / 64)
  int nterm;      // 1 (bf16) or 3 (bf16x3)
  int msub;       // 128-row sub-tiles per CTA tile
  int a_batched;  // A's batch coordinate is l (else 0)
  int b_batched;
  int n_blocks;
  int m_blocks;
  int nchunk;
  int chunks_per_row;
  int num_tasks;
  int use_tma_store;   // tmOut is valid (all modes; LINEAR: the fp32 output)
  int use_tma_store2;  // tmOut2 is valid (LINEAR: the bf16 operand output)
  int lo_col_offset;  // EPI_BF16_SPLIT: column offset of the lo half in the output rows
  int write_lo;
  void* out;          // direct-store path
  long long out_ld;   // elements per output row
  long long out_batch_stride;
  const uint32_t* lut;  // [L, kRankLutEntries]
  const float* affine;  // [L, 2]
  // ---- streamed-A mode (K too large for the resident A buffer): A and B panels share the stage ring, msub = 1
  int stream_a;
  int k_pad;  // elements between the hi and lo halves of an operand row (bf16x3), = padded K
  // ---- EPI_LINEAR (the fusion encoder's nn.Linear layers):  y = act(acc + bias) [+ residual]
  const float* bias;      // [cols] or NULL
  const float* residual;  // fp32 [rows, res_ld] or NULL (may alias out_f32)
  long long res_ld;
  float* out_f32;  // fp32 [rows, out_ld] or NULL
  __nv_bfloat16* out_bf16;  // bf16 [rows, bf16_ld] (hi at column n, lo at column bf16_lo_off + n) or NULL
  long long bf16_ld;
  int bf16_lo_off;
  int act;  // 0 none, 1 relu, 2 exact-erf gelu
  int wide_store;  // EPI_BF16_SPLIT: tmOut has 64-column boxes (128-byte rows, 128B swizzle): one 4 KB store per two chunks
  int res_tma;  // the residual IS the fp32 output tensor (in-place residual stream) and it is TMA-addressable: the
                // residual tile is fetched with a TMA load through tmOut into the staging tile the result is stored from
  // ---- EPI_TOPK: append every score >= topk_thresh[l] to the outcome's candidate list (no dense output at all)
  const float* topk_thresh;       // [L]
  unsigned int* topk_count;       // [L] atomic counters (may exceed topk_cap: overflow is detected by the caller)
  unsigned long long* topk_cand;  // [L, topk_cap]  (score bits << 32 | row * cols + col)
  int topk_cap;
  unsigned int* sched_counter;  // zeroed by the host before the launch: dynamic task scheduler (NULL: static deal)
  int mirror;      // EPI_RANK_U16 with lower_only: also write each rank at [col, row]; diagonal = 0
  int lower_only;  // keep only row > col (unordered pairs of one catalogue) and skip column blocks above the diagonal
  int a_reuse;     // A is shared by all outcomes (GEMM 1: z . W_l): tasks are ordered row-block-major and dealt in
                   // contiguous per-CTA ranges so that the resident A panels are loaded once per row block, not per task
  int debug_epi;   // measurement knob of the pipelined normaliser-layout epilogue: 1 = no look-ups, 2 = no stores (invalid output)
  int store_evict_first;  // pipelined normaliser-layout epilogue: L2 evict-first hint on the rank stores (long rows)
  int packed;      // normaliser layout without the mirror image: every 32x32 chunk with row >= col goes ONCE, as a
                   // contiguous 2 KB tile, to tile slot bi*(bi+1)/2 + bj of the outcome (MDG_PAIRS_PACKED_TILES)
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// Rank lookup against the LUT staged in shared memory (same arithmetic as rank_lookup_raw).  mul.hi keeps the
// bucket extraction on the (idle) FMA pipe instead of the ALU pipe the epilogue is bound by.
__device__ __forceinline__ uint32_t rank_lookup_smem(uint32_t lut_smem, float x, float scale, float bias) {
  const uint32_t kb = rank_key_bits(x, scale, bias);
  uint32_t bucket;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(bucket) : "r"(kb), "r"(1u << 28));  // kb >> 4
  return rank_finish(lds_u32(lut_smem + bucket * 4), kb);
}

// Histogram-CDF table (rank_pwl_raw): lane c reads copy c of the 256-entry table, i.e. always bank c.
// `lut_lane` = table base + 4 * lane.  Piece extraction and addressing stay on the FMA pipe (mul.hi / mad).
__device__ __forceinline__ uint32_t rank_lookup_pwl_smem(uint32_t lut_lane, float x, float scale, float bias) {
  const uint32_t kb = rank_key_bits(x, scale, bias);
  uint32_t addr;
  asm("{\n\t"
      ".reg .u32 pc;\n\t"
      "mul.hi.u32 pc, %1, %2;\n\t"       // kb >> 9
      "mad.lo.u32 %0, pc, 128, %3;\n\t"  // 32 copies x 4 B per piece
      "}\n"
      : "=r"(addr)
      : "r"(kb), "r"(1u << (32 - kRankPwlSubBits)), "r"(lut_lane));
  return rank_pwl_finish(lds_u32(addr), kb);
}
template <bool PWL>
__device__ __forceinline__ uint32_t rank_lookup_epi(uint32_t lut_smem, uint32_t lut_lane, float x, float scale,
                                                    float bias) {
  if constexpr (PWL) return rank_lookup_pwl_smem(lut_lane, x, scale, bias);
  else return rank_lookup_smem(lut_smem, x, scale, bias);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct TaskCoord {
  int l, m0, nb0, nb1;
};
__device__ __forceinline__ TaskCoord decode_task(const PairScoreParams& p, int t) {
  TaskCoord c;
  int mb, ch;
  if (p.a_reuse) {  // row-block-major: consecutive tasks share the A operand
    const int per_mb = p.L * p.chunks_per_row;
    mb = t / per_mb;
    const int rem = t - mb * per_mb;
    c.l = rem / p.chunks_per_row;
    ch = rem - c.l * p.chunks_per_row;
  } else {
    const int per_l = p.m_blocks * p.chunks_per_row;
    c.l = t / per_l;
    const int rem = t - c.l * per_l;
    mb = rem / p.chunks_per_row;
    ch = rem - mb * p.chunks_per_row;
  }
  if (p.lower_only) mb = p.m_blocks - 1 - mb;  // largest row blocks (most tiles below the diagonal) first
  c.m0 = mb * kBM * p.msub;
  c.nb0 = ch * p.nchunk;
  c.nb1 = min(c.nb0 + p.nchunk, p.n_blocks);
  if (p.lower_only) {  // column blocks starting at or beyond the last row of this row block hold no row > col pair
    const int last_row = c.m0 + kBM * p.msub - 1;
    c.nb1 = min(c.nb1, last_row / kBN + 1);
    if (c.nb1 < c.nb0) c.nb1 = c.nb0;
  }
  return c;
}

// One warp's 32 rows x 64 bytes: registers -> 64B-swizzled staging -> TMA store (one bulk group per chunk).  Two
// staging buffers alternate so a store only waits for the store before the previous one to have read its buffer.
struct StagedStore {
  uint32_t buf[2];  // this warp's 2 KB staging buffers (buf[1] == buf[0]: single-buffered)
  uint32_t row_off;  // lane * 64
  uint32_t swz;      // (lane >> 1) & 3
  int cur;
  int pending;       // committed bulk groups not yet known to have been read
  // wait until buf[cur] may be overwritten; returns its shared-memory address
  __device__ __forceinline__ uint32_t acquire(int lane) {
    const bool dbl = buf[0] != buf[1];
    if (pending >= (dbl ? 2 : 1)) {
      if (lane == 0) {
        if (dbl) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
      }
      __syncwarp();
      pending = dbl ? 1 : 0;
    }
    return buf[cur];
  }
  // the warp has filled buf[cur]: hand it to the TMA store engine
  __device__ __forceinline__ void commit(const CUtensorMap* tm, int c0, int c1, int c2, int lane) {
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_store_3d(tm, buf[cur], c0, c1, c2);
      tma_store_commit();
    }
    ++pending;
    if (buf[0] != buf[1]) cur ^= 1;
  }
  __device__ __forceinline__ void store(const CUtensorMap* tm, const uint32_t (&w)[16], int c0, int c1, int c2,
                                        int lane) {
    const uint32_t base = acquire(lane) + row_off;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      st_shared_v4(base + ((static_cast<uint32_t>(ch) ^ swz) << 4), w[4 * ch], w[4 * ch + 1], w[4 * ch + 2],
                   w[4 * ch + 3]);
    commit(tm, c0, c1, c2, lane);
  }
};

// Two 2 KB staging tiles per warp used alternately: a tile is rewritten two bulk stores after the one that read it.
struct StagingRing2 {
  uint32_t buf[2];
  int cur;
  __device__ __forceinline__ uint32_t acquire(int lane) {  // the store before the previous one has read its tile
    if (elect_one()) tma_store_wait_read<1>();
    __syncwarp();
    return buf[cur];
  }
  __device__ __forceinline__ void commit(const CUtensorMap* tm, int c0, int c1, int c2, int lane) {
    fence_proxy_async_smem();
    __syncwarp();
    if (elect_one()) {
      tma_store_3d(tm, buf[cur], c0, c1, c2);
      tma_store_commit();
    }
    cur ^= 1;
  }
};

template <int EPI, int NE>
__global__ void __launch_bounds__(PairSmem<NE>::kThreads, 1)
pair_score_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                  const __grid_constant__ PairScoreParams p) {
  using SM = PairSmem<NE, epi_staging_bufs(EPI, NE), epi_needs_aux32(EPI)>;
  constexpr int kBStages = SM::kBStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw_addr);

  const uint32_t sA = base + SM::kA, sB = base + SM::kB, sStaging = base + SM::kStaging, sLut = base + SM::kLut,
                 sBar = base + SM::kBar;
  const uint32_t bar_a_full = sBar, bar_a_empty = sBar + 8;
  auto bar_b_full = [&](int i) { return sBar + 512 + 8 * i; };   // up to 4 stages each (behind the residual barriers)
  auto bar_b_empty = [&](int i) { return sBar + 544 + 8 * i; };
  auto bar_t_full = [&](int i) { return sBar + 64 + 8 * i; };
  auto bar_t_empty = [&](int i) { return sBar + 80 + 8 * i; };
  const uint32_t tmem_slot = sBar + 96;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = lane_id();

  if (threadIdx.x == 0) {
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, 1);
    for (int i = 0; i < kBStages; ++i) {
      mbar_init(bar_b_full(i), 1);
      mbar_init(bar_b_empty(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_t_full(i), 1);
      mbar_init(bar_t_empty(i), NE);
    }
    for (int i = 0; i < 4; ++i) {  // dynamic-scheduler task ring: full (producer -> consumers), empty (1 MMA + NE)
      mbar_init(sBar + 128 + 8 * i, 1);
      mbar_init(sBar + 160 + 8 * i, 1 + NE);
    }
    mbar_init(sBar + 224, 1);  // rank-table bulk load (pipelined normaliser-layout epilogue)
    for (int i = 0; i < 2 * NE; ++i) mbar_init(sBar + 256 + 8 * i, 1);  // EPI_LINEAR: residual tile loads, two per warp
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.use_tma_store) tma_prefetch_desc(&tmOut);
    if (p.use_tma_store2) tma_prefetch_desc(&tmOut2);
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + SM::kBar + 96);
  // PDL: everything above overlapped the predecessor's tail; from here on its output (operands, task counter) is read
  pdl_wait();
  pdl_launch_dependents();

  // Task order: outcome-major.  Static mode deals tasks round-robin; dynamic mode (sched_counter != NULL) hands them
  // out in the same order through an atomic counter (needed when tasks are uneven, e.g. lower-triangle mode).
  // Either way the CTAs running at any moment cover a few adjacent outcomes, so their output tiles stay within a
  // compact address range (measured: contiguous per-CTA task ranges are 20% slower) and they share z_cols via L2.
  const uint32_t bar_task_full0 = sBar + 128, bar_task_empty0 = sBar + 160, task_ids0 = sBar + 192;
  const bool dyn = p.sched_counter != nullptr;
  struct TaskIter {
    bool dyn;
    int t, step, end, n;  // static: t advances by step; n = tasks seen (ring position in dynamic mode)
    uint32_t full0, empty0, ids0;
    // consumer side: next task id or -1
    __device__ __forceinline__ int next_consumer(int lane) {
      if (!dyn) {
        const int cur = t;
        t += step;
        return cur < end ? cur : -1;
      }
      const int slot = n % kTaskRing;
      mbar_wait(full0 + 8 * slot, (n / kTaskRing) & 1, 7);
      uint32_t id;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(id) : "r"(ids0 + 4 * slot) : "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
      ++n;
      return static_cast<int>(id) < end ? static_cast<int>(id) : -1;
    }
    // producer side (whole warp): fetch the next task from the global counter and publish it to the consumers
    __device__ __forceinline__ int next_producer(unsigned int* counter) {
      if (!dyn) {
        const int cur = t;
        t += step;
        return cur < end ? cur : -1;
      }
      const int slot = n % kTaskRing;
      mbar_wait(empty0 + 8 * slot, ((n / kTaskRing) & 1) ^ 1, 8);
      unsigned int id = 0;
      if (elect_one()) {
        id = atomicAdd(counter, 1u);
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(ids0 + 4 * slot), "r"(id) : "memory");
        mbar_arrive(full0 + 8 * slot);
      }
      id = __shfl_sync(0xffffffffu, id, 0);  // elect.sync picks lane 0 of a converged warp
      ++n;
      return static_cast<int>(id) < end ? static_cast<int>(id) : -1;
    }
  };
  TaskIter tasks;
  tasks.dyn = dyn;
  tasks.t = static_cast<int>(blockIdx.x);
  tasks.step = static_cast<int>(gridDim.x);
  tasks.end = p.num_tasks;
  if (p.a_reuse && !dyn) {  // contiguous range per CTA
    const int per = (p.num_tasks + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    tasks.t = static_cast<int>(blockIdx.x) * per;
    tasks.step = 1;
    tasks.end = min(tasks.t + per, p.num_tasks);
  }
  tasks.n = 0;
  tasks.full0 = bar_task_full0;
  tasks.empty0 = bar_task_empty0;
  tasks.ids0 = task_ids0;

  const int kb = p.kb;
  const int n_apanels = (p.nterm == 1) ? p.msub * kb : 2 * kb;
  const int kb_b = (p.nterm == 1) ? kb : 2 * kb;

  if (warp == 0) {
    // ============================================================ TMA producer (whole warp loops, one lane issues)
    {
      int stage = 0;
      uint32_t b_phase = 0;
      int it = 0;  // executed tasks
      int na = 0;  // tasks so far that (re)loaded the A panels: phase counter of the two A barriers
      int last_m0 = -1;
      if (p.stream_a) {
        const int ksteps = (p.nterm == 1) ? kb : 3 * kb;
        for (int t = tasks.next_producer(p.sched_counter); t >= 0; t = tasks.next_producer(p.sched_counter)) {
          const TaskCoord c = decode_task(p, t);
          for (int nb = c.nb0; nb < c.nb1; ++nb) {
            for (int s = 0; s < ksteps; ++s) {
              const int term = s / kb, k = s - term * kb;  // 0: hi*hi, 1: lo*hi, 2: hi*lo
              mbar_wait(bar_b_empty(stage), b_phase ^ 1, 2);
              if (elect_one()) {
                mbar_arrive_expect_tx(bar_b_full(stage), 2 * kPanelBytes);
                tma_load_3d(sA + stage * kPanelBytes, &tmA, bar_b_full(stage), (term == 1 ? p.k_pad : 0) + k * kBK,
                            c.m0, p.a_batched ? c.l : 0);
                tma_load_3d(sB + stage * kPanelBytes, &tmB, bar_b_full(stage), (term == 2 ? p.k_pad : 0) + k * kBK,
                            nb * kBN, p.b_batched ? c.l : 0);
              }
              __syncwarp();
              if (++stage == kBStages) {
                stage = 0;
                b_phase ^= 1;
              }
            }
          }
        }
      } else
      for (int t = tasks.next_producer(p.sched_counter); t >= 0; t = tasks.next_producer(p.sched_counter), ++it) {
        const TaskCoord c = decode_task(p, t);
        if (c.nb0 >= c.nb1) {  // a column chunk entirely above the diagonal (normaliser layout): nothing to load
          --it;
          continue;
        }
        // GEMM 1 (a_reuse): consecutive tasks of a CTA share the row block, so its z panels stay in shared memory and
        // neither warp touches the A barriers — the B ring keeps streaming across the task boundary instead of
        // draining at every task (one tile per task there: the drain was 2/3 of GEMM 1's time)
        const bool a_resident = p.a_reuse && !p.a_batched && it > 0 && c.m0 == last_m0;
        last_m0 = c.m0;
        if (!a_resident) {
          mbar_wait(bar_a_empty, (na & 1) ^ 1, 1);  // every MMA that read the previous panels has completed
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_a_full, n_apanels * kPanelBytes);
            for (int pn = 0; pn < n_apanels; ++pn) {
              int row, kc;
              if (p.nterm == 1) {
                int ms = pn / kb;
                row = c.m0 + ms * kBM;
                kc = (pn - ms * kb) * kBK;
              } else {
                row = c.m0;
                kc = pn * kBK;
              }
              tma_load_3d(sA + pn * kPanelBytes, &tmA, bar_a_full, kc, row, p.a_batched ? c.l : 0);
            }
          }
          ++na;
        }
        __syncwarp();
        for (int nb = c.nb0; nb < c.nb1; ++nb) {
          for (int kbi = 0; kbi < kb_b; ++kbi) {
            mbar_wait(bar_b_empty(stage), b_phase ^ 1, 2);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar_b_full(stage), kPanelBytes);
              tma_load_3d(sB + stage * kPanelBytes, &tmB, bar_b_full(stage), kbi * kBK, nb * kBN,
                          p.b_batched ? c.l : 0);
            }
            __syncwarp();
            if (++stage == kBStages) {
              stage = 0;
              b_phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ UMMA issuer
    // The whole warp runs the (warp-uniform) loop; ONE elected lane — the same every time — issues tcgen05.mma and
    // tcgen05.commit.  Descriptors are uniform values: the constant high word and a per-panel low word (start
    // address >> 4) that only needs "+ 2" per K step of 16 elements.
    {
      const uint32_t idesc = umma_idesc_bf16_f32(kBM, kBN);
      const uint64_t desc_hi = umma_desc_kmajor_sw128(0) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_kmajor_sw128(0));  // LBO field
      auto desc_of = [&](uint32_t smem_addr) {
        return desc_hi | static_cast<uint64_t>(desc_lo0 | ((smem_addr & 0x3FFFFu) >> 4));
      };
      const uint32_t desc_hi32 = static_cast<uint32_t>(desc_hi >> 32);
      const uint32_t a_lo0 = desc_lo0 | ((sA & 0x3FFFFu) >> 4), b_lo0 = desc_lo0 | ((sB & 0x3FFFFu) >> 4);
      int stage = 0;
      uint32_t b_phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      int it = 0;
      int na = 0, last_m0 = -1;  // A-barrier phase counter / row block of the previous task (as in the producer)
      if (p.stream_a) {
        const int ksteps = (p.nterm == 1) ? kb : 3 * kb;
        for (int t = tasks.next_consumer(lane); t >= 0; t = tasks.next_consumer(lane)) {
          const TaskCoord c = decode_task(p, t);
          for (int nb = c.nb0; nb < c.nb1; ++nb) {
            mbar_wait(bar_t_empty(acc_stage), acc_phase ^ 1, 4);
            tc_fence_after_sync();
            const uint32_t d = tmem_base + static_cast<uint32_t>(acc_stage * 2 * kBN);
            for (int s = 0; s < ksteps; ++s) {
              mbar_wait(bar_b_full(stage), b_phase, 5);
              tc_fence_after_sync();
              const uint64_t adesc = desc_of(sA + stage * kPanelBytes);
              const uint64_t bdesc = desc_of(sB + stage * kPanelBytes);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k)
                  umma_bf16(d, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                            (s > 0 || k > 0) ? 1u : 0u);
                umma_commit(bar_b_empty(stage));
              }
              __syncwarp();
              if (++stage == kBStages) {
                stage = 0;
                b_phase ^= 1;
              }
            }
            if (elect_one()) umma_commit(bar_t_full(acc_stage));
            __syncwarp();
            acc_stage ^= 1;
            if (acc_stage == 0) acc_phase ^= 1;
          }
        }
      } else
      for (int t = tasks.next_consumer(lane); t >= 0; t = tasks.next_consumer(lane), ++it) {
        const TaskCoord c = decode_task(p, t);
        if (c.nb0 >= c.nb1) {  // empty task: skipped by the producer as well
          --it;
          continue;
        }
        const bool a_resident = p.a_reuse && !p.a_batched && it > 0 && c.m0 == last_m0;  // as in the producer
        last_m0 = c.m0;
        if (!a_resident) {
          if (it > 0) {  // hand the old panels back: tracks every MMA issued so far
            if (elect_one()) umma_commit(bar_a_empty);
            __syncwarp();
          }
          mbar_wait(bar_a_full, na & 1, 3);
          ++na;
        }
        for (int nb = c.nb0; nb < c.nb1; ++nb) {
          mbar_wait(bar_t_empty(acc_stage), acc_phase ^ 1, 4);
          tc_fence_after_sync();
          const uint32_t d0 = tmem_base + static_cast<uint32_t>(acc_stage * 2 * kBN);
          for (int kbi = 0; kbi < kb_b; ++kbi) {
            mbar_wait(bar_b_full(stage), b_phase, 5);
            tc_fence_after_sync();
            // descriptor low words (start address >> 4 in the low 14 bits): one panel = kPanelBytes >> 4 = 1024 units
            const uint32_t b_lo = b_lo0 + static_cast<uint32_t>(stage) * (kPanelBytes >> 4);
            if (p.nterm == 1) {
              // one bf16 term: msub row sub-tiles share the B panel, one accumulator each
              const uint32_t a0 = a_lo0 + static_cast<uint32_t>(kbi) * (kPanelBytes >> 4);
              const uint32_t a1 = a0 + static_cast<uint32_t>(kb) * (kPanelBytes >> 4);
              const uint32_t acc = kbi == 0 ? 0u : 1u;
              if (elect_one()) {
                umma_bf16_lo(d0, a0, b_lo, desc_hi32, idesc, acc);
                umma_bf16_lo(d0, a0 + 2, b_lo + 2, desc_hi32, idesc, 1u);
                umma_bf16_lo(d0, a0 + 4, b_lo + 4, desc_hi32, idesc, 1u);
                umma_bf16_lo(d0, a0 + 6, b_lo + 6, desc_hi32, idesc, 1u);
                if (p.msub == 2) {
                  umma_bf16_lo(d0 + kBN, a1, b_lo, desc_hi32, idesc, acc);
                  umma_bf16_lo(d0 + kBN, a1 + 2, b_lo + 2, desc_hi32, idesc, 1u);
                  umma_bf16_lo(d0 + kBN, a1 + 4, b_lo + 4, desc_hi32, idesc, 1u);
                  umma_bf16_lo(d0 + kBN, a1 + 6, b_lo + 6, desc_hi32, idesc, 1u);
                }
                umma_commit(bar_b_empty(stage));  // frees this B stage once the MMAs above have read it
              }
            } else {
              // bf16x3: B = hi -> A_hi[k] then A_lo[k]; B = lo -> A_hi[k]; all into the same accumulator
              const bool b_is_hi = kbi < kb;
              const uint32_t a0 = a_lo0 + static_cast<uint32_t>(b_is_hi ? kbi : kbi - kb) * (kPanelBytes >> 4);
              const uint32_t a1 = a0 + static_cast<uint32_t>(kb) * (kPanelBytes >> 4);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k)
                  umma_bf16_lo(d0, a0 + 2 * k, b_lo + 2 * k, desc_hi32, idesc, (kbi == 0 && k == 0) ? 0u : 1u);
                if (b_is_hi) {
#pragma unroll
                  for (int k = 0; k < kBK / kUmmaK; ++k) umma_bf16_lo(d0, a1 + 2 * k, b_lo + 2 * k, desc_hi32, idesc, 1u);
                }
                umma_commit(bar_b_empty(stage));
              }
            }
            __syncwarp();
            if (++stage == kBStages) {
              stage = 0;
              b_phase ^= 1;
            }
          }
          if (elect_one()) umma_commit(bar_t_full(acc_stage));  // accumulators of this tile complete
          __syncwarp();
          acc_stage ^= 1;
          if (acc_stage == 0) acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ============================================================ epilogue warps
    const int ew = warp - kFirstEpiWarp;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int grp = ew >> 2;    // which slice of the quadrant's msub*128 accumulator columns
    constexpr int kGroups = NE / 4;
    const int cols_per_warp = p.msub * kBN / kGroups;
    const int ms = (grp * cols_per_warp) / kBN;
    const int col_begin = (grp * cols_per_warp) % kBN;
    const int col_end = col_begin + cols_per_warp;
    if constexpr (epi_is_pipelined_mirror(EPI, NE)) {
      // ---------------------------------------------------------------------------------------------------------
      // Normaliser-layout epilogue (normalize_scores.py:67-70), software-pipelined.  8 epilogue warps; a warp owns a
      // strip of 32 rows x (2 or 4) 32-column chunks per tile.  While the 1024 look-ups of chunk c run, the
      // accumulator fragment of chunk c+1 (the next chunk of the strip, or the first one of the next tile of the task)
      // is already being fetched from TMEM into a second register set, and the two 2 KB staging tiles of the warp
      // alternate, so neither the tcgen05.ld round trip nor the bulk store's shared-memory read is ever waited for
      // in steady state: the shared-memory pipe (LUT gathers + stmatrix) is the only resource the loop is bound by.
      // A chunk with row > col is ranked once and stored twice (plain tile at [row, col], stmatrix.trans tile at
      // [col, row]); the diagonal chunk builds both masked tiles, ORs them in shared memory and is stored once.
      // p.packed: no mirror image — every chunk goes once to its slot of the packed tile array.
      // ---------------------------------------------------------------------------------------------------------
      StagingRing2 ring;
      ring.buf[0] = sStaging + ew * 2 * kStagingBytesPerWarp;
      ring.buf[1] = ring.buf[0] + kStagingBytesPerWarp;
      ring.cur = 0;
      const uint32_t sLutLane = sLut + static_cast<uint32_t>(lane) * 4;
      const int fr = lane >> 2, fc = (lane & 3) * 2;  // fragment row / first column of this thread
      const int n_chunks = cols_per_warp / 32;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      int cur_l = -1;
      float scale = 0.f, bias = 0.f;
      const uint32_t bar_lut = sBar + 224;
      uint32_t lut_phase = 0;
      // L2 eviction priority of the rank stores.  With long rows (N >= 8192: 32 box rows of a tile are >= 0.5 MB apart)
      // lines that sit in L2 until capacity evicts them go out to DRAM in an order that scatters over pages; marking
      // them evict-first drains them while their neighbours are still arriving (measured, 12 GB of output: N = 20,000
      // 3.09 -> 2.91 ms, N = 16,384 2.91 -> 2.83 ms; N <= 6,144 unchanged, hence the threshold on the host side)
      const uint64_t store_policy = p.store_evict_first ? kL2EvictFirst : kL2EvictNormal;

      // ranks of one chunk: P[hf][i][s] = (row 16hf + 8s + fr, cols 8i + fc, +1) packed as b16x2
      auto lookup_chunk = [&](const uint32_t (&v)[2][16], uint32_t (&P)[2][4][2]) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const uint32_t ra = rank_lookup_epi<epi_is_pwl(EPI)>(sLut, sLutLane, __uint_as_float(v[hf][4 * i + 2 * s2]), scale, bias);
              const uint32_t rb = rank_lookup_epi<epi_is_pwl(EPI)>(sLut, sLutLane, __uint_as_float(v[hf][4 * i + 2 * s2 + 1]), scale, bias);
              P[hf][i][s2] = __byte_perm(ra, rb, 0x5410);
            }
      };
      auto fill_plain = [&](uint32_t dst, const uint32_t (&P)[2][4][2]) {  // staging row = chunk row, 64B swizzle
        const int i = lane >> 3, k = lane & 7;  // this thread addresses row k of matrix i (= column group)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int R = 16 * hf + 8 * s2 + k;
            stmatrix_x4(dst + R * 64 + ((i ^ ((R >> 1) & 3)) << 4), P[hf][0][s2], P[hf][1][s2], P[hf][2][s2], P[hf][3][s2]);
          }
      };
      auto fill_trans = [&](uint32_t dst, const uint32_t (&P)[2][4][2]) {  // staging row = chunk column
        const int m = lane >> 3, k = lane & 7;  // stored-row k of matrix m (= rows 8m .. 8m+7 of the chunk)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 8 * i + k;
          stmatrix_x4_trans(dst + c * 64 + ((m ^ ((c >> 1) & 3)) << 4), P[0][i][0], P[0][i][1], P[1][i][0], P[1][i][1]);
        }
      };
      // look-ups + stores of one chunk whose accumulator fragment is in v
      auto process = [&](const uint32_t (&v)[2][16], int l, int row0, int n0) {
        uint32_t P[2][4][2];
        if (p.debug_epi & 1) {  // measurement only: the store path without the look-ups
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int s2 = 0; s2 < 2; ++s2) P[hf][i][s2] = __byte_perm(v[hf][4 * i + 2 * s2], v[hf][4 * i + 2 * s2 + 1], 0x7632);
        } else {
          lookup_chunk(v, P);
        }
        if (p.debug_epi & 2) {  // measurement only: the look-ups without the store path
          uint32_t acc = 0;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc ^= P[hf][i][0] ^ P[hf][i][1];
          if (acc == 0x12345678u && lane == 33) st_shared_u32(ring.buf[0], acc);  // keeps the look-ups alive
          return;
        }
        const int bi = row0 >> 5, bj = n0 >> 5;
        const int slot_row = (bi * (bi + 1) / 2 + bj) * 32;  // packed layout: first row of this chunk's tile
        if (n0 != row0) {
          if (!p.packed) {
            // both tiles of the chunk under ONE proxy fence and ONE bulk group: the previous chunk's group was
            // committed before this chunk's 1024 look-ups, so waiting for its reads costs nothing here
            if (elect_one()) tma_store_wait_read<0>();
            __syncwarp();
            fill_trans(ring.buf[0], P);
            fill_plain(ring.buf[1], P);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              tma_store_3d_hint(&tmOut2, ring.buf[0], row0, n0, l, store_policy);  // [l, n0.., row0..]
              tma_store_3d_hint(&tmOut, ring.buf[1], n0, row0, l, store_policy);
              tma_store_commit();
            }
          } else {
            fill_plain(ring.acquire(lane), P);
            ring.commit(&tmOut, 0, slot_row, l, lane);
          }
        } else {
          // diagonal chunk: keep col < row only (diagonal = 0, normalize_scores.py:69)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int s2 = 0; s2 < 2; ++s2) {
                const int r = 16 * hf + 8 * s2 + fr, c0 = 8 * i + fc;
                uint32_t w = P[hf][i][s2];
                if (c0 >= r) w &= 0xFFFF0000u;
                if (c0 + 1 >= r) w &= 0x0000FFFFu;
                P[hf][i][s2] = w;
              }
          if (p.packed) {
            fill_plain(ring.acquire(lane), P);
            ring.commit(&tmOut, 0, slot_row, l, lane);
          } else {
            // both staging tiles must be free: lower triangle in one, its transpose in the other, OR, one store
            if (elect_one()) tma_store_wait_read<0>();
            __syncwarp();
            const uint32_t d = ring.buf[ring.cur], t = ring.buf[ring.cur ^ 1];
            fill_plain(d, P);
            fill_trans(t, P);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q) {  // lane = staging row; the swizzle is the same in both tiles
              uint32_t a0, a1, a2, a3, b0, b1, b2, b3;
              ld_shared_v4(d + lane * 64 + q * 16, a0, a1, a2, a3);
              ld_shared_v4(t + lane * 64 + q * 16, b0, b1, b2, b3);
              st_shared_v4(d + lane * 64 + q * 16, a0 | b0, a1 | b1, a2 | b2, a3 | b3);
            }
            ring.commit(&tmOut, n0, row0, l, lane);
          }
        }
      };

      for (int t = tasks.next_consumer(lane); t >= 0; t = tasks.next_consumer(lane)) {
        const TaskCoord c = decode_task(p, t);
        if (c.nb0 >= c.nb1) continue;  // empty task (column chunk above the diagonal)
        if (c.l != cur_l) {
          // new outcome: one bulk copy (TMA, 32 KB) of its table once every warp is done with the old one
          named_bar_sync(1, NE * 32);
          if (ew == 0 && elect_one()) {
            mbar_arrive_expect_tx(bar_lut, kRankLutEntries * 4);
            bulk_load_1d(sLut, p.lut + static_cast<size_t>(c.l) * kRankLutEntries, kRankLutEntries * 4, bar_lut);
          }
          scale = __ldg(p.affine + 2 * c.l);
          bias = __ldg(p.affine + 2 * c.l + 1);
          cur_l = c.l;
          mbar_wait(bar_lut, lut_phase, 9);
          lut_phase ^= 1;
        }
        const int row0 = c.m0 + ms * kBM + quad * 32;  // first of this warp's 32 rows
        const bool rows_ok = row0 < p.rows;
        // cursor over (tile nb, chunk k) of this task; a tile is opened by waiting for its accumulators and closed
        // (handed back to the MMA warp) once every chunk of it that this warp needs has been fetched AND awaited —
        // fetch() is only called right after tcgen05.wait::ld, so nothing is in flight when it closes a tile
        int nb = c.nb0, k = 0;
        bool open = false;
        auto fetch = [&](uint32_t (&v)[2][16], int& n0_out) -> bool {
          while (nb < c.nb1) {
            if (!open) {
              mbar_wait(bar_t_full(acc_stage), acc_phase, 6);
              tc_fence_after_sync();
              open = true;
              k = 0;
            }
            if (k < n_chunks && rows_ok) {
              const int cc = col_begin + 32 * k;
              const int n0 = nb * kBN + cc;
              if (n0 < p.cols && n0 <= row0 + 31) {  // else: this and the later chunks of the tile hold no row >= col
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                       static_cast<uint32_t>((acc_stage * 2 + ms) * kBN + cc);
                tmem_ld_16x256b_x4(taddr, v[0]);
                tmem_ld_16x256b_x4(taddr + (16u << 16), v[1]);
                n0_out = n0;
                ++k;
                return true;
              }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (elect_one()) mbar_arrive(bar_t_empty(acc_stage));
            acc_stage ^= 1;
            if (acc_stage == 0) acc_phase ^= 1;
            open = false;
            ++nb;
          }
          return false;
        };
        uint32_t va[2][16], vb[2][16];
        int n0a = 0, n0b = 0;
        bool have = fetch(va, n0a);
        while (have) {
          tmem_ld_wait();
          const bool have_b = fetch(vb, n0b);
          process(va, c.l, row0, n0a);
          if (!have_b) break;
          tmem_ld_wait();
          have = fetch(va, n0a);
          process(vb, c.l, row0, n0b);
        }
      }
      if (lane == 0) tma_store_wait_all<0>();
      __syncwarp();
    } else {
    StagedStore ss;
    ss.buf[0] = sStaging + ew * kStagingBytesPerWarp;
    // the LUT region is idle outside the rank epilogue: use it as a second staging buffer per warp
    ss.buf[1] = epi_is_rank(EPI) ? ss.buf[0] : sLut + ew * kStagingBytesPerWarp;
    ss.row_off = static_cast<uint32_t>(lane) * 64;
    ss.swz = static_cast<uint32_t>((lane >> 1) & 3);  // 64-byte swizzle: chunk ^= (row >> 1) & 3
    ss.cur = 0;
    ss.pending = 0;
    const uint32_t res_bar = sBar + 256 + 16 * ew;  // EPI_LINEAR: this warp's two residual-tile barriers
    uint32_t res_phase = 0;                         // bit hf = phase of barrier hf
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    int cur_l = -1;
    float scale = 0.f, bias = 0.f;

    for (int t = tasks.next_consumer(lane); t >= 0; t = tasks.next_consumer(lane)) {
      const TaskCoord c = decode_task(p, t);
      if (c.nb0 >= c.nb1) continue;  // empty task (column chunk above the diagonal)
      if (epi_is_rank(EPI) && c.l != cur_l) {
        named_bar_sync(1, NE * 32);  // everyone is done with the previous outcome's LUT
        const uint4* src = reinterpret_cast<const uint4*>(p.lut + static_cast<size_t>(c.l) * kRankLutEntries);
        const int tid = ew * 32 + lane;
#pragma unroll 4
        for (int i = tid; i < kRankLutEntries / 4; i += NE * 32) {
          uint4 v = __ldg(src + i);
          st_shared_v4(sLut + i * 16, v.x, v.y, v.z, v.w);
        }
        scale = __ldg(p.affine + 2 * c.l);
        bias = __ldg(p.affine + 2 * c.l + 1);
        cur_l = c.l;
        named_bar_sync(1, NE * 32);
      }
      float topk_thr = 0.f;
      if constexpr (EPI == EPI_TOPK) topk_thr = __ldg(p.topk_thresh + c.l);
      const int row0 = c.m0 + ms * kBM + quad * 32;  // first of this warp's 32 rows
      const int my_row = row0 + lane;
      for (int nb = c.nb0; nb < c.nb1; ++nb) {
        mbar_wait(bar_t_full(acc_stage), acc_phase, 6);
        tc_fence_after_sync();
        if constexpr (epi_is_mirror(EPI)) {
          // Normaliser-layout epilogue (normalize_scores.py:67-70).  Each 32x32 chunk with row > col is ranked once and
          // written twice: at [row, col] and, transposed, at [col, row].  The accumulator is read in the mma fragment
          // layout (tcgen05.ld.16x256b), because that is what stmatrix stores: four stmatrix.x4 put the chunk into the
          // 64B-swizzled staging tile of the plain TMA store and four stmatrix.x4.trans build the transposed tile —
          // 32 conflict-free shared-memory wavefronts per chunk instead of 16 + 64 for st.shared.v4 + 32 sub-word
          // st.shared.u16 from the row-per-lane layout (the shared-memory pipe is this kernel's bound).
          const uint32_t sLutLane = sLut + static_cast<uint32_t>(lane) * 4;
          const int fr = lane >> 2, fc = (lane & 3) * 2;  // fragment row / first column of this thread
          if (row0 < p.rows) {
            for (int cc = col_begin; cc < col_end; cc += 32) {
              const int n0 = nb * kBN + cc;
              if (n0 >= p.cols) break;
              if (n0 > row0 + 31) break;  // every (row, col) of this chunk has col > row
              uint32_t v[2][16];
              const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                     static_cast<uint32_t>((acc_stage * 2 + ms) * kBN + cc);
              tmem_ld_16x256b_x4(taddr, v[0]);
              tmem_ld_16x256b_x4(taddr + (16u << 16), v[1]);
              tmem_ld_wait();
              // P[hf][i][s]: ranks of (row 16hf + 8s + fr, cols 8i + fc, +1) packed as b16x2
              uint32_t P[2][4][2];
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                  for (int s = 0; s < 2; ++s) {
                    const uint32_t ra = rank_lookup_epi<epi_is_pwl(EPI)>(sLut, sLutLane, __uint_as_float(v[hf][4 * i + 2 * s]), scale, bias);
                    const uint32_t rb = rank_lookup_epi<epi_is_pwl(EPI)>(sLut, sLutLane, __uint_as_float(v[hf][4 * i + 2 * s + 1]), scale, bias);
                    P[hf][i][s] = __byte_perm(ra, rb, 0x5410);
                  }
              const bool direct = !p.use_tma_store || n0 == row0;  // diagonal chunk (or unaligned output): masked stores
              if (!direct) {
                // transposed tile: staging row = chunk column c, 16-byte piece m = rows 8m .. 8m+7, piece position
                // swizzled like the plain tile (64B swizzle: piece ^ ((row >> 1) & 3))
                {
                  const uint32_t tb = ss.acquire(lane);
                  const int m = lane >> 3, k = lane & 7;  // this thread addresses stored-row k of matrix m
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const int c = 8 * i + k;
                    stmatrix_x4_trans(tb + c * 64 + ((m ^ ((c >> 1) & 3)) << 4), P[0][i][0], P[0][i][1], P[1][i][0], P[1][i][1]);
                  }
                  ss.commit(&tmOut2, row0, n0, c.l, lane);  // [l, n0.., row0..]
                }
                {
                  const uint32_t nbuf = ss.acquire(lane);
                  const int i = lane >> 3, k = lane & 7;  // matrix i = column group, row k
#pragma unroll
                  for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                      const int R = 16 * hf + 8 * s + k;
                      stmatrix_x4(nbuf + R * 64 + ((i ^ ((R >> 1) & 3)) << 4), P[hf][0][s], P[hf][1][s], P[hf][2][s], P[hf][3][s]);
                    }
                  ss.commit(&tmOut, n0, row0, c.l, lane);
                }
              } else {
                uint16_t* ob = reinterpret_cast<uint16_t*>(p.out) + c.l * p.out_batch_stride;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                  for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                      for (int e = 0; e < 2; ++e) {
                        const int row = row0 + 16 * hf + 8 * s + fr, col = n0 + 8 * i + fc + e;
                        const uint16_t r = static_cast<uint16_t>(e ? (P[hf][i][s] >> 16) : (P[hf][i][s] & 0xFFFFu));
                        if (row < p.rows && col < p.cols) {
                          if (col < row) {
                            ob[static_cast<long long>(row) * p.out_ld + col] = r;
                            ob[static_cast<long long>(col) * p.out_ld + row] = r;
                          } else if (col == row) {
                            ob[static_cast<long long>(row) * p.out_ld + col] = 0;  // diagonal (normalize_scores.py:69)
                          }
                        }
                      }
              }
            }
          }
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_t_empty(acc_stage));
          acc_stage ^= 1;
          if (acc_stage == 0) acc_phase ^= 1;
          continue;
        }
        if constexpr (EPI == EPI_TOPK) {
          // Top-k epilogue.  (1) A tcgen05.ld issued while MMAs are queued only completes once the queue drains, so the
          // warp's whole slice of the tile (up to 4 chunks) is fetched with ALL loads in flight and ONE wait, and the
          // accumulator stage goes back to the MMA warp before anything is examined.  (2) Candidates are ~1e-4 of the
          // scores but ~20 % of the 1024-score chunks hold one, so the scan is a warp vote per column (uniform, almost
          // never taken branch) and the rare append is warp-aggregated: one atomicAdd per vote, not one scan loop
          // plus one atomic round trip per chunk.
          if (row0 < p.rows && p.topk_cap >= 0) {  // topk_cap < 0: mainloop-only timing (debug)
            uint32_t a[4][32];
            bool want[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int cc = col_begin + 32 * q, n0 = nb * kBN + cc;
              want[q] = cc < col_end && n0 < p.cols && !(p.lower_only && n0 > row0 + 31);
              if (want[q])
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                  static_cast<uint32_t>((acc_stage * 2 + ms) * kBN + cc), a[q]);
            }
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t_empty(acc_stage));
            const float thr = my_row < p.rows ? topk_thr : CUDART_INF_F;  // rows beyond the matrix never qualify
            // per-lane hit masks of the (up to) 4 chunks: 2 instructions per score, no branches
            uint32_t hm[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              hm[q] = 0;
              if (!want[q]) continue;
              const int n0 = nb * kBN + col_begin + 32 * q;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (__uint_as_float(a[q][j]) >= thr) hm[q] |= 1u << j;
              // columns this lane may report: inside the matrix and, for unordered pairs, strictly left of its row
              int jlim = p.cols - n0;
              if (p.lower_only) jlim = min(jlim, my_row - n0);
              if (jlim < 32) hm[q] &= jlim > 0 ? (1u << jlim) - 1u : 0u;
            }
            // Append loop: ONE copy of the code for the whole tile (128 inlined copies of an append block, or four
            // unrolled per-chunk loops, thrash the 32 KB instruction cache: 3-4x slower kernel).  Every pass takes
            // each lane's next hit; hits are ~1 per warp per tile, so this almost always runs once or not at all.
            while (__any_sync(0xffffffffu, (hm[0] | hm[1] | hm[2] | hm[3]) != 0)) {
              int idx = -1;
#pragma unroll
              for (int q = 3; q >= 0; --q)
                if (hm[q] != 0) idx = q * 32 + (__ffs(hm[q]) - 1);  // lowest chunk wins
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if ((idx >> 5) == q) hm[q] &= hm[q] - 1u;  // clear the bit just taken (idx = -1 matches no q)
              uint32_t bits = 0;
#pragma unroll
              for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (idx == q * 32 + j) bits = a[q][j];               // 128-way register select
              const bool h = idx >= 0;
              const uint32_t bal = __ballot_sync(0xffffffffu, h);
              const int leader = __ffs(bal) - 1;
              unsigned int base_slot = 0;
              if (lane == leader) base_slot = atomicAdd(p.topk_count + c.l, static_cast<unsigned int>(__popc(bal)));
              base_slot = __shfl_sync(0xffffffffu, base_slot, leader);
              if (h) {
                const unsigned int slot = base_slot + static_cast<unsigned int>(__popc(bal & ((1u << lane) - 1u)));
                const int col = nb * kBN + col_begin + idx;
                if (slot < static_cast<unsigned int>(p.topk_cap))
                  p.topk_cand[static_cast<size_t>(c.l) * p.topk_cap + slot] =
                      (static_cast<unsigned long long>(bits) << 32) |
                      static_cast<unsigned long long>(static_cast<unsigned int>(my_row) *
                                                      static_cast<unsigned int>(p.cols) + col);
              }
            }
          } else {
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t_empty(acc_stage));
          }
          acc_stage ^= 1;
          if (acc_stage == 0) acc_phase ^= 1;
          continue;
        }
        bool released = false;
        if (row0 < p.rows) {
          // A tcgen05.ld issued while MMAs are queued only completes once the queue drains (~1.2k cycles), so the
          // epilogues with registers to spare fetch kBurst chunks of the warp's strip per tcgen05.wait::ld instead of
          // paying one round trip per chunk, and hand the accumulator stage back to the MMA warp as soon as the last
          // burst has landed — before the stores.
          constexpr int kBurst = epi_burst(EPI, NE);
          for (int cc0 = col_begin; cc0 < col_end; cc0 += 32 * kBurst) {
            uint32_t vv[kBurst][32];
            if constexpr (kBurst > 1) {
#pragma unroll
              for (int q = 0; q < kBurst; ++q) {
                const int ccq = cc0 + 32 * q, n0q = nb * kBN + ccq;
                if (ccq < col_end && n0q < p.cols && !(p.lower_only && n0q > row0 + 31))
                  tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                    static_cast<uint32_t>((acc_stage * 2 + ms) * kBN + ccq), vv[q]);
              }
              tmem_ld_wait();
              if (cc0 + 32 * kBurst >= col_end) {  // every chunk of the tile this warp needs is in registers
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_t_empty(acc_stage));
                released = true;
              }
            }
            if constexpr (EPI == EPI_BF16_SPLIT && kBurst == 4) {
              if (p.wide_store) {
                // GEMM 1's output path was 40 % of its time (one fill + proxy fence + 2 KB store + commit per 32-column
                // chunk, serialised per warp): with the whole strip in registers two chunks form one 128-byte-row tile
                // (32 rows x 64 bf16, 128B swizzle) in the warp's 4 KB slice of the (idle) table region
                const uint32_t wbuf = sLut + static_cast<uint32_t>(ew) * 4096u;
                for (int part = 0; part < (p.write_lo ? 2 : 1); ++part) {
#pragma unroll
                  for (int pr = 0; pr < 2; ++pr) {
                    const int cc = cc0 + 64 * pr, n0 = nb * kBN + cc;
                    if (cc >= col_end || n0 >= p.cols) break;
                    if (lane == 0) tma_store_wait_read<0>();  // the previous tile has left the buffer
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                      const uint32_t (&vh)[32] = vv[2 * pr + h];
                      uint32_t pk[16];
#pragma unroll
                      for (int j = 0; j < 16; ++j) {
                        float a = __uint_as_float(vh[2 * j]), b = __uint_as_float(vh[2 * j + 1]);
                        if (part == 1) {
                          a -= __bfloat162float(__float2bfloat16_rn(a));
                          b -= __bfloat162float(__float2bfloat16_rn(b));
                        }
                        pk[j] = pack_bf16x2(a, b);
                      }
#pragma unroll
                      for (int c4 = 0; c4 < 4; ++c4)
                        st_shared_v4(wbuf + static_cast<uint32_t>(lane) * 128u +
                                         ((static_cast<uint32_t>(4 * h + c4) ^ static_cast<uint32_t>(lane & 7)) << 4),
                                     pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                      tma_store_3d(&tmOut, wbuf, n0 + part * p.lo_col_offset, row0, c.l);
                      tma_store_commit();
                    }
                    ss.pending = 1;  // the kernel's final wait covers these groups
                  }
                }
                continue;
              }
            }
#pragma unroll
          for (int q = 0; q < kBurst; ++q) {
            const int cc = cc0 + 32 * q;
            if (cc >= col_end) break;
            const int n0 = nb * kBN + cc;
            if (n0 >= p.cols) break;
            if (p.lower_only && n0 > row0 + 31) break;  // every (row, col) of this chunk has col > row
            uint32_t (&v)[32] = vv[q];
            if constexpr (kBurst == 1) {
              const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                     static_cast<uint32_t>((acc_stage * 2 + ms) * kBN + cc);
              tmem_ld_32x32(taddr, v);
              if constexpr (EPI != EPI_LINEAR) tmem_ld_wait();  // LINEAR overlaps its global loads with the TMEM load
            }

            if constexpr (epi_is_rank(EPI)) {  // full layout (the normaliser layout has its own branch above)
              uint32_t pk[16];
              const uint32_t sLutLane = sLut + static_cast<uint32_t>(lane) * 4;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t r0 = rank_lookup_epi<epi_is_pwl(EPI)>(sLut, sLutLane, __uint_as_float(v[2 * j]), scale, bias);
                const uint32_t r1 = rank_lookup_epi<epi_is_pwl(EPI)>(sLut, sLutLane, __uint_as_float(v[2 * j + 1]), scale, bias);
                pk[j] = __byte_perm(r0, r1, 0x5410);
              }
              if (p.use_tma_store) {
                ss.store(&tmOut, pk, n0, row0, c.l, lane);
              } else if (my_row < p.rows) {
                uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + c.l * p.out_batch_stride +
                              static_cast<long long>(my_row) * p.out_ld + n0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n0 + j < p.cols) o[j] = static_cast<uint16_t>((j & 1) ? (pk[j >> 1] >> 16) : (pk[j >> 1] & 0xFFFFu));
              }
            } else if constexpr (EPI == EPI_LINEAR) {
              // y = act(acc + bias) (+ residual); fp32 and/or bf16 (hi | lo) outputs, guarded direct stores
              const bool full = (n0 + 32 <= p.cols);
              float y[32];
              // bias first, as independent vector loads (one exposed latency instead of 32 dependent ones)
              if (p.bias != nullptr) {
                if (full && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0) {
                  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 q = __ldg(b4 + j);
                    y[4 * j] = q.x; y[4 * j + 1] = q.y; y[4 * j + 2] = q.z; y[4 * j + 3] = q.w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) y[j] = (n0 + j < p.cols) ? __ldg(p.bias + n0 + j) : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = 0.f;
              }
              // residual, likewise issued before the accumulator is needed
              float rs[32];
              const bool row_ok = my_row < p.rows;
              const bool res_tma = p.res_tma != 0;
              if (res_tma) {
                // in-place residual stream: the two 16-column halves of this chunk's residual tile are TMA-loaded into
                // the warp's two staging tiles (the tiles the result is stored from, same tensor map and coordinates)
                // while the accumulator is still on its way — instead of 8 per-lane 16-byte row loads that touch 32
                // different lines each (the LSU-bound pattern of the round-1 epilogue)
                if (lane == 0) tma_store_wait_read<0>();
                __syncwarp();
                ss.pending = 0;
                if (lane == 0) {
#pragma unroll
                  for (int hf = 0; hf < 2; ++hf) {
                    if (n0 + hf * 16 >= p.cols) break;
                    mbar_arrive_expect_tx(res_bar + 8 * hf, kStagingBytesPerWarp);
                    tma_load_3d(ss.buf[hf], &tmOut, res_bar + 8 * hf, n0 + hf * 16, row0, 0);
                  }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) rs[j] = 0.f;
              } else if (p.residual != nullptr && row_ok) {
                const float* r = p.residual + static_cast<long long>(my_row) * p.res_ld + n0;
                if (full && (p.res_ld & 3) == 0) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 q = *reinterpret_cast<const float4*>(r + 4 * j);
                    rs[4 * j] = q.x; rs[4 * j + 1] = q.y; rs[4 * j + 2] = q.z; rs[4 * j + 3] = q.w;
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j) rs[j] = (n0 + j < p.cols) ? r[j] : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) rs[j] = 0.f;
              }
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) y[j] += __uint_as_float(v[j]);
              if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
              } else if (p.act == 2) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) {  // exact-erf GELU on the packed fp32x2 pipe (|error| < 1e-6, fused_encoder.cuh)
                  const float2 r = fe_gelu2(make_float2(y[j], y[j + 1]));
                  y[j] = r.x;
                  y[j + 1] = r.y;
                }
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) y[j] += rs[j];
              if (res_tma) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  if (n0 + hf * 16 >= p.cols) break;
                  mbar_wait(res_bar + 8 * hf, (res_phase >> hf) & 1u, 10);
                  res_phase ^= 1u << hf;
                  const uint32_t base = ss.buf[hf] + ss.row_off;  // this lane's row of the residual tile (64B swizzle)
#pragma unroll
                  for (int ch = 0; ch < 4; ++ch) {
                    const uint32_t addr = base + ((static_cast<uint32_t>(ch) ^ ss.swz) << 4);
                    uint32_t r0, r1, r2, r3;
                    ld_shared_v4(addr, r0, r1, r2, r3);
                    y[hf * 16 + 4 * ch] += __uint_as_float(r0);
                    y[hf * 16 + 4 * ch + 1] += __uint_as_float(r1);
                    y[hf * 16 + 4 * ch + 2] += __uint_as_float(r2);
                    y[hf * 16 + 4 * ch + 3] += __uint_as_float(r3);
                    st_shared_v4(addr, __float_as_uint(y[hf * 16 + 4 * ch]), __float_as_uint(y[hf * 16 + 4 * ch + 1]),
                                 __float_as_uint(y[hf * 16 + 4 * ch + 2]), __float_as_uint(y[hf * 16 + 4 * ch + 3]));
                  }
                  fence_proxy_async_smem();
                  __syncwarp();
                  if (lane == 0) {
                    tma_store_3d(&tmOut, ss.buf[hf], n0 + hf * 16, row0, 0);
                    tma_store_commit();
                  }
                  ++ss.pending;
                }
              } else if (p.out_f32 != nullptr) {
                if (p.use_tma_store) {
#pragma unroll
                  for (int hf = 0; hf < 2; ++hf) {  // two 16-column (64-byte) fills
                    if (n0 + hf * 16 >= p.cols) break;
                    uint32_t w[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) w[j] = __float_as_uint(y[hf * 16 + j]);
                    ss.store(&tmOut, w, n0 + hf * 16, row0, 0, lane);
                  }
                } else if (row_ok) {
                  float* o = p.out_f32 + static_cast<long long>(my_row) * p.out_ld + n0;
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (n0 + j < p.cols) o[j] = y[j];
                }
              }
              if (p.out_bf16 != nullptr) {
                for (int part = 0; part < (p.write_lo ? 2 : 1); ++part) {
                  if (p.use_tma_store2) {
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                      float a = y[2 * j], b = y[2 * j + 1];
                      if (part == 1) {
                        a -= __bfloat162float(__float2bfloat16_rn(a));
                        b -= __bfloat162float(__float2bfloat16_rn(b));
                      }
                      pk[j] = pack_bf16x2(a, b);
                    }
                    ss.store(&tmOut2, pk, n0 + part * p.bf16_lo_off, row0, 0, lane);
                  } else if (row_ok) {
                    __nv_bfloat16* op = p.out_bf16 + static_cast<long long>(my_row) * p.bf16_ld + n0 + part * p.bf16_lo_off;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                      if (n0 + j < p.cols) {
                        float a = y[j];
                        if (part == 1) a -= __bfloat162float(__float2bfloat16_rn(a));
                        op[j] = __float2bfloat16_rn(a);
                      }
                  }
                }
              }
            } else if constexpr (EPI == EPI_BF16_SPLIT) {
              // hi = bf16(y), lo = bf16(y - hi): the A operand of the second GEMM, K-major rows [hi | lo]
              for (int part = 0; part < (p.write_lo ? 2 : 1); ++part) {
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float a = __uint_as_float(v[2 * j]), b = __uint_as_float(v[2 * j + 1]);
                  if (part == 1) {
                    a -= __bfloat162float(__float2bfloat16_rn(a));
                    b -= __bfloat162float(__float2bfloat16_rn(b));
                  }
                  pk[j] = pack_bf16x2(a, b);
                }
                const int ncol = n0 + part * p.lo_col_offset;
                if (p.use_tma_store) {
                  ss.store(&tmOut, pk, ncol, row0, c.l, lane);
                } else if (my_row < p.rows) {
                  uint32_t* o = reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(p.out) +
                                                            c.l * p.out_batch_stride + my_row * p.out_ld + ncol);
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (n0 + 2 * j < p.cols) o[j] = pk[j];  // cols and offsets are even in this mode
                }
              }
            } else {
              // fp32 logits or sigmoid
              if constexpr (EPI == EPI_SIGMOID) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  v[j] = __float_as_uint(__fdividef(1.0f, 1.0f + __expf(-__uint_as_float(v[j]))));
              }
              if (p.use_tma_store) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {  // two 16-column (64-byte) fills
                  if (n0 + hf * 16 >= p.cols) break;
                  uint32_t w[16];
#pragma unroll
                  for (int j = 0; j < 16; ++j) w[j] = v[hf * 16 + j];
                  ss.store(&tmOut, w, n0 + hf * 16, row0, c.l, lane);
                }
              } else if (my_row < p.rows) {
                float* o = reinterpret_cast<float*>(p.out) + c.l * p.out_batch_stride + my_row * p.out_ld + n0;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (n0 + j < p.cols) o[j] = __uint_as_float(v[j]);
              }
            }
          }
          }
        }
        if (!released) {
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_t_empty(acc_stage));
        }
        acc_stage ^= 1;
        if (acc_stage == 0) acc_phase ^= 1;
      }
    }
    if (ss.pending > 0 && lane == 0) tma_store_wait_all<0>();
    __syncwarp();
    }  // legacy (non-pipelined) epilogues
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ triple gather
// Scores of a LIST of (outcome, head, tail) triples without the dense [L, Nh, Nt] tensor (reference: the full forward
// followed by pred[labels, heads, tails], train_ddi_batch.py:285-286, evaluate.py:191-195):
//   out[t] = Y[l_t, h_t, :] . z_cols[t_t, :]   with Y = z_rows . W_l from GEMM 1 (bf16 operand rows, [hi | lo] in the
// fp32-parity mode) and z_cols in the same operand format.  One warp per triple, 16-byte loads, shuffle reduction.
__global__ void __launch_bounds__(256) gather_dot_kernel(const __nv_bfloat16* __restrict__ y,
                                                         const __nv_bfloat16* __restrict__ zc, long long nr_pad,
                                                         int ka, int D, int split, int L, int Nr, int Nc,
                                                         const int* __restrict__ labels, const int* __restrict__ heads,
                                                         const int* __restrict__ tails, long long n, int sigmoid,
                                                         float* __restrict__ out) {
  const long long t = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= n) return;
  const int lane = threadIdx.x & 31;
  const int l = labels[t], h = heads[t], c = tails[t];
  if (l < 0 || l >= L || h < 0 || h >= Nr || c < 0 || c >= Nc) {  // out-of-range index: poison, do not touch memory
    if (lane == 0) out[t] = __uint_as_float(0x7fc00000u);
    return;
  }
  const uint4* yr = reinterpret_cast<const uint4*>(y + (static_cast<long long>(l) * nr_pad + h) * ka);
  const uint4* zr = reinterpret_cast<const uint4*>(zc + static_cast<long long>(c) * ka);
  float acc = 0.f;
  for (int k8 = lane; k8 < D / 8; k8 += 32) {  // 8 bf16 per 16-byte load; D is a multiple of 64
    const uint4 a = yr[k8], b = zr[k8];
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    float alo[8] = {0, 0, 0, 0, 0, 0, 0, 0}, blo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (split) {
      const uint4 a2 = yr[D / 8 + k8], b2 = zr[D / 8 + k8];
      const uint32_t aw2[4] = {a2.x, a2.y, a2.z, a2.w}, bw2[4] = {b2.x, b2.y, b2.z, b2.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw2[i]));
        const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw2[i]));
        alo[2 * i] = p.x; alo[2 * i + 1] = p.y; blo[2 * i] = q.x; blo[2 * i + 1] = q.y;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw[i]));
      const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&bw[i]));
      acc = fmaf(p.x + alo[2 * i], q.x + blo[2 * i], acc);
      acc = fmaf(p.y + alo[2 * i + 1], q.y + blo[2 * i + 1], acc);
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) out[t] = sigmoid ? 1.0f / (1.0f + expf(-acc)) : acc;
}

// Ensemble reductions over K same-shaped tensors (reference: mean over checkpoints of sigmoid scores,
// predict.py:493, 612; geometric mean of the checkpoints' normalised ranks, generate_embeddings.ipynb cell 18 =
// scipy.stats.mstats.gmean = exp(mean(log x)) in the input's float32).
struct EnsemblePtrs {
  const void* p[16];
};
template <int MODE>  // 0: mean of fp32, 1: gmean of fp32, 2: gmean of uint16 ranks * scale
__global__ void __launch_bounds__(256) ensemble_reduce_kernel(EnsemblePtrs in, int K, long long n, float scale,
                                                              float* __restrict__ out) {
  const float invk = 1.0f / static_cast<float>(K);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      float x;
      if (MODE == 2) x = static_cast<float>(static_cast<const uint16_t*>(in.p[k])[i]) * scale;
      else x = static_cast<const float*>(in.p[k])[i];
      acc += (MODE == 0) ? x : logf(x);
    }
    out[i] = (MODE == 0) ? acc * invk : expf(acc * invk);
  }
}

// Ensemble of FUSED quantile ranks (reference: geometric mean of the checkpoints' normalised ranks, then the same rank
// normalisation again — generate_embeddings.ipynb cells 18 + 20) without ever leaving uint16:
//   g[i]   = sum_k ilog[ r_k[i] ]              ilog = caller-supplied fixed-point log2 table [Q + 1] (the gmean's log-sum)
//   out[i] = #{ t in ens thresholds[l] : t <= float(g[i]) }   through the ensemble table's LUT (rank_lookup_raw), or
//   outf[i] = float(g[i])                      (builder mode: feeds mdg_lower_triangle_quantiles -> the ensemble table)
// An element whose member ranks are all 0 (the diagonal of the normaliser layout) stays 0.  One block column per
// outcome slice; the ilog table and the outcome's LUT sit in shared memory; 8 elements per thread and 16-byte accesses.
template <bool TO_RANK>
__global__ void __launch_bounds__(256) ensemble_rank_kernel(EnsemblePtrs in, int K, long long n,
                                                            const uint16_t* __restrict__ ilog, int Q,
                                                            const uint32_t* __restrict__ lut_all,
                                                            const float* __restrict__ affine,
                                                            uint16_t* __restrict__ out, float* __restrict__ outf) {
  extern __shared__ __align__(16) uint8_t ens_smem[];
  uint16_t* s_ilog = reinterpret_cast<uint16_t*>(ens_smem);                                   // [Q + 1] (padded to 16 B)
  uint32_t* s_lut = reinterpret_cast<uint32_t*>(ens_smem + ((static_cast<size_t>(Q) + 1) * 2 + 15) / 16 * 16);
  const int l = blockIdx.y;
  for (int i = threadIdx.x; i <= Q; i += blockDim.x) s_ilog[i] = ilog[i];
  float scale = 0.f, bias = 0.f;
  if (TO_RANK) {
    const uint32_t* lut = lut_all + static_cast<size_t>(l) * kRankLutEntries;
    for (int i = threadIdx.x; i < kRankLutEntries; i += blockDim.x) s_lut[i] = lut[i];
    scale = affine[2 * l];
    bias = affine[2 * l + 1];
  }
  __syncthreads();
  const long long base = static_cast<long long>(l) * n;
  const long long n8 = n / 8;
  for (long long i8 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i8 < n8;
       i8 += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t g[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t any[4] = {0, 0, 0, 0};
    for (int k = 0; k < K; ++k) {
      const uint4 v = *reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(in.p[k]) + base + i8 * 8);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        any[j] |= w[j];
        g[2 * j] += s_ilog[min(w[j] & 0xFFFFu, static_cast<uint32_t>(Q))];
        g[2 * j + 1] += s_ilog[min(w[j] >> 16, static_cast<uint32_t>(Q))];
      }
    }
    if (TO_RANK) {
      uint32_t r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t lo = rank_lookup_raw(s_lut, static_cast<float>(g[2 * j]), scale, bias) & 0xFFFFu;
        uint32_t hi = rank_lookup_raw(s_lut, static_cast<float>(g[2 * j + 1]), scale, bias) & 0xFFFFu;
        if ((any[j] & 0xFFFFu) == 0) lo = 0;
        if ((any[j] >> 16) == 0) hi = 0;
        r[j] = lo | (hi << 16);
      }
      *reinterpret_cast<uint4*>(out + base + i8 * 8) = make_uint4(r[0], r[1], r[2], r[3]);
    } else {
      float4* o = reinterpret_cast<float4*>(outf + base + i8 * 8);
      o[0] = make_float4(static_cast<float>(g[0]), static_cast<float>(g[1]), static_cast<float>(g[2]), static_cast<float>(g[3]));
      o[1] = make_float4(static_cast<float>(g[4]), static_cast<float>(g[5]), static_cast<float>(g[6]), static_cast<float>(g[7]));
    }
  }
  // tail (n % 8 elements), one thread each
  for (long long i = n8 * 8 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint32_t g = 0, any = 0;
    for (int k = 0; k < K; ++k) {
      const uint32_t r = static_cast<const uint16_t*>(in.p[k])[base + i];
      any |= r;
      g += s_ilog[min(r, static_cast<uint32_t>(Q))];
    }
    if (TO_RANK) out[base + i] = any ? static_cast<uint16_t>(rank_lookup_raw(s_lut, static_cast<float>(g), scale, bias) & 0xFFFFu) : 0;
    else outf[base + i] = static_cast<float>(g);
  }
}

// ------------------------------------------------------------------------------------------------ operand prep
// z [N, D] fp32 -> bf16 [Npad, Ka]  with Ka = D (hi only) or 2D ([hi | lo]); rows >= N are zero.
// Optional row L2 normalisation (F.normalize: x / max(||x||_2, 1e-12), models.py:947-949).  One warp per row.
// Also zeroes the dynamic tile scheduler's task counter for the N^2 GEMM two launches later (`sched_zero`, may be
// NULL), which keeps a memset node out of the PDL chain convert_z -> GEMM 1 -> GEMM 2.
__global__ void __launch_bounds__(256) convert_z_kernel(const float* __restrict__ z, int N, int Npad, int D,
                                                        int split, int normalize, __nv_bfloat16* __restrict__ out,
                                                        unsigned int* __restrict__ sched_zero) {
  pdl_wait();  // z is the predecessor's output; the counter may still be in use by the previous step's GEMM
  pdl_launch_dependents();
  if (sched_zero != nullptr && blockIdx.x == 0 && threadIdx.x < 8) sched_zero[threadIdx.x] = 0u;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= Npad) return;
  const int Ka = split ? 2 * D : D;
  __nv_bfloat16* o = out + static_cast<size_t>(row) * Ka;
  float vals[8];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int d = lane + 32 * i;
    float x = (row < N && d < D) ? z[static_cast<size_t>(row) * D + d] : 0.f;
    vals[i] = x;
    ss += x * x;
  }
  if (normalize) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
    float denom = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
    for (int i = 0; i < 8; ++i) vals[i] = vals[i] / denom;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int d = lane + 32 * i;
    if (d < D) {
      __nv_bfloat16 hi = __float2bfloat16_rn(vals[i]);
      o[d] = hi;
      if (split) o[D + d] = __float2bfloat16_rn(vals[i] - __bfloat162float(hi));
    }
  }
}

// ------------------------------------------------------------------------------------------------ peer all-gather
// The path's only exchange step: every GPU replicates the fused-embedding table z [N, D].  Instead
// of a library collective, each rank PUSHES its row shard into every rank's copy of the table through NVLink
// peer-mapped pointers (16-byte stores, all G destinations from one load), then the last CTA to finish signals an
// epoch flag on every peer (st.release.sys) and waits for the peers' flags (ld.acquire.sys): when the kernel retires,
// the local table holds every rank's rows.  One launch, no intermediate buffers, no host synchronisation.
constexpr int kMaxPeers = 8;
struct PeerTable {
  float* buf[kMaxPeers];           // rank r's table [N, D] as mapped into THIS process
  unsigned int* flags[kMaxPeers];  // rank r's flag words: [0, 8) one slot per sender, [8] local CTA counter
};
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256) peer_allgather_kernel(const float4* __restrict__ shard, long long n16,
                                                             long long dst_off16, PeerTable pt, int world, int rank,
                                                             unsigned int epoch) {
  pdl_wait();  // the shard is the predecessor's (the encoder's) output
  pdl_launch_dependents();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = shard[i];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < world) reinterpret_cast<float4*>(pt.buf[q])[dst_off16 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  unsigned int* mine = pt.flags[rank];
  if (threadIdx.x == 0) s_last = (atomicAdd(mine + 8, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) mine[8] = 0u;  // for the next launch (stream-ordered after this one)
  __threadfence_system();
  if (threadIdx.x < world) {
    st_release_sys(pt.flags[threadIdx.x] + rank, epoch);  // "rank's rows of epoch `epoch` have landed on you"
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(mine + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > (1ll << 35)) {  // ~15 s: a peer never arrived; fail the launch instead of hanging the GPU
        g_hang_code = 0x90000000u | threadIdx.x;
        __threadfence_system();
        asm volatile("trap;");
      }
    }
  }
}

// F.normalize(x, p=2, dim=-1): x / max(||x||_2, 1e-12) per row (reference: models.py:849-850, 861-862, 890-891 — the
// token normalisation of the unimodal bypass / raw-encoder-output / 'mean'-'add' fusion paths).  One warp per row.
__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ x, long long rows, int dim,
                                                                float* __restrict__ out) {
  const long long row = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + row * dim;
  float ss = 0.f;
  for (int d = lane; d < dim; d += 32) ss = fmaf(xr[d], xr[d], ss);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, s);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int d = lane; d < dim; d += 32) out[row * dim + d] = xr[d] / denom;
}

// W [L, D, D] fp32 (W[l][a][b]) -> Wt bf16 [L, D, Ka] with Wt[l][b][a] (+ lo half at K offset D): the B operand of
// GEMM 1 in K-major form.  32x32 shared-memory tile transpose.
__global__ void __launch_bounds__(256) convert_w_kernel(const float* __restrict__ W, int D, int split,
                                                        __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int l = blockIdx.z;
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* w = W + static_cast<size_t>(l) * D * D;
  const int Ka = split ? 2 * D : D;
  __nv_bfloat16* o = out + static_cast<size_t>(l) * D * Ka;
  for (int r = ty; r < 32; r += 8) {
    int a = a0 + r, b = b0 + tx;
    tile[r][tx] = (a < D && b < D) ? w[static_cast<size_t>(a) * D + b] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int b = b0 + r, a = a0 + tx;
    if (a < D && b < D) {
      float x = tile[tx][r];
      __nv_bfloat16 hi = __float2bfloat16_rn(x);
      o[static_cast<size_t>(b) * Ka + a] = hi;
      if (split) o[static_cast<size_t>(b) * Ka + D + a] = __float2bfloat16_rn(x - __bfloat162float(hi));
    }
  }
}

}  // namespace mdg
