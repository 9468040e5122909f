/*
 * madrigal_b200.h — C ABI of the B200-native (sm_100a) drug-pair scoring path.
 *
 * This is the drop-in boundary for ONE path of biopharmaai/Madrigal (the reference):
 *
 *     modality tokens -> fusion transformer -> fused drug embedding z
 *                     -> per-outcome bilinear decoder  z_A . W_k . z_B^T
 *                     -> per-outcome rank normalisation
 *
 * The reference is pure Python/PyTorch and has no FFI layer; its boundary for this path is a set of
 * Python call signatures.  Each entry point below names the reference symbol (file:line under the
 * reference root) whose arithmetic it replaces.  madrigal_b200/*.py holds Python modules with the reference's
 * class names / constructor arguments / state_dict keys that bind these entry points through ctypes; the stub a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns every buffer, the library
 *     never allocates, frees or retains device memory beyond the call;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no hidden synchronisation;
 *   - return value 0 = success, non-zero = MdgStatus; mdg_last_error() returns a thread-local message;
 *   - there is NO CPU fallback and NO alternative backend: unsupported shapes fail with MDG_ERR_UNSUPPORTED;
 *   - row-major ("C") layouts throughout, fp32 inputs exactly as the reference holds them.
 */
#ifndef MADRIGAL_B200_H_
#define MADRIGAL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDG_ABI_VERSION 5

typedef enum MdgStatus {
  MDG_OK = 0,
  MDG_ERR_INVALID_ARGUMENT = 1, /* NULL pointer, negative size, inconsistent shapes */
  MDG_ERR_UNSUPPORTED = 2,      /* shape/config outside what the sm_100a kernels implement */
  MDG_ERR_WORKSPACE = 3,        /* workspace too small */
  MDG_ERR_CUDA = 4,             /* CUDA runtime/driver error (message in mdg_last_error) */
  MDG_ERR_NO_DEVICE = 5         /* no sm_100 device visible */
} MdgStatus;

/* Thread-local, never NULL. */
const char* mdg_last_error(void);
int mdg_abi_version(void);
/* 0 if device `device` (or the current one if <0) is compute capability 10.x, else MDG_ERR_NO_DEVICE. */
int mdg_check_device(int device);

/* ------------------------------------------------------------------------------------------------------------------
 * Bilinear decoder, all pairs  (reference: BilinearDDIScorer.bilinear/forward, madrigal/models/models.py:537-547;
 *                               caller-side sigmoid, madrigal/evaluate/predict.py:235,358)
 *
 *   S[l, i, j] = sum_{a,b} z_rows[i, a] * W[l, a, b] * z_cols[j, b]        l < L, i < Nr, j < Nc
 *
 * `W` is the weight AS THE REFERENCE READS IT (through the Symmetric parametrisation if one is registered,
 * models.py:522-524,922) — i.e. `label_range` slicing is done by the caller by offsetting W and `L`.
 * ------------------------------------------------------------------------------------------------------------------ */

typedef enum MdgPrecision {
  MDG_PREC_BF16 = 0, /* operands rounded to bf16 once, fp32 accumulation in TMEM (north-star bar: 1e-2)          */
  MDG_PREC_FP32 = 1  /* bf16x3 split operands (hi*hi + lo*hi + hi*lo), fp32 accumulation (north-star bar: 1e-3) */
} MdgPrecision;

typedef enum MdgOutMode {
  MDG_OUT_LOGIT_F32 = 0,   /* out: float   [L, Nr, Nc]  raw logits                                        */
  MDG_OUT_SIGMOID_F32 = 1, /* out: float   [L, Nr, Nc]  1/(1+exp(-logit))          (predict.py:358)       */
  MDG_OUT_RANK_U16 = 2     /* out: uint16  [L, Nr, Nc]  searchsorted(table[l], logit, side='right')       */
} MdgOutMode;

typedef enum MdgPairs {
  MDG_PAIRS_FULL = 0,     /* every (i, j)                                                                          */
  MDG_PAIRS_SYMMETRIC = 1, /* z_rows == z_cols only: only row > col is computed — the pair set the reference normaliser
                             ranks (notebooks/normalize_scores.py:67).  MDG_OUT_RANK_U16: each rank is written at
                             [row, col] AND [col, row], diagonal 0 (the normaliser's output layout, :69-70);
                             mdg_pair_topk: unordered pairs.  Not available for the fp32 logit / sigmoid outputs. */
  MDG_PAIRS_PACKED_TILES = 2 /* MDG_OUT_RANK_U16 only: the same ranks WITHOUT the mirror image (half the bytes to write
                             and to copy to the host).  out: uint16 [L, T, 32, 32] with T = mdg_packed_tiles_per_outcome(N)
                             = nb (nb + 1) / 2, nb = ceil(N / 32): tile (bi, bj), bj <= bi, sits at index bi (bi + 1) / 2 +
                             bj and holds ranks[l, 32 bi + r, 32 bj + c] at [r, c]; entries with col >= row (diagonal
                             tiles) are 0, entries beyond N are unspecified.  The host rebuilds the normaliser's
                             [L, N, N] layout by scattering the tiles and adding the transpose. */
} MdgPairs;
int64_t mdg_packed_tiles_per_outcome(int64_t N);
/* Host half of a packed device-to-host transfer: scatters MDG_PAIRS_PACKED_TILES tiles that have been copied to host
   memory ([L, T, 32, 32] uint16) into the normaliser's layout out_host [L, N, N] — every rank at [i, j] and [j, i], zero
   diagonal (the mirror step of notebooks/normalize_scores.py:67-70, applied to ranks the GPU has already computed).
   Pure data movement on `threads` host threads (<= 0: all hardware threads); HOST pointers, no CUDA call, blocks until
   done.  scoring.score_all_pairs_to_host uses it so that only half of the rank tensor crosses PCIe. */
int mdg_host_mirror_tiles(const uint16_t* tiles_host, int64_t L, int64_t N, uint16_t* out_host, int32_t threads);

/* Prepared per-outcome reference-quantile table for the fused rank epilogue (see mdg_rank_table_build). */
#define MDG_RANK_BUCKET_BITS 13
#define MDG_RANK_SUB_BITS 4
#define MDG_RANK_LUT_ENTRIES (1 << MDG_RANK_BUCKET_BITS) /* uint32 entries per outcome (32 KB) */
#define MDG_RANK_MAX_Q 65535

typedef enum MdgRankKind {
  MDG_RANK_LUT = 0, /* exact bucket LUT: every threshold moves by <= ~3 grid cells (mdg_rank_table_build)          */
  MDG_RANK_PWL = 1  /* 256-bin histogram CDF, linear inside a bin, exact at the bin edges; bank-conflict-free in
                       shared memory (mdg_rank_table_build_pwl).  Thresholds move by up to max_dev ranks.          */
} MdgRankKind;

typedef struct MdgRankTable {
  const float* thresholds; /* [L, Q] ascending, snapped to the lookup grid (what np.searchsorted is run against) */
  const uint32_t* lut;     /* [L, MDG_RANK_LUT_ENTRIES]                                                          */
  const float* affine;     /* [L, 2]  (scale, bias) of the logit -> grid map                                     */
  int32_t L;
  int32_t Q;
  int32_t kind;            /* MdgRankKind of `lut`                                                               */
} MdgRankTable;

/*
 * Snap ascending reference quantiles onto the epilogue's lookup grid and build the lookup structure.
 *   quantiles      [L, Q] fp32, ascending per row (e.g. order statistics of the strict-lower-triangle logits of a
 *                  reference panel — the distribution notebooks/normalize_scores.py:36-60 ranks against)
 *   thresholds_out [L, Q] fp32: the snapped table.  |thresholds_out - quantiles| <= ~2 grid cells
 *                  (cell = 1.02*(q_max-q_min)/2^17), strictly ascending wherever fp32 allows.
 *   lut_out        [L, MDG_RANK_LUT_ENTRIES] uint32, affine_out [L, 2] fp32.
 * Contract: for every finite fp32 x, the fused epilogue's rank for outcome l equals
 *           np.searchsorted(thresholds_out[l], x, side='right') bit for bit.
 */
int mdg_rank_table_build(const float* quantiles, int32_t L, int32_t Q, float* thresholds_out, uint32_t* lut_out,
                         float* affine_out, void* stream);

/*
 * Histogram-CDF variant (MDG_RANK_PWL).  Same inputs / outputs / contract as mdg_rank_table_build (the fused epilogue
 * equals np.searchsorted(thresholds_out[l], x, side='right') bit for bit), but the lookup structure is a 256-bin
 * piecewise-linear CDF on the same grid: exact at the 257 bin edges, interpolated inside a bin.  Because the 256
 * entries are replicated once per shared-memory bank, the epilogue's lookups are conflict-free (the exact LUT costs
 * ~4.9 shared-memory wavefronts per warp lookup).  The price: thresholds_out[i] may sit up to max_dev ranks away from
 * quantiles[i]; max_dev_out [L] (optional, device) reports max_i |table_rank(quantiles[i]) - (i + 1)| per outcome
 * (about 1 for smooth score distributions at Q = 16384).  thresholds_out[i] = +inf where no finite score reaches
 * rank i + 1.
 */
int mdg_rank_table_build_pwl(const float* quantiles, int32_t L, int32_t Q, float* thresholds_out, uint32_t* lut_out,
                             float* affine_out, float* max_dev_out, void* stream);

/* Stand-alone lookup of already materialised logits through a prepared table (same device function as the fused
 * epilogue).  logits [L, n] -> ranks [L, n] uint16. */
int mdg_rank_lookup(const float* logits, int64_t n_per_outcome, const MdgRankTable* table, uint16_t* ranks,
                    void* stream);

size_t mdg_pair_score_workspace_bytes(int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision);

/*
 * All-pairs decoder with fused epilogue.
 *   z_rows [Nr, D], z_cols [Nc, D], W [L, D, D] fp32.  D in {64, 128, 192, 256}.
 *   normalize_rows != 0: L2-normalise each z row first (F.normalize, models.py:947-949).
 *   table: required for MDG_OUT_RANK_U16 (table->L must be >= L; row l of the table is used for W[l]).
 *   out: [L, Nr, Nc] of the mode's element type.
 *   workspace: >= mdg_pair_score_workspace_bytes(...) bytes, 256-byte aligned.
 */
int mdg_pair_score(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                   int64_t L, int precision, int out_mode, int pairs, int normalize_rows,
                   const MdgRankTable* table, void* out, void* workspace, size_t workspace_bytes, void* stream);

/*
 * Prepared decoder: the outcome weights converted ONCE to the GEMM-operand form (bf16, K-major, [hi | lo] in the
 * fp32-parity mode) — W is constant between checkpoints loads, the reference re-reads it through the Symmetric
 * parametrisation on every decoder call (models.py:537-547, 922).  `prepared`: caller-owned device buffer of
 * mdg_pair_prepared_bytes(D, L, precision) bytes, 256-byte aligned, for ALL L outcomes of the decoder.
 * mdg_pair_score_prepared is mdg_pair_score for outcomes [first_outcome, first_outcome + L) of that buffer (the
 * reference's `label_range`, predict.py:420-429) without the per-call weight conversion.
 */
size_t mdg_pair_prepared_bytes(int64_t D, int64_t L, int precision);
int mdg_pair_prepare(const float* W, int64_t D, int64_t L, int precision, void* prepared, size_t prepared_bytes,
                     void* stream);
int mdg_pair_score_prepared(const float* z_rows, const float* z_cols, const void* prepared, int64_t first_outcome,
                            int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision, int out_mode, int pairs,
                            int normalize_rows, const MdgRankTable* table, void* out, void* workspace,
                            size_t workspace_bytes, void* stream);

/*
 * The path's only exchange step (the reference is single-GPU): replicate the fused-embedding table
 * z [N, D] on every GPU of one node.  Each rank pushes its row shard into EVERY rank's copy of the table through
 * NVLink peer-mapped pointers and the ranks exchange one epoch flag; when the kernel retires on a rank, that rank's
 * table holds all rows.  No library collective, no staging buffers, nothing on the host.
 *   shard             [shard_rows, D] fp32, this rank's rows (16-byte aligned); lands at rows [row_offset, +shard_rows)
 *   peer_tables_host  HOST array of `world` device pointers: rank r's table [N, D] as mapped into this process
 *                     (e.g. torch symmetric memory / cudaIpc mappings); entry `rank` is the local table
 *   peer_flags_host   HOST array of `world` device pointers to >= 16 zero-initialised uint32 per rank, same mapping
 *   epoch             strictly increasing over successive calls, identical on all ranks (all ranks call in lockstep);
 *                     a table must not be rewritten while a peer may still read it: alternate two tables (the Python
 *                     wrapper does) or separate the calls by another synchronisation.
 * A peer that never arrives trips a ~15 s in-kernel timeout (the launch fails) instead of hanging the device.
 */
int mdg_peer_allgather(const float* shard, int64_t shard_rows, int64_t row_offset, int32_t D,
                       void* const* peer_tables_host, void* const* peer_flags_host, int32_t world, int32_t rank,
                       uint32_t epoch, void* stream);

/* F.normalize(x, p=2, dim=-1) with eps 1e-12 on rows of x [rows, dim] fp32 -> out (may alias x).  The reference applies
 * it to single tokens on its unimodal-bypass, raw-encoder-output and 'mean'/'add' fusion paths (models.py:849-850,
 * 861-862, 870-878, 890-891); the transformer path's token normalisation is fused into mdg_assemble_tokens and the
 * decoder's into mdg_pair_score (normalize_rows). */
int mdg_l2_normalize_rows(const float* x, int64_t rows, int32_t dim, float* out, void* stream);

/*
 * Per-outcome top-k pairs without materialising any dense output (BASELINE config 4: "per-outcome top-1000").
 * The decoder's epilogue appends every score >= thresholds[l] to outcome l's candidate list (capacity `cap`); the
 * lists are then sorted (score descending, ties by pair index ascending) and the first k are written out.
 *   pairs = MDG_PAIRS_SYMMETRIC: unordered pairs of one catalogue — only row > col is kept and column blocks above
 *           the diagonal are never computed (the set normalize_scores.py ranks); MDG_PAIRS_FULL: every (row, col).
 *   thresholds [L] fp32 (device): e.g. a high quantile of the rank table, so that k <= #candidates <= cap.
 *   scores_out [L, k] fp32, rows_out / cols_out [L, k] int32 (-1 / -inf padding when fewer than k candidates)
 *   status_out [L] int32: 0 ok, 1 fewer than k candidates (threshold too high), 2 more than cap (too low: the list is
 *   then an arbitrary subset and must be recomputed).
 */
size_t mdg_pair_topk_workspace_bytes(int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision, int32_t cap);
int mdg_pair_topk(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                  int64_t L, int precision, int pairs, int normalize_rows, const float* thresholds, int32_t k,
                  int32_t cap, float* scores_out, int32_t* rows_out, int32_t* cols_out, int32_t* status_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/*
 * Scores of a LIST of triples without the dense tensor  (reference: `pred = sigmoid(model(...))` followed by
 * `pred[ddi_labels, head_idx, tail_idx]`, train_ddi_batch.py:285-286 and evaluate.py:191-195 — the reference
 * materialises all of [L, Nh, Nt] and then gathers; data.py:938 explains why).
 *   out[t] = z_rows[heads[t]] . W[labels[t]] . z_cols[tails[t]]      t < n     (sigmoid applied for MDG_OUT_SIGMOID_F32)
 * GEMM 1 (z_rows . W_l for every outcome) runs on the tensor cores exactly as in mdg_pair_score; the N^2 GEMM is
 * replaced by one dot product per listed triple.  labels/heads/tails: int32 device arrays; an out-of-range index
 * yields NaN in out[t].  Workspace: mdg_pair_score_workspace_bytes(Nr, Nc, D, L, precision).
 */
int mdg_pair_score_gather(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                          int64_t L, int precision, int normalize_rows, const int32_t* labels, const int32_t* heads,
                          const int32_t* tails, int64_t n, int out_mode, float* out, void* workspace,
                          size_t workspace_bytes, void* stream);

/*
 * Ensemble reductions over K same-shaped device tensors of n elements (K <= MDG_MAX_ENSEMBLE):
 *   MDG_ENS_MEAN_F32        out = mean_k x_k                   (mean over checkpoints of sigmoid scores,
 *                                                               madrigal/evaluate/predict.py:493, 612)
 *   MDG_ENS_GMEAN_F32       out = exp(mean_k log x_k), fp32    (geometric mean of the checkpoints' normalised ranks,
 *                                                               notebooks/generate_embeddings.ipynb cell 18:
 *                                                               scipy.stats.mstats.gmean on float32)
 *   MDG_ENS_GMEAN_RANK_U16  the same on uint16 quantile ranks scaled by rank_scale (= 1/Q); 0 stays 0.
 * members_host: HOST array of K device pointers.  The reference then re-ranks the result with the same normaliser
 * (ipynb cell 20): feed `out` to mdg_exact_rank.
 */
#define MDG_MAX_ENSEMBLE 16
typedef enum MdgEnsembleMode { MDG_ENS_MEAN_F32 = 0, MDG_ENS_GMEAN_F32 = 1, MDG_ENS_GMEAN_RANK_U16 = 2 } MdgEnsembleMode;
int mdg_ensemble_reduce(const void* const* members_host, int32_t K, int64_t n, int mode, float rank_scale, float* out,
                        void* stream);

/*
 * Ensemble of FUSED quantile ranks (reference: scipy gmean of the checkpoints' normalised ranks followed by the same
 * rank normalisation, notebooks/generate_embeddings.ipynb:434, 634-649) in the quantile-table formulation, uint16 in,
 * uint16 out:
 *     g[l, i]   = sum_k ilog_table[ members[k][l, i] ]                     (the gmean's log-sum, fixed point)
 *     ranks_out = np.searchsorted(ens_table->thresholds[l], float32(g), side='right'), 0 where every member rank is 0
 *                 (the diagonal of the normaliser layout)
 *  or logsum_out = float32(g)  — builder mode: feed a reference panel's g to mdg_lower_triangle_quantiles and
 *                 mdg_rank_table_build to obtain `ens_table` (exact-LUT kind).
 * members_host: HOST array of K device pointers to uint16 [L, n_per_outcome] (16-byte aligned; n % 8 == 0 unless L == 1);
 * ilog_table: device uint16 [Q + 1], e.g. round(1024 * (log2(max(r, 0.5)) + 1)) — it DEFINES the fixed-point log, so a
 * host restatement with the same table is bit-exact.  Exactly one of ranks_out / logsum_out is non-NULL.
 */
int mdg_ensemble_rank_u16(const void* const* members_host, int32_t K, int64_t L, int64_t n_per_outcome,
                          const uint16_t* ilog_table, int32_t Q, const MdgRankTable* ens_table, uint16_t* ranks_out,
                          float* logsum_out, void* stream);

/* Number of kernel launches the last successful mdg_pair_score call on this thread enqueued (for bench accounting). */
int mdg_last_launch_count(void);

/* Measurement hooks (bench.py's roofline line).  mdg_profile_enable(n > 0): the next n mdg_pair_score calls on this
 * thread record a CUDA event pair, on the caller's stream, around their dominant kernel (the N^2 GEMM + epilogue);
 * n = 0 disables.  mdg_profile_read synchronises on those events (the only call in this library that does), writes
 * the per-launch durations in ms to a HOST array, resets the counter and returns how many were written (-1: error). */
int mdg_profile_enable(int max_records);
int mdg_profile_read(float* ms_out_host, int max_records);

/* ------------------------------------------------------------------------------------------------------------------
 * Exact in-sample normalised rank  (reference: classwise_normalized_rank_3d_numpy + run_slice,
 *                                   notebooks/normalize_scores.py:36-74)
 *
 * For each outcome l: rank the strict lower triangle (i > j) of scores[l] (1-based, ascending), divide by
 * N(N-1)/2, write the value at [i,j] and [j,i], zero diagonal.  Ties: the reference's order among equal scores
 * is numpy-introsort-defined; this implementation gives equal scores the rank of the first of them in (i,j)
 * row-major order + their stable position (== np.argsort(kind='stable')), which satisfies
 * searchsorted_left + 1 <= rank <= searchsorted_right.
 *   scores [L, N, N] fp32 -> out [L, N, N] fp32.
 *   workspace: >= mdg_exact_rank_workspace_bytes(N) bytes.
 * ------------------------------------------------------------------------------------------------------------------ */
size_t mdg_exact_rank_workspace_bytes(int64_t N);
int mdg_exact_rank(const float* scores, int64_t L, int64_t N, float* out, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Reference-quantile builder: Q order statistics (ranks ceil(i*M/Q), i = 1..Q, M = N(N-1)/2) of each outcome's
 * strict-lower-triangle scores — the distribution normalize_scores.py ranks against — as ascending fp32 rows, the
 * input of mdg_rank_table_build.   scores [L, N, N] -> quantiles_out [L, Q].  Same workspace as mdg_exact_rank. */
int mdg_lower_triangle_quantiles(const float* scores, int64_t L, int64_t N, int32_t Q, float* quantiles_out,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fusion encoder  (reference: TransformerFusion.forward, madrigal/models/models.py:401-455, with
 *                  nn.TransformerEncoderLayer / nn.MultiheadAttention eval-mode semantics of torch 1.13)
 *
 * Weight pointers are the reference state_dict tensors, unmodified (fp32, torch [out,in] layout).
 * ------------------------------------------------------------------------------------------------------------------ */

typedef enum MdgAgg { MDG_AGG_CLS = 0, MDG_AGG_XATTN = 1, MDG_AGG_MEAN = 2, MDG_AGG_MAX = 3 } MdgAgg;
typedef enum MdgActn { MDG_ACTN_RELU = 0, MDG_ACTN_GELU = 1 } MdgActn;

#define MDG_MAX_LAYERS 8
#define MDG_MAX_TOKENS 32

typedef struct MdgFusionLayer {
  const float* in_proj_weight;  /* [3*Dl, Dl]  self_attn.in_proj_weight  */
  const float* in_proj_bias;    /* [3*Dl]                                */
  const float* out_proj_weight; /* [Dl, Dl]    self_attn.out_proj.weight */
  const float* out_proj_bias;   /* [Dl]                                  */
  const float* linear1_weight;  /* [F, Dl]                               */
  const float* linear1_bias;    /* [F]                                   */
  const float* linear2_weight;  /* [Dl, F]                               */
  const float* linear2_bias;    /* [Dl]                                  */
  const float* norm1_weight;    /* [Dl] */
  const float* norm1_bias;
  const float* norm2_weight;
  const float* norm2_bias;
} MdgFusionLayer;

typedef struct MdgFusionWeights {
  const float* embed2latent_weight; /* [Dl, E] */
  const float* embed2latent_bias;   /* [Dl]    */
  const float* latent2embed_weight; /* [E, Dl] */
  const float* latent2embed_bias;   /* [E]     */
  MdgFusionLayer layers[MDG_MAX_LAYERS];
  /* x-attn pooling (models.py:370-385, 422-443); NULL unless agg == MDG_AGG_XATTN */
  const float* x_attn_query;           /* [1, Dl] */
  const float* x_attn_kv_norm_weight;  /* [Dl] */
  const float* x_attn_kv_norm_bias;
  const float* x_attn_query_norm_weight;
  const float* x_attn_query_norm_bias;
  const float* x_attn_in_proj_weight;  /* [3*Dl, Dl] */
  const float* x_attn_in_proj_bias;
  const float* x_attn_out_proj_weight; /* [Dl, Dl] */
  const float* x_attn_out_proj_bias;
} MdgFusionWeights;

typedef struct MdgFusionCfg {
  int32_t embed_dim;  /* E  */
  int32_t num_layers; /* <= MDG_MAX_LAYERS */
  int32_t num_heads;  /* H  */
  int32_t head_dim;   /* Dl = H * head_dim */
  int32_t ffn_dim;    /* F  */
  int32_t actn;       /* MdgActn */
  int32_t norm_first; /* 0/1 */
  int32_t agg;        /* MdgAgg */
  int32_t num_tokens; /* T <= MDG_MAX_TOKENS */
} MdgFusionCfg;

/*
 *   tokens        [B, T, E] fp32, position-encoded (what models.py:853 passes as `pos_enc_sequence`)
 *   key_mask      [B, T] uint8, non-zero = token missing  (`fusion_mask`, True = masked key)
 *   src_mask      [T, T] uint8 or NULL, non-zero = query row may not attend key column (`src_mask`)
 *   pool_key_mask [T] uint8 or NULL: x-attn pooling's constant key mask (`x_attn_key_padding_mask`, models.py:382-385)
 *   z_out         [B, E] fp32
 *   precision     MdgPrecision of the nn.Linear GEMMs (LayerNorm, softmax, residual stream are always fp32)
 *   workspace     >= mdg_fusion_workspace_bytes(cfg, B, precision) bytes, 256-byte aligned
 */
size_t mdg_fusion_workspace_bytes(const MdgFusionCfg* cfg, int64_t B, int precision);
/* Optional: convert the weights to the GEMM operand format ONCE (caller-owned buffer of mdg_fusion_prepared_bytes,
 * 256-byte aligned) and pass it to every mdg_fusion_encode call.  With prepared == NULL the conversion is redone on
 * every call and the workspace must be mdg_fusion_workspace_bytes + mdg_fusion_prepared_bytes large. */
size_t mdg_fusion_prepared_bytes(const MdgFusionCfg* cfg, int precision);
int mdg_fusion_prepare(const MdgFusionWeights* w, const MdgFusionCfg* cfg, int precision, void* prepared,
                       size_t prepared_bytes, void* stream);
int mdg_fusion_encode(const MdgFusionWeights* w, const MdgFusionCfg* cfg, const void* prepared, const float* tokens,
                      const uint8_t* key_mask, const uint8_t* src_mask, const uint8_t* pool_key_mask, float* z_out,
                      int64_t B, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Measurement hook for the fused encoder kernel: with the environment variable MDG_FUSION_TRACE set, CTA 0 of every
 * fused launch records clock64() at each phase boundary (records [0,512): first epilogue warp — after each wait for
 * the MMA warp and before each hand-over; [512,1024): MMA warp — after each wait for the epilogue and after each
 * commit).  Synchronises the device, copies up to max_records values to a HOST array, returns the count (-1: error). */
int mdg_fusion_trace_read(uint64_t* clocks_out_host, int max_records);

/* Token assembly feeding mdg_fusion_encode  (reference: NovelDDIEncoder.encode, models.py:772-852).
 *   embeds [B, M, E] stacked modality embeddings in the order [non-TX..., TX...] (models.py:772), masks [B, M] uint8
 *   (non-zero = modality missing).  Sequence: [cls?][non-TX][bottleneck tokens][TX]; bottleneck / CLS tokens are
 *   never masked (models.py:805, 823).  cls_token NULL = no CLS.  normalize: L2-normalise each token (models.py:849).
 *   pos_enc [pos_len, E] is added to the first pos_len tokens (sinusoidal buffer or learnable parameter).
 *   seq_out [B, T, E], seq_mask_out [B, T] with T = M + num_bottlenecks + (cls ? 1 : 0). */
int mdg_assemble_tokens(const float* embeds, const uint8_t* masks, int64_t B, int32_t M, int32_t E, int32_t n_non_tx,
                        int32_t num_bottlenecks, const float* bottleneck_tokens, const float* cls_token,
                        const float* pos_enc, int32_t pos_len, int32_t normalize, float* seq_out,
                        uint8_t* seq_mask_out, void* stream);

/* Masked pooling over tokens: mode 0 = mean, 1 = max, 2 = sum of the un-masked tokens of each drug (reference:
 * fusion='mean' / 'add', models.py:870-878, and the 'mean'/'max' aggregations' scatter_mean/scatter_max).
 *   tokens [B, T, E], masks [B, T] (non-zero = missing) -> z_out [B, E]; a drug with no visible token gives 0. */
int mdg_masked_pool(const float* tokens, const uint8_t* masks, int64_t B, int32_t T, int32_t E, int32_t mode,
                    float* z_out, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Unimodal MLP bypass  (reference: MLPAdaptor as `uni_fuser`, models.py:459-518, 855-865; norm='ln', order='nd')
 *   x [B, dims[0]] -> y [B, dims[n_linear]];  linear i: weight [dims[i+1], dims[i]], bias [dims[i+1]];
 *   LayerNorm (ln_weight[i], ln_bias[i], over dims[i]) precedes linear i for 1 <= i < n_linear-1 when non-NULL;
 *   activation after every linear but the last.
 * ------------------------------------------------------------------------------------------------------------------ */
#define MDG_MAX_MLP_LINEAR 8
typedef struct MdgMlp {
  int32_t n_linear;
  int32_t dims[MDG_MAX_MLP_LINEAR + 1];
  int32_t actn; /* MdgActn */
  const float* weight[MDG_MAX_MLP_LINEAR];
  const float* bias[MDG_MAX_MLP_LINEAR];
  const float* ln_weight[MDG_MAX_MLP_LINEAR];
  const float* ln_bias[MDG_MAX_MLP_LINEAR];
} MdgMlp;
size_t mdg_mlp_workspace_bytes(const MdgMlp* mlp, int64_t B, int precision);
int mdg_mlp_forward(const MdgMlp* mlp, const float* x, float* y, int64_t B, int precision, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * chemCPA transcriptomic token  (reference: TxAdaptingComPert.predict, madrigal/chemcpa/chemCPA/model.py:678-697, the
 * `tx` modality encoder NovelDDIEncoder.encode calls at models.py:761-769).  The encoder MLP itself (Linear ->
 * BatchNorm1d(eval) -> ReLU chains, model.py:161-231) runs through mdg_mlp_forward with the batch-norm statistics
 * folded into the preceding linear; this call finishes the token:
 *   out[b,:] = (basal[b,:] + scale_b * drug_latent[b,:]) + cov_table[cov_idx[b],:]
 *   doser: 0 scale_b = dosage[b]; 1 'sigm', 2 'logsigm' (GeneralizedSigmoid, model.py:259-271, parameters
 *          doser_beta / doser_bias [num_drugs] indexed by drug_idx[b]); 3 scale_b = dosage[b] holds a precomputed
 *          scale (the 'amortized' doser MLP's output).
 *   drug_latent [B, dim] or NULL (use_drugs=False); cov_table [C, dim] + cov_idx [B] (int64) or both NULL
 *   (latent_basal + drug only).  All device pointers, fp32; out may alias basal. */
int mdg_tx_latent_combine(const float* basal, const float* drug_latent, const float* dosage, const int64_t* drug_idx,
                          const float* doser_beta, const float* doser_bias, int32_t doser, const float* cov_table,
                          const int64_t* cov_idx, int64_t B, int32_t dim, float* out, void* stream);

/* Per-drug 'mlp' dosers (reference: TxAdaptingComPert, chemCPA/model.py:405-416 construction, :609-621 use):
 *   scale_out[b] = sigmoid(MLP_i(dosage[b])),  i = drug_idx[b],
 *   MLP_i = Linear(1, width) -> ReLU -> [Linear(width, width) -> ReLU] x (depth - 1) -> Linear(width, 1)
 * (chemCPA `MLP([1] + [width] * depth + [1], batch_norm=False)`, one per drug).  Parameters stacked over drugs:
 *   w_in, b_in [num_drugs, width]; w_hid [num_drugs, depth - 1, width, width] (out, in), b_hid [num_drugs, depth - 1,
 *   width] (both NULL when depth == 1); w_out [num_drugs, width]; b_out [num_drugs].
 * The result is the precomputed dose scale mdg_tx_latent_combine takes with doser = 3.  width <= 256.  A drug index
 * outside [0, num_drugs) gives NaN for that sample (the reference raises IndexError).  fp32 device pointers. */
int mdg_doser_mlp(const float* dosage, const int64_t* drug_idx, int64_t B, int32_t num_drugs, int32_t width,
                  int32_t depth, const float* w_in, const float* b_in, const float* w_hid, const float* b_hid,
                  const float* w_out, const float* b_out, float* scale_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MADRIGAL_B200_H_ */
