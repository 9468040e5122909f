// C-ABI entry points of libmadrigal_b200.so (see include/madrigal_b200.h for the contract and the reference
// symbols each entry point replaces).  Host-side glue only: argument checking, workspace carving, TMA descriptor
// encoding and kernel launches on the caller's stream.  No allocation, no synchronisation, no CPU compute path (the one
// host-side routine, mdg_host_mirror_tiles in host_mirror.inl, moves ranks the GPU computed; it computes nothing).
#include <cuda.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../../include/madrigal_b200.h"
#include "exact_rank.cuh"
#include "fused_encoder.cuh"
#include "fusion_encode.cuh"
#include "pair_score.cuh"
#include "rank_table.cuh"

namespace {

thread_local char g_err[512] = "";
thread_local int g_last_launches = 0;

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define MDG_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t e_ = (expr);                                                                        \
    if (e_ != cudaSuccess) return fail(MDG_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---- optional per-launch timing of the dominant kernel (GEMM 2 of mdg_pair_score) for bench.py's roofline line
constexpr int kMaxProfile = 256;
thread_local cudaEvent_t g_prof_start[kMaxProfile];
thread_local cudaEvent_t g_prof_stop[kMaxProfile];
thread_local int g_prof_cap = 0;    // 0 = disabled
thread_local int g_prof_count = 0;  // records taken since enable/reset
thread_local int g_prof_created = 0;

// ---- driver entry point for cuTensorMapEncodeTiled (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 3-D tensor map over a row-major [batch, rows, inner] array; box = [1, box_rows, box_inner].
int make_map_3d(CUtensorMap* m, CUtensorMapDataType dt, int elem_bytes, const void* ptr, uint64_t inner,
                uint64_t rows, uint64_t batch, uint64_t row_stride_elems, uint64_t batch_stride_elems,
                uint32_t box_inner, uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(MDG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * elem_bytes, batch_stride_elems * elem_bytes};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MDG_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu batch=%llu strides=%llu,%llu box=%u,%u",
                (int)r, (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batch,
                (unsigned long long)strides[0], (unsigned long long)strides[1], box_inner, box_rows);
  return MDG_OK;
}

constexpr int kMaxDevices = 64;

// Per-device lazily initialised state is guarded by one std::once_flag per device: the library may be entered by one
// host thread per GPU at the same time (include/madrigal_b200.h, threading contract).
int num_sms() {
  static int n[kMaxDevices];
  static std::once_flag once[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  std::call_once(once[dev], [dev] {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  });
  return n[dev];
}

// Programmatic dependent launch on/off (MDG_NO_PDL=1 restores plain stream-ordered launches: A/B knob).
bool use_pdl() {
  static const bool on = getenv("MDG_NO_PDL") == nullptr;
  return on;
}

// Launch `kernel` on `stream`, optionally with the programmatic-stream-serialization attribute (the kernel MUST call
// mdg::pdl_wait() before its first global-memory access).
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                      Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && use_pdl()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

struct PairWorkspace {
  __nv_bfloat16* zr;
  __nv_bfloat16* zc;
  __nv_bfloat16* wt;
  __nv_bfloat16* y;
  unsigned int* sched;  // task counters of the dynamic scheduler (zeroed before each launch)
  int64_t nr_pad, nc_pad, ka;
  size_t total;
};

PairWorkspace carve_pair_ws(void* ws, int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision) {
  PairWorkspace w;
  w.ka = (precision == MDG_PREC_FP32) ? 2 * D : D;
  w.nr_pad = round_up(Nr, 128);
  w.nc_pad = round_up(Nc, 128);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 1023) / 1024 * 1024;
    return o;
  };
  size_t o_zr = take(static_cast<size_t>(w.nr_pad) * w.ka * 2);
  size_t o_zc = take(static_cast<size_t>(w.nc_pad) * w.ka * 2);
  size_t o_wt = take(static_cast<size_t>(L) * D * w.ka * 2);
  size_t o_y = take(static_cast<size_t>(L) * w.nr_pad * w.ka * 2);
  size_t o_sched = take(256);
  uint8_t* b = static_cast<uint8_t*>(ws);
  w.sched = reinterpret_cast<unsigned int*>(b + o_sched);
  w.zr = reinterpret_cast<__nv_bfloat16*>(b + o_zr);
  w.zc = reinterpret_cast<__nv_bfloat16*>(b + o_zc);
  w.wt = reinterpret_cast<__nv_bfloat16*>(b + o_wt);
  w.y = reinterpret_cast<__nv_bfloat16*>(b + o_y);
  w.total = off;
  return w;
}

template <int EPI, int NE>
int launch_pair_instance(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                         const CUtensorMap& tmOut2, const mdg::PairScoreParams& p, int grid, cudaStream_t stream) {
  using SM = mdg::PairSmem<NE, mdg::epi_staging_bufs(EPI, NE), mdg::epi_needs_aux32(EPI)>;
  static std::once_flag attr_once[kMaxDevices];
  static cudaError_t attr_err[kMaxDevices];
  int dev = 0;
  MDG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(MDG_ERR_UNSUPPORTED, "device index %d out of range", dev);
  std::call_once(attr_once[dev], [dev] {
    attr_err[dev] = cudaFuncSetAttribute(mdg::pair_score_kernel<EPI, NE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SM::kBytes);
  });
  MDG_CUDA(attr_err[dev]);
  MDG_CUDA(launch_ex(mdg::pair_score_kernel<EPI, NE>, dim3(grid), dim3(SM::kThreads), SM::kBytes, stream, true, tmA, tmB,
                     tmOut, tmOut2, p));
  return MDG_OK;
}

// Which instance runs the normaliser-layout rank epilogue: the software-pipelined one (8 epilogue warps, two staging
// tiles per warp; needs TMA stores) or the first-generation one (16 warps; also the path for unaligned outputs, which
// use guarded direct stores).  MDG_MIRROR_EPI=legacy|pipelined pins the choice for A/B measurements.
bool mirror_pipelined(const mdg::PairScoreParams& p) {
  static const int pin = [] {
    const char* e = getenv("MDG_MIRROR_EPI");
    if (!e) return 0;
    return strcmp(e, "legacy") == 0 ? 1 : (strcmp(e, "pipelined") == 0 ? 2 : 0);
  }();
  if (!p.use_tma_store || !p.use_tma_store2) return p.packed != 0;
  if (p.packed) return true;
  return pin != 1;
}

int launch_pair_kernel(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut,
                       mdg::PairScoreParams& p, int epi_mode, cudaStream_t stream,
                       const CUtensorMap* tmOut2_opt = nullptr) {
  const CUtensorMap& tmOut2 = tmOut2_opt ? *tmOut2_opt : tmOut;
  // task decomposition: enough tasks for ~16 per CTA (tail <= ~6%), chunks of >= 4 column blocks
  const int sms = num_sms();
  p.n_blocks = static_cast<int>((p.cols + mdg::kBN - 1) / mdg::kBN);
  p.m_blocks = static_cast<int>((p.rows + mdg::kBM * p.msub - 1) / (mdg::kBM * p.msub));
  int64_t row_tasks = static_cast<int64_t>(p.L) * p.m_blocks;
  static const int tasks_per_cta = [] {
    const char* e = getenv("MDG_TASKS_PER_CTA");  // tuning knob
    int v = e ? atoi(e) : 0;
    return v > 0 ? v : 16;
  }();
  int want_chunks = static_cast<int>((static_cast<int64_t>(tasks_per_cta) * sms + row_tasks - 1) / row_tasks);
  if (want_chunks < 1) want_chunks = 1;
  int nchunk = (p.n_blocks + want_chunks - 1) / want_chunks;
  // small problems (fewer tiles than ~4 per SM): one tile per task so that every SM gets work and the critical path
  // of a CTA is a single A load + one tile; large problems: chunks of >= 4 column blocks amortise the A load
  const bool small = row_tasks * p.n_blocks <= 4LL * sms;
  if (nchunk < 4 && !small) nchunk = 4;
  if (nchunk > p.n_blocks) nchunk = p.n_blocks;
  if (nchunk < 1) nchunk = 1;
  p.nchunk = nchunk;
  p.chunks_per_row = (p.n_blocks + nchunk - 1) / nchunk;
  int64_t tasks = row_tasks * p.chunks_per_row;
  if (tasks > 0x7fffffffLL) return fail(MDG_ERR_UNSUPPORTED, "too many tasks (%lld)", (long long)tasks);
  p.num_tasks = static_cast<int>(tasks);
  if (p.num_tasks == 0) return MDG_OK;
  const int grid = p.num_tasks < sms ? p.num_tasks : sms;
  static const int rank_warps = [] {
    const char* e = getenv("MDG_RANK_EPI_WARPS");  // tuning knob: 8 or 16 epilogue warps for the rank epilogue
    return (e && atoi(e) == 8) ? 8 : 16;
  }();
  static const int linear_warps = [] {
    const char* e = getenv("MDG_LINEAR_EPI_WARPS");  // tuning knob
    return (e && atoi(e) == 16) ? 16 : 8;
  }();
  int rc;
  switch (epi_mode) {
    case mdg::EPI_F32: rc = launch_pair_instance<mdg::EPI_F32, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream); break;
    case mdg::EPI_SIGMOID: rc = launch_pair_instance<mdg::EPI_SIGMOID, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream); break;
    case mdg::EPI_BF16_SPLIT: rc = launch_pair_instance<mdg::EPI_BF16_SPLIT, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream); break;
    case mdg::EPI_RANK_U16_MIRROR:
      rc = mirror_pipelined(p) ? launch_pair_instance<mdg::EPI_RANK_U16_MIRROR, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream)
                               : launch_pair_instance<mdg::EPI_RANK_U16_MIRROR, 16>(tmA, tmB, tmOut, tmOut2, p, grid, stream);
      break;
    case mdg::EPI_RANK_U16_PWL:
      rc = launch_pair_instance<mdg::EPI_RANK_U16_PWL, 16>(tmA, tmB, tmOut, tmOut2, p, grid, stream);
      break;
    case mdg::EPI_RANK_U16_MIRROR_PWL:
      rc = mirror_pipelined(p)
               ? launch_pair_instance<mdg::EPI_RANK_U16_MIRROR_PWL, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream)
               : launch_pair_instance<mdg::EPI_RANK_U16_MIRROR_PWL, 16>(tmA, tmB, tmOut, tmOut2, p, grid, stream);
      break;
    case mdg::EPI_TOPK: rc = launch_pair_instance<mdg::EPI_TOPK, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream); break;
    case mdg::EPI_LINEAR:
      rc = (linear_warps == 8) ? launch_pair_instance<mdg::EPI_LINEAR, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream)
                               : launch_pair_instance<mdg::EPI_LINEAR, 16>(tmA, tmB, tmOut, tmOut2, p, grid, stream);
      break;
    case mdg::EPI_RANK_U16:
      rc = (rank_warps == 8) ? launch_pair_instance<mdg::EPI_RANK_U16, 8>(tmA, tmB, tmOut, tmOut2, p, grid, stream)
                             : launch_pair_instance<mdg::EPI_RANK_U16, 16>(tmA, tmB, tmOut, tmOut2, p, grid, stream);
      break;
    default: return fail(MDG_ERR_INVALID_ARGUMENT, "bad epilogue mode %d", epi_mode);
  }
  if (rc) return rc;
  ++g_last_launches;
  return MDG_OK;
}

}  // namespace

extern "C" {

const char* mdg_last_error(void) { return g_err; }
int mdg_abi_version(void) { return MDG_ABI_VERSION; }
int mdg_last_launch_count(void) { return g_last_launches; }

int mdg_check_device(int device) {
  int dev = device;
  if (dev < 0) MDG_CUDA(cudaGetDevice(&dev));
  int major = 0;
  MDG_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(MDG_ERR_NO_DEVICE, "device %d is sm_%d0, this library is sm_100a only", dev, major);
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ rank table
int mdg_rank_table_build(const float* quantiles, int32_t L, int32_t Q, float* thresholds_out, uint32_t* lut_out,
                         float* affine_out, void* stream) {
  if (!quantiles || !thresholds_out || !lut_out || !affine_out)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_table_build: NULL pointer");
  if (L <= 0 || Q <= 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_table_build: L=%d Q=%d", L, Q);
  if (Q > MDG_RANK_MAX_Q) return fail(MDG_ERR_UNSUPPORTED, "Q=%d exceeds %d (uint16 ranks)", Q, MDG_RANK_MAX_Q);
  if (static_cast<const void*>(quantiles) == static_cast<const void*>(thresholds_out))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_table_build: quantiles and thresholds_out must not alias");
  mdg::rank_table_build_kernel<<<L, 256, 0, static_cast<cudaStream_t>(stream)>>>(quantiles, Q, thresholds_out, lut_out,
                                                                                  affine_out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

int mdg_rank_table_build_pwl(const float* quantiles, int32_t L, int32_t Q, float* thresholds_out, uint32_t* lut_out,
                             float* affine_out, float* max_dev_out, void* stream) {
  if (!quantiles || !thresholds_out || !lut_out || !affine_out)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_table_build_pwl: NULL pointer");
  if (L <= 0 || Q <= 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_table_build_pwl: L=%d Q=%d", L, Q);
  if (Q > MDG_RANK_MAX_Q) return fail(MDG_ERR_UNSUPPORTED, "Q=%d exceeds %d (uint16 ranks)", Q, MDG_RANK_MAX_Q);
  if (static_cast<const void*>(quantiles) == static_cast<const void*>(thresholds_out))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_table_build_pwl: quantiles and thresholds_out must not alias");
  mdg::rank_table_build_pwl_kernel<<<L, 256, 0, static_cast<cudaStream_t>(stream)>>>(quantiles, Q, thresholds_out,
                                                                                      lut_out, affine_out, max_dev_out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

int mdg_rank_lookup(const float* logits, int64_t n_per_outcome, const MdgRankTable* table, uint16_t* ranks,
                    void* stream) {
  if (!logits || !table || !ranks || !table->lut || !table->affine)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_lookup: NULL pointer");
  if (n_per_outcome < 0 || table->L <= 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_lookup: bad sizes");
  if (n_per_outcome == 0) return MDG_OK;
  int64_t blocks = (n_per_outcome + 255) / 256;
  if (blocks > 4 * 148) blocks = 4 * 148;
  dim3 grid(static_cast<unsigned>(blocks), static_cast<unsigned>(table->L));
  if (table->kind != MDG_RANK_LUT && table->kind != MDG_RANK_PWL)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_rank_lookup: table kind %d", table->kind);
  if (table->kind == MDG_RANK_PWL)
    mdg::rank_lookup_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, n_per_outcome, table->lut,
                                                                                        table->affine, ranks);
  else
    mdg::rank_lookup_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, n_per_outcome, table->lut,
                                                                                         table->affine, ranks);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ decoder
int64_t mdg_packed_tiles_per_outcome(int64_t N) {
  if (N <= 0) return 0;
  const int64_t nb32 = (N + 31) / 32;
  return nb32 * (nb32 + 1) / 2;
}

size_t mdg_pair_score_workspace_bytes(int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision) {
  if (Nr < 0 || Nc < 0 || D <= 0 || L < 0) return 0;
  return carve_pair_ws(nullptr, Nr, Nc, D, L, precision).total;
}

struct TopkArgs {
  const float* thresh;
  unsigned int* counts;
  unsigned long long* cand;
  int cap;
};
constexpr int kOutTopk = 100;    // internal out_mode of mdg_pair_topk
constexpr int kOutGather = 101;  // internal out_mode of mdg_pair_score_gather
struct GatherArgs {
  const int32_t *labels, *heads, *tails;
  int64_t n;
  int sigmoid;
};

// `prepared_wt` != NULL: the decoder weights already in GEMM-operand form (mdg_pair_prepare), W is not read.
static int pair_score_impl(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                           int64_t L, int precision, int out_mode, int pairs, int normalize_rows,
                           const MdgRankTable* table, void* out, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream, const TopkArgs* topk, const GatherArgs* gather = nullptr,
                           const __nv_bfloat16* prepared_wt = nullptr) {
  if (!z_rows || !z_cols || (!W && !prepared_wt) || (!out && out_mode != kOutTopk))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: NULL pointer");
  if (Nr < 0 || Nc < 0 || L < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: negative size");
  if (D != 64 && D != 128 && D != 192 && D != 256)
    return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_score: D=%lld (supported: 64, 128, 192, 256)", (long long)D);
  if (precision != MDG_PREC_BF16 && precision != MDG_PREC_FP32)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: precision=%d", precision);
  if (out_mode != kOutTopk && out_mode != kOutGather && (out_mode < MDG_OUT_LOGIT_F32 || out_mode > MDG_OUT_RANK_U16))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: out_mode=%d", out_mode);
  if (pairs != MDG_PAIRS_FULL && out_mode != kOutTopk && out_mode != MDG_OUT_RANK_U16)
    return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_score: pairs=%d is implemented for the rank and top-k outputs only", pairs);
  if (pairs != MDG_PAIRS_FULL && pairs != MDG_PAIRS_SYMMETRIC && pairs != MDG_PAIRS_PACKED_TILES)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: pairs=%d", pairs);
  if (pairs != MDG_PAIRS_FULL && (z_rows != z_cols || Nr != Nc))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: pairs=%d needs z_rows == z_cols (one catalogue)", pairs);
  if (pairs == MDG_PAIRS_PACKED_TILES && out_mode != MDG_OUT_RANK_U16)
    return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_score: MDG_PAIRS_PACKED_TILES is a rank-output layout");
  if (pairs == MDG_PAIRS_PACKED_TILES && reinterpret_cast<uintptr_t>(out) % 16 != 0)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: packed-tile output must be 16-byte aligned");
  if (Nr > (1 << 30) || Nc > (1 << 30) || L > (1 << 24))
    return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_score: sizes too large");
  if (out_mode == MDG_OUT_RANK_U16) {
    if (!table || !table->lut || !table->affine)
      return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: MDG_OUT_RANK_U16 needs a rank table");
    if (table->L < L) return fail(MDG_ERR_INVALID_ARGUMENT, "rank table has %d outcomes, need %lld", table->L, (long long)L);
    if (table->kind != MDG_RANK_LUT && table->kind != MDG_RANK_PWL)
      return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: rank table kind %d", table->kind);
  }
  if (Nr == 0 || Nc == 0 || L == 0) return MDG_OK;
  if (!workspace) return fail(MDG_ERR_WORKSPACE, "mdg_pair_score: NULL workspace");
  if (reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return fail(MDG_ERR_WORKSPACE, "mdg_pair_score: workspace must be 256-byte aligned");
  PairWorkspace ws = carve_pair_ws(workspace, Nr, Nc, D, L, precision);
  if (workspace_bytes < ws.total)
    return fail(MDG_ERR_WORKSPACE, "mdg_pair_score: workspace %zu < required %zu", workspace_bytes, ws.total);
  int rc = mdg_check_device(-1);
  if (rc) return rc;

  const int split = precision == MDG_PREC_FP32;
  const int nterm = split ? 3 : 1;
  static const int msub_pin = [] {
    const char* e = getenv("MDG_MSUB");  // tuning knob: 1 = 128-row tasks (half the resident A operand)
    return (e && atoi(e) == 1) ? 1 : 0;
  }();
  const int msub = (split || msub_pin == 1) ? 1 : 2;
  const int kb = static_cast<int>(D / 64);
  const int64_t ka = ws.ka;

  // ---- operand preparation: fp32 -> bf16 (hi | lo), W transposed to K-major (skipped for a prepared decoder).
  //      One catalogue scored against itself (z_rows == z_cols) is converted once.  The LAST conversion kernel also
  //      zeroes the dynamic scheduler's task counter, so that convert_z -> GEMM 1 -> GEMM 2 is an unbroken PDL chain.
  static const bool static_sched = getenv("MDG_STATIC_SCHED") != nullptr;  // tuning/debug knob
  const bool dyn_sched = !static_sched && out_mode != kOutGather;
  const bool same_z = z_rows == z_cols && Nr == Nc;
  if (same_z) ws.zc = ws.zr;
  {
    if (!prepared_wt) {
      dim3 g(static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>(L));
      mdg::convert_w_kernel<<<g, 256, 0, stream>>>(W, static_cast<int>(D), split, ws.wt);
      MDG_CUDA(cudaGetLastError());
      ++g_last_launches;
    }
    const int wpb = 8;
    if (!same_z) {
      mdg::convert_z_kernel<<<static_cast<unsigned>((ws.nc_pad + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
          z_cols, static_cast<int>(Nc), static_cast<int>(ws.nc_pad), static_cast<int>(D), split, normalize_rows, ws.zc,
          nullptr);
      MDG_CUDA(cudaGetLastError());
      ++g_last_launches;
    }
    MDG_CUDA(launch_ex(mdg::convert_z_kernel, dim3(static_cast<unsigned>((ws.nr_pad + wpb - 1) / wpb)), dim3(wpb * 32), 0,
                       stream, /*pdl=*/true, z_rows, static_cast<int>(Nr),
                       static_cast<int>(ws.nr_pad), static_cast<int>(D), split, normalize_rows, ws.zr,
                       dyn_sched ? ws.sched : static_cast<unsigned int*>(nullptr)));
    ++g_last_launches;
  }
  const __nv_bfloat16* wt = prepared_wt ? prepared_wt : ws.wt;

  CUtensorMap tmA, tmB, tmOut;
  // ---- GEMM 1:  Y[l] = z_rows . W[l]      A = zr [1, nr_pad, ka], B = wt [L, D, ka], out = y [L, nr_pad, ka]
  {
    rc = make_map_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ws.zr, ka, ws.nr_pad, 1, ka, ws.nr_pad * ka, 64, 128,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_3d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, ka, D, L, ka, D * ka, 64, 128,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static const bool narrow_store = getenv("MDG_GEMM1_NARROW_STORE") != nullptr;  // A/B knob: 32-column (64-byte-row) tiles
    if (narrow_store)
      rc = make_map_3d(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ws.y, ka, ws.nr_pad, L, ka, ws.nr_pad * ka, 32, 32,
                       CU_TENSOR_MAP_SWIZZLE_64B);
    else
      rc = make_map_3d(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ws.y, ka, ws.nr_pad, L, ka, ws.nr_pad * ka, 64, 32,
                       CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    mdg::PairScoreParams p;
    memset(&p, 0, sizeof(p));
    p.wide_store = narrow_store ? 0 : 1;
    p.L = static_cast<int>(L);
    p.rows = static_cast<int>(ws.nr_pad);
    p.cols = static_cast<int>(D);
    p.kb = kb;
    p.nterm = nterm;
    p.msub = msub;
    p.a_batched = 0;
    p.b_batched = 1;
    p.use_tma_store = 1;
    p.lo_col_offset = static_cast<int>(D);
    p.write_lo = split;
    p.a_reuse = getenv("MDG_GEMM1_NO_REUSE") ? 0 : 1;  // z row blocks stay resident across outcomes
    p.out = ws.y;
    p.out_ld = ka;
    p.out_batch_stride = ws.nr_pad * ka;
    rc = launch_pair_kernel(tmA, tmB, tmOut, p, mdg::EPI_BF16_SPLIT, stream);
    if (rc) return rc;
  }
  if (out_mode == kOutGather) {  // listed triples only: out[t] = Y[l_t, h_t, :] . zc[t_t, :]
    const long long blocks = (gather->n + 7) / 8;
    mdg::gather_dot_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        ws.y, ws.zc, ws.nr_pad, static_cast<int>(ka), static_cast<int>(D), split, static_cast<int>(L),
        static_cast<int>(Nr), static_cast<int>(Nc), gather->labels, gather->heads, gather->tails, gather->n,
        gather->sigmoid, static_cast<float*>(out));
    MDG_CUDA(cudaGetLastError());
    ++g_last_launches;
    return MDG_OK;
  }
  // ---- GEMM 2:  S[l] = Y[l] . z_cols^T    A = y [L, nr_pad, ka], B = zc [1, nc_pad, ka], out [L, Nr, Nc]
  {
    rc = make_map_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ws.y, ka, ws.nr_pad, L, ka, ws.nr_pad * ka, 64, 128,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_map_3d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ws.zc, ka, ws.nc_pad, 1, ka, ws.nc_pad * ka, 64, 128,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    mdg::PairScoreParams p;
    memset(&p, 0, sizeof(p));
    p.L = static_cast<int>(L);
    p.rows = static_cast<int>(Nr);
    p.cols = static_cast<int>(Nc);
    p.kb = kb;
    p.nterm = nterm;
    p.msub = msub;
    p.a_batched = 1;
    p.b_batched = 0;
    p.out = out;
    p.out_ld = Nc;
    p.out_batch_stride = Nr * Nc;
    if (dyn_sched) p.sched_counter = ws.sched;  // zeroed by convert_z_kernel above
    int elem = 4;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    int epi = mdg::EPI_F32;
    if (out_mode == kOutTopk) {
      epi = mdg::EPI_TOPK;
      p.topk_thresh = topk->thresh;
      p.topk_count = topk->counts;
      p.topk_cand = topk->cand;
      p.topk_cap = topk->cap;
      if (getenv("MDG_DEBUG_MAINLOOP_ONLY")) p.topk_cap = -1;  // debug: time the TMA/MMA mainloop without the epilogue
      p.lower_only = pairs == MDG_PAIRS_SYMMETRIC;
    } else if (out_mode == MDG_OUT_SIGMOID_F32) {
      epi = mdg::EPI_SIGMOID;
    } else if (out_mode == MDG_OUT_RANK_U16) {
      const bool pwl = table->kind == MDG_RANK_PWL;
      epi = pwl ? mdg::EPI_RANK_U16_PWL : mdg::EPI_RANK_U16;
      p.lut = table->lut;
      p.affine = table->affine;
      elem = 2;
      dt = CU_TENSOR_MAP_DATA_TYPE_UINT16;
      if (pairs == MDG_PAIRS_SYMMETRIC || pairs == MDG_PAIRS_PACKED_TILES) {
        epi = pwl ? mdg::EPI_RANK_U16_MIRROR_PWL : mdg::EPI_RANK_U16_MIRROR;
        p.lower_only = 1;
        p.mirror = 1;
        p.packed = pairs == MDG_PAIRS_PACKED_TILES;
        // measurement knob (results invalid): 1 = skip the table look-ups, 2 = skip the staging fills and stores
        static const int dbg_epi = [] { const char* e = getenv("MDG_DEBUG_EPI"); return e ? atoi(e) : 0; }();
        p.debug_epi = dbg_epi;
        // long rows: evict-first hint on the rank stores (pair_score.cuh); MDG_STORE_HINT=normal|evict_first pins it
        static const int hint_pin = [] {
          const char* e = getenv("MDG_STORE_HINT");
          return !e ? 0 : (strcmp(e, "evict_first") == 0 ? 2 : (strcmp(e, "normal") == 0 ? 1 : 0));
        }();
        p.store_evict_first = hint_pin ? (hint_pin == 2) : (Nc >= 8192);
      }
    }
    // TMA store needs 16-byte aligned base and row pitch; otherwise guarded direct stores from registers
    const bool tma_ok = out_mode != kOutTopk && (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                        ((Nc * elem) % 16 == 0) && (getenv("MDG_FORCE_DIRECT_STORE") == nullptr);
    p.use_tma_store = tma_ok ? 1 : 0;
    CUtensorMap tmOutT = tmA;  // mirror mode: the same tensor and box shape, used for the transposed tiles
    if (p.packed) {
      // packed tiles: per outcome an array of T = nb32 (nb32 + 1) / 2 contiguous 32 x 32 uint16 tiles = a [T * 32, 32]
      // matrix with 64-byte rows; one TMA store per tile
      const int64_t nb32 = (Nr + 31) / 32, tiles = nb32 * (nb32 + 1) / 2;
      rc = make_map_3d(&tmOut, dt, elem, out, 32, tiles * 32, L, 32, tiles * 1024, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      p.use_tma_store = 1;
      p.out_ld = 32;
      p.out_batch_stride = tiles * 1024;
    } else if (tma_ok) {
      rc = make_map_3d(&tmOut, dt, elem, out, Nc, Nr, L, Nc, Nr * Nc, 64 / elem, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      if (p.mirror) {
        rc = make_map_3d(&tmOutT, dt, elem, out, Nc, Nr, L, Nc, Nr * Nc, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
        p.use_tma_store2 = 1;
      }
    } else {
      tmOut = tmA;  // unused
    }
    const bool prof = g_prof_cap > 0 && g_prof_count < g_prof_cap;
    if (prof) MDG_CUDA(cudaEventRecord(g_prof_start[g_prof_count], stream));
    rc = launch_pair_kernel(tmA, tmB, tmOut, p, epi, stream, &tmOutT);
    if (rc) return rc;
    if (prof) {
      MDG_CUDA(cudaEventRecord(g_prof_stop[g_prof_count], stream));
      ++g_prof_count;
    }
  }
  return MDG_OK;
}

int mdg_pair_score(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                   int64_t L, int precision, int out_mode, int pairs, int normalize_rows, const MdgRankTable* table,
                   void* out, void* workspace, size_t workspace_bytes, void* stream_v) {
  g_last_launches = 0;
  if (out_mode == kOutTopk) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score: out_mode=%d", out_mode);
  return pair_score_impl(z_rows, z_cols, W, Nr, Nc, D, L, precision, out_mode, pairs, normalize_rows, table, out,
                         workspace, workspace_bytes, static_cast<cudaStream_t>(stream_v), nullptr);
}

// ------------------------------------------------------------------------------------------------ prepared decoder
static size_t prepared_wt_bytes(int64_t D, int64_t L, int precision) {
  const int64_t ka = (precision == MDG_PREC_FP32) ? 2 * D : D;
  return static_cast<size_t>((L * D * ka * 2 + 1023) / 1024 * 1024);
}

size_t mdg_pair_prepared_bytes(int64_t D, int64_t L, int precision) {
  if (D <= 0 || L < 0 || (precision != MDG_PREC_BF16 && precision != MDG_PREC_FP32)) return 0;
  return prepared_wt_bytes(D, L > 0 ? L : 1, precision);
}

int mdg_pair_prepare(const float* W, int64_t D, int64_t L, int precision, void* prepared, size_t prepared_bytes,
                     void* stream_v) {
  if (!W || !prepared) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_prepare: NULL pointer");
  if (D != 64 && D != 128 && D != 192 && D != 256)
    return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_prepare: D=%lld (supported: 64, 128, 192, 256)", (long long)D);
  if (precision != MDG_PREC_BF16 && precision != MDG_PREC_FP32)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_prepare: precision=%d", precision);
  if (L < 0 || L > (1 << 24)) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_prepare: L=%lld", (long long)L);
  if (L == 0) return MDG_OK;
  if (reinterpret_cast<uintptr_t>(prepared) % 256 != 0 || prepared_bytes < prepared_wt_bytes(D, L, precision))
    return fail(MDG_ERR_WORKSPACE, "mdg_pair_prepare: buffer (%zu B) too small or misaligned, need %zu", prepared_bytes,
                prepared_wt_bytes(D, L, precision));
  dim3 g(static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>(L));
  mdg::convert_w_kernel<<<g, 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      W, static_cast<int>(D), precision == MDG_PREC_FP32, static_cast<__nv_bfloat16*>(prepared));
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

int mdg_pair_score_prepared(const float* z_rows, const float* z_cols, const void* prepared, int64_t first_outcome,
                            int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision, int out_mode, int pairs,
                            int normalize_rows, const MdgRankTable* table, void* out, void* workspace,
                            size_t workspace_bytes, void* stream_v) {
  g_last_launches = 0;
  if (out_mode == kOutTopk || out_mode == kOutGather)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_prepared: out_mode=%d", out_mode);
  if (!prepared || first_outcome < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_prepared: bad prepared handle");
  if (D <= 0 || (precision != MDG_PREC_BF16 && precision != MDG_PREC_FP32))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_prepared: D=%lld precision=%d", (long long)D, precision);
  const int64_t ka = (precision == MDG_PREC_FP32) ? 2 * D : D;
  const __nv_bfloat16* wt = static_cast<const __nv_bfloat16*>(prepared) + first_outcome * D * ka;
  return pair_score_impl(z_rows, z_cols, nullptr, Nr, Nc, D, L, precision, out_mode, pairs, normalize_rows, table, out,
                         workspace, workspace_bytes, static_cast<cudaStream_t>(stream_v), nullptr, nullptr, wt);
}

// ------------------------------------------------------------------------------------------------ peer all-gather
int mdg_peer_allgather(const float* shard, int64_t shard_rows, int64_t row_offset, int32_t D,
                       void* const* peer_tables_host, void* const* peer_flags_host, int32_t world, int32_t rank,
                       uint32_t epoch, void* stream_v) {
  if (!peer_tables_host || !peer_flags_host) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_peer_allgather: NULL pointer table");
  if (world < 1 || world > mdg::kMaxPeers || rank < 0 || rank >= world)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_peer_allgather: world=%d rank=%d (1..%d ranks)", world, rank, mdg::kMaxPeers);
  if (shard_rows < 0 || row_offset < 0 || D <= 0 || (D % 4) != 0)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_peer_allgather: rows=%lld offset=%lld D=%d (D must be a multiple of 4)",
                (long long)shard_rows, (long long)row_offset, D);
  if (shard_rows > 0 && (!shard || reinterpret_cast<uintptr_t>(shard) % 16 != 0))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_peer_allgather: shard must be a 16-byte aligned device pointer");
  mdg::PeerTable pt;
  memset(&pt, 0, sizeof(pt));
  for (int r = 0; r < world; ++r) {
    if (!peer_tables_host[r] || !peer_flags_host[r] || reinterpret_cast<uintptr_t>(peer_tables_host[r]) % 16 != 0)
      return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_peer_allgather: table/flag pointer of rank %d is NULL or misaligned", r);
    pt.buf[r] = static_cast<float*>(peer_tables_host[r]);
    pt.flags[r] = static_cast<unsigned int*>(peer_flags_host[r]);
  }
  const long long n16 = shard_rows * (D / 4), off16 = row_offset * (D / 4);
  long long blocks = (n16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  MDG_CUDA(launch_ex(mdg::peer_allgather_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0,
                     static_cast<cudaStream_t>(stream_v), /*pdl=*/true, reinterpret_cast<const float4*>(shard), n16, off16,
                     pt, static_cast<int>(world), static_cast<int>(rank), static_cast<unsigned int>(epoch)));
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ row normalisation
int mdg_l2_normalize_rows(const float* x, int64_t rows, int32_t dim, float* out, void* stream_v) {
  if (!x || !out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_l2_normalize_rows: NULL pointer");
  if (rows < 0 || dim <= 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_l2_normalize_rows: rows=%lld dim=%d", (long long)rows, dim);
  if (rows == 0) return MDG_OK;
  const long long blocks = (rows + 7) / 8;  // one warp per row
  if (blocks > 0x7fffffffLL) return fail(MDG_ERR_UNSUPPORTED, "mdg_l2_normalize_rows: too many rows");
  mdg::l2_normalize_rows_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      x, rows, dim, out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ triple gather
int mdg_pair_score_gather(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                          int64_t L, int precision, int normalize_rows, const int32_t* labels, const int32_t* heads,
                          const int32_t* tails, int64_t n, int out_mode, float* out, void* workspace,
                          size_t workspace_bytes, void* stream_v) {
  g_last_launches = 0;
  if (n < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_gather: n=%lld", (long long)n);
  if (out_mode != MDG_OUT_LOGIT_F32 && out_mode != MDG_OUT_SIGMOID_F32)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_gather: out_mode=%d (logit or sigmoid)", out_mode);
  if (n == 0) return MDG_OK;
  if (!labels || !heads || !tails || !out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_gather: NULL pointer");
  if (n > (1LL << 34)) return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_score_gather: too many triples");
  if (Nr == 0 || Nc == 0 || L == 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_score_gather: triples over an empty tensor");
  GatherArgs g{labels, heads, tails, n, out_mode == MDG_OUT_SIGMOID_F32};
  return pair_score_impl(z_rows, z_cols, W, Nr, Nc, D, L, precision, kOutGather, MDG_PAIRS_FULL, normalize_rows, nullptr,
                         out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream_v), nullptr, &g);
}

// ------------------------------------------------------------------------------------------------ ensembles
int mdg_ensemble_reduce(const void* const* members_host, int32_t K, int64_t n, int mode, float rank_scale, float* out,
                        void* stream_v) {
  if (!members_host || !out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_reduce: NULL pointer");
  if (K < 1 || K > MDG_MAX_ENSEMBLE) return fail(MDG_ERR_UNSUPPORTED, "mdg_ensemble_reduce: K=%d (1..%d)", K, MDG_MAX_ENSEMBLE);
  if (n < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_reduce: n < 0");
  if (mode < MDG_ENS_MEAN_F32 || mode > MDG_ENS_GMEAN_RANK_U16)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_reduce: mode=%d", mode);
  if (n == 0) return MDG_OK;
  mdg::EnsemblePtrs ptrs;
  for (int k = 0; k < 16; ++k) ptrs.p[k] = k < K ? members_host[k] : nullptr;
  for (int k = 0; k < K; ++k)
    if (!ptrs.p[k]) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_reduce: member %d is NULL", k);
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const unsigned g = static_cast<unsigned>(blocks);
  if (mode == MDG_ENS_MEAN_F32) mdg::ensemble_reduce_kernel<0><<<g, 256, 0, stream>>>(ptrs, K, n, rank_scale, out);
  else if (mode == MDG_ENS_GMEAN_F32) mdg::ensemble_reduce_kernel<1><<<g, 256, 0, stream>>>(ptrs, K, n, rank_scale, out);
  else mdg::ensemble_reduce_kernel<2><<<g, 256, 0, stream>>>(ptrs, K, n, rank_scale, out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

int mdg_ensemble_rank_u16(const void* const* members_host, int32_t K, int64_t L, int64_t n_per_outcome,
                          const uint16_t* ilog_table, int32_t Q, const MdgRankTable* ens_table, uint16_t* ranks_out,
                          float* logsum_out, void* stream_v) {
  if (!members_host || !ilog_table) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_rank_u16: NULL pointer");
  if ((ranks_out == nullptr) == (logsum_out == nullptr))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_rank_u16: give exactly one of ranks_out and logsum_out");
  if (K < 1 || K > MDG_MAX_ENSEMBLE) return fail(MDG_ERR_UNSUPPORTED, "mdg_ensemble_rank_u16: K=%d (1..%d)", K, MDG_MAX_ENSEMBLE);
  if (L < 0 || n_per_outcome < 0 || L > 65535) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_rank_u16: bad sizes");
  if (Q < 1 || Q > MDG_RANK_MAX_Q) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_rank_u16: Q=%d", Q);
  if (ranks_out) {
    if (!ens_table || !ens_table->lut || !ens_table->affine)
      return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_rank_u16: rank output needs the ensemble rank table");
    if (ens_table->kind != MDG_RANK_LUT) return fail(MDG_ERR_UNSUPPORTED, "mdg_ensemble_rank_u16: exact-LUT tables only");
    if (ens_table->L < L) return fail(MDG_ERR_INVALID_ARGUMENT, "ensemble table has %d outcomes, need %lld", ens_table->L, (long long)L);
  }
  if ((n_per_outcome % 8) != 0 && L > 1)
    return fail(MDG_ERR_UNSUPPORTED, "mdg_ensemble_rank_u16: n_per_outcome must be a multiple of 8 when L > 1 (16-byte accesses)");
  if (L == 0 || n_per_outcome == 0) return MDG_OK;
  mdg::EnsemblePtrs ptrs;
  for (int k = 0; k < 16; ++k) ptrs.p[k] = k < K ? members_host[k] : nullptr;
  for (int k = 0; k < K; ++k)
    if (!ptrs.p[k] || reinterpret_cast<uintptr_t>(ptrs.p[k]) % 16 != 0)
      return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_ensemble_rank_u16: member %d is NULL or not 16-byte aligned", k);
  const size_t smem = ((static_cast<size_t>(Q) + 1) * 2 + 15) / 16 * 16 + (ranks_out ? MDG_RANK_LUT_ENTRIES * 4 : 0);
  static std::once_flag once[kMaxDevices][2];
  static cudaError_t err[kMaxDevices][2];
  int dev = 0;
  MDG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(MDG_ERR_UNSUPPORTED, "device index %d out of range", dev);
  const int which = ranks_out ? 1 : 0;
  std::call_once(once[dev][which], [dev, which] {
    err[dev][which] = which ? cudaFuncSetAttribute(mdg::ensemble_rank_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                            : cudaFuncSetAttribute(mdg::ensemble_rank_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  });
  MDG_CUDA(err[dev][which]);
  if (smem > 200 * 1024) return fail(MDG_ERR_UNSUPPORTED, "mdg_ensemble_rank_u16: Q too large for shared memory");
  long long bx = (n_per_outcome / 8 + 255) / 256;
  const long long cap = (static_cast<long long>(num_sms()) * 4 + L - 1) / L;  // ~4 blocks per SM in total
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(L));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (ranks_out)
    mdg::ensemble_rank_kernel<true><<<grid, 256, smem, stream>>>(ptrs, K, n_per_outcome, ilog_table, Q, ens_table->lut,
                                                                 ens_table->affine, ranks_out, nullptr);
  else
    mdg::ensemble_rank_kernel<false><<<grid, 256, smem, stream>>>(ptrs, K, n_per_outcome, ilog_table, Q, nullptr, nullptr,
                                                                  nullptr, logsum_out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ top-k output
struct TopkWs {
  unsigned int* counts;
  int* offsets;
  unsigned long long *cand, *keys, *sorted;
  void* cub_temp;
  size_t cub_bytes, pair_off, total;
};

static TopkWs plan_topk(void* ws, int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision, int cap) {
  TopkWs t;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) / 256 * 256;
    return o;
  };
  const size_t n = static_cast<size_t>(L > 0 ? L : 1) * cap;
  size_t cub_bytes = 0;
  cub::DeviceSegmentedRadixSort::SortKeys(nullptr, cub_bytes, static_cast<const unsigned long long*>(nullptr),
                                          static_cast<unsigned long long*>(nullptr), static_cast<long long>(n),
                                          static_cast<int>(L), static_cast<const int*>(nullptr),
                                          static_cast<const int*>(nullptr));
  const size_t o_counts = take((L + 1) * 4), o_off = take((L + 2) * 4), o_cand = take(n * 8), o_keys = take(n * 8),
               o_sorted = take(n * 8), o_cub = take(cub_bytes ? cub_bytes : 1);
  t.pair_off = off;
  off += carve_pair_ws(nullptr, Nr, Nc, D, L, precision).total;
  uint8_t* b = static_cast<uint8_t*>(ws);
  t.counts = reinterpret_cast<unsigned int*>(b + o_counts);
  t.offsets = reinterpret_cast<int*>(b + o_off);
  t.cand = reinterpret_cast<unsigned long long*>(b + o_cand);
  t.keys = reinterpret_cast<unsigned long long*>(b + o_keys);
  t.sorted = reinterpret_cast<unsigned long long*>(b + o_sorted);
  t.cub_temp = b + o_cub;
  t.cub_bytes = cub_bytes;
  t.total = off;
  return t;
}

size_t mdg_pair_topk_workspace_bytes(int64_t Nr, int64_t Nc, int64_t D, int64_t L, int precision, int32_t cap) {
  if (Nr < 0 || Nc < 0 || D <= 0 || L < 0 || cap <= 0) return 0;
  return plan_topk(nullptr, Nr, Nc, D, L, precision, cap).total;
}

int mdg_pair_topk(const float* z_rows, const float* z_cols, const float* W, int64_t Nr, int64_t Nc, int64_t D,
                  int64_t L, int precision, int pairs, int normalize_rows, const float* thresholds, int32_t k,
                  int32_t cap, float* scores_out, int32_t* rows_out, int32_t* cols_out, int32_t* status_out,
                  void* workspace, size_t workspace_bytes, void* stream_v) {
  g_last_launches = 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!thresholds || !scores_out || !rows_out || !cols_out || !status_out)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_topk: NULL pointer");
  if (k <= 0 || cap < k) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_pair_topk: need 0 < k <= cap (k=%d cap=%d)", k, cap);
  if (Nr * Nc > 0xFFFFFFFFLL) return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_topk: Nr*Nc exceeds 32-bit pair indices");
  if (static_cast<int64_t>(L) * cap > 0x7FFFFFFFLL) return fail(MDG_ERR_UNSUPPORTED, "mdg_pair_topk: L*cap too large");
  if (L == 0) return MDG_OK;
  if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return fail(MDG_ERR_WORKSPACE, "mdg_pair_topk: workspace must be non-NULL and 256-byte aligned");
  TopkWs t = plan_topk(workspace, Nr, Nc, D, L, precision, cap);
  if (workspace_bytes < t.total) return fail(MDG_ERR_WORKSPACE, "mdg_pair_topk: workspace %zu < required %zu", workspace_bytes, t.total);
  MDG_CUDA(cudaMemsetAsync(t.counts, 0, static_cast<size_t>(L) * 4, stream));
  TopkArgs args{thresholds, t.counts, t.cand, cap};
  int rc = pair_score_impl(z_rows, z_cols, W, Nr, Nc, D, L, precision, kOutTopk, pairs, normalize_rows, nullptr, nullptr,
                           static_cast<uint8_t*>(workspace) + t.pair_off, workspace_bytes - t.pair_off, stream, &args);
  if (rc) return rc;
  if (Nr == 0 || Nc == 0) {
    // no pairs: every list is empty
  }
  dim3 g(static_cast<unsigned>((cap + 255) / 256 > 64 ? 64 : (cap + 255) / 256), static_cast<unsigned>(L));
  mdg::topk_make_keys_kernel<<<g, 256, 0, stream>>>(t.cand, t.counts, cap, t.keys, t.offsets, static_cast<int>(L));
  MDG_CUDA(cudaGetLastError());
  size_t tb = t.cub_bytes;
  MDG_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(t.cub_temp, tb, t.keys, t.sorted, static_cast<long long>(L) * cap,
                                                   static_cast<int>(L), t.offsets, t.offsets + 1, 0, 64, stream));
  dim3 g2(static_cast<unsigned>((k + 255) / 256), static_cast<unsigned>(L));
  mdg::topk_emit_kernel<<<g2, 256, 0, stream>>>(t.sorted, t.counts, cap, k, static_cast<int>(Nc), scores_out, rows_out,
                                                cols_out, status_out);
  MDG_CUDA(cudaGetLastError());
  g_last_launches += 3;
  return MDG_OK;
}

int mdg_profile_enable(int max_records) {
  if (max_records < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_profile_enable: max_records=%d", max_records);
  if (max_records > kMaxProfile) max_records = kMaxProfile;
  for (; g_prof_created < max_records; ++g_prof_created) {
    MDG_CUDA(cudaEventCreate(&g_prof_start[g_prof_created]));
    MDG_CUDA(cudaEventCreate(&g_prof_stop[g_prof_created]));
  }
  g_prof_cap = max_records;
  g_prof_count = 0;
  return MDG_OK;
}

int mdg_profile_read(float* ms_out_host, int max_records) {
  if (!ms_out_host || max_records < 0) return -1;
  int n = g_prof_count < max_records ? g_prof_count : max_records;
  for (int i = 0; i < n; ++i) {
    if (cudaEventSynchronize(g_prof_stop[i]) != cudaSuccess) return -1;
    if (cudaEventElapsedTime(&ms_out_host[i], g_prof_start[i], g_prof_stop[i]) != cudaSuccess) return -1;
  }
  g_prof_count = 0;
  return n;
}

#include "capi_fusion_fwd.inl"
#include "host_mirror.inl"
}  // extern "C"

#include "capi_fusion.inl"
