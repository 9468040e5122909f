#include <type_traits>
// C-ABI entry points of the fusion encoder, the unimodal MLP bypass and token assembly (host orchestration).
// Every nn.Linear is one launch of the tcgen05 GEMM kernel (EPI_LINEAR); see fusion_encode.cuh for the glue kernels.

namespace {

inline int kpad_of(int K) { return static_cast<int>(round_up(K, 64)); }
inline long long ka_of(int K, int split) { return static_cast<long long>(kpad_of(K)) * (split ? 2 : 1); }

struct WsPlanner {
  size_t off = 0;
  uint8_t* base = nullptr;
  template <typename T>
  T* take(size_t count) {
    size_t o = off;
    off += (count * sizeof(T) + 255) / 256 * 256;
    return base ? reinterpret_cast<T*>(base + o) : nullptr;
  }
};

// y[rows, N] = act(A . W^T + bias) (+ residual), A = bf16 operand rows [hi | lo](k_pad), W likewise ([N] rows).
int run_linear(const __nv_bfloat16* A, long long rows, const __nv_bfloat16* W, int N, int K, int split,
               const float* bias, int act, const float* residual, long long res_ld, float* out_f32, long long out_ld,
               __nv_bfloat16* out_bf16, int out_bf16_K, cudaStream_t stream) {
  if (rows == 0 || N == 0) return MDG_OK;
  const int k_pad = kpad_of(K);
  const long long ka = ka_of(K, split);
  CUtensorMap tmA, tmB;
  int rc = make_map_3d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, ka, rows, 1, ka, rows * ka, 64, 128,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_map_3d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, W, ka, N, 1, ka, static_cast<long long>(N) * ka, 64,
                   128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  mdg::PairScoreParams p;
  memset(&p, 0, sizeof(p));
  p.L = 1;
  p.rows = static_cast<int>(rows);
  p.cols = N;
  p.kb = k_pad / 64;
  p.nterm = split ? 3 : 1;
  p.k_pad = k_pad;
  const int panels_needed = split ? 2 * p.kb : p.kb;  // for one 128-row sub-tile
  if (panels_needed <= mdg::kMaxAPanels) {
    p.stream_a = 0;
    // two 128-row sub-tiles per CTA halve the B traffic, but a small GEMM wants more, smaller tasks instead
    const long long tiles256 = ((rows + 255) / 256) * ((N + 127) / 128);
    p.msub = (!split && 2 * p.kb <= mdg::kMaxAPanels && tiles256 > 2LL * num_sms()) ? 2 : 1;
  } else {
    p.stream_a = 1;
    p.msub = 1;
  }
  p.bias = bias;
  p.act = act;
  p.residual = residual;
  p.res_ld = res_ld;
  p.out_f32 = out_f32;
  p.out_ld = out_ld;
  p.out_bf16 = out_bf16;
  p.bf16_ld = out_bf16 ? ka_of(out_bf16_K, split) : 0;
  p.bf16_lo_off = out_bf16 ? kpad_of(out_bf16_K) : 0;
  p.write_lo = split;
  // outputs go through swizzled staging + TMA stores whenever base and row pitch are 16-byte aligned
  CUtensorMap tmO1 = tmA, tmO2 = tmA;
  p.use_tma_store = 0;
  p.use_tma_store2 = 0;
  if (out_f32 && reinterpret_cast<uintptr_t>(out_f32) % 16 == 0 && (out_ld * 4) % 16 == 0) {
    rc = make_map_3d(&tmO1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out_f32, N, rows, 1, out_ld, rows * out_ld, 16, 32,
                     CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    p.use_tma_store = 1;
  }
  if (out_bf16 && reinterpret_cast<uintptr_t>(out_bf16) % 16 == 0) {
    rc = make_map_3d(&tmO2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out_bf16, p.bf16_ld, rows, 1, p.bf16_ld,
                     rows * p.bf16_ld, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    p.use_tma_store2 = 1;
  }
  // in-place residual stream (h += linear(...)): fetch the residual tile by TMA through the output's tensor map
  static const bool no_res_tma = getenv("MDG_LINEAR_RES_DIRECT") != nullptr;  // A/B knob: per-lane residual row loads
  p.res_tma = (!no_res_tma && p.use_tma_store && residual != nullptr && residual == out_f32 && res_ld == out_ld) ? 1 : 0;
  return launch_pair_kernel(tmA, tmB, tmO1, p, mdg::EPI_LINEAR, stream, &tmO2);
}

int convert_rows(const float* x, long long rows, int K, long long ld_in, long long row_stride, int split,
                 __nv_bfloat16* out, cudaStream_t stream) {
  if (rows == 0) return MDG_OK;
  const int wpb = 8;
  mdg::convert_rows_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      x, rows, K, ld_in, row_stride, kpad_of(K), split, out);
  MDG_CUDA(cudaGetLastError());
  ++g_last_launches;
  return MDG_OK;
}

int ln_convert(const float* h, long long rows, int D, const float* addvec, const float* w, const float* b, int do_ln,
               float* out_f32, __nv_bfloat16* out_bf16, int split, cudaStream_t stream) {
  if (rows == 0) return MDG_OK;
  const int wpb = 8;
  mdg::ln_convert_kernel<<<static_cast<unsigned>((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(
      h, rows, D, addvec, w, b, do_ln, out_f32, out_bf16, kpad_of(D), split);
  MDG_CUDA(cudaGetLastError());
  ++g_last_launches;
  return MDG_OK;
}

struct FusionPlan {
  int E, Dl, F, H, hd, T, layers, split;
  long long chunk_drugs, chunk_rows;
  // workspace pieces
  __nv_bfloat16 *w_e2l, *w_l2e, *w_xin, *w_xout;
  __nv_bfloat16 *w_in[MDG_MAX_LAYERS], *w_out[MDG_MAX_LAYERS], *w_l1[MDG_MAX_LAYERS], *w_l2[MDG_MAX_LAYERS];
  // fused kernel only: LayerNorm weight/bias folded into the linear that consumes the normalised rows
  __nv_bfloat16 *w_in_f[MDG_MAX_LAYERS], *w_l1_f[MDG_MAX_LAYERS], *w_xin_f;
  float *in_bias_f[MDG_MAX_LAYERS], *l1_bias_f[MDG_MAX_LAYERS], *xin_bias_f;
  __nv_bfloat16 *xb, *nb, *ob, *fb, *pb;
  float *h, *qkv, *p32, *q_res, *q_proj, *pend, *xo_qres;
  size_t ob_bytes, fb_bytes, pb_bytes;
  size_t total, prepared_total;
};

int plan_fusion(const MdgFusionCfg* cfg, long long B, int precision, void* ws, void* prepared, FusionPlan* pl) {
  if (!cfg) return fail(MDG_ERR_INVALID_ARGUMENT, "fusion: NULL cfg");
  pl->E = cfg->embed_dim;
  pl->H = cfg->num_heads;
  pl->hd = cfg->head_dim;
  pl->Dl = cfg->num_heads * cfg->head_dim;
  pl->F = cfg->ffn_dim;
  pl->T = cfg->num_tokens;
  pl->layers = cfg->num_layers;
  pl->split = precision == MDG_PREC_FP32;
  if (pl->E <= 0 || pl->H <= 0 || pl->hd <= 0 || pl->F <= 0 || pl->T <= 0 || pl->layers < 0)
    return fail(MDG_ERR_INVALID_ARGUMENT, "fusion: non-positive dimension in cfg");
  if (pl->T > MDG_MAX_TOKENS) return fail(MDG_ERR_UNSUPPORTED, "fusion: T=%d > %d tokens", pl->T, MDG_MAX_TOKENS);
  if (pl->layers > MDG_MAX_LAYERS) return fail(MDG_ERR_UNSUPPORTED, "fusion: %d layers > %d", pl->layers, MDG_MAX_LAYERS);
  if (pl->Dl > 4096 || pl->F > 8192 || pl->E > 4096) return fail(MDG_ERR_UNSUPPORTED, "fusion: dimension too large");
  if (cfg->agg < MDG_AGG_CLS || cfg->agg > MDG_AGG_MAX) return fail(MDG_ERR_UNSUPPORTED, "fusion: unknown agg %d", cfg->agg);
  if (cfg->actn != MDG_ACTN_RELU && cfg->actn != MDG_ACTN_GELU)
    return fail(MDG_ERR_UNSUPPORTED, "fusion: unsupported activation %d", cfg->actn);
  const int s = pl->split;
  static const long long max_rows_cfg = [] {
    const char* e = getenv("MDG_FUSION_CHUNK_ROWS");  // tuning knob: token rows processed per internal chunk
    long long v = e ? atoll(e) : 0;
    return v >= 128 ? v : 262144LL;
  }();
  long long max_rows = max_rows_cfg;
  long long cd = max_rows / pl->T;
  if (cd < 1) cd = 1;
  if (cd > B) cd = B > 0 ? B : 1;
  pl->chunk_drugs = cd;
  pl->chunk_rows = cd * pl->T;
  const int E = pl->E, Dl = pl->Dl, F = pl->F;
  // (a) prepared weights: bf16 operand copies + the x-attn query vectors (depend on the parameters only)
  WsPlanner pw;
  pw.base = static_cast<uint8_t*>(prepared);
  pl->w_e2l = pw.take<__nv_bfloat16>(static_cast<size_t>(Dl) * ka_of(E, s));
  pl->w_l2e = pw.take<__nv_bfloat16>(static_cast<size_t>(E) * ka_of(Dl, s));
  pl->w_xin = pw.take<__nv_bfloat16>(static_cast<size_t>(2 * Dl) * ka_of(Dl, s));
  pl->w_xout = pw.take<__nv_bfloat16>(static_cast<size_t>(Dl) * ka_of(Dl, s));
  for (int i = 0; i < pl->layers; ++i) {
    pl->w_in[i] = pw.take<__nv_bfloat16>(static_cast<size_t>(3 * Dl) * ka_of(Dl, s));
    pl->w_out[i] = pw.take<__nv_bfloat16>(static_cast<size_t>(Dl) * ka_of(Dl, s));
    pl->w_l1[i] = pw.take<__nv_bfloat16>(static_cast<size_t>(F) * ka_of(Dl, s));
    pl->w_l2[i] = pw.take<__nv_bfloat16>(static_cast<size_t>(Dl) * ka_of(F, s));
    pl->w_in_f[i] = pw.take<__nv_bfloat16>(static_cast<size_t>(3 * Dl) * ka_of(Dl, 0));
    pl->w_l1_f[i] = pw.take<__nv_bfloat16>(static_cast<size_t>(F) * ka_of(Dl, 0));
  }
  pl->w_xin_f = pw.take<__nv_bfloat16>(static_cast<size_t>(2 * Dl) * ka_of(Dl, 0));
  for (int i = 0; i < pl->layers; ++i) {
    pl->in_bias_f[i] = pw.take<float>(static_cast<size_t>(3 * Dl));
    pl->l1_bias_f[i] = pw.take<float>(static_cast<size_t>(F));
  }
  pl->xin_bias_f = pw.take<float>(static_cast<size_t>(3 * Dl));
  pl->q_res = pw.take<float>(Dl);
  pl->q_proj = pw.take<float>(Dl);
  pl->pend = pw.take<float>(static_cast<size_t>(2 * pl->layers + 1) * Dl);  // fused kernel: biases owed to H per stage
  pl->xo_qres = pw.take<float>(Dl);  // fused kernel: x_attn out_proj bias + residual query
  pl->prepared_total = pw.off;
  // (b) per-call activation workspace
  WsPlanner w;
  w.base = static_cast<uint8_t*>(ws);
  const size_t R = static_cast<size_t>(pl->chunk_rows), C = static_cast<size_t>(pl->chunk_drugs);
  pl->xb = w.take<__nv_bfloat16>(R * ka_of(E, s));
  pl->nb = w.take<__nv_bfloat16>(R * ka_of(Dl, s));
  pl->ob_bytes = R * ka_of(Dl, s) * 2;
  pl->ob = w.take<__nv_bfloat16>(R * ka_of(Dl, s));
  pl->fb_bytes = R * ka_of(F, s) * 2;
  pl->fb = w.take<__nv_bfloat16>(R * ka_of(F, s));
  pl->pb_bytes = C * ka_of(Dl, s) * 2;
  pl->pb = w.take<__nv_bfloat16>(C * ka_of(Dl, s));
  pl->h = w.take<float>(R * Dl);
  pl->qkv = w.take<float>(R * 3 * Dl > R * static_cast<size_t>(E) ? R * 3 * Dl : R * static_cast<size_t>(E));
  pl->p32 = w.take<float>(C * Dl);
  pl->total = w.off;
  return MDG_OK;
}


// ---- fused single-kernel encoder (fused_encoder.cuh): eligibility + launch
bool fused_encoder_eligible(const MdgFusionCfg* cfg, const FusionPlan& pl) {
  if (pl.split || !cfg->norm_first) return false;
  if (pl.Dl % 64 != 0 || pl.Dl > 256) return false;
  if (pl.hd != 16 && pl.hd != 32) return false;
  if (pl.E % 16 != 0 || pl.E > 256) return false;
  if (pl.T > 32 || pl.T < 1) return false;
  if (pl.F % 4 != 0) return false;  // bias vectors are staged in 16-byte pieces
  return true;
}

// the fused kernel reads bias / LayerNorm vectors with 16-byte loads
bool fused_encoder_aligned(const MdgFusionWeights* w, const MdgFusionCfg* cfg, const float* tokens, const float* z_out) {
  auto ok = [](const void* p) { return reinterpret_cast<uintptr_t>(p) % 16 == 0; };
  if (!ok(tokens) || !ok(z_out) || !ok(w->latent2embed_bias)) return false;
  if (cfg->agg == MDG_AGG_XATTN && !ok(w->x_attn_out_proj_bias)) return false;
  return true;
}

template <int HD>
int launch_fused_instance(const CUtensorMap* tm, const mdg::FusedEncParams& p, int grid, cudaStream_t stream) {
  static std::once_flag attr_once[kMaxDevices];
  static cudaError_t attr_err[kMaxDevices];
  int dev = 0;
  MDG_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return fail(MDG_ERR_UNSUPPORTED, "device index %d out of range", dev);
  std::call_once(attr_once[dev], [dev] {
    attr_err[dev] = cudaFuncSetAttribute(mdg::fused_encoder_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         mdg::kFeSmemBytes);
  });
  MDG_CUDA(attr_err[dev]);
  // programmatic dependent launch: barrier init, TMEM allocation and descriptor prefetch overlap the previous
  // kernel's tail; the kernel waits (griddepcontrol.wait) before it reads the tokens
  MDG_CUDA(launch_ex(mdg::fused_encoder_kernel<HD>, dim3(grid), dim3(mdg::kFeThreads), mdg::kFeSmemBytes, stream, true,
                     tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], p));
  ++g_last_launches;
  return MDG_OK;
}

int run_fused_encoder(const MdgFusionWeights* w, const MdgFusionCfg* cfg, const FusionPlan& pl, const float* tokens,
                      const uint8_t* key_mask, const uint8_t* src_mask, const uint8_t* pool_key_mask, float* z_out,
                      long long B, cudaStream_t stream) {
  const int E = pl.E, Dl = pl.Dl, F = pl.F, hd = pl.hd;
  const long long ke = kpad_of(E), kd = kpad_of(Dl), kf = kpad_of(F);
  const int f_pad = static_cast<int>(kf);
  const int FC = (f_pad % 256 == 0) ? 256 : (f_pad % 128 == 0 ? 128 : 64);
  const int nl = pl.layers > 0 ? pl.layers : 1;
  // the per-layer weight blocks are laid out identically, so the layer index is the maps' batch coordinate
  const long long lstride = pl.layers > 1 ? (pl.w_in[1] - pl.w_in[0]) : 0;
  CUtensorMap tm[8];
  int rc;
  const CUtensorMapDataType bf = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if ((rc = make_map_3d(&tm[0], bf, 2, pl.w_e2l, ke, Dl, 1, ke, Dl * ke, 64, Dl, sw))) return rc;
  if (pl.layers > 0) {
    if ((rc = make_map_3d(&tm[1], bf, 2, pl.w_in_f[0], kd, 3 * Dl, nl, kd, lstride ? lstride : 3 * Dl * kd, 64, hd, sw))) return rc;
    if ((rc = make_map_3d(&tm[2], bf, 2, pl.w_out[0], kd, Dl, nl, kd, lstride ? lstride : Dl * kd, 64, Dl, sw))) return rc;
    if ((rc = make_map_3d(&tm[3], bf, 2, pl.w_l1_f[0], kd, F, nl, kd, lstride ? lstride : F * kd, 64, FC, sw))) return rc;
    if ((rc = make_map_3d(&tm[4], bf, 2, pl.w_l2[0], kf, Dl, nl, kf, lstride ? lstride : Dl * kf, 64, Dl, sw))) return rc;
  } else {
    tm[1] = tm[2] = tm[3] = tm[4] = tm[0];
  }
  if ((rc = make_map_3d(&tm[5], bf, 2, pl.w_l2e, kd, E, 1, kd, E * kd, 64, E, sw))) return rc;
  if (cfg->agg == MDG_AGG_XATTN) {
    if ((rc = make_map_3d(&tm[6], bf, 2, pl.w_xin_f, kd, 2 * Dl, 1, kd, 2 * Dl * kd, 64, hd, sw))) return rc;
    if ((rc = make_map_3d(&tm[7], bf, 2, pl.w_xout, kd, Dl, 1, kd, Dl * kd, 64, Dl, sw))) return rc;
  } else {
    tm[6] = tm[7] = tm[0];
  }
  mdg::FusedEncParams p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.T = pl.T; p.E = E; p.Dl = Dl; p.F = F; p.H = pl.H; p.hd = hd; p.layers = pl.layers;
  p.act = cfg->actn == MDG_ACTN_GELU ? 2 : 1;
  p.agg = cfg->agg;
  p.kp_e = static_cast<int>(ke / 64);
  p.kp_d = static_cast<int>(kd / 64);
  p.f_pad = f_pad;
  p.fc = FC;
  p.tokens = tokens;
  p.key_mask = key_mask;
  p.src_mask = src_mask;
  p.pool_mask = pool_key_mask;
  p.z_out = z_out;
  p.pend = pl.pend;
  for (int i = 0; i < pl.layers; ++i) {
    const MdgFusionLayer& L = w->layers[i];
    (void)L;
    p.in_bias[i] = pl.in_bias_f[i];  // LayerNorm-folded biases (fusion_prepare_impl)
    p.l1_bias[i] = pl.l1_bias_f[i];
  }
  p.l2e_bias = w->latent2embed_bias;
  if (cfg->agg == MDG_AGG_XATTN) {
    p.xin_bias = pl.xin_bias_f;
    p.xout_bias = w->x_attn_out_proj_bias;
    p.xq_nw = w->x_attn_query_norm_weight; p.xq_nb = w->x_attn_query_norm_bias;
    p.q_res = pl.q_res;
    p.q_proj = pl.q_proj;
    p.xo_qres = pl.xo_qres;
  }
  p.trace = getenv("MDG_FUSION_TRACE") != nullptr;  // measurement hook: phase timeline of CTA 0 (mdg_fusion_trace_read)
  p.drugs_per_tile = mdg::kFeRows / pl.T;
  p.num_tiles = (B + p.drugs_per_tile - 1) / p.drugs_per_tile;
  const int sms = num_sms();
  const int grid = static_cast<int>(p.num_tiles < sms ? p.num_tiles : sms);
  return hd == 16 ? launch_fused_instance<16>(tm, p, grid, stream) : launch_fused_instance<32>(tm, p, grid, stream);
}

// qkv rows: fp32 [R, 3 Dl] (fp32-parity mode) or bf16 [R, kpad(3 Dl)] (bf16 mode)
template <typename TIn>
int launch_attention_t(const FusionPlan& pl, const TIn* qkv, long long ld, const uint8_t* key_mask,
                       const uint8_t* src_mask, long long Bc, cudaStream_t stream) {
  const char* mma_knob = getenv("MDG_ATTENTION_MMA");
  const bool mma_all = std::is_same<TIn, __nv_bfloat16>::value && mma_knob != nullptr && mma_knob[0] == 'a' && (pl.hd == 64 || pl.hd == 128 || pl.hd == 256) && !pl.split;
  if (pl.T <= 8 && pl.hd <= 64 && !mma_all) {  // few tokens: head dimension on lanes, registers only
    const long long items = Bc * pl.H;
    long long blocks = (items + 7) / 8;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    const int kp = kpad_of(pl.Dl);
#define MDG_ATT_SMALL(TT, DPT)                                                                              \
  mdg::attention_small_kernel<TT, DPT, TIn><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(             \
      qkv, ld, key_mask, src_mask, Bc, pl.T, pl.H, pl.hd, pl.ob, kp, pl.split)
    if (pl.T <= 4 && pl.hd <= 32) MDG_ATT_SMALL(4, 1);
    else if (pl.T <= 4) MDG_ATT_SMALL(4, 2);
    else if (pl.hd <= 32) MDG_ATT_SMALL(8, 1);
    else MDG_ATT_SMALL(8, 2);
#undef MDG_ATT_SMALL
    MDG_CUDA(cudaGetLastError());
    ++g_last_launches;
    return MDG_OK;
  }
  if constexpr (std::is_same<TIn, __nv_bfloat16>::value) {
    // bf16 rows, head_dim 64, T <= 32: warp-level tensor-core attention (one (drug, head) per warp)
    const char* knob = getenv("MDG_ATTENTION_MMA");  // "0": FMA kernels; "all": also for T <= 8
    const bool off = knob != nullptr && knob[0] == '0';
    const bool all = knob != nullptr && knob[0] == 'a';
    if ((pl.hd == 64 || pl.hd == 128 || pl.hd == 256) && pl.T <= 32 && !pl.split && !off && (pl.T > 8 || all) && ld % 8 == 0) {
      const int warps = 4;
      const size_t smem = static_cast<size_t>(warps) * 3 * 32 * pl.hd * 2;  // q, k, v tiles of 32 rows per warp
      static bool attr_mma[64] = {false};
      int dev = 0;
      MDG_CUDA(cudaGetDevice(&dev));
      if (dev >= 0 && dev < 64 && !attr_mma[dev]) {
        MDG_CUDA(cudaFuncSetAttribute(mdg::attention_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        MDG_CUDA(cudaFuncSetAttribute(mdg::attention_mma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024));
        attr_mma[dev] = true;
      }
      const long long items = Bc * pl.H;
      long long blocks = (items + warps - 1) / warps;
      const long long cap = static_cast<long long>(num_sms()) * 16;
      if (blocks > cap) blocks = cap;
      const unsigned nb = static_cast<unsigned>(blocks);
      if (pl.hd == 64)
        mdg::attention_mma_kernel<64><<<nb, warps * 32, smem, stream>>>(qkv, ld, key_mask, src_mask, Bc, pl.T, pl.H, pl.ob, kpad_of(pl.Dl));
      else if (pl.hd == 128)
        mdg::attention_mma_kernel<128><<<nb, warps * 32, smem, stream>>>(qkv, ld, key_mask, src_mask, Bc, pl.T, pl.H, pl.ob, kpad_of(pl.Dl));
      else
        mdg::attention_mma_kernel<256><<<nb, warps * 32, smem, stream>>>(qkv, ld, key_mask, src_mask, Bc, pl.T, pl.H, pl.ob, kpad_of(pl.Dl));
      MDG_CUDA(cudaGetLastError());
      ++g_last_launches;
      return MDG_OK;
    }
  }
  if ((pl.hd == 32 || pl.hd == 64) && getenv("MDG_ATTENTION_GENERIC") == nullptr) {
    // 8 < T <= 32: K rows / V columns in registers (the production shapes T = 21 / 23, head_dim 64)
    const int warps = 4;
    const size_t smem = static_cast<size_t>(warps) * pl.T * (2 * pl.hd + 4) * sizeof(float);
    static bool attr_rows[64] = {false};
    int dev = 0;
    MDG_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_rows[dev]) {
      MDG_CUDA(cudaFuncSetAttribute(mdg::attention_rows_kernel<32, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      MDG_CUDA(cudaFuncSetAttribute(mdg::attention_rows_kernel<64, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_rows[dev] = true;
    }
    const long long items = Bc * pl.H;
    long long blocks = (items + warps - 1) / warps;
    const long long cap = static_cast<long long>(num_sms()) * 12;
    if (blocks > cap) blocks = cap;
    if (pl.hd == 32)
      mdg::attention_rows_kernel<32, TIn><<<static_cast<unsigned>(blocks), warps * 32, smem, stream>>>(
          qkv, ld, key_mask, src_mask, Bc, pl.T, pl.H, pl.ob, kpad_of(pl.Dl), pl.split);
    else
      mdg::attention_rows_kernel<64, TIn><<<static_cast<unsigned>(blocks), warps * 32, smem, stream>>>(
          qkv, ld, key_mask, src_mask, Bc, pl.T, pl.H, pl.ob, kpad_of(pl.Dl), pl.split);
    MDG_CUDA(cudaGetLastError());
    ++g_last_launches;
    return MDG_OK;
  }
  int TP = 1;
  while (TP < pl.T) TP <<= 1;
  const int G = 32 / TP;
  const size_t per_warp = static_cast<size_t>(G) * (2 * static_cast<size_t>(pl.T) * (pl.hd + 1) + pl.hd) * sizeof(float);
  int warps = static_cast<int>((96 * 1024) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) return fail(MDG_ERR_UNSUPPORTED, "attention tile T=%d head_dim=%d needs %zu B of shared memory", pl.T, pl.hd, per_warp);
  const size_t smem = per_warp * warps;
  static bool attr_set[64] = {false};
  int dev = 0;
  MDG_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    MDG_CUDA(cudaFuncSetAttribute(mdg::attention_kernel<TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set[dev] = true;
  }
  const long long groups = (Bc * pl.H + G - 1) / G;
  long long blocks = (groups + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  mdg::attention_kernel<TIn><<<static_cast<unsigned>(blocks), warps * 32, smem, stream>>>(
      qkv, ld, key_mask, src_mask, Bc, pl.T, TP, pl.H, pl.hd, pl.ob, kpad_of(pl.Dl), pl.split);
  MDG_CUDA(cudaGetLastError());
  ++g_last_launches;
  return MDG_OK;
}

}  // namespace

extern "C" {

size_t mdg_fusion_workspace_bytes(const MdgFusionCfg* cfg, int64_t B, int precision) {
  FusionPlan pl;
  if (plan_fusion(cfg, B, precision, nullptr, nullptr, &pl) != MDG_OK) return 0;
  return pl.total;
}

size_t mdg_fusion_prepared_bytes(const MdgFusionCfg* cfg, int precision) {
  FusionPlan pl;
  if (plan_fusion(cfg, 1, precision, nullptr, nullptr, &pl) != MDG_OK) return 0;
  return pl.prepared_total;
}

static int check_fusion_weights(const MdgFusionWeights* w, const MdgFusionCfg* cfg) {
  if (!w->embed2latent_weight || !w->embed2latent_bias || !w->latent2embed_weight || !w->latent2embed_bias)
    return fail(MDG_ERR_INVALID_ARGUMENT, "fusion: NULL embed2latent/latent2embed weights");
  for (int i = 0; i < cfg->num_layers; ++i) {
    const MdgFusionLayer& L = w->layers[i];
    if (!L.in_proj_weight || !L.in_proj_bias || !L.out_proj_weight || !L.out_proj_bias || !L.linear1_weight ||
        !L.linear1_bias || !L.linear2_weight || !L.linear2_bias || !L.norm1_weight || !L.norm1_bias ||
        !L.norm2_weight || !L.norm2_bias)
      return fail(MDG_ERR_INVALID_ARGUMENT, "fusion: NULL weight in layer %d", i);
  }
  if (cfg->agg == MDG_AGG_XATTN &&
      (!w->x_attn_query || !w->x_attn_kv_norm_weight || !w->x_attn_kv_norm_bias || !w->x_attn_query_norm_weight ||
       !w->x_attn_query_norm_bias || !w->x_attn_in_proj_weight || !w->x_attn_in_proj_bias ||
       !w->x_attn_out_proj_weight || !w->x_attn_out_proj_bias))
    return fail(MDG_ERR_INVALID_ARGUMENT, "fusion: agg=x-attn needs the x_attn_* weights");
  return MDG_OK;
}

// weights: fp32 [out, in] (already K-major) -> bf16 GEMM operand rows; x-attn query path
static int fusion_prepare_impl(const MdgFusionWeights* w, const MdgFusionCfg* cfg, const FusionPlan& pl,
                               cudaStream_t stream) {
  int rc;
  const int E = pl.E, Dl = pl.Dl, F = pl.F, s = pl.split;
  if ((rc = convert_rows(w->embed2latent_weight, Dl, E, E, 1, s, pl.w_e2l, stream))) return rc;
  if ((rc = convert_rows(w->latent2embed_weight, E, Dl, Dl, 1, s, pl.w_l2e, stream))) return rc;
  for (int i = 0; i < pl.layers; ++i) {
    const MdgFusionLayer& L = w->layers[i];
    if ((rc = convert_rows(L.in_proj_weight, 3 * Dl, Dl, Dl, 1, s, pl.w_in[i], stream))) return rc;
    if ((rc = convert_rows(L.out_proj_weight, Dl, Dl, Dl, 1, s, pl.w_out[i], stream))) return rc;
    if ((rc = convert_rows(L.linear1_weight, F, Dl, Dl, 1, s, pl.w_l1[i], stream))) return rc;
    if ((rc = convert_rows(L.linear2_weight, Dl, F, F, 1, s, pl.w_l2[i], stream))) return rc;
  }
  if (fused_encoder_eligible(cfg, pl)) {
    mdg::FusedPendArgs a;
    memset(&a, 0, sizeof(a));
    a.e2l_bias = w->embed2latent_bias;
    for (int i = 0; i < pl.layers; ++i) {
      a.out_bias[i] = w->layers[i].out_proj_bias;
      a.l2_bias[i] = w->layers[i].linear2_bias;
    }
    a.layers = pl.layers;
    a.Dl = Dl;
    mdg::fused_pend_kernel<<<(Dl + 127) / 128, 128, 0, stream>>>(a, pl.pend);
    MDG_CUDA(cudaGetLastError());
    ++g_last_launches;
    // LayerNorm affine folded into the next linear:  LN(x) . W^T + b  =  xhat . (W * ln_w)^T + (b + W . ln_b)
    auto fold = [&](const float* W, const float* bias, const float* lnw, const float* lnb, int N, __nv_bfloat16* wout,
                    float* bout) -> int {
      mdg::fold_ln_linear_kernel<<<(N + 7) / 8, 256, 0, stream>>>(W, bias, lnw, lnb, N, Dl, kpad_of(Dl), wout, bout);
      MDG_CUDA(cudaGetLastError());
      ++g_last_launches;
      return MDG_OK;
    };
    for (int i = 0; i < pl.layers; ++i) {
      const MdgFusionLayer& L = w->layers[i];
      if ((rc = fold(L.in_proj_weight, L.in_proj_bias, L.norm1_weight, L.norm1_bias, 3 * Dl, pl.w_in_f[i], pl.in_bias_f[i]))) return rc;
      if ((rc = fold(L.linear1_weight, L.linear1_bias, L.norm2_weight, L.norm2_bias, F, pl.w_l1_f[i], pl.l1_bias_f[i]))) return rc;
    }
    if (cfg->agg == MDG_AGG_XATTN) {  // k | v rows of the pooling MHA, LN = x_attn_kv_norm; biases keep the [3 Dl] indexing
      if ((rc = fold(w->x_attn_in_proj_weight + static_cast<size_t>(Dl) * Dl, w->x_attn_in_proj_bias + Dl,
                     w->x_attn_kv_norm_weight, w->x_attn_kv_norm_bias, 2 * Dl, pl.w_xin_f, pl.xin_bias_f + Dl))) return rc;
    }
  }
  if (cfg->agg == MDG_AGG_XATTN) {
    if ((rc = convert_rows(w->x_attn_in_proj_weight + static_cast<size_t>(Dl) * Dl, 2 * Dl, Dl, Dl, 1, s, pl.w_xin, stream))) return rc;
    if ((rc = convert_rows(w->x_attn_out_proj_weight, Dl, Dl, Dl, 1, s, pl.w_xout, stream))) return rc;
    mdg::xattn_query_kernel<<<1, 256, 0, stream>>>(w->x_attn_query, w->x_attn_query_norm_weight,
                                                   w->x_attn_query_norm_bias, cfg->norm_first,
                                                   w->x_attn_in_proj_weight, w->x_attn_in_proj_bias, Dl, pl.hd,
                                                   pl.q_res, pl.q_proj);
    MDG_CUDA(cudaGetLastError());
    ++g_last_launches;
    mdg::vec_add_kernel<<<(Dl + 127) / 128, 128, 0, stream>>>(w->x_attn_out_proj_bias, pl.q_res, pl.xo_qres, Dl);
    MDG_CUDA(cudaGetLastError());
    ++g_last_launches;
  }
  return MDG_OK;
}

int mdg_fusion_prepare(const MdgFusionWeights* w, const MdgFusionCfg* cfg, int precision, void* prepared,
                       size_t prepared_bytes, void* stream_v) {
  g_last_launches = 0;
  if (!w || !cfg || !prepared) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_fusion_prepare: NULL pointer");
  if (precision != MDG_PREC_BF16 && precision != MDG_PREC_FP32)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_fusion_prepare: precision=%d", precision);
  FusionPlan pl;
  int rc = plan_fusion(cfg, 1, precision, nullptr, prepared, &pl);
  if (rc) return rc;
  if (reinterpret_cast<uintptr_t>(prepared) % 256 != 0 || prepared_bytes < pl.prepared_total)
    return fail(MDG_ERR_WORKSPACE, "mdg_fusion_prepare: buffer (%zu B) too small or misaligned, need %zu", prepared_bytes, pl.prepared_total);
  if ((rc = check_fusion_weights(w, cfg))) return rc;
  if ((rc = mdg_check_device(-1))) return rc;
  return fusion_prepare_impl(w, cfg, pl, static_cast<cudaStream_t>(stream_v));
}

int mdg_fusion_encode(const MdgFusionWeights* w, const MdgFusionCfg* cfg, const void* prepared, const float* tokens,
                      const uint8_t* key_mask, const uint8_t* src_mask, const uint8_t* pool_key_mask, float* z_out,
                      int64_t B, int precision, void* workspace, size_t workspace_bytes, void* stream_v) {
  g_last_launches = 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!w || !cfg) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_fusion_encode: NULL pointer");
  if (B < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_fusion_encode: B=%lld", (long long)B);
  if (precision != MDG_PREC_BF16 && precision != MDG_PREC_FP32)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_fusion_encode: precision=%d", precision);
  // without a prepared-weights buffer the operand copies live at the head of the workspace (converted every call)
  FusionPlan pl;
  int rc = plan_fusion(cfg, B, precision, nullptr, nullptr, &pl);
  if (rc) return rc;
  if (B == 0) return MDG_OK;
  if (!tokens || !key_mask || !z_out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_fusion_encode: NULL pointer");
  const size_t head = prepared ? 0 : pl.prepared_total;
  if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256 != 0)
    return fail(MDG_ERR_WORKSPACE, "mdg_fusion_encode: workspace must be non-NULL and 256-byte aligned");
  if (workspace_bytes < pl.total + head)
    return fail(MDG_ERR_WORKSPACE, "mdg_fusion_encode: workspace %zu < required %zu", workspace_bytes, pl.total + head);
  if (prepared && reinterpret_cast<uintptr_t>(prepared) % 256 != 0)
    return fail(MDG_ERR_WORKSPACE, "mdg_fusion_encode: prepared buffer must be 256-byte aligned");
  rc = plan_fusion(cfg, B, precision, static_cast<uint8_t*>(workspace) + head,
                   prepared ? const_cast<void*>(prepared) : workspace, &pl);
  if (rc) return rc;
  if ((rc = check_fusion_weights(w, cfg))) return rc;
  rc = mdg_check_device(-1);
  if (rc) return rc;

  const int E = pl.E, Dl = pl.Dl, F = pl.F, T = pl.T, s = pl.split;
  const int act = cfg->actn == MDG_ACTN_GELU ? 2 : 1;
  if (!prepared && (rc = fusion_prepare_impl(w, cfg, pl, stream))) return rc;
  // MDG_FUSION_GENERIC (debug/test knob) forces the multi-kernel path
  if (fused_encoder_eligible(cfg, pl) && fused_encoder_aligned(w, cfg, tokens, z_out) &&
      getenv("MDG_FUSION_GENERIC") == nullptr)
    return run_fused_encoder(w, cfg, pl, tokens, key_mask, src_mask, pool_key_mask, z_out, B, stream);
  // zero the K-padding columns of operand buffers that kernels fill only up to their logical width
  if (kpad_of(Dl) != Dl) {
    MDG_CUDA(cudaMemsetAsync(pl.ob, 0, pl.ob_bytes, stream));
    MDG_CUDA(cudaMemsetAsync(pl.pb, 0, pl.pb_bytes, stream));
  }
  if (kpad_of(F) != F) MDG_CUDA(cudaMemsetAsync(pl.fb, 0, pl.fb_bytes, stream));

  for (long long b0 = 0; b0 < B; b0 += pl.chunk_drugs) {
    const long long Bc = (B - b0 < pl.chunk_drugs) ? (B - b0) : pl.chunk_drugs;
    const long long R = Bc * T;
    const float* tok = tokens + b0 * T * E;
    const uint8_t* km = key_mask + b0 * T;
    // embed2latent (models.py:411)
    if ((rc = convert_rows(tok, R, E, E, 1, s, pl.xb, stream))) return rc;
    if ((rc = run_linear(pl.xb, R, pl.w_e2l, Dl, E, s, w->embed2latent_bias, 0, nullptr, 0, pl.h, Dl, nullptr, 0, stream))) return rc;
    // transformer encoder layers (models.py:412; nn.TransformerEncoderLayer._sa_block/_ff_block, no final norm)
    for (int i = 0; i < pl.layers; ++i) {
      const MdgFusionLayer& L = w->layers[i];
      if (cfg->norm_first) {
        if ((rc = ln_convert(pl.h, R, Dl, nullptr, L.norm1_weight, L.norm1_bias, 1, nullptr, pl.nb, s, stream))) return rc;
      } else {
        if ((rc = ln_convert(pl.h, R, Dl, nullptr, nullptr, nullptr, 0, nullptr, pl.nb, s, stream))) return rc;
      }
      if (s) {  // fp32-parity mode: fp32 q|k|v rows
        if ((rc = run_linear(pl.nb, R, pl.w_in[i], 3 * Dl, Dl, s, L.in_proj_bias, 0, nullptr, 0, pl.qkv, 3 * Dl, nullptr, 0, stream))) return rc;
        if ((rc = launch_attention_t<float>(pl, pl.qkv, 3LL * Dl, km, src_mask, Bc, stream))) return rc;
      } else {  // bf16 mode: the in_proj epilogue writes bf16 rows (pitch kpad(3 Dl)) into the same buffer
        __nv_bfloat16* qkv16 = reinterpret_cast<__nv_bfloat16*>(pl.qkv);
        if ((rc = run_linear(pl.nb, R, pl.w_in[i], 3 * Dl, Dl, 0, L.in_proj_bias, 0, nullptr, 0, nullptr, 0, qkv16, 3 * Dl, stream))) return rc;
        if ((rc = launch_attention_t<__nv_bfloat16>(pl, qkv16, kpad_of(3 * Dl), km, src_mask, Bc, stream))) return rc;
      }
      if ((rc = run_linear(pl.ob, R, pl.w_out[i], Dl, Dl, s, L.out_proj_bias, 0, pl.h, Dl, pl.h, Dl, nullptr, 0, stream))) return rc;
      if (cfg->norm_first) {
        if ((rc = ln_convert(pl.h, R, Dl, nullptr, L.norm2_weight, L.norm2_bias, 1, nullptr, pl.nb, s, stream))) return rc;
      } else {
        if ((rc = ln_convert(pl.h, R, Dl, nullptr, L.norm1_weight, L.norm1_bias, 1, pl.h, pl.nb, s, stream))) return rc;
      }
      if ((rc = run_linear(pl.nb, R, pl.w_l1[i], F, Dl, s, L.linear1_bias, act, nullptr, 0, nullptr, 0, pl.fb, F, stream))) return rc;
      if ((rc = run_linear(pl.fb, R, pl.w_l2[i], Dl, F, s, L.linear2_bias, 0, pl.h, Dl, pl.h, Dl, nullptr, 0, stream))) return rc;
      if (!cfg->norm_first) {
        if ((rc = ln_convert(pl.h, R, Dl, nullptr, L.norm2_weight, L.norm2_bias, 1, pl.h, nullptr, s, stream))) return rc;
      }
    }
    float* z = z_out + b0 * E;
    if (cfg->agg == MDG_AGG_CLS) {
      // latent2embed of token 0 only (models.py:415-421)
      if ((rc = convert_rows(pl.h, Bc, Dl, Dl, T, s, pl.pb, stream))) return rc;
      if ((rc = run_linear(pl.pb, Bc, pl.w_l2e, E, Dl, s, w->latent2embed_bias, 0, nullptr, 0, z, E, nullptr, 0, stream))) return rc;
    } else if (cfg->agg == MDG_AGG_XATTN) {
      // models.py:422-443
      if ((rc = ln_convert(pl.h, R, Dl, nullptr, w->x_attn_kv_norm_weight, w->x_attn_kv_norm_bias, 1, nullptr, pl.nb, s, stream))) return rc;
      if ((rc = run_linear(pl.nb, R, pl.w_xin, 2 * Dl, Dl, s, w->x_attn_in_proj_bias + Dl, 0, nullptr, 0, pl.qkv, 2 * Dl, nullptr, 0, stream))) return rc;
      {
        long long blocks = (Bc * pl.H + 7) / 8;
        const long long cap = static_cast<long long>(num_sms()) * 16;
        if (blocks > cap) blocks = cap;
        mdg::pool_attention_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
            pl.qkv, pl.q_proj, pool_key_mask, Bc, T, pl.H, pl.hd, pl.pb, kpad_of(Dl), s);
        MDG_CUDA(cudaGetLastError());
        ++g_last_launches;
      }
      if ((rc = run_linear(pl.pb, Bc, pl.w_xout, Dl, Dl, s, w->x_attn_out_proj_bias, 0, nullptr, 0, pl.p32, Dl, nullptr, 0, stream))) return rc;
      if ((rc = ln_convert(pl.p32, Bc, Dl, pl.q_res, w->x_attn_query_norm_weight, w->x_attn_query_norm_bias,
                           cfg->norm_first ? 0 : 1, nullptr, pl.pb, s, stream))) return rc;
      if ((rc = run_linear(pl.pb, Bc, pl.w_l2e, E, Dl, s, w->latent2embed_bias, 0, nullptr, 0, z, E, nullptr, 0, stream))) return rc;
    } else {
      // mean / max over the unmasked tokens of latent2embed(h) (models.py:415, 444-451)
      if ((rc = ln_convert(pl.h, R, Dl, nullptr, nullptr, nullptr, 0, nullptr, pl.nb, s, stream))) return rc;
      if ((rc = run_linear(pl.nb, R, pl.w_l2e, E, Dl, s, w->latent2embed_bias, 0, nullptr, 0, pl.qkv, E, nullptr, 0, stream))) return rc;
      const long long n = Bc * E;
      mdg::masked_pool_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
          pl.qkv, km, Bc, T, E, cfg->agg == MDG_AGG_MAX, z);
      MDG_CUDA(cudaGetLastError());
      ++g_last_launches;
    }
  }
  return MDG_OK;
}

int mdg_fusion_trace_read(uint64_t* clocks_out_host, int max_records) {
  if (!clocks_out_host || max_records < 0) return -1;
  static unsigned long long buf[2 * mdg::kFeTraceLen];
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(buf, mdg::g_fe_trace, sizeof(buf)) != cudaSuccess) return -1;
  const int n = max_records < 2 * mdg::kFeTraceLen ? max_records : 2 * mdg::kFeTraceLen;
  for (int i = 0; i < n; ++i) clocks_out_host[i] = buf[i];
  return n;
}

// ------------------------------------------------------------------------------------------------ unimodal MLP
static int plan_mlp(const MdgMlp* m, long long B, int precision, void* ws, __nv_bfloat16** wts, __nv_bfloat16** act_b,
                    float** act_f, size_t* total) {
  if (!m || m->n_linear < 1 || m->n_linear > MDG_MAX_MLP_LINEAR) return fail(MDG_ERR_INVALID_ARGUMENT, "mlp: bad n_linear");
  const int s = precision == MDG_PREC_FP32;
  WsPlanner w;
  w.base = static_cast<uint8_t*>(ws);
  int maxd = 0;
  for (int i = 0; i <= m->n_linear; ++i) {
    if (m->dims[i] <= 0 || m->dims[i] > 8192) return fail(MDG_ERR_UNSUPPORTED, "mlp: dims[%d]=%d", i, m->dims[i]);
    if (m->dims[i] > maxd) maxd = m->dims[i];
  }
  for (int i = 0; i < m->n_linear; ++i)
    wts[i] = w.take<__nv_bfloat16>(static_cast<size_t>(m->dims[i + 1]) * ka_of(m->dims[i], s));
  *act_b = w.take<__nv_bfloat16>(static_cast<size_t>(B > 0 ? B : 1) * ka_of(maxd, s));
  *act_f = w.take<float>(static_cast<size_t>(B > 0 ? B : 1) * maxd);
  *total = w.off;
  return MDG_OK;
}

size_t mdg_mlp_workspace_bytes(const MdgMlp* mlp, int64_t B, int precision) {
  __nv_bfloat16* wts[MDG_MAX_MLP_LINEAR];
  __nv_bfloat16* ab;
  float* af;
  size_t total = 0;
  if (plan_mlp(mlp, B, precision, nullptr, wts, &ab, &af, &total) != MDG_OK) return 0;
  return total;
}

int mdg_mlp_forward(const MdgMlp* mlp, const float* x, float* y, int64_t B, int precision, void* workspace,
                    size_t workspace_bytes, void* stream_v) {
  g_last_launches = 0;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!mlp || !x || !y) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_mlp_forward: NULL pointer");
  if (B < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_mlp_forward: B < 0");
  if (mlp->actn != MDG_ACTN_RELU && mlp->actn != MDG_ACTN_GELU) return fail(MDG_ERR_UNSUPPORTED, "mlp: activation %d", mlp->actn);
  __nv_bfloat16* wts[MDG_MAX_MLP_LINEAR];
  __nv_bfloat16* ab;
  float* af;
  size_t total = 0;
  int rc = plan_mlp(mlp, B, precision, workspace, wts, &ab, &af, &total);
  if (rc) return rc;
  if (B == 0) return MDG_OK;
  if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256 != 0 || workspace_bytes < total)
    return fail(MDG_ERR_WORKSPACE, "mdg_mlp_forward: workspace (%zu B) too small or misaligned, need %zu", workspace_bytes, total);
  if ((rc = mdg_check_device(-1))) return rc;
  const int s = precision == MDG_PREC_FP32;
  const int act = mlp->actn == MDG_ACTN_GELU ? 2 : 1;
  const int n = mlp->n_linear;
  for (int i = 0; i < n; ++i) {
    if (!mlp->weight[i] || !mlp->bias[i]) return fail(MDG_ERR_INVALID_ARGUMENT, "mlp: NULL weight %d", i);
    if ((rc = convert_rows(mlp->weight[i], mlp->dims[i + 1], mlp->dims[i], mlp->dims[i], 1, s, wts[i], stream))) return rc;
  }
  // layer i input: x (i == 0) or the fp32 activation of layer i-1 (optionally LayerNorm'ed first, models.py:499-514)
  for (int i = 0; i < n; ++i) {
    const int K = mlp->dims[i], N = mlp->dims[i + 1];
    const float* in = (i == 0) ? x : af;
    const bool ln = i > 0 && mlp->ln_weight[i] != nullptr;
    if ((rc = ln_convert(in, B, K, nullptr, mlp->ln_weight[i], mlp->ln_bias[i], ln ? 1 : 0, nullptr, ab, s, stream))) return rc;
    const bool last = (i == n - 1);
    if ((rc = run_linear(ab, B, wts[i], N, K, s, mlp->bias[i], last ? 0 : act, nullptr, 0, last ? y : af, N, nullptr, 0, stream))) return rc;
  }
  return MDG_OK;
}

// ------------------------------------------------------------------------------------ chemCPA transcriptomic token
int mdg_tx_latent_combine(const float* basal, const float* drug_latent, const float* dosage, const int64_t* drug_idx,
                          const float* doser_beta, const float* doser_bias, int32_t doser, const float* cov_table,
                          const int64_t* cov_idx, int64_t B, int32_t dim, float* out, void* stream_v) {
  if (!basal || !out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_tx_latent_combine: NULL pointer");
  if (B < 0 || dim <= 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_tx_latent_combine: bad sizes");
  if (doser < 0 || doser > 3) return fail(MDG_ERR_UNSUPPORTED, "mdg_tx_latent_combine: doser %d", doser);
  if (drug_latent && !dosage) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_tx_latent_combine: NULL dose scale");
  if (drug_latent && (doser == 1 || doser == 2) && (!drug_idx || !doser_beta || !doser_bias))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_tx_latent_combine: sigmoid doser needs drug_idx, beta, bias");
  if ((cov_table == nullptr) != (cov_idx == nullptr))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_tx_latent_combine: cov_table and cov_idx go together");
  if (B == 0) return MDG_OK;
  const long long n = static_cast<long long>(B) * dim;
  mdg::tx_latent_combine_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      basal, drug_latent, dosage, reinterpret_cast<const long long*>(drug_idx), doser_beta, doser_bias, doser, cov_table,
      reinterpret_cast<const long long*>(cov_idx), B, dim, out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

int mdg_doser_mlp(const float* dosage, const int64_t* drug_idx, int64_t B, int32_t num_drugs, int32_t width,
                  int32_t depth, const float* w_in, const float* b_in, const float* w_hid, const float* b_hid,
                  const float* w_out, const float* b_out, float* scale_out, void* stream_v) {
  if (B < 0 || num_drugs <= 0 || width <= 0 || depth < 1)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_doser_mlp: bad sizes");
  if (width > mdg::kDoserMaxWidth) return fail(MDG_ERR_UNSUPPORTED, "mdg_doser_mlp: width %d > %d", width, mdg::kDoserMaxWidth);
  if (B == 0) return MDG_OK;
  if (!dosage || !drug_idx || !w_in || !b_in || !w_out || !b_out || !scale_out || (depth > 1 && (!w_hid || !b_hid)))
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_doser_mlp: NULL pointer");
  mdg::doser_mlp_kernel<<<static_cast<unsigned>((B + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream_v)>>>(
      dosage, reinterpret_cast<const long long*>(drug_idx), B, num_drugs, width, depth, w_in, b_in, w_hid, b_hid, w_out,
      b_out, scale_out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ token assembly
int mdg_assemble_tokens(const float* embeds, const uint8_t* masks, int64_t B, int32_t M, int32_t E, int32_t n_non_tx,
                        int32_t num_bottlenecks, const float* bottleneck_tokens, const float* cls_token,
                        const float* pos_enc, int32_t pos_len, int32_t normalize, float* seq_out,
                        uint8_t* seq_mask_out, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!embeds || !masks || !seq_out || !seq_mask_out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_assemble_tokens: NULL pointer");
  if (B < 0 || M <= 0 || E <= 0 || n_non_tx < 0 || n_non_tx > M || num_bottlenecks < 0 || pos_len < 0)
    return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_assemble_tokens: bad sizes");
  if (num_bottlenecks > 0 && !bottleneck_tokens) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_assemble_tokens: NULL bottleneck tokens");
  const int has_cls = cls_token != nullptr;
  const int T = M + num_bottlenecks + has_cls;
  if (T > MDG_MAX_TOKENS) return fail(MDG_ERR_UNSUPPORTED, "mdg_assemble_tokens: T=%d > %d", T, MDG_MAX_TOKENS);
  if (pos_len > T) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_assemble_tokens: pos_len %d > T %d", pos_len, T);
  if (B == 0) return MDG_OK;
  const long long rows = static_cast<long long>(B) * T;
  mdg::assemble_tokens_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      embeds, masks, B, M, E, n_non_tx, num_bottlenecks, has_cls, bottleneck_tokens, cls_token, pos_enc, pos_len,
      normalize, seq_out, seq_mask_out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

int mdg_masked_pool(const float* tokens, const uint8_t* masks, int64_t B, int32_t T, int32_t E, int32_t mode,
                    float* z_out, void* stream_v) {
  if (!tokens || !masks || !z_out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_masked_pool: NULL pointer");
  if (B < 0 || T <= 0 || E <= 0 || mode < 0 || mode > 2) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_masked_pool: bad arguments");
  if (B == 0) return MDG_OK;
  const long long n = static_cast<long long>(B) * E;
  mdg::masked_pool_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_v)>>>(
      tokens, masks, B, T, E, mode, z_out);
  MDG_CUDA(cudaGetLastError());
  return MDG_OK;
}

// ------------------------------------------------------------------------------------------------ exact rank
size_t mdg_exact_rank_workspace_bytes(int64_t N) {
  if (N < 2 || N > 92000) return 0;
  return mdg::plan_exact_rank(nullptr, N).total;
}

static int exact_rank_sort(const float* scores_l, long long N, const mdg::ExactRankWs& w, cudaStream_t stream) {
  const unsigned long long M = static_cast<unsigned long long>(N) * (N - 1) / 2;
  long long blocks = static_cast<long long>((M + 255) / 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  mdg::tri_gather_keys_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(scores_l, static_cast<int>(N), M,
                                                                                  w.keys_in, w.idx_in);
  MDG_CUDA(cudaGetLastError());
  size_t tb = w.cub_bytes;
  MDG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.keys_in, w.keys_out, w.idx_in, w.idx_out,
                                           static_cast<long long>(M), 0, 32, stream));
  return MDG_OK;
}

int mdg_exact_rank(const float* scores, int64_t L, int64_t N, float* out, void* workspace, size_t workspace_bytes,
                   void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!scores || !out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_exact_rank: NULL pointer");
  if (L < 0 || N < 0) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_exact_rank: negative size");
  if (L == 0 || N == 0) return MDG_OK;
  if (N == 1) {
    MDG_CUDA(cudaMemsetAsync(out, 0, static_cast<size_t>(L) * sizeof(float), stream));
    return MDG_OK;
  }
  if (N > 92000) return fail(MDG_ERR_UNSUPPORTED, "mdg_exact_rank: N=%lld (pair index exceeds 32 bits)", (long long)N);
  mdg::ExactRankWs w = mdg::plan_exact_rank(workspace, N);
  if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256 != 0 || workspace_bytes < w.total)
    return fail(MDG_ERR_WORKSPACE, "mdg_exact_rank: workspace (%zu B) too small or misaligned, need %zu", workspace_bytes, w.total);
  const unsigned long long M = static_cast<unsigned long long>(N) * (N - 1) / 2;
  long long blocks = static_cast<long long>((M + 255) / 256);
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  // small inputs: direct scatter; otherwise order the pairs by the high bits of their index first (exact_rank.cuh)
  const bool direct = M < (1ull << 22) || getenv("MDG_EXACT_RANK_DIRECT") != nullptr;
  int bits = 1;
  while ((1ull << bits) < M) ++bits;
  const int tiles = static_cast<int>((N + 31) / 32);
  for (int64_t l = 0; l < L; ++l) {
    int rc = exact_rank_sort(scores + static_cast<size_t>(l) * N * N, N, w, stream);
    if (rc) return rc;
    float* out_l = out + static_cast<size_t>(l) * N * N;
    if (direct) {
      mdg::tri_scatter_rank_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(w.idx_out, static_cast<int>(N), M, out_l);
      MDG_CUDA(cudaGetLastError());
      continue;
    }
    // keys = sorted pair indices (position = rank), values = positions (idx_in still holds 0..M-1: SortPairs keeps
    // its inputs); outputs reuse the two key buffers, which the rank path no longer needs
    size_t tb = w.cub_bytes;
    MDG_CUDA(cub::DeviceRadixSort::SortPairs(w.cub_temp, tb, w.idx_out, w.keys_in, w.idx_in, w.keys_out,
                                             static_cast<long long>(M), mdg::kRankSliceBits, bits, stream));
    mdg::tri_place_rank_kernel<<<static_cast<unsigned>((M + mdg::kRankSlice - 1) / mdg::kRankSlice), 256, 0, stream>>>(
        w.keys_in, w.keys_out, static_cast<int>(N), M, out_l);
    MDG_CUDA(cudaGetLastError());
    mdg::tri_mirror_kernel<<<static_cast<unsigned>(static_cast<long long>(tiles) * (tiles + 1) / 2), dim3(32, 8), 0, stream>>>(
        out_l, static_cast<int>(N));
    MDG_CUDA(cudaGetLastError());
  }
  return MDG_OK;
}

int mdg_lower_triangle_quantiles(const float* scores, int64_t L, int64_t N, int32_t Q, float* quantiles_out,
                                 void* workspace, size_t workspace_bytes, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (!scores || !quantiles_out) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_lower_triangle_quantiles: NULL pointer");
  if (L < 0 || N < 2 || Q < 1) return fail(MDG_ERR_INVALID_ARGUMENT, "mdg_lower_triangle_quantiles: bad sizes");
  if (N > 92000) return fail(MDG_ERR_UNSUPPORTED, "mdg_lower_triangle_quantiles: N too large");
  const unsigned long long M = static_cast<unsigned long long>(N) * (N - 1) / 2;
  if (static_cast<unsigned long long>(Q) > M) return fail(MDG_ERR_INVALID_ARGUMENT, "Q=%d exceeds the %llu pairs", Q, M);
  if (L == 0) return MDG_OK;
  mdg::ExactRankWs w = mdg::plan_exact_rank(workspace, N);
  if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 256 != 0 || workspace_bytes < w.total)
    return fail(MDG_ERR_WORKSPACE, "mdg_lower_triangle_quantiles: workspace too small or misaligned, need %zu", w.total);
  for (int64_t l = 0; l < L; ++l) {
    int rc = exact_rank_sort(scores + static_cast<size_t>(l) * N * N, N, w, stream);
    if (rc) return rc;
    mdg::pick_quantiles_kernel<<<(Q + 255) / 256, 256, 0, stream>>>(w.keys_out, M, Q,
                                                                     quantiles_out + static_cast<size_t>(l) * Q);
    MDG_CUDA(cudaGetLastError());
  }
  return MDG_OK;
}

}  // extern "C"
