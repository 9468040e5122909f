// Temporary: entry points declared in the header but not implemented yet fail loudly.
extern "C" {
size_t mdg_exact_rank_workspace_bytes(int64_t) { return 0; }
int mdg_exact_rank(const float*, int64_t, int64_t, float*, void*, size_t, void*) {
  return fail(MDG_ERR_UNSUPPORTED, "mdg_exact_rank: not implemented yet");
}
size_t mdg_fusion_workspace_bytes(const MdgFusionCfg*, int64_t) { return 0; }
int mdg_fusion_encode(const MdgFusionWeights*, const MdgFusionCfg*, const float*, const uint8_t*, const uint8_t*,
                      const uint8_t*, float*, int64_t, void*, size_t, void*) {
  return fail(MDG_ERR_UNSUPPORTED, "mdg_fusion_encode: not implemented yet");
}
int mdg_mlp_forward(const MdgMlp*, const float*, float*, int64_t, void*) {
  return fail(MDG_ERR_UNSUPPORTED, "mdg_mlp_forward: not implemented yet");
}
}
