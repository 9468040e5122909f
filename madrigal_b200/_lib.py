"""ctypes binding of libmadrigal_b200.so — the C ABI declared in include/madrigal_b200.h.

There is deliberately no fallback: if the library is missing or the device is not sm_100, every call raises.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint8, c_uint16, c_uint32, c_void_p

from .build import LIB_PATH

# A/B measurements of compile-time variants: MDG_LIB_PATH points at another build of the SAME sources/ABI.
LIB_PATH = os.environ.get("MDG_LIB_PATH", LIB_PATH)

EXPECTED_ABI = 5  # MDG_ABI_VERSION of include/madrigal_b200.h these ctypes structs/signatures were written for
MDG_MAX_LAYERS = 8
MDG_MAX_TOKENS = 32
MDG_MAX_MLP_LINEAR = 8
MDG_RANK_LUT_ENTRIES = 1 << 13
MDG_RANK_MAX_Q = 65535

MDG_PREC_BF16, MDG_PREC_FP32 = 0, 1
MDG_OUT_LOGIT_F32, MDG_OUT_SIGMOID_F32, MDG_OUT_RANK_U16 = 0, 1, 2
MDG_PAIRS_FULL, MDG_PAIRS_SYMMETRIC, MDG_PAIRS_PACKED_TILES = 0, 1, 2
MDG_RANK_KIND = {"lut": 0, "pwl": 1}
MDG_ENS_MEAN_F32, MDG_ENS_GMEAN_F32, MDG_ENS_GMEAN_RANK_U16 = 0, 1, 2
MDG_MAX_ENSEMBLE = 16
MDG_AGG = {"cls": 0, "x-attn": 1, "mean": 2, "max": 3}
MDG_ACTN = {"relu": 0, "gelu": 1}


class MdgRankTable(Structure):
    _fields_ = [("thresholds", c_void_p), ("lut", c_void_p), ("affine", c_void_p), ("L", c_int32), ("Q", c_int32),
                ("kind", c_int32)]


class MdgFusionLayer(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "in_proj_weight", "in_proj_bias", "out_proj_weight", "out_proj_bias", "linear1_weight", "linear1_bias",
        "linear2_weight", "linear2_bias", "norm1_weight", "norm1_bias", "norm2_weight", "norm2_bias")]


class MdgFusionWeights(Structure):
    _fields_ = ([(n, c_void_p) for n in ("embed2latent_weight", "embed2latent_bias", "latent2embed_weight",
                                         "latent2embed_bias")]
                + [("layers", MdgFusionLayer * MDG_MAX_LAYERS)]
                + [(n, c_void_p) for n in (
                    "x_attn_query", "x_attn_kv_norm_weight", "x_attn_kv_norm_bias", "x_attn_query_norm_weight",
                    "x_attn_query_norm_bias", "x_attn_in_proj_weight", "x_attn_in_proj_bias",
                    "x_attn_out_proj_weight", "x_attn_out_proj_bias")])


class MdgFusionCfg(Structure):
    _fields_ = [(n, c_int32) for n in ("embed_dim", "num_layers", "num_heads", "head_dim", "ffn_dim", "actn",
                                       "norm_first", "agg", "num_tokens")]


class MdgMlp(Structure):
    _fields_ = [("n_linear", c_int32), ("dims", c_int32 * (MDG_MAX_MLP_LINEAR + 1)), ("actn", c_int32),
                ("weight", c_void_p * MDG_MAX_MLP_LINEAR), ("bias", c_void_p * MDG_MAX_MLP_LINEAR),
                ("ln_weight", c_void_p * MDG_MAX_MLP_LINEAR), ("ln_bias", c_void_p * MDG_MAX_MLP_LINEAR)]


# every symbol include/madrigal_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "mdg_last_error": (c_char_p, []),
    "mdg_abi_version": (c_int, []),
    "mdg_check_device": (c_int, [c_int]),
    "mdg_rank_table_build": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mdg_rank_table_build_pwl": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mdg_rank_lookup": (c_int, [c_void_p, c_int64, POINTER(MdgRankTable), c_void_p, c_void_p]),
    "mdg_packed_tiles_per_outcome": (c_int64, [c_int64]),
    "mdg_host_mirror_tiles": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int32]),
    "mdg_pair_score_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int]),
    "mdg_pair_score": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                               c_int, POINTER(MdgRankTable), c_void_p, c_void_p, c_size_t, c_void_p]),
    "mdg_pair_prepared_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "mdg_pair_prepare": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "mdg_pair_score_prepared": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_int,
                                        c_int, c_int, c_int, POINTER(MdgRankTable), c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "mdg_peer_allgather": (c_int, [c_void_p, c_int64, c_int64, c_int32, POINTER(c_void_p), POINTER(c_void_p), c_int32,
                                   c_int32, c_uint32, c_void_p]),
    "mdg_l2_normalize_rows": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "mdg_pair_topk_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int, c_int32]),
    "mdg_pair_topk": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int,
                              c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                              c_void_p]),
    "mdg_pair_score_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_size_t,
                                      c_void_p]),
    "mdg_ensemble_reduce": (c_int, [POINTER(c_void_p), c_int32, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "mdg_ensemble_rank_u16": (c_int, [POINTER(c_void_p), c_int32, c_int64, c_int64, c_void_p, c_int32,
                                      POINTER(MdgRankTable), c_void_p, c_void_p, c_void_p]),
    "mdg_last_launch_count": (c_int, []),
    "mdg_profile_enable": (c_int, [c_int]),
    "mdg_profile_read": (c_int, [POINTER(c_float), c_int]),
    "mdg_exact_rank_workspace_bytes": (c_size_t, [c_int64]),
    "mdg_exact_rank": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mdg_lower_triangle_quantiles": (c_int, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_size_t,
                                             c_void_p]),
    "mdg_fusion_workspace_bytes": (c_size_t, [POINTER(MdgFusionCfg), c_int64, c_int]),
    "mdg_fusion_prepared_bytes": (c_size_t, [POINTER(MdgFusionCfg), c_int]),
    "mdg_fusion_prepare": (c_int, [POINTER(MdgFusionWeights), POINTER(MdgFusionCfg), c_int, c_void_p, c_size_t,
                                   c_void_p]),
    "mdg_fusion_encode": (c_int, [POINTER(MdgFusionWeights), POINTER(MdgFusionCfg), c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "mdg_fusion_trace_read": (c_int, [c_void_p, c_int]),
    "mdg_assemble_tokens": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                    c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p]),
    "mdg_masked_pool": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "mdg_mlp_workspace_bytes": (c_size_t, [POINTER(MdgMlp), c_int64, c_int]),
    "mdg_mlp_forward": (c_int, [POINTER(MdgMlp), c_void_p, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p]),
    "mdg_tx_latent_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                      c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "mdg_doser_mlp": (c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_LIB = None


def lib():
    """Load the shared library (once).  Raises if it has not been built — there is no other implementation."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m madrigal_b200.build` (nvcc, sm_100a). "
                "madrigal_b200 has no CPU or PyTorch fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        handle.mdg_abi_version.restype = c_int
        abi = handle.mdg_abi_version()
        if abi != EXPECTED_ABI:  # structs are passed by pointer: a stale .so would corrupt memory silently
            raise RuntimeError(f"{LIB_PATH} has ABI version {abi}, this binding expects {EXPECTED_ABI}: rebuild it "
                               f"with `python -m madrigal_b200.build --force`")
        from .build import library_is_current
        if not library_is_current():
            import warnings
            warnings.warn(f"{LIB_PATH} was built from different sources than the ones in this tree "
                          f"(`python -m madrigal_b200.build` rebuilds it)", RuntimeWarning)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = handle
    return _LIB


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mdg_last_error().decode(errors="replace")
        raise RuntimeError(f"madrigal_b200 {what} failed (status {rc}): {msg}")
