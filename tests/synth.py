"""Seeded synthetic inputs shared by the golden generator, the CPU tests, the GPU parity tests and bench.py.

Everything comes from numpy's PCG64 `default_rng(seed)` so the same arrays can be regenerated anywhere without the
reference (random-init weights with torch's default init *distributions*, N(0,1) tokens,
Bernoulli(0.5) missing-modality masks with token 0 always present).
"""
import numpy as np

F32 = np.float32


def decoder_inputs(N: int, D: int, L: int, seed: int = 0, symmetric: bool = True, unit_scale: bool = True):
    """z [N,D] ~ N(0,1)/sqrt(D) (SURVEY §8d decoder-only timing), P [L,D,D] ~ U(+-1/sqrt(D)) (nn.Bilinear init)."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((N, D), dtype=F32)
    if unit_scale:
        z /= F32(np.sqrt(D))
    bound = F32(1.0 / np.sqrt(D))
    P = rng.uniform(-bound, bound, size=(L, D, D)).astype(F32)
    if symmetric:
        P = np.triu(P) + np.swapaxes(np.triu(P, 1), -1, -2)
    return z, P


# ------------------------------------------------------------------------------------------------ fusion encoder
CELL_LINES = ['a375', 'a549', 'asc', 'ha1e', 'hcc515', 'hec108', 'hela', 'hepg2', 'ht29', 'huvec', 'mcf7', 'npc',
              'pc3', 'thp1', 'vcap', 'yapc']  # madrigal/utils.py:28 (ORDERED)
NUM_NON_TX = 3  # str, kg, cv (madrigal/utils.py:30-36)


def _uniform(rng, shape, fan_in):
    b = 1.0 / np.sqrt(fan_in)
    return rng.uniform(-b, b, size=shape).astype(F32)


def fusion_state_dict(cfg: dict, seed: int = 0):
    """Random-init TransformerFusion state_dict (reference key names, models.py:365-381) from a numpy seed.

    cfg: embed_dim, num_layers, num_heads, head_dim, ffn_dim, agg.  Linear weights/biases ~ U(+-1/sqrt(fan_in))
    (nn.Linear's default distribution); LayerNorm affine params perturbed away from (1, 0) so they are exercised.
    """
    rng = np.random.default_rng(seed)
    E, Dl, Fd = cfg["embed_dim"], cfg["num_heads"] * cfg["head_dim"], cfg["ffn_dim"]
    sd = {"embed2latent.weight": _uniform(rng, (Dl, E), E), "embed2latent.bias": _uniform(rng, (Dl,), E)}

    def ln(prefix):
        sd[prefix + ".weight"] = (1.0 + 0.1 * rng.standard_normal(Dl)).astype(F32)
        sd[prefix + ".bias"] = (0.1 * rng.standard_normal(Dl)).astype(F32)

    def mha(prefix):
        sd[prefix + ".in_proj_weight"] = _uniform(rng, (3 * Dl, Dl), Dl)
        sd[prefix + ".in_proj_bias"] = _uniform(rng, (3 * Dl,), Dl)
        sd[prefix + ".out_proj.weight"] = _uniform(rng, (Dl, Dl), Dl)
        sd[prefix + ".out_proj.bias"] = _uniform(rng, (Dl,), Dl)

    for i in range(cfg["num_layers"]):
        p = f"transformer_encoder.layers.{i}"
        mha(p + ".self_attn")
        sd[p + ".linear1.weight"] = _uniform(rng, (Fd, Dl), Dl)
        sd[p + ".linear1.bias"] = _uniform(rng, (Fd,), Dl)
        sd[p + ".linear2.weight"] = _uniform(rng, (Dl, Fd), Fd)
        sd[p + ".linear2.bias"] = _uniform(rng, (Dl,), Fd)
        ln(p + ".norm1")
        ln(p + ".norm2")
    sd["latent2embed.weight"] = _uniform(rng, (E, Dl), Dl)
    sd["latent2embed.bias"] = _uniform(rng, (E,), Dl)
    if cfg["agg"] == "x-attn":
        ln("x_attn_kv_norm")
        ln("x_attn_query_norm")
        mha("x_attn_mha_layer")
        sd["x_attn_query"] = rng.standard_normal((1, Dl)).astype(F32)
    return sd


def fusion_inputs(B: int, T: int, E: int, seed: int = 0, always_visible=(0,), p_missing: float = 0.5):
    """tokens [B,T,E] ~ N(0,1); mask [B,T] True = missing, Bernoulli(p_missing), `always_visible` never masked;
    masked slots overwritten with a different finite draw (their content must not matter)."""
    rng = np.random.default_rng(seed + 1000)
    tokens = rng.standard_normal((B, T, E)).astype(F32)
    mask = rng.random((B, T)) < p_missing
    mask[:, list(always_visible)] = False
    garbage = (3.0 * rng.standard_normal((B, T, E))).astype(F32)
    tokens = np.where(mask[:, :, None], garbage, tokens)
    return tokens, mask


def mlp_adaptor_params(in_dim, hidden_dims, out_dim, seed=0):
    """MLPAdaptor (norm='ln', order='nd') parameters in nn.Sequential order as a list of op dicts (see
    oracle.mlp_adaptor): Linear, act, [LN, Linear, act]*, Linear  (models.py:471-481, 499-514)."""
    rng = np.random.default_rng(seed + 2000)
    ops = [{"op": "linear", "w": _uniform(rng, (hidden_dims[0], in_dim), in_dim),
            "b": _uniform(rng, (hidden_dims[0],), in_dim)}, {"op": "act"}]
    for i in range(len(hidden_dims) - 1):
        ops.append({"op": "ln", "w": (1.0 + 0.1 * rng.standard_normal(hidden_dims[i])).astype(F32),
                    "b": (0.1 * rng.standard_normal(hidden_dims[i])).astype(F32)})
        ops.append({"op": "linear", "w": _uniform(rng, (hidden_dims[i + 1], hidden_dims[i]), hidden_dims[i]),
                    "b": _uniform(rng, (hidden_dims[i + 1],), hidden_dims[i])})
        ops.append({"op": "act"})
    ops.append({"op": "linear", "w": _uniform(rng, (out_dim, hidden_dims[-1]), hidden_dims[-1]),
                "b": _uniform(rng, (out_dim,), hidden_dims[-1])})
    return ops


def params_checksum(arrays) -> float:
    """Order-dependent float64 checksum used by the golden fixtures to detect RNG-stream drift."""
    tot = 0.0
    for k, a in enumerate(arrays):
        a = np.asarray(a, dtype=np.float64).reshape(-1)
        tot += float((a * np.cos(np.arange(a.size) * 0.37 + k)).sum())
    return tot


# ------------------------------------------------------------------------------------- chemCPA tx encoder (f-4)
CHEMCPA_CASES = [
    dict(name="no_drugs_production_shape", num_genes=978, num_drugs=7, n_cell=18, use_drugs=False, doser_type="amortized",
         hparams=dict(dim=32, autoencoder_width=64, autoencoder_depth=2, dosers_width=8, dosers_depth=2,
                      embedding_encoder_width=16, embedding_encoder_depth=0), emb_dim=24, B=11, seed=3),
    dict(name="drugs_logsigm", num_genes=120, num_drugs=9, n_cell=5, use_drugs=True, doser_type="logsigm",
         hparams=dict(dim=48, autoencoder_width=96, autoencoder_depth=3, dosers_width=8, dosers_depth=2,
                      embedding_encoder_width=16, embedding_encoder_depth=0), emb_dim=24, B=13, seed=4),
    dict(name="drugs_amortized_enc1", num_genes=64, num_drugs=6, n_cell=4, use_drugs=True, doser_type="amortized",
         hparams=dict(dim=32, autoencoder_width=48, autoencoder_depth=1, dosers_width=16, dosers_depth=2,
                      embedding_encoder_width=40, embedding_encoder_depth=1), emb_dim=20, B=9, seed=5),
    dict(name="drugs_sigm_depth0", num_genes=50, num_drugs=5, n_cell=3, use_drugs=True, doser_type="sigm",
         hparams=dict(dim=16, autoencoder_width=48, autoencoder_depth=0, dosers_width=16, dosers_depth=2,
                      embedding_encoder_width=40, embedding_encoder_depth=0), emb_dim=12, B=8, seed=6),
    dict(name="drugs_mlp_dosers", num_genes=72, num_drugs=7, n_cell=4, use_drugs=True, doser_type="mlp",
         hparams=dict(dim=24, autoencoder_width=40, autoencoder_depth=1, dosers_width=24, dosers_depth=3,
                      embedding_encoder_width=16, embedding_encoder_depth=0), emb_dim=10, B=21, seed=8),
]


def _chemcpa_mlp_state(rng, prefix, sizes, batch_norm=True):
    """state_dict entries of chemCPA's MLP(sizes) (model.py:176-224) with non-trivial BatchNorm1d statistics."""
    sd, pos = {}, 0
    n = len(sizes) - 1
    for s in range(n):
        sd[f"{prefix}network.{pos}.weight"] = _uniform(rng, (sizes[s + 1], sizes[s]), sizes[s])
        sd[f"{prefix}network.{pos}.bias"] = _uniform(rng, (sizes[s + 1],), sizes[s])
        pos += 1
        if batch_norm and s < n - 1:
            w = sizes[s + 1]
            sd[f"{prefix}network.{pos}.weight"] = (1.0 + 0.2 * rng.standard_normal(w)).astype(F32)
            sd[f"{prefix}network.{pos}.bias"] = (0.1 * rng.standard_normal(w)).astype(F32)
            sd[f"{prefix}network.{pos}.running_mean"] = (0.2 * rng.standard_normal(w)).astype(F32)
            sd[f"{prefix}network.{pos}.running_var"] = (0.5 + rng.random(w)).astype(F32)
            sd[f"{prefix}network.{pos}.num_batches_tracked"] = np.asarray(7, dtype=np.int64)
            pos += 1
        pos += 1  # ReLU (dropped after the last Linear)
    return sd


def chemcpa_case(case):
    """(state_dict of numpy arrays in the reference's key names, drug embedding table, inputs) for one CHEMCPA_CASES
    entry; the same arrays are loaded into the reference module (golden generator), the oracle and the CUDA path."""
    rng = np.random.default_rng(case["seed"] + 7000)
    hp = case["hparams"]
    sd = _chemcpa_mlp_state(rng, "encoder.", [case["num_genes"]] + [hp["autoencoder_width"]] * hp["autoencoder_depth"] + [hp["dim"]])
    table = rng.standard_normal((case["num_drugs"], case["emb_dim"])).astype(F32)
    if case["use_drugs"]:
        sd.update(_chemcpa_mlp_state(rng, "drug_embedding_encoder.", [case["emb_dim"]] + [hp["embedding_encoder_width"]] * hp["embedding_encoder_depth"] + [hp["dim"]]))
        if case["doser_type"] == "amortized":
            sd.update(_chemcpa_mlp_state(rng, "dosers.", [case["emb_dim"] + 1] + [hp["dosers_width"]] * hp["dosers_depth"] + [1]))
        elif case["doser_type"] == "mlp":  # one MLP([1] + [w] * depth + [1], batch_norm=False) per drug (model.py:405-416)
            for d in range(case["num_drugs"]):
                sd.update(_chemcpa_mlp_state(rng, f"dosers.{d}.", [1] + [hp["dosers_width"]] * hp["dosers_depth"] + [1],
                                             batch_norm=False))
        else:
            sd["dosers.beta"] = (1.0 + 0.3 * rng.standard_normal((1, case["num_drugs"]))).astype(F32)
            sd["dosers.bias"] = (0.3 * rng.standard_normal((1, case["num_drugs"]))).astype(F32)
    sd["covariates_embeddings.0.weight"] = rng.standard_normal((case["n_cell"], hp["dim"])).astype(F32)
    B = case["B"]
    inputs = dict(genes=rng.standard_normal((B, case["num_genes"])).astype(F32),
                  drugs_idx=rng.integers(0, case["num_drugs"], B).astype(np.int64),
                  dosages=(rng.random(B) * 3.0).astype(F32),
                  cov_idx=rng.integers(0, case["n_cell"], B).astype(np.int64))
    return sd, table, inputs


# ------------------------------------------------------------------------- MLPEncoder (cv / tx 'mlp' encoders, f-4)
MLPENCODER_CASES = [
    dict(name="cv_default_no_norm", in_dim=96, hidden=[64, 48], out_dim=32, p=0.1, norm=None, actn="relu", order="nd", B=9, seed=11),
    dict(name="tx_ln_gelu", in_dim=978, hidden=[128, 64, 64], out_dim=128, p=0.0, norm="ln", actn="gelu", order="nd", B=7, seed=12),
    dict(name="ln_dropout_first", in_dim=40, hidden=[32, 24], out_dim=16, p=0.2, norm="ln", actn="relu", order="dn", B=5, seed=13),
    dict(name="bn_eval_running_stats", in_dim=56, hidden=[48, 40, 24], out_dim=32, p=0.1, norm="bn", actn="gelu", order="nd", B=6, seed=14),
]


def mlp_encoder_ops(case):
    """Op list (see oracle.mlp_adaptor) of MLPEncoder(in_dim, hidden, out_dim, p, norm, actn, order) in eval mode:
    Linear, act, [norm?, Linear, act]*, Linear (models.py:133-143, 146-176); Dropout is the identity."""
    ops = mlp_adaptor_params(case["in_dim"], case["hidden"], case["out_dim"], case["seed"])
    for o in ops:
        if o["op"] == "act":
            o["actn"] = case["actn"]
    if case["norm"] is None:
        ops = [o for o in ops if o["op"] != "ln"]
    if case["norm"] == "bn":  # eval-mode BatchNorm1d in the norm slots: same affine parameters + running statistics
        rng = np.random.default_rng(case["seed"] + 3000)
        for o in ops:
            if o["op"] == "ln":
                o["op"] = "bn"
                o["mean"] = (0.3 * rng.standard_normal(o["w"].shape[0])).astype(F32)
                o["var"] = (0.5 + rng.random(o["w"].shape[0])).astype(F32)
    return ops
