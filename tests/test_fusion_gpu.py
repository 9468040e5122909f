"""GPU parity: fusion encoder (mdg_fusion_encode), unimodal MLP (mdg_mlp_forward) and token assembly
(mdg_assemble_tokens) through the C ABI vs the CPU oracle and the committed reference goldens.

Tolerances: precision='fp32' (bf16x3 split GEMMs, fp32 everything else): |got - ref| <= 1e-3 * max(|ref|, rms(ref));
precision='bf16': 3e-2 of rms (bf16 operands through 2 transformer layers; the north-star states no encoder bar for
bf16, the decoder's 1e-2 applies to single GEMM chains).
"""
import json
import os

import numpy as np
import pytest
import torch

import synth
from oracle import oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "golden_meta.json")))


@pytest.fixture(scope="module")
def mb(cuda_device):
    import madrigal_b200
    return madrigal_b200


def gpu(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def assert_close(got, ref, tol, what=""):
    ref = np.asarray(ref, np.float64)
    err = np.abs(np.asarray(got, np.float64) - ref)
    rms = np.sqrt(np.mean(ref ** 2))
    ratio = (err / (tol * np.maximum(np.abs(ref), rms))).max()
    assert np.isfinite(got).all() and ratio <= 1.0, f"{what}: max err/bound = {ratio:.3f} (max|err| {err.max():.3e}, rms {rms:.3e})"


def make_module(mb, case, dev, precision="fp32"):
    mod = mb.TransformerFusion(case["embed_dim"], case["nb"], case["num_layers"], case["num_heads"], case["head_dim"],
                               case["ffn_dim"], transformer_actn=case["actn"], transformer_norm_first=case["norm_first"],
                               transformer_batch_first=False, transformer_agg=case["agg"], precision=precision)
    sd = synth.fusion_state_dict(case, case["seed"])
    mod.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)  # reference key names
    return mod.to(dev).eval(), sd


@pytest.mark.parametrize("case", META["fusion"], ids=lambda c: c["name"])
def test_fusion_vs_reference_golden(mb, cuda_device, case):
    g = np.load(os.path.join(HERE, "golden", "golden_fusion.npz"))
    mod, sd = make_module(mb, case, cuda_device)
    tokens, mask = synth.fusion_inputs(case["B"], case["T"], case["embed_dim"], case["seed"],
                                       always_visible=tuple(case["always_visible"]))
    if not np.isclose(synth.params_checksum([sd[k] for k in sorted(sd)] + [tokens]), case["checksum"], rtol=1e-9):
        pytest.skip("numpy Generator stream drift")
    name = case["name"]
    src = gpu(g[f"{name}.src_mask"], cuda_device) if f"{name}.src_mask" in g.files else None
    if f"{name}.pool_mask" in g.files:
        mod.x_attn_key_padding_mask = torch.from_numpy(g[f"{name}.pool_mask"])[None, :]
    with torch.no_grad():
        z = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), src).cpu().numpy()
    assert_close(z, g[f"{name}.z"], 1e-3, name)


@pytest.mark.parametrize("agg,norm_first,actn,T,nb,B,dims", [
    ("mean", True, "gelu", 4, 0, 300, (128, 8, 32, 512)),      # BASELINE config 5 shape: T=4, Dl=256, F=2*Dl
    ("cls", False, "relu", 5, 0, 257, (128, 8, 64, 256)),
    ("x-attn", True, "gelu", 23, 4, 40, (128, 8, 64, 256)),    # production DrugBank shape
    ("max", True, "relu", 21, 2, 33, (64, 2, 256, 512)),       # head_dim 256 (TWOSIDES)
    ("x-attn", False, "gelu", 19, 0, 64, (256, 8, 32, 1024)),
])
def test_fusion_vs_oracle(mb, cuda_device, agg, norm_first, actn, T, nb, B, dims):
    E, H, hd, F = dims
    case = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn=actn, norm_first=norm_first,
                agg=agg, nb=nb, seed=17 + T)
    mod, sd = make_module(mb, case, cuda_device)
    tokens, mask = synth.fusion_inputs(B, T, E, case["seed"], always_visible=(0,) + tuple(range(3, 3 + nb)))
    src = None
    if nb > 0:
        src = np.zeros((T, T), bool)
        src[:3, T - 16:] = True
        src[T - 16:, :3] = True
    pool = None
    if agg == "x-attn":
        pool = np.zeros(T, bool)
        if nb > 0:
            pool[:3] = True
            pool[-16:] = True
        mod.x_attn_key_padding_mask = torch.from_numpy(pool)[None, :]
    ref = oracle.fusion_forward(sd, case, tokens, mask, src, pool, dtype=np.float64)
    with torch.no_grad():
        z = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), None if src is None else gpu(src, cuda_device))
        assert_close(z.cpu().numpy(), ref, 1e-3, "fp32 mode")
        # masked-slot contents are don't-care (only for aggregations that never read masked tokens' outputs)
        if agg in ("mean", "max") or nb > 0:
            tok2 = np.where(mask[:, :, None], np.float32(9.0), tokens)
            z2 = mod(gpu(tok2, cuda_device), gpu(mask, cuda_device), None if src is None else gpu(src, cuda_device))
            assert (z2 - z).abs().max().item() <= 1e-5 * max(1.0, z.abs().max().item())
        mod.precision = "bf16"
        zb = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), None if src is None else gpu(src, cuda_device))
    rms = np.sqrt(np.mean(ref ** 2))
    assert np.abs(zb.cpu().numpy() - ref).max() <= 3e-2 * max(rms, np.abs(ref).max() * 0.1)


def test_fusion_chunking_and_empty(mb, cuda_device):
    """More drugs than one internal chunk (32768 token rows) and B == 0."""
    case = dict(embed_dim=64, num_layers=1, num_heads=2, head_dim=32, ffn_dim=64, actn="relu", norm_first=True,
                agg="mean", nb=0, seed=5)
    mod, sd = make_module(mb, case, cuda_device)
    B, T = 9000, 4  # 36000 rows -> 2 chunks
    tokens, mask = synth.fusion_inputs(B, T, 64, 5)
    with torch.no_grad():
        z = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device)).cpu().numpy()
        idx = np.r_[0:50, 8150:8250, 8950:9000]
        ref = oracle.fusion_forward(sd, case, tokens[idx], mask[idx], dtype=np.float64)
        assert_close(z[idx], ref, 1e-3, "chunked")
        z0 = mod(torch.empty((0, T, 64), device=cuda_device), torch.empty((0, T), dtype=torch.bool, device=cuda_device))
        assert z0.shape == (0, 64)


def test_unsupported_configs_fail_loudly(mb, cuda_device):
    with pytest.raises(NotImplementedError):
        mb.TransformerFusion(32, 0, 1, 2, 16, 32, transformer_agg="median")
    with pytest.raises(NotImplementedError):
        mb.TransformerFusion(32, 0, 1, 2, 16, 32, transformer_actn="tanh")
    mod = mb.TransformerFusion(32, 0, 1, 2, 16, 32, transformer_agg="mean").to(cuda_device)
    with pytest.raises(RuntimeError, match="T=40"):
        mod(torch.zeros(2, 40, 32, device=cuda_device), torch.zeros(2, 40, dtype=torch.bool, device=cuda_device))
    with pytest.raises(RuntimeError, match="CUDA"):
        mod(torch.zeros(2, 4, 32), torch.zeros(2, 4, dtype=torch.bool))


@pytest.mark.parametrize("case", [c for c in META["posenc_mlp"] if c["name"].startswith("mlp")], ids=lambda c: c["name"])
def test_mlp_adaptor_vs_reference_golden(mb, cuda_device, case):
    g = np.load(os.path.join(HERE, "golden", "golden_posenc_mlp.npz"))
    mod = mb.MLPAdaptor(case["E"], case["hidden"], case["E"], case["p"], "ln", case["actn"], "nd")
    assert list(mod.state_dict().keys()) == case["keys"]  # same `fc.N.*` key layout as the reference module
    ops = synth.mlp_adaptor_params(case["E"], case["hidden"], case["E"], case["seed"])
    lin = [o for o in ops if o["op"] in ("linear", "ln")]
    mods = [x for x in mod.fc if isinstance(x, (torch.nn.Linear, torch.nn.LayerNorm))]
    for o, x in zip(lin, mods):
        x.weight.data = torch.from_numpy(o["w"])
        x.bias.data = torch.from_numpy(o["b"])
    mod = mod.to(cuda_device)
    x = np.random.default_rng(case["seed"]).standard_normal((7, case["E"])).astype(np.float32)
    with torch.no_grad():
        y = mod(gpu(x, cuda_device)).cpu().numpy()
    assert_close(y, g[f"{case['name']}.y"], 1e-3, case["name"])


def _encode_case(mb, case, cuda_device, precision="fp32"):
    """FusionEncoder + inputs of one golden `encode` case (tests/golden/make_golden.py: encode_goldens)."""
    g = np.load(os.path.join(HERE, "golden", "golden_encode.npz"))
    name, E, seed, B = case["name"], case["E"], case["seed"], case["B"]
    rng = np.random.default_rng(seed)
    embeds = rng.standard_normal((B, 19, E)).astype(np.float32)
    masks = rng.random((B, 19)) < 0.55
    masks[:, 0] = False
    if "uni_proj" in case["fusion"]:
        masks[1, :] = True
        masks[1, 0] = False
        masks[4, :] = True
        masks[4, 2] = False
    tf = case.get("tf", dict(num_heads=4, head_dim=8, ffn_dim=64))
    hp = dict(transformer_num_layers=2, transformer_att_heads=tf["num_heads"], transformer_head_dim=tf["head_dim"],
              transformer_ffn_dim=tf["ffn_dim"], transformer_dropout=0.1, transformer_actn="gelu",
              transformer_norm_first=True, transformer_batch_first=False, transformer_agg=case["agg"])
    proj = dict(proj_hidden_dims=[48, 40], proj_dropout=0.2, proj_norm="ln", proj_actn="relu", proj_order="nd")
    enc = mb.FusionEncoder(E, case["nb"], 0.1, hp, proj, fusion=case["fusion"], normalize=case["normalize"],
                           pos_emb_type="learnable" if case["pos"] == "learnable" else "sinusoidal", precision=precision)
    cfg = dict(embed_dim=E, num_layers=2, num_heads=tf["num_heads"], head_dim=tf["head_dim"], ffn_dim=tf["ffn_dim"],
               agg=case["agg"])
    sd = synth.fusion_state_dict(cfg, seed)
    if not np.isclose(synth.params_checksum([sd[k] for k in sorted(sd)] + [embeds]), case["checksum"], rtol=1e-9):
        pytest.skip("numpy Generator stream drift")
    enc.transformer.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    get = lambda k: torch.from_numpy(g[f"{name}.{k}"]) if f"{name}.{k}" in g.files else None
    with torch.no_grad():
        if case["nb"] > 0:
            enc.tx_bottleneck_tokens.copy_(get("tx_bottleneck_tokens"))
        if case["agg"] == "cls":
            enc.cls.copy_(get("cls"))
        if case["pos"] == "learnable":
            enc.pos_encoder.pe.copy_(get("pos_encoder.pe"))
        if case["fusion"] == "transformer_uni_proj":
            ops = synth.mlp_adaptor_params(E, [48, 40], E, seed)
            lin = [o for o in ops if o["op"] in ("linear", "ln")]
            for o, x in zip(lin, [x for x in enc.uni_fuser.fc if isinstance(x, (torch.nn.Linear, torch.nn.LayerNorm))]):
                x.weight.copy_(torch.from_numpy(o["w"]))
                x.bias.copy_(torch.from_numpy(o["b"]))
    return enc.to(cuda_device).eval(), embeds, masks, g[f"{name}.z"]


@pytest.mark.parametrize("case", META["encode"], ids=lambda c: c["name"])
def test_fusion_encoder_vs_reference_encode_golden(mb, cuda_device, case):
    """FusionEncoder (assembly kernel + encoder + unimodal bypass) vs the reference's own NovelDDIEncoder.encode."""
    enc, embeds, masks, want = _encode_case(mb, case, cuda_device)
    with torch.no_grad():
        z = enc(gpu(embeds, cuda_device), gpu(masks, cuda_device)).cpu().numpy()
    assert_close(z, want, 1e-3, case["name"])


@pytest.mark.parametrize("case", META["encode"], ids=lambda c: c["name"])
def test_novel_ddi_encoder_dropin_vs_reference_encode_golden(mb, cuda_device, case):
    """madrigal_b200.NovelDDIEncoder.encode — the reference's call signature, driven exactly as the golden generator
    drove the reference's own encode (stub modality encoders returning the seeded embeddings, KG rows in a permuted
    node order behind `drug_index_map`) — against the reference's output."""
    fe, embeds, masks, want = _encode_case(mb, case, cuda_device)
    B = embeds.shape[0]
    e = gpu(embeds, cuda_device)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).to(cuda_device)

    class Mols:
        node_feature = torch.zeros(1)

    class KG:
        x_dict = None
        edge_index_dict = None

    enc = mb.NovelDDIEncoder(fe, str_encoder=lambda mols, feats: {"graph_feature": e[:, 0]},
                             kg_encoder=lambda x, ei: {"drug": e[:, 1][perm]}, cv_encoder=lambda cv: e[:, 2],
                             tx_encoder_dict={cl: (lambda sigs: sigs) for cl in synth.CELL_LINES}, kg_encoder_name="hgt")
    tx = {cl: {"sigs": e[:, 3 + i]} for i, cl in enumerate(synth.CELL_LINES)}
    kg = {"data": KG(), "drug_index_map": perm}
    with torch.no_grad():
        z = enc.encode(torch.arange(B, device=cuda_device), gpu(masks, cuda_device), Mols(), kg, None, tx)
        raw = enc(torch.arange(B, device=cuda_device), gpu(masks, cuda_device), Mols(), kg, None, tx,
                  raw_encoder_output=True)
    assert_close(z.cpu().numpy(), want, 1e-3, case["name"])
    assert raw.shape == (int((~masks).sum()), embeds.shape[2]) and torch.isfinite(raw).all()


def test_model_wrapper_end_to_end(mb, cuda_device):
    """NovelDDIMultilabel drop-in: tokens -> z (FusionEncoder) -> logits, vs the oracle chain."""
    E, L, B = 128, 5, 96
    hp = dict(transformer_num_layers=2, transformer_att_heads=8, transformer_head_dim=32, transformer_ffn_dim=256,
              transformer_dropout=0.1, transformer_actn="gelu", transformer_norm_first=True,
              transformer_batch_first=False, transformer_agg="x-attn")
    proj = dict(proj_hidden_dims=[64, 64], proj_dropout=0.0, proj_norm="ln", proj_actn="relu", proj_order="nd")
    torch.manual_seed(0)
    enc = mb.FusionEncoder(E, 4, 0.1, hp, proj, fusion="transformer", normalize=False, pos_emb_type="sinusoidal")
    model = mb.NovelDDIMultilabel(mb.PrecomputedEmbeddingEncoder(enc), E, L, normalize=True).to(cuda_device).eval()
    keys = set(model.state_dict().keys())
    assert {"decoder.parametrizations.weight.original", "decoder.bias", "encoder.fusion_encoder.tx_bottleneck_tokens",
            "encoder.fusion_encoder.transformer.x_attn_query"} <= keys
    rng = np.random.default_rng(3)
    embeds = rng.standard_normal((B, 19, E)).astype(np.float32)
    masks = rng.random((B, 19)) < 0.5
    masks[:, 0] = False
    batch = {"drugs": None, "strs": None, "cv": None, "tx": {"all_embeds": gpu(embeds, cuda_device)}}
    with torch.no_grad():
        logits = model(batch, batch, gpu(masks, cuda_device), gpu(masks, cuda_device), None, label_range=(1, 4))
    # oracle chain on the same parameters
    sd = {k[len("encoder.fusion_encoder.transformer."):]: v.cpu().numpy() for k, v in model.state_dict().items()
          if k.startswith("encoder.fusion_encoder.transformer.")}
    cfg = dict(num_layers=2, num_heads=8, head_dim=32, ffn_dim=256, actn="gelu", norm_first=True, agg="x-attn")
    seq, fmask, src = oracle.assemble_fusion_inputs(
        embeds, masks, n_non_tx=3, num_tx_bottlenecks=4, agg="x-attn",
        tx_bottleneck_tokens=model.encoder.fusion_encoder.tx_bottleneck_tokens.detach().cpu().numpy(),
        pe=oracle.sinusoidal_pe(E, 3, 23), pos_emb_type="sinusoidal")
    pool = np.zeros(23, bool)
    pool[:3] = True
    pool[-16:] = True
    z = oracle.fusion_forward(sd, cfg, seq, fmask, src, pool, dtype=np.float64)
    zn = z / np.maximum(np.linalg.norm(z, axis=1, keepdims=True), 1e-12)
    W = oracle.symmetric(model.decoder.parametrizations.weight.original.detach().cpu().numpy().astype(np.float64))
    ref = oracle.bilinear_scores(zn, zn, W, (1, 4), dtype=np.float64)
    assert logits.shape == (3, B, B)
    assert_close(logits.cpu().numpy(), ref, 2e-3, "tokens -> logits")


# ---------------------------------------------------------------------------------------------------------------
# Fused single-kernel encoder (csrc/fused_encoder.cuh): taken for precision='bf16', pre-LN, latent <= 256.
# Checked against the fp64 oracle (bf16-operand tolerance) and against the multi-kernel bf16 path (same operand
# rounding points except k/v, which the fused kernel exchanges in bf16).
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("agg,actn,T,nb,B,dims", [
    ("x-attn", "gelu", 4, 0, 4096, (256, 8, 32, 512)),   # bench.py's encoder (BASELINE config 2)
    ("mean", "gelu", 4, 0, 1000, (128, 8, 32, 512)),     # BASELINE config 5 shape
    ("cls", "relu", 5, 0, 257, (128, 8, 16, 256)),       # T does not divide 128; head_dim 16, two attention phases
    ("max", "relu", 1, 0, 130, (64, 4, 16, 64)),         # one token per drug; head_dim 16
    ("x-attn", "gelu", 23, 4, 77, (128, 4, 32, 200)),    # production token layout (src_mask, bottlenecks), F % 64 != 0
    ("mean", "gelu", 32, 0, 9, (48, 2, 32, 1024)),       # T = 32, latent 64, 4 FFN chunks, E % 64 != 0
    ("cls", "gelu", 19, 0, 50, (256, 6, 32, 128)),       # latent 192, three attention phases
])
def test_fused_encoder_kernel(mb, cuda_device, agg, actn, T, nb, B, dims):
    from madrigal_b200 import _lib
    E, H, hd, F = dims
    case = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn=actn, norm_first=True,
                agg=agg, nb=nb, seed=31 + T)
    mod, sd = make_module(mb, case, cuda_device, precision="bf16")
    tokens, mask = synth.fusion_inputs(B, T, E, case["seed"], always_visible=(0,) + tuple(range(3, 3 + nb)))
    src = None
    if nb > 0:
        src = np.zeros((T, T), bool)
        src[:3, T - 16:] = True
        src[T - 16:, :3] = True
    pool = None
    if agg == "x-attn":
        pool = np.zeros(T, bool)
        if nb > 0:
            pool[:3] = True
            pool[-16:] = True
        mod.x_attn_key_padding_mask = torch.from_numpy(pool)[None, :]
    ref = oracle.fusion_forward(sd, case, tokens, mask, src, pool, dtype=np.float64)
    args = (gpu(tokens, cuda_device), gpu(mask, cuda_device), None if src is None else gpu(src, cuda_device))
    with torch.no_grad():
        z = mod(*args).cpu().numpy()
        assert mod.last_launch_count == 1, "the fused kernel must be ONE launch"
        os.environ["MDG_FUSION_GENERIC"] = "1"
        try:
            zg = mod(*args).cpu().numpy()
            assert mod.last_launch_count > 1
        finally:
            del os.environ["MDG_FUSION_GENERIC"]
    rms = np.sqrt(np.mean(ref ** 2))
    assert np.isfinite(z).all()
    assert np.abs(z - ref).max() <= 3e-2 * max(rms, np.abs(ref).max() * 0.1), "fused vs fp64 oracle"
    # two bf16 evaluations with different rounding points (the fused kernel folds the LayerNorm affine into the next
    # linear's weights and exchanges k/v in bf16): each is within the bf16 bar of the fp64 oracle, so is their difference
    assert np.abs(z - zg).max() <= 3e-2 * max(rms, np.abs(ref).max() * 0.1), "fused vs multi-kernel bf16 path"
    # masked-slot contents are don't-care
    if agg in ("mean", "max") or nb > 0:
        tok2 = np.where(mask[:, :, None], np.float32(9.0), tokens)
        with torch.no_grad():
            z2 = mod(gpu(tok2, cuda_device), *args[1:]).cpu().numpy()
        assert np.abs(z2 - z).max() <= 1e-5 * max(1.0, np.abs(z).max())


def test_mlp_encoder_is_the_tabular_modality_encoder(mb, cuda_device):
    """reference MLPEncoder (models.py:121-180, the cv / tx 'mlp' modality encoder) == MLPAdaptor line for line: same
    kernel chain, same `fc.*` keys; checked against the oracle's restatement with norm=None (the cv default) and 'ln'."""
    for norm in (None, "ln"):
        mod = mb.MLPEncoder(96, [64, 48], 128, 0.1, norm, "relu", "nd")
        ops = []
        for m in mod.fc:
            if isinstance(m, torch.nn.Linear):
                ops.append({"op": "linear", "w": m.weight.detach().numpy().copy(), "b": m.bias.detach().numpy().copy()})
            elif isinstance(m, torch.nn.LayerNorm):
                ops.append({"op": "ln", "w": m.weight.detach().numpy().copy(), "b": m.bias.detach().numpy().copy()})
            elif isinstance(m, (torch.nn.ReLU, torch.nn.GELU)):
                ops.append({"op": "act", "actn": "relu"})
        x = np.random.default_rng(4).standard_normal((33, 96)).astype(np.float32)
        ref = oracle.mlp_adaptor(ops, x, dtype=np.float64)
        y = mod.to(cuda_device)(gpu(x, cuda_device)).detach().cpu().numpy()
        assert_close(y, ref, 1e-3, f"MLPEncoder norm={norm}")


def _chemcpa_module(mb, case, sd, table, dev, precision):
    emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(table), freeze=True)
    mod = mb.chemcpa.TxAdaptingComPert(num_genes=case["num_genes"], num_drugs=case["num_drugs"],
                                       covariate_names_unique={"cell_iname": [f"C{i}" for i in range(case["n_cell"])]},
                                       doser_type=case["doser_type"], hparams=dict(case["hparams"]),
                                       drug_embeddings=emb, append_layer_width=None, use_drugs=case["use_drugs"],
                                       disable_adv=True, precision=precision)
    res = mod.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys and [k for k in res.missing_keys if k != "drug_embeddings.weight"] == []
    return mod.to(dev).eval()


@pytest.mark.parametrize("case", synth.CHEMCPA_CASES, ids=lambda c: c["name"])
def test_chemcpa_tx_encoder_vs_reference_golden(mb, cuda_device, case):
    """chemcpa.TxAdaptingComPert.predict (mdg_mlp_forward with folded batch-norm + mdg_tx_latent_combine) vs the
    reference module's latents (golden) and the oracle; same state_dict keys as the reference."""
    g = np.load(os.path.join(HERE, "golden", "golden_chemcpa.npz"))
    sd, table, inp = synth.chemcpa_case(case)
    ref_b, ref_t = oracle.chemcpa_tx_latents(sd, inp["genes"], [inp["cov_idx"]], use_drugs=case["use_drugs"],
                                             doser_type=case["doser_type"], drug_table=table,
                                             drugs_idx=inp["drugs_idx"], dosages=inp["dosages"], dtype=np.float64)
    onehot = torch.nn.functional.one_hot(torch.from_numpy(inp["cov_idx"]), case["n_cell"]).long()
    for precision, tol in (("fp32", 1e-3), ("bf16", 3e-2)):
        mod = _chemcpa_module(mb, case, sd, table, cuda_device, precision)
        out = mod.predict(genes=gpu(inp["genes"], cuda_device), drugs_idx=gpu(inp["drugs_idx"], cuda_device),
                          dosages=gpu(inp["dosages"], cuda_device), covariates=[onehot.to(cuda_device)],
                          return_latent_basal=True, return_latent_treated=True)
        assert out[0] is None and out[1] is None and len(out) == 4
        basal, treated = out[2].cpu().numpy(), out[3].cpu().numpy()
        assert_close(basal, ref_b, tol, f"{case['name']} basal {precision}")
        assert_close(treated, ref_t, tol, f"{case['name']} treated {precision}")
        if precision == "fp32":
            assert_close(basal, g[f"{case['name']}.basal"], tol, "basal vs golden")
            assert_close(treated, g[f"{case['name']}.treated"], tol, "treated vs golden")
    # the way NovelDDIEncoder.encode takes its tx tokens (models.py:761-769): one latent, split per cell line
    only = mod.predict(genes=gpu(inp["genes"], cuda_device), drugs_idx=gpu(inp["drugs_idx"], cuda_device),
                       dosages=gpu(inp["dosages"], cuda_device), covariates=[onehot.to(cuda_device)],
                       return_latent_basal=False, return_latent_treated=True)
    assert len(only) == 3 and torch.equal(only[2], out[3])


def test_chemcpa_mlp_rejects_training_mode_and_cpu(mb, cuda_device):
    m = mb.chemcpa.MLP([16, 32, 8])
    with pytest.raises(RuntimeError):
        m(torch.zeros(2, 16, device=cuda_device))  # training mode: batch statistics are not an inference path
    with pytest.raises((RuntimeError, ValueError, TypeError)):
        m.eval()(torch.zeros(2, 16))  # CPU tensor: no fallback
    assert list(mb.chemcpa.MLP([4, 5, 6], append_layer_width=3, append_layer_position="first").state_dict())[:2] == \
        ["network.append_linear.weight", "network.append_linear.bias"]


@pytest.mark.parametrize("agg,T,nb,B,dims", [
    ("x-attn", 23, 4, 70, (128, 8, 64, 256)),    # production DrugBank shape: 8 heads of 64
    ("x-attn", 21, 2, 50, (128, 2, 256, 512)),   # production TWOSIDES shape: 2 heads of 256
    ("mean", 9, 0, 33, (64, 3, 128, 128)),       # head_dim 128, one query tile of 16 rows, two key tiles
    ("cls", 16, 0, 40, (64, 2, 64, 96)),         # T = 16 exactly (no second query tile)
    ("max", 32, 0, 21, (96, 1, 128, 64)),        # T = 32: every key tile full
])
def test_tensor_core_attention_vs_fma_attention(mb, cuda_device, agg, T, nb, B, dims):
    """bf16 mode, 8 < T <= 32, head_dim 64/128/256: attention_mma_kernel (warp-level MMAs) vs the FMA attention kernels
    (MDG_ATTENTION_MMA=0) and the fp64 oracle; missing-modality masks, src_mask and bottleneck tokens included."""
    E, H, hd, F = dims
    case = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True,
                agg=agg, nb=nb, seed=51 + T)
    mod, sd = make_module(mb, case, cuda_device, precision="bf16")
    tokens, mask = synth.fusion_inputs(B, T, E, case["seed"], always_visible=(0,) + tuple(range(3, 3 + nb)))
    src = None
    if nb > 0:
        src = np.zeros((T, T), bool)
        src[:3, T - 16:] = True
        src[T - 16:, :3] = True
    pool = None
    if agg == "x-attn":
        pool = np.zeros(T, bool)
        if nb > 0:
            pool[:3] = True
            pool[-16:] = True
        mod.x_attn_key_padding_mask = torch.from_numpy(pool)[None, :]
    ref = oracle.fusion_forward(sd, case, tokens, mask, src, pool, dtype=np.float64)
    args = (gpu(tokens, cuda_device), gpu(mask, cuda_device), None if src is None else gpu(src, cuda_device))
    os.environ["MDG_FUSION_GENERIC"] = "1"   # latent <= 256 shapes would otherwise take the fused encoder kernel
    try:
        with torch.no_grad():
            z = mod(*args).cpu().numpy()
            os.environ["MDG_ATTENTION_MMA"] = "0"
            try:
                zf = mod(*args).cpu().numpy()
            finally:
                del os.environ["MDG_ATTENTION_MMA"]
    finally:
        del os.environ["MDG_FUSION_GENERIC"]
    rms = np.sqrt(np.mean(ref ** 2))
    bound = 3e-2 * max(rms, np.abs(ref).max() * 0.1)
    assert np.isfinite(z).all() and np.isfinite(zf).all()
    assert np.abs(z - ref).max() <= bound, "tensor-core attention vs fp64 oracle"
    assert np.abs(zf - ref).max() <= bound, "FMA attention vs fp64 oracle"
    assert np.abs(z - zf).max() <= bound and not np.array_equal(z, zf), "the two attention kernels both ran"


@pytest.mark.parametrize("agg,nb,T", [("x-attn", 4, 23), ("mean", 0, 4)])
def test_torch_op_fusion_encode_matches_module(mb, cuda_device, agg, nb, T):
    """torch.ops.madrigal_b200.fusion_encode (functional form: state_dict tensors in, z out) == TransformerFusion.forward;
    the parameter tensors are adopted without a copy; CPU tensors have no kernel."""
    E, H, hd, F = 64, 4, 16, 96
    case = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True,
                agg=agg, nb=nb, seed=91)
    for precision in ("fp32", "bf16"):
        mod, _ = make_module(mb, case, cuda_device, precision=precision)
        tokens, mask = synth.fusion_inputs(37, T, E, 91, always_visible=(0,) + tuple(range(3, 3 + nb)))
        src = None
        if nb > 0:
            src = np.zeros((T, T), bool)
            src[:3, T - 16:] = True
            src[T - 16:, :3] = True
            src = gpu(src, cuda_device)
        params = list(mod.state_dict().values())
        cfg = [E, nb, 2, H, hd, F, 1]
        with torch.no_grad():
            want = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), src)
        got = torch.ops.madrigal_b200.fusion_encode(gpu(tokens, cuda_device), gpu(mask, cuda_device), src, None, params,
                                                    cfg, "gelu", agg, precision)
        assert torch.equal(got, want)
        if agg == "x-attn":  # explicit pooling key mask (the 4-token configurations of BASELINE use this)
            pool = torch.zeros(T, dtype=torch.bool, device=cuda_device)
            pool[1] = True
            mod.x_attn_key_padding_mask = pool.cpu()[None, :]
            with torch.no_grad():
                want2 = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), src)
            got2 = torch.ops.madrigal_b200.fusion_encode(gpu(tokens, cuda_device), gpu(mask, cuda_device), src, pool,
                                                         params, cfg, "gelu", agg, precision)
            assert torch.equal(got2, want2) and not torch.equal(got2, got)
    with pytest.raises(NotImplementedError):
        torch.ops.madrigal_b200.fusion_encode(torch.zeros(2, T, E), torch.zeros(2, T, dtype=torch.bool), None, None,
                                              [p.cpu() for p in params], cfg, "gelu", agg, "fp32")


# ---------------------------------------------------------------------------------------------------------------
# Shipped production shapes in bf16 mode, incl. latent 2048 (8 heads x 256, TWOSIDES hardy_sweep_321.yaml:31-34): the
# generic path's streamed-A GEMMs (K = 2048 does not fit the resident operand buffer) and its workspace plan
# ---------------------------------------------------------------------------------------------------------------
PRODUCTION = [c for c in META["fusion"] if c["name"].startswith("production_")]


@pytest.mark.parametrize("case", PRODUCTION, ids=lambda c: c["name"])
def test_production_shapes_bf16_vs_reference_golden(mb, cuda_device, case):
    """bf16-operand / fp32-accumulate encoder on the reference's shipped shapes against the reference's own fp32
    output: |dz| <= 1e-2 * max|z_ref| element-wise (north-star bf16 bar, taken against the scale of the tensor because
    z has entries near zero), and <= 3e-3 * max|z_ref| on average."""
    g = np.load(os.path.join(HERE, "golden", "golden_fusion.npz"))
    mod, sd = make_module(mb, case, cuda_device, precision="bf16")
    tokens, mask = synth.fusion_inputs(case["B"], case["T"], case["embed_dim"], case["seed"],
                                       always_visible=tuple(case["always_visible"]))
    if not np.isclose(synth.params_checksum([sd[k] for k in sorted(sd)] + [tokens]), case["checksum"], rtol=1e-9):
        pytest.skip("numpy Generator stream drift")
    name = case["name"]
    src = gpu(g[f"{name}.src_mask"], cuda_device)
    mod.x_attn_key_padding_mask = torch.from_numpy(g[f"{name}.pool_mask"])[None, :]
    with torch.no_grad():
        z = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), src).cpu().numpy()
    ref = g[f"{name}.z"].astype(np.float64)
    err = np.abs(z - ref)
    assert np.isfinite(z).all()
    assert err.max() <= 1e-2 * np.abs(ref).max(), f"{name}: {err.max():.3e} vs max|ref| {np.abs(ref).max():.3e}"
    assert err.mean() <= 3e-3 * np.abs(ref).max(), name


def test_latent_2048_uniproj_encode_bf16_vs_reference_golden(mb, cuda_device):
    """The whole fusion section of the shipped TWOSIDES hardy_sweep_321 configuration (latent 2048, FFN 1024, 2
    bottlenecks, sinusoidal positions, fusion='transformer_uni_proj') in bf16 mode against the reference's own
    NovelDDIEncoder.encode (fp32 mode is covered by test_fusion_encoder_vs_reference_encode_golden)."""
    case = [c for c in META["encode"] if c["name"] == "enc_twosides_hd256x8_uniproj"][0]
    enc, embeds, masks, want = _encode_case(mb, case, cuda_device, precision="bf16")
    with torch.no_grad():
        z = enc(gpu(embeds, cuda_device), gpu(masks, cuda_device)).cpu().numpy()
    err = np.abs(z - want.astype(np.float64))
    assert np.isfinite(z).all() and err.max() <= 1e-2 * np.abs(want).max(), err.max()


@pytest.mark.parametrize("path", ["fused", "generic"])
def test_bf16_logit_level_chain_T23_xattn(mb, cuda_device, path):
    """North-star bf16 bar at the LOGIT level through the production token layout: 23 tokens (3 non-TX + 4 bottleneck
    + 16 TX, src_mask, x-attn pooling over the bottlenecks) -> encoder (fused one-launch kernel / generic multi-kernel
    path) -> bilinear decoder, all bf16-operand / fp32-accumulate, against the fp64 oracle chain:
    max |logit - ref| <= 1e-2 * max|ref|."""
    E, H, hd, F, T, nb, B, L = 128, 8, 32, 512, 23, 4, 512, 6
    case = dict(embed_dim=E, num_layers=2, num_heads=H, head_dim=hd, ffn_dim=F, actn="gelu", norm_first=True,
                agg="x-attn", nb=nb, seed=77)
    mod, sd = make_module(mb, case, cuda_device, precision="bf16")
    tokens, mask = synth.fusion_inputs(B, T, E, case["seed"], always_visible=(0,) + tuple(range(3, 3 + nb)))
    src = np.zeros((T, T), bool)
    src[:3, T - 16:] = True
    src[T - 16:, :3] = True
    pool = np.zeros(T, bool)
    pool[:3] = True
    pool[-16:] = True
    mod.x_attn_key_padding_mask = torch.from_numpy(pool)[None, :]
    _, W = synth.decoder_inputs(1, E, L, seed=78)
    z_ref = oracle.fusion_forward(sd, case, tokens, mask, src, pool, dtype=np.float64)
    lg_ref = oracle.bilinear_scores(z_ref, z_ref, W, dtype=np.float64)
    if path == "generic":
        os.environ["MDG_FUSION_GENERIC"] = "1"
    try:
        with torch.no_grad():
            z = mod(gpu(tokens, cuda_device), gpu(mask, cuda_device), gpu(src, cuda_device))
            assert (mod.last_launch_count == 1) == (path == "fused")
            lg = mb.pair_score(z, z, gpu(W, cuda_device), precision="bf16", out="logit").cpu().numpy()
    finally:
        os.environ.pop("MDG_FUSION_GENERIC", None)
    err = np.abs(lg - lg_ref)
    assert np.isfinite(lg).all()
    assert err.max() <= 1e-2 * np.abs(lg_ref).max(), f"{path}: {err.max():.3e} vs {np.abs(lg_ref).max():.3e}"
    assert err.mean() <= 2e-3 * np.abs(lg_ref).max(), path


def test_normalize_paths_use_the_row_normalise_kernel(mb, cuda_device):
    """fusion='mean' with normalize=True and the unimodal bypass with normalize=True go through mdg_l2_normalize_rows
    (no eager PyTorch arithmetic): compare with the oracle restatement (models.py:849-850, 861-862, 870-873)."""
    E, B = 64, 37
    rng = np.random.default_rng(9)
    embeds = rng.standard_normal((B, 19, E)).astype(np.float32)
    masks = rng.random((B, 19)) < 0.5
    masks[:, 0] = False
    hp = dict(transformer_num_layers=1, transformer_att_heads=2, transformer_head_dim=32, transformer_ffn_dim=64,
              transformer_dropout=0.0, transformer_actn="gelu", transformer_norm_first=True,
              transformer_batch_first=False, transformer_agg="x-attn")
    proj = dict(proj_hidden_dims=[32], proj_dropout=0.0, proj_norm="ln", proj_actn="relu", proj_order="nd")
    enc = mb.FusionEncoder(E, 0, 0.0, hp, proj, fusion="mean", normalize=True, pos_emb_type="sinusoidal").to(cuda_device)
    with torch.no_grad():
        z = enc(gpu(embeds, cuda_device), gpu(masks, cuda_device)).cpu().numpy()
    x = embeds / np.maximum(np.sqrt((embeds.astype(np.float64) ** 2).sum(-1, keepdims=True)), 1e-12)
    keep = ~masks
    ref = (x * keep[:, :, None]).sum(1) / keep.sum(1, keepdims=True)
    assert np.abs(z - ref).max() <= 1e-5


@pytest.mark.parametrize("case", synth.MLPENCODER_CASES, ids=lambda c: c["name"])
def test_mlp_encoder_vs_reference_golden(mb, cuda_device, case):
    """madrigal_b200.MLPEncoder against the reference `MLPEncoder` outputs (golden_mlpencoder.npz), norm None / 'ln' /
    'bn' (eval-mode BatchNorm1d folded into the following Linear), both dropout orders; fp32-parity mode, 1e-3."""
    g = np.load(os.path.join(HERE, "golden", "golden_mlpencoder.npz"))
    ops = synth.mlp_encoder_ops(case)
    params = [o for o in ops if o["op"] in ("linear", "ln", "bn")]
    mod = mb.MLPEncoder(case["in_dim"], case["hidden"], case["out_dim"], case["p"], case["norm"], case["actn"], case["order"])
    layers = [x for x in mod.fc if isinstance(x, (torch.nn.Linear, torch.nn.LayerNorm, torch.nn.BatchNorm1d))]
    assert len(layers) == len(params)
    with torch.no_grad():
        for x, o in zip(layers, params):
            x.weight.copy_(torch.from_numpy(o["w"]))
            x.bias.copy_(torch.from_numpy(o["b"]))
            if o["op"] == "bn":
                x.running_mean.copy_(torch.from_numpy(o["mean"]))
                x.running_var.copy_(torch.from_numpy(o["var"]))
    mod = mod.to(cuda_device)
    xin = np.random.default_rng(case["seed"]).standard_normal((case["B"], case["in_dim"])).astype(np.float32)
    if case["norm"] == "bn":
        with pytest.raises(RuntimeError, match="inference"):
            mod.train()(gpu(xin, cuda_device))
    y = mod.eval()(gpu(xin, cuda_device)).detach().cpu().numpy()
    assert_close(y, g[f"{case['name']}.y"], 1e-3, case["name"])
    if case["norm"] == "bn":  # the fold is refreshed when a running statistic changes
        with torch.no_grad():
            [m for m in mod.fc if isinstance(m, torch.nn.BatchNorm1d)][0].running_mean.add_(0.5)
        y2 = mod(gpu(xin, cuda_device)).detach().cpu().numpy()
        assert np.abs(y2 - y).max() > 1e-4
