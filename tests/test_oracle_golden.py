"""CPU: oracle/oracle.py against the golden outputs of the UNMODIFIED reference (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

import synth
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
META = json.load(open(os.path.join(G, "golden_meta.json")))


def _load(name):
    return np.load(os.path.join(G, f"golden_{name}.npz"))


def _check_stream(meta, arrays):
    """The fixtures hold outputs only; parameters/inputs are regenerated from seeds.  Guard against RNG drift."""
    got = synth.params_checksum(arrays)
    if not np.isclose(got, meta["checksum"], rtol=1e-9, atol=1e-9):
        pytest.skip("numpy Generator stream differs from the one the fixtures were generated with")


def _close(a, b, tol):
    a64, b64 = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.sqrt(np.mean(b64 ** 2)), 1e-30)
    err = np.abs(a64 - b64).max() / scale
    assert err <= tol, f"max err / rms = {err:.3e} > {tol}"


@pytest.mark.parametrize("case", META["fusion"], ids=lambda c: c["name"])
def test_fusion_forward_matches_reference(case):
    g = _load("fusion")
    sd = synth.fusion_state_dict(case, case["seed"])
    tokens, mask = synth.fusion_inputs(case["B"], case["T"], case["embed_dim"], case["seed"],
                                       always_visible=tuple(case["always_visible"]))
    _check_stream(case, [sd[k] for k in sorted(sd)] + [tokens])
    name = case["name"]
    src = g[f"{name}.src_mask"] if f"{name}.src_mask" in g.files else None
    pool = g[f"{name}.pool_mask"] if f"{name}.pool_mask" in g.files else None
    z = oracle.fusion_forward(sd, case, tokens, mask, src, pool)
    _close(z, g[f"{name}.z"], 2e-5)  # fp32 vs fp32, different summation order
    z64 = oracle.fusion_forward(sd, case, tokens, mask, src, pool, dtype=np.float64)
    _close(z64, g[f"{name}.z"], 2e-5)


def test_fusion_masked_slots_are_dont_care():
    case = META["fusion"][0]
    g = _load("fusion")
    sd = synth.fusion_state_dict(case, case["seed"])
    tokens, mask = synth.fusion_inputs(case["B"], case["T"], case["embed_dim"], case["seed"],
                                       always_visible=tuple(case["always_visible"]))
    tokens2 = np.where(mask[:, :, None], np.float32(-7.5), tokens)
    name = case["name"]
    z1 = oracle.fusion_forward(sd, case, tokens, mask, g[f"{name}.src_mask"], g[f"{name}.pool_mask"])
    z2 = oracle.fusion_forward(sd, case, tokens2, mask, g[f"{name}.src_mask"], g[f"{name}.pool_mask"])
    # x-attn pooling reads only bottleneck tokens; masked slots are only ever masked keys
    _close(z1, z2, 1e-6)


@pytest.mark.parametrize("case", META["decoder"], ids=lambda c: c["name"])
def test_decoder_matches_reference(case):
    g = _load("decoder")
    z1, P = synth.decoder_inputs(case["N1"], case["D"], case["L"], case["seed"], symmetric=False, unit_scale=False)
    z2, _ = synth.decoder_inputs(case["N2"], case["D"], 1, case["seed"] + 50, symmetric=False, unit_scale=False)
    _check_stream(case, [z1, z2, P])
    W = oracle.symmetric(P)
    assert np.isclose(synth.params_checksum([W]), float(g[f"{case['name']}.W_sym_checksum"]), rtol=1e-9)
    assert (W == np.swapaxes(W, 1, 2)).all()
    if case["normalize"]:
        z1, z2 = oracle.l2_normalize(z1), oracle.l2_normalize(z2)
    lr = tuple(case["label_range"]) if case["label_range"] else None
    s = oracle.bilinear_scores(z1, z2, W, lr)
    assert s.shape == g[f"{case['name']}.scores"].shape
    _close(s, g[f"{case['name']}.scores"], 1e-5)


@pytest.mark.parametrize("case", META["normalizer"], ids=lambda c: c["name"])
def test_normalizer_matches_reference(case):
    g = _load("normalizer")
    rng = np.random.default_rng(case["seed"])
    raw = rng.standard_normal((case["L"], case["N"], case["N"])).astype(np.float32)
    if case["ties"]:
        raw = np.round(raw * 4) / 4
    _check_stream(case, [raw])
    ref = g[f"{case['name']}.out"]
    out = oracle.normalize_scores(raw)
    N = case["N"]
    M = N * (N - 1) // 2
    assert out.dtype == np.float32 and out.shape == ref.shape
    # structural invariants visible in the reference's outputs
    assert (out == out.swapaxes(1, 2)).all() and (np.diagonal(out, axis1=1, axis2=2) == 0).all()
    if not case["ties"]:
        assert (out == ref).all()  # bit-identical float32
        assert (oracle.normalize_scores(raw, kind="stable") == ref).all()
        cw = oracle.classwise_normalized_rank(raw.copy())
        assert np.array_equal(cw, g[f"{case['name']}.classwise"])
    else:
        # tie order is numpy-implementation-defined in the reference: check the well-defined envelope
        for l in range(case["L"]):
            v = np.sort(oracle.lower_triangle_values(raw[l]))
            i, j = np.tril_indices(N, -1)
            lo = np.searchsorted(v, raw[l][i, j], "left") + 1
            hi = np.searchsorted(v, raw[l][i, j], "right")
            for arr in (ref[l], out[l], oracle.normalize_scores(raw, kind="stable")[l]):
                r = np.rint(arr[i, j].astype(np.float64) * M).astype(np.int64)
                assert (r >= lo).all() and (r <= hi).all()
                assert sorted(r.tolist()) == list(range(1, M + 1))  # a permutation of 1..M


def test_quantile_rank_brackets_exact_rank():
    """searchsorted_left + 1 <= exact rank <= searchsorted_right with the full sorted sample as the table (Q = M)."""
    rng = np.random.default_rng(7)
    N = 20
    raw = np.round(rng.standard_normal((2, N, N)) * 8).astype(np.float32) / 8
    M = N * (N - 1) // 2
    table = oracle.reference_quantiles(raw, M)
    ref = oracle.normalize_scores(raw)
    i, j = np.tril_indices(N, -1)
    right = oracle.quantile_rank(table, raw, "right")
    left = oracle.quantile_rank(table, raw, "left")
    for l in range(2):
        r = np.rint(ref[l][i, j].astype(np.float64) * M).astype(np.int64)
        assert (left[l][i, j] + 1 <= r).all() and (r <= right[l][i, j]).all()


def test_reference_quantiles_error_bound():
    rng = np.random.default_rng(8)
    N, Q = 64, 128
    raw = rng.standard_normal((1, N, N)).astype(np.float32)
    M = N * (N - 1) // 2
    table = oracle.reference_quantiles(raw, Q)
    assert (np.diff(table, axis=1) >= 0).all()
    ref = oracle.normalize_scores(raw)
    approx = oracle.quantile_rank(table, raw, "right").astype(np.float64) / Q
    i, j = np.tril_indices(N, -1)
    assert np.abs(approx[0][i, j] - ref[0][i, j]).max() <= 1.0 / Q + 1e-9


@pytest.mark.parametrize("case", [c for c in META["posenc_mlp"] if c["name"].startswith("pe")], ids=lambda c: c["name"])
def test_sinusoidal_pe_matches_reference(case):
    g = _load("posenc_mlp")
    ref = g[f"{case['name']}.pe"]
    pe = oracle.sinusoidal_pe(case["E"], case["max_len"], ref.shape[1])
    assert pe.shape == ref.shape
    assert np.abs(pe - ref).max() <= 2e-6


@pytest.mark.parametrize("case", [c for c in META["posenc_mlp"] if c["name"].startswith("mlp")], ids=lambda c: c["name"])
def test_mlp_adaptor_matches_reference(case):
    g = _load("posenc_mlp")
    ops = synth.mlp_adaptor_params(case["E"], case["hidden"], case["E"], case["seed"])
    for o in ops:
        if o["op"] == "act":
            o["actn"] = case["actn"]
    x = np.random.default_rng(case["seed"]).standard_normal((7, case["E"])).astype(np.float32)
    _check_stream(case, [o["w"] for o in ops if o["op"] != "act"] + [x])
    _close(oracle.mlp_adaptor(ops, x), g[f"{case['name']}.y"], 1e-5)


@pytest.mark.parametrize("case", META["encode"], ids=lambda c: c["name"])
def test_encode_assembly_matches_reference(case):
    """Token assembly + fusion + unimodal bypass restated vs the reference's own NovelDDIEncoder.encode."""
    g = _load("encode")
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
    cfg = dict(embed_dim=E, num_layers=2, num_heads=tf["num_heads"], head_dim=tf["head_dim"], ffn_dim=tf["ffn_dim"],
               agg=case["agg"], actn="gelu", norm_first=True)
    sd = synth.fusion_state_dict(cfg, seed)
    _check_stream(case, [sd[k] for k in sorted(sd)] + [embeds])
    get = lambda k: g[f"{name}.{k}"] if f"{name}.{k}" in g.files else None
    ref = g[f"{name}.z"]
    if case["fusion"] == "mean":  # models.py:870-873
        x = embeds
        if case["normalize"]:
            x = x / np.maximum(np.sqrt((x ** 2).sum(-1, keepdims=True)), 1e-12)
        keep = ~masks
        z = (x * keep[:, :, None]).sum(1) / keep.sum(1, keepdims=True)
        _close(z, ref, 1e-5)
        return
    multi = np.ones(B, bool)
    if case["fusion"] == "transformer_uni_proj":
        multi, uni_idx = oracle.split_unimodal(masks)
        assert (~multi).sum() == 2
    if case["pos"] == "sinusoidal":
        T = 19 + case["nb"] + (1 if case["agg"] == "cls" else 0)
        pe = oracle.sinusoidal_pe(E, case["max_len"], T)
    else:
        pe = get("pos_encoder.pe")
    seq, fmask, src = oracle.assemble_fusion_inputs(
        embeds[multi], masks[multi], n_non_tx=3, num_tx_bottlenecks=case["nb"], agg=case["agg"],
        tx_bottleneck_tokens=get("tx_bottleneck_tokens"), cls=get("cls"), pe=pe, pos_emb_type=case["pos"],
        pos_max_len=case["max_len"], normalize=case["normalize"])
    pool = None
    if case["agg"] == "x-attn":  # models.py:382-385
        pool = np.zeros(19 + case["nb"], bool)
        if case["nb"] > 0:
            pool[:3] = True
            pool[-16:] = True
    z = np.empty((B, E), np.float32)
    z[multi] = oracle.fusion_forward(sd, cfg, seq, fmask, src, pool)
    if case["fusion"] == "transformer_uni_proj":
        ops = synth.mlp_adaptor_params(E, [48, 40], E, seed)
        for o in ops:
            if o["op"] == "act":
                o["actn"] = "relu"
        uni = embeds[~multi][np.arange((~multi).sum()), uni_idx]
        z[~multi] = oracle.mlp_adaptor(ops, uni)
    _close(z, ref, 3e-5)


def test_gmean_restatement_matches_scipy_mstats():
    """generate_embeddings.ipynb cell 18 calls scipy.stats.mstats.gmean on the stacked float32 rank tensors; the
    oracle restates it as exp(mean(log a)).  scipy is importable here, so pin the restatement against the real call."""
    from scipy.stats.mstats import gmean
    rng = np.random.default_rng(3)
    members = [oracle.normalize_scores(rng.standard_normal((2, 40, 40)).astype(np.float32)) for _ in range(5)]
    with np.errstate(divide="ignore"):
        ref = np.asarray(gmean(np.stack(members, axis=-1), axis=-1))
    got = oracle.gmean_normalized_ranks(members)
    assert got.dtype == np.float32 and np.allclose(got, ref, rtol=1e-6, atol=0)
    assert (np.diagonal(got, axis1=1, axis2=2) == 0).all()


@pytest.mark.parametrize("case", synth.CHEMCPA_CASES, ids=lambda c: c["name"])
def test_chemcpa_tx_latents_match_reference(case):
    """oracle.chemcpa_tx_latents vs the reference TxAdaptingComPert.predict latents (golden_chemcpa.npz, generated by
    tests/golden/make_golden_chemcpa.py from chemcpa/chemCPA/model.py)."""
    g = np.load(os.path.join(G, "golden_chemcpa.npz"))
    sd, table, inp = synth.chemcpa_case(case)
    assert abs(synth.params_checksum([sd[k] for k in sd] + [table]) - float(g[f"{case['name']}.checksum"])) < 1e-6
    basal, treated = oracle.chemcpa_tx_latents(sd, inp["genes"], [inp["cov_idx"]], use_drugs=case["use_drugs"],
                                               doser_type=case["doser_type"], drug_table=table,
                                               drugs_idx=inp["drugs_idx"], dosages=inp["dosages"])
    np.testing.assert_allclose(basal, g[f"{case['name']}.basal"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(treated, g[f"{case['name']}.treated"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("case", synth.MLPENCODER_CASES, ids=lambda c: c["name"])
def test_mlp_encoder_matches_reference(case):
    """oracle.mlp_adaptor on MLPEncoder's op list vs the reference `MLPEncoder` (models.py:121-180; goldens from
    tests/golden/make_golden_mlpencoder.py), and the drop-in's `fc.*` keys vs the reference module's."""
    g = np.load(os.path.join(G, "golden_mlpencoder.npz"))
    ops = synth.mlp_encoder_ops(case)
    params = [o for o in ops if o["op"] in ("linear", "ln", "bn")]
    chk = synth.params_checksum([o["w"] for o in params] + [o["b"] for o in params])
    assert abs(chk - float(g[f"{case['name']}.checksum"])) < 1e-6
    x = np.random.default_rng(case["seed"]).standard_normal((case["B"], case["in_dim"])).astype(np.float32)
    np.testing.assert_allclose(oracle.mlp_adaptor(ops, x), g[f"{case['name']}.y"], rtol=2e-5, atol=2e-5)
    import madrigal_b200 as mb
    mod = mb.MLPEncoder(case["in_dim"], case["hidden"], case["out_dim"], case["p"], case["norm"], case["actn"], case["order"])
    assert list(mod.state_dict().keys()) == [str(k) for k in g[f"{case['name']}.keys"]]
