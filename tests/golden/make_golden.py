"""Generate the golden fixtures in this directory from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports biopharmaai/Madrigal's own modules from /root/reference (oracle/ref_import.py), loads seeded synthetic
parameters (tests/synth.py) into them, runs the reference code on CPU in fp32 and stores only configs, seeds, parameter
checksums and OUTPUTS (small).  tests/test_oracle_golden.py regenerates the same parameters/inputs from the seeds and
checks oracle/oracle.py against these outputs; the GPU tests then check the CUDA path against the oracle.
Nothing here is read on the GPU box except the generated golden_*.npz/json files.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import synth  # noqa: E402
from oracle.ref_import import load_reference_models, load_reference_normalizer  # noqa: E402

torch.set_grad_enabled(False)
m = load_reference_models()

FUSION_CASES = [
    # name, cfg, B, T, nb, n_tx
    dict(name="xattn_prenorm_gelu_nb4", embed_dim=32, num_layers=2, num_heads=4, head_dim=8, ffn_dim=64, actn="gelu",
         norm_first=True, agg="x-attn", nb=4, n_tx=16, B=6),
    dict(name="xattn_postnorm_relu_nb2", embed_dim=32, num_layers=2, num_heads=2, head_dim=16, ffn_dim=48, actn="relu",
         norm_first=False, agg="x-attn", nb=2, n_tx=16, B=5),
    dict(name="xattn_prenorm_nb0", embed_dim=16, num_layers=1, num_heads=2, head_dim=8, ffn_dim=32, actn="gelu",
         norm_first=True, agg="x-attn", nb=0, n_tx=16, B=4),
    dict(name="cls_prenorm_nb4", embed_dim=32, num_layers=2, num_heads=4, head_dim=8, ffn_dim=64, actn="gelu",
         norm_first=True, agg="cls", nb=4, n_tx=16, B=5),
    dict(name="cls_postnorm_nb0", embed_dim=16, num_layers=2, num_heads=2, head_dim=8, ffn_dim=32, actn="relu",
         norm_first=False, agg="cls", nb=0, n_tx=16, B=4),
    dict(name="mean_T4", embed_dim=32, num_layers=2, num_heads=4, head_dim=8, ffn_dim=64, actn="relu",
         norm_first=False, agg="mean", nb=0, n_tx=1, B=9),
    dict(name="max_T4", embed_dim=32, num_layers=2, num_heads=4, head_dim=8, ffn_dim=64, actn="gelu",
         norm_first=True, agg="max", nb=0, n_tx=1, B=9),
    dict(name="xattn_T4_custom_poolmask", embed_dim=32, num_layers=2, num_heads=8, head_dim=4, ffn_dim=64, actn="gelu",
         norm_first=True, agg="x-attn", nb=0, n_tx=1, B=7),
    dict(name="production_drugbank", embed_dim=128, num_layers=2, num_heads=8, head_dim=64, ffn_dim=256, actn="gelu",
         norm_first=True, agg="x-attn", nb=4, n_tx=16, B=3),  # configs/ddi_finetune/DrugBank/*elated_sweep_163.yaml
    dict(name="production_twosides_hd256", embed_dim=128, num_layers=2, num_heads=2, head_dim=256, ffn_dim=512,
         actn="gelu", norm_first=True, agg="x-attn", nb=2, n_tx=16, B=2),  # TWOSIDES/good_sweep_105.yaml
    dict(name="production_twosides_hd256x8", embed_dim=128, num_layers=2, num_heads=8, head_dim=256, ffn_dim=1024,
         actn="gelu", norm_first=True, agg="x-attn", nb=2, n_tx=16, B=3),  # TWOSIDES/hardy_sweep_321.yaml:17,31-34: latent 2048
]


def build_src_mask(T_wo_cls, n_non_tx, n_tx, nb, with_cls):
    """Exactly the construction at models.py:813-842 (run through torch to keep it the reference's own logic)."""
    if nb == 0:
        return None
    src = torch.zeros((T_wo_cls, T_wo_cls), dtype=torch.bool)
    sub = torch.ones((n_non_tx, n_tx), dtype=torch.bool)
    src[:n_non_tx, -n_tx:] = sub
    src[-n_tx:, :n_non_tx] = sub.T
    if with_cls:
        src = torch.cat([torch.zeros((1, src.shape[1]), dtype=torch.bool), src], dim=0)
        src = torch.cat([torch.zeros((src.shape[0], 1), dtype=torch.bool), src], dim=1)
    return src


def fusion_goldens():
    out, meta = {}, []
    for idx, c in enumerate(FUSION_CASES):
        seed = 100 + idx
        n_non_tx = 3
        T = n_non_tx + c["nb"] + c["n_tx"] + (1 if c["agg"] == "cls" else 0)
        mod = m.TransformerFusion(c["embed_dim"], c["nb"], c["num_layers"], c["num_heads"], c["head_dim"],
                                  c["ffn_dim"], transformer_dropout=0.1, transformer_actn=c["actn"],
                                  transformer_norm_first=c["norm_first"], transformer_batch_first=False,
                                  transformer_agg=c["agg"]).eval()
        sd = synth.fusion_state_dict(c, seed)
        missing, unexpected = mod.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        always = tuple(range((1 if c["agg"] == "cls" else 0), (1 if c["agg"] == "cls" else 0) + 1))  # str token
        if c["agg"] == "cls":
            always = (0, 1)
        if c["nb"] > 0:  # bottleneck tokens are never masked (models.py:805)
            off = (1 if c["agg"] == "cls" else 0) + n_non_tx
            always = always + tuple(range(off, off + c["nb"]))
        tokens, mask = synth.fusion_inputs(c["B"], T, c["embed_dim"], seed, always_visible=always)
        src = build_src_mask(T - (1 if c["agg"] == "cls" else 0), n_non_tx, c["n_tx"], c["nb"], c["agg"] == "cls")
        pool_mask = None
        if c["agg"] == "x-attn":
            if c["n_tx"] == 16:
                pool_mask = mod.x_attn_key_padding_mask[0].numpy().copy()  # the module's own constant (models.py:382)
            else:  # synthetic T: the reference hard-wires 19+nb keys, overwrite its constant with a T-long one
                pool_mask = np.zeros(T, dtype=bool)
                pool_mask[-1] = True
                mod.x_attn_key_padding_mask = torch.from_numpy(pool_mask)[None, :]
        z = mod(torch.from_numpy(tokens), torch.from_numpy(mask), src).numpy()
        assert z.shape == (c["B"], c["embed_dim"]) and np.isfinite(z).all(), (c["name"], z.shape)
        out[f"{c['name']}.z"] = z.astype(np.float32)
        if pool_mask is not None:
            out[f"{c['name']}.pool_mask"] = pool_mask
        if src is not None:
            out[f"{c['name']}.src_mask"] = src.numpy()
        meta.append(dict(c, seed=seed, T=T, always_visible=list(always),
                         checksum=synth.params_checksum([sd[k] for k in sorted(sd)] + [tokens])))
    return out, meta


def decoder_goldens():
    out, meta = {}, []
    for idx, (N1, N2, D, L, lr, norm) in enumerate([(16, 24, 32, 5, None, False), (16, 24, 32, 5, (1, 4), False),
                                                    (20, 20, 64, 3, None, True)]):
        seed = 200 + idx
        z1, P = synth.decoder_inputs(N1, D, L, seed, symmetric=False, unit_scale=False)
        z2, _ = synth.decoder_inputs(N2, D, 1, seed + 50, symmetric=False, unit_scale=False)

        class StubEncoder(nn.Module):  # returns the fused embeddings directly (the modality encoders are out of scope)
            def forward(self, drugs, masks, mols, kg, cv, tx, **kw):
                return torch.from_numpy(z1 if drugs == "head" else z2)

        model = m.NovelDDIMultilabel(StubEncoder(), feat_dim=D, prediction_dim=L, normalize=norm).eval()  # models.py:914
        model.decoder.parametrizations.weight.original.data = torch.from_numpy(P)
        W = model.decoder.weight.numpy().copy()  # through Symmetric (models.py:522-524, 922)
        bh = {"drugs": "head", "strs": None, "cv": None, "tx": None}
        bt = {"drugs": "tail", "strs": None, "cv": None, "tx": None}
        s = model(bh, bt, None, None, None, label_range=lr).numpy()
        out[f"dec{idx}.scores"] = s.astype(np.float32)
        out[f"dec{idx}.W_sym_checksum"] = np.array(synth.params_checksum([W]))
        meta.append(dict(name=f"dec{idx}", N1=N1, N2=N2, D=D, L=L, label_range=lr, normalize=norm, seed=seed,
                         checksum=synth.params_checksum([z1, z2, P])))
    return out, meta


def normalizer_goldens():
    classwise, make_run_slice = load_reference_normalizer()
    out, meta = {}, []
    for idx, (L, N, ties) in enumerate([(3, 24, False), (2, 33, False), (2, 16, True)]):
        seed = 300 + idx
        rng = np.random.default_rng(seed)
        raw = rng.standard_normal((L, N, N)).astype(np.float32)
        if ties:
            raw = np.round(raw * 4) / 4  # many equal scores
        norm = np.zeros_like(raw)
        run_slice = make_run_slice(raw, norm)
        for l in range(L):
            run_slice((l, l + 1))  # normalize_scores.py:78-85 maps run_slice over single-outcome slices
        out[f"norm{idx}.out"] = norm
        # also the bare ranking function on an untouched multi-class tensor (the shape[0] > 1 branch, :45-46)
        out[f"norm{idx}.classwise"] = classwise(raw.copy()).astype(np.float64)
        meta.append(dict(name=f"norm{idx}", L=L, N=N, ties=ties, seed=seed, checksum=synth.params_checksum([raw])))
    return out, meta


def posenc_mlp_goldens():
    out, meta = {}, []
    # positional encodings (models.py:551-603)
    for idx, (E, max_len, nb, agg) in enumerate([(32, 19, 0, "x-attn"), (32, 3, 4, "x-attn"), (16, 4, 2, "cls"),
                                                 (16, 20, 0, "cls")]):
        pe = m.PositionEncodingSinusoidal(d_model=E, dropout=0.0, max_len=max_len, num_tx_bottlenecks=nb,
                                          transformer_agg=agg)
        out[f"pe{idx}.pe"] = pe.pe.numpy().copy()
        meta.append(dict(name=f"pe{idx}", E=E, max_len=max_len, nb=nb, agg=agg))
    # MLPAdaptor (models.py:459-518) with defaults proj_norm='ln', proj_order='nd'
    for idx, (E, hidden, actn, p) in enumerate([(32, [64, 48], "relu", 0.2), (16, [32, 32, 24], "gelu", 0.0),
                                                (32, [40], "relu", 0.1)]):
        seed = 400 + idx
        mod = m.MLPAdaptor(E, hidden, E, p, "ln", actn, "nd").eval()
        ops = synth.mlp_adaptor_params(E, hidden, E, seed)
        lin = [o for o in ops if o["op"] in ("linear", "ln")]
        mods = [x for x in mod.fc if isinstance(x, (nn.Linear, nn.LayerNorm))]
        assert len(lin) == len(mods)
        for o, x in zip(lin, mods):
            x.weight.data = torch.from_numpy(o["w"])
            x.bias.data = torch.from_numpy(o["b"])
        rng = np.random.default_rng(seed)
        x = rng.standard_normal((7, E)).astype(np.float32)
        out[f"mlp{idx}.y"] = mod(torch.from_numpy(x.copy())).numpy()
        meta.append(dict(name=f"mlp{idx}", E=E, hidden=hidden, actn=actn, p=p, seed=seed,
                         keys=list(mod.state_dict().keys()), checksum=synth.params_checksum([o["w"] for o in lin] + [x])))
    return out, meta


def encode_goldens():
    """The real NovelDDIEncoder.encode (models.py:717-896) driven with stub modality encoders that return seeded
    modality embeddings: pins token order, bottleneck insertion, src_mask, CLS, pos-enc, normalisation and the
    transformer_uni_proj unimodal/multimodal routing (models.py:772-865)."""
    out, meta = {}, []
    cases = [
        dict(name="enc_xattn_nb4_sin", agg="x-attn", nb=4, pos="sinusoidal", fusion="transformer", normalize=False, B=8),
        dict(name="enc_xattn_nb2_uniproj", agg="x-attn", nb=2, pos="learnable", fusion="transformer_uni_proj",
             normalize=False, B=10),
        dict(name="enc_cls_nb2_norm", agg="cls", nb=2, pos="learnable", fusion="transformer", normalize=True, B=6),
        dict(name="enc_cls_nb0_sin", agg="cls", nb=0, pos="sinusoidal", fusion="transformer_uni_proj", normalize=False,
             B=7),
        dict(name="enc_mean_fusion", agg="x-attn", nb=0, pos="sinusoidal", fusion="mean", normalize=True, B=5),
        # the shipped TWOSIDES hardy_sweep_321 configuration at full size (sweep_config_hardy_sweep_321.yaml:17, 21,
        # 31-34): feature_dim 128, 8 heads x 256 = latent 2048, FFN 1024, 2 bottlenecks, sinusoidal positions,
        # fusion 'transformer_uni_proj' (unimodal drugs bypass the transformer through uni_fuser)
        dict(name="enc_twosides_hd256x8_uniproj", agg="x-attn", nb=2, pos="sinusoidal", fusion="transformer_uni_proj",
             normalize=False, B=6, E=128, tf=dict(num_heads=8, head_dim=256, ffn_dim=1024)),
    ]
    for idx, c in enumerate(cases):
        E = c.get("E", 32)
        tf = c.get("tf", dict(num_heads=4, head_dim=8, ffn_dim=64))
        hp = dict(transformer_num_layers=2, transformer_att_heads=tf["num_heads"], transformer_head_dim=tf["head_dim"],
                  transformer_ffn_dim=tf["ffn_dim"], transformer_dropout=0.1, transformer_actn="gelu",
                  transformer_norm_first=True, transformer_batch_first=False)
        seed = 500 + idx
        rng = np.random.default_rng(seed)
        B = c["B"]
        embeds = rng.standard_normal((B, 19, E)).astype(np.float32)
        masks = rng.random((B, 19)) < 0.55
        masks[:, 0] = False
        if "uni_proj" in c["fusion"]:  # make some rows unimodal (single visible modality), one of them not `str`
            masks[1, :] = True
            masks[1, 0] = False
            masks[4, :] = True
            masks[4, 2] = False
        enc = object.__new__(m.NovelDDIEncoder)
        nn.Module.__init__(enc)
        enc.embed_dim, enc.fusion, enc.normalize = E, c["fusion"], c["normalize"]
        enc.adapt_before_fusion, enc.use_tx_basal = False, False
        enc.str_encoder = lambda mols, feats: {"graph_feature": torch.from_numpy(embeds[:, 0])}
        enc.kg_encoder_name = "hgt"
        enc.kg_encoder = lambda x, e: {"drug": torch.from_numpy(embeds[:, 1])}
        enc.cv_encoder = lambda cv: torch.from_numpy(embeds[:, 2])
        enc.tabular_mod_encoders = {}
        enc.tx_encoder_dict = {cl: (lambda sigs: sigs) for cl in synth.CELL_LINES}
        enc.num_tx_bottlenecks, enc.transformer_agg = c["nb"], c["agg"]
        max_len = (19 if c["nb"] == 0 else 3) + (1 if c["agg"] == "cls" else 0)  # models.py:668-676
        cfg = dict(embed_dim=E, num_layers=2, num_heads=tf["num_heads"], head_dim=tf["head_dim"], ffn_dim=tf["ffn_dim"],
                   agg=c["agg"])
        sd = synth.fusion_state_dict(cfg, seed)
        extra = {}
        if c["nb"] > 0:
            extra["tx_bottleneck_tokens"] = rng.standard_normal((c["nb"], E)).astype(np.float32)
            enc.tx_bottleneck_tokens = nn.Parameter(torch.from_numpy(extra["tx_bottleneck_tokens"]))
        if c["pos"] == "sinusoidal":
            enc.pos_encoder = m.PositionEncodingSinusoidal(d_model=E, dropout=0.1, max_len=max_len,
                                                           num_tx_bottlenecks=c["nb"], transformer_agg=c["agg"])
        else:
            enc.pos_encoder = m.PositionEncodingLearnable(d_model=E, dropout=0.1, max_len=max_len,
                                                          num_tx_bottlenecks=c["nb"], transformer_agg=c["agg"])
            extra["pos_encoder.pe"] = rng.standard_normal((1, max_len, E)).astype(np.float32)
            enc.pos_encoder.pe.data = torch.from_numpy(extra["pos_encoder.pe"])
        enc.transformer = m.TransformerFusion(E, c["nb"], transformer_agg=c["agg"], **hp)
        enc.transformer.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        if c["agg"] == "cls":
            extra["cls"] = rng.standard_normal((1, E)).astype(np.float32)
            enc.cls = nn.Parameter(torch.from_numpy(extra["cls"]))
        ops = synth.mlp_adaptor_params(E, [48, 40], E, seed)
        enc.uni_fuser = m.MLPAdaptor(E, [48, 40], E, 0.2, "ln", "relu", "nd")
        lin = [o for o in ops if o["op"] in ("linear", "ln")]
        for o, x in zip(lin, [x for x in enc.uni_fuser.fc if isinstance(x, (nn.Linear, nn.LayerNorm))]):
            x.weight.data = torch.from_numpy(o["w"])
            x.bias.data = torch.from_numpy(o["b"])
        enc.eval()

        class Mols:
            node_feature = torch.zeros(1)

        drugs = torch.arange(B)
        kg = {"data": type("KG", (), {"x_dict": None, "edge_index_dict": None})(), "drug_index_map": torch.arange(B)}
        tx = {cl: {"sigs": torch.from_numpy(embeds[:, 3 + i])} for i, cl in enumerate(synth.CELL_LINES)}
        z = enc.encode(drugs, torch.from_numpy(masks), Mols(), kg, None, tx).numpy()
        assert z.shape == (B, E) and np.isfinite(z).all()
        out[f"{c['name']}.z"] = z.astype(np.float32)
        for k, v in extra.items():
            out[f"{c['name']}.{k}"] = v
        meta.append(dict(c, seed=seed, E=E, tf=tf, max_len=max_len,
                         checksum=synth.params_checksum([sd[k] for k in sorted(sd)] + [embeds])))
    return out, meta


if __name__ == "__main__":
    groups = {"fusion": fusion_goldens, "decoder": decoder_goldens, "normalizer": normalizer_goldens,
              "posenc_mlp": posenc_mlp_goldens, "encode": encode_goldens}
    all_meta = {"numpy": np.__version__, "torch": torch.__version__,
                "reference": "biopharmaai/Madrigal @ /root/reference (unmodified, imported via oracle/ref_import.py)"}
    for name, fn in groups.items():
        arrays, meta = fn()
        np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **arrays)
        all_meta[name] = meta
        print(name, len(arrays), "arrays", sum(a.nbytes for a in arrays.values()), "bytes")
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(all_meta, f, indent=1, default=lambda o: list(o) if isinstance(o, tuple) else str(o))
