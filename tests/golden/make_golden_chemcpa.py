"""Golden fixtures for the chemCPA transcriptomic encoder's latent path, from the UNMODIFIED reference module
(/root/reference/madrigal/chemcpa/chemCPA/model.py, loaded by file path: it imports only numpy/torch).

    python tests/golden/make_golden_chemcpa.py        # build container only; writes golden_chemcpa.npz

Seeded parameters (tests/synth.py: chemcpa_case) are loaded into `TxAdaptingComPert` (constructed the way
models.py:278-288 does: disable_adv=True, pretrained-style frozen drug embedding table), `.eval()`, and
`predict(..., return_latent_basal=True, return_latent_treated=True)` is run on CPU in fp32.  Only outputs are stored.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_chemcpa_model", "/root/reference/madrigal/chemcpa/chemCPA/model.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

torch.set_grad_enabled(False)
out = {}
for case in synth.CHEMCPA_CASES:
    sd, table, inp = synth.chemcpa_case(case)
    emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(table), freeze=True)
    # optimiser settings the reference constructor insists on (model.py:476-515); no effect on the forward pass
    train_hp = dict(autoencoder_lr=1e-3, autoencoder_wd=0.0, adversary_lr=1e-3, adversary_wd=0.0, dosers_lr=1e-3,
                    dosers_wd=0.0, step_size_lr=45, adversary_width=8, adversary_depth=1)
    model = ref.TxAdaptingComPert(num_genes=case["num_genes"], num_drugs=case["num_drugs"],
                                  covariate_names_unique={"cell_iname": [f"C{i}" for i in range(case["n_cell"])]},
                                  doser_type=case["doser_type"], hparams=dict(case["hparams"], **train_hp), drug_embeddings=emb,
                                  append_layer_width=None, use_drugs=case["use_drugs"], disable_adv=True)
    res = model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    missing = [k for k in res.missing_keys if not k.startswith("decoder.") and k != "drug_embeddings.weight"]
    assert not missing, missing
    model.eval()
    onehot = torch.nn.functional.one_hot(torch.from_numpy(inp["cov_idx"]), case["n_cell"]).long()
    _, _, basal, treated = model.predict(genes=torch.from_numpy(inp["genes"]), drugs_idx=torch.from_numpy(inp["drugs_idx"]),
                                         dosages=torch.from_numpy(inp["dosages"]), covariates=[onehot],
                                         return_latent_basal=True, return_latent_treated=True)
    out[f"{case['name']}.basal"] = basal.numpy()
    out[f"{case['name']}.treated"] = treated.numpy()
    out[f"{case['name']}.keys"] = np.asarray([k for k in model.state_dict().keys()
                                              if not k.startswith(("decoder.", "adversary_"))])
    out[f"{case['name']}.checksum"] = np.asarray(synth.params_checksum([sd[k] for k in sd] + [table]))
    print(case["name"], basal.shape, float(np.abs(treated.numpy()).max()))
np.savez_compressed(os.path.join(HERE, "golden_chemcpa.npz"), **out)
