"""Golden fixtures for the tabular modality encoder `MLPEncoder` (madrigal/models/models.py:121-180: the `cv` and
`tx: mlp` encoders), from the UNMODIFIED reference class.

    python tests/golden/make_golden_mlpencoder.py       # build container only; writes golden_mlpencoder.npz

Seeded parameters (tests/synth.py: mlp_adaptor_params) are copied into the reference module's Linear / LayerNorm
layers in nn.Sequential order; eval-mode forward on CPU in fp32.  Only outputs and a parameter checksum are stored.
"""
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
from oracle.ref_import import load_reference_models  # noqa: E402

torch.set_grad_enabled(False)
m = load_reference_models()
out = {}
for case in synth.MLPENCODER_CASES:
    ops = synth.mlp_encoder_ops(case)
    mod = m.MLPEncoder(case["in_dim"], case["hidden"], case["out_dim"], case["p"], case["norm"], case["actn"], case["order"])
    layers = [x for x in mod.fc if isinstance(x, (nn.Linear, nn.LayerNorm, nn.BatchNorm1d))]
    params = [o for o in ops if o["op"] in ("linear", "ln", "bn")]
    assert len(layers) == len(params), (len(layers), len(params))
    for x, o in zip(layers, params):
        assert tuple(x.weight.shape) == o["w"].shape
        x.weight.data = torch.from_numpy(o["w"])
        x.bias.data = torch.from_numpy(o["b"])
        if o["op"] == "bn":
            x.running_mean.data = torch.from_numpy(o["mean"])
            x.running_var.data = torch.from_numpy(o["var"])
    mod.eval()
    x = np.random.default_rng(case["seed"]).standard_normal((case["B"], case["in_dim"])).astype(np.float32)
    y = mod(torch.from_numpy(x)).numpy()
    out[f"{case['name']}.y"] = y
    out[f"{case['name']}.keys"] = np.asarray(list(mod.state_dict().keys()))
    out[f"{case['name']}.checksum"] = np.asarray(synth.params_checksum([o["w"] for o in params] + [o["b"] for o in params]))
    print(case["name"], y.shape, float(np.abs(y).max()))
np.savez_compressed(os.path.join(HERE, "golden_mlpencoder.npz"), **out)
