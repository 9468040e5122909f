"""Import the UNMODIFIED reference (biopharmaai/Madrigal) modules for the drug-pair scoring path.

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference exists) to
  * validate the CPU restatement in oracle/oracle.py against the reference's own modules, and
  * generate the golden fixtures under tests/golden/ (see tests/golden/make_golden.py).
Nothing on the GPU box may call this: /root/reference does not exist there.

The reference imports torch_geometric / torch_scatter / torchdrug (and a few small helpers) at module
top (madrigal/models/models.py:14-16, madrigal/utils.py:15-17); none of them is installed here and none
is on the scoring path except torch_scatter.scatter_{mean,max,add} (models.py:447,451,873,878), which we
provide as 10-line pure-torch functions with torch_scatter 2.0.9 semantics.
"""
import ast
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("MADRIGAL_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "madrigal"))


class _Anything:
    """Placeholder class/callable for symbols that are imported but never used on the scoring path."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed third-party symbol called on the scoring path")


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []  # behave like a package so `import a.b` works
    def _any(attr):  # any other symbol; dunder lookups (inspect's __file__ probes ...) must fail normally
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything

    mod.__getattr__ = _any
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def _scatter_index(src, index, dim_size):
    if dim_size is None:
        dim_size = int(index.max().item()) + 1
    return index.view(-1, *([1] * (src.dim() - 1))).expand_as(src), dim_size


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    idx, n = _scatter_index(src, index, dim_size)
    return torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    idx, n = _scatter_index(src, index, dim_size)
    total = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).scatter_add_(0, idx, src)
    count = torch.zeros(n, dtype=src.dtype, device=src.device).scatter_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return total / count.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))


def scatter_max(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0
    idx, n = _scatter_index(src, index, dim_size)
    res = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    res = res.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    res = torch.where(torch.isinf(res), torch.zeros_like(res), res)  # torch_scatter fills empty bins with 0
    return res, None


_MODELS = None


def load_reference_models():
    """Return the reference's `madrigal.models.models` module (imported from REFERENCE_ROOT, unmodified)."""
    global _MODELS
    if _MODELS is not None:
        return _MODELS
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_ROOT}")
    for name in ("torch_geometric", "torch_geometric.nn", "torch_geometric.data", "torch_geometric.loader",
                 "torchdrug", "torchdrug.models", "torchdrug.data", "torchdrug.layers", "torchdrug.core"):
        if name not in sys.modules:
            _stub_module(name)
    sys.modules["torchdrug"].models = sys.modules["torchdrug.models"]
    sys.modules["torchdrug"].data = sys.modules["torchdrug.data"]
    if "torch_scatter" not in sys.modules:
        _stub_module("torch_scatter", scatter_mean=scatter_mean, scatter_add=scatter_add, scatter_max=scatter_max)
    for name in ("dotenv", "jsonpickle", "seml", "sacred", "scanpy", "anndata", "rdkit", "dgl", "wandb", "umap"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub_module(name, load_dotenv=lambda *a, **k: None)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference's chemCPA subpackage pulls in a large optional stack; stub only that one submodule
    try:
        _MODELS = importlib.import_module("madrigal.models.models")
    except Exception:
        for k in [k for k in sys.modules if k.startswith("madrigal")]:
            del sys.modules[k]
        _stub_module("madrigal.chemcpa")
        _stub_module("madrigal.chemcpa.chemCPA")
        _stub_module("madrigal.chemcpa.chemCPA.model")
        _stub_module("madrigal.chemcpa.chemcpa_config_utils")
        pkg = types.ModuleType("madrigal")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "madrigal")]
        pkg.__spec__ = importlib.machinery.ModuleSpec("madrigal", loader=None, is_package=True)
        sys.modules["madrigal"] = pkg
        _MODELS = importlib.import_module("madrigal.models.models")
    return _MODELS


def load_reference_normalizer():
    """AST-extract `classwise_normalized_rank_3d_numpy` and `run_slice` from notebooks/normalize_scores.py.

    The script does file I/O at import time (normalize_scores.py:19-32), so it cannot be imported.  The two
    function bodies are compiled from the reference file where it lies; nothing is copied into this repo.
    Returns (classwise_normalized_rank_3d_numpy, make_run_slice) where make_run_slice(raw, out) binds the
    globals `raw_scores`, `raw_scores_norm`, `mask_indices` exactly as normalize_scores.py:26-33 does.
    """
    path = os.path.join(REFERENCE_ROOT, "notebooks", "normalize_scores.py")
    tree = ast.parse(open(path).read(), filename=path)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef)
              and n.name in ("classwise_normalized_rank_3d_numpy", "run_slice")]
    assert len(wanted) == 2
    code = compile(ast.Module(body=wanted, type_ignores=[]), path, "exec")

    def make_env(raw_scores, raw_scores_norm):
        from time import time
        env = {"np": np, "time": time, "print": lambda *a, **k: None}
        exec(code, env)
        env["raw_scores"] = raw_scores
        env["raw_scores_norm"] = raw_scores_norm
        # normalize_scores.py:33
        env["mask_indices"] = np.vstack(np.triu_indices(raw_scores.shape[1], k=0, m=raw_scores.shape[2]))
        return env

    env0 = {"np": np}
    exec(code, env0)
    return env0["classwise_normalized_rank_3d_numpy"], (lambda raw, out: make_env(raw, out)["run_slice"])
