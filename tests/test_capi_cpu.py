"""CPU: the C-ABI library builds, loads, and exports every symbol include/madrigal_b200.h declares.
No compute call is made (there is no GPU here and the library has no CPU path)."""
import ctypes
import os
import re

import pytest

from madrigal_b200 import _lib
from madrigal_b200.build import LIB_PATH, REPO_ROOT, build_library


@pytest.fixture(scope="module")
def handle():
    build_library()
    assert os.path.exists(LIB_PATH)
    return _lib.lib()


def _declared_functions():
    text = open(os.path.join(REPO_ROOT, "include", "madrigal_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mdg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(handle):
    names = _declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"


def test_abi_version_and_error_string(handle):
    assert handle.mdg_abi_version() == _lib.EXPECTED_ABI == 5
    assert isinstance(handle.mdg_last_error(), bytes)


def test_invalid_arguments_fail_without_touching_the_gpu(handle):
    # argument validation happens before any CUDA call
    rc = handle.mdg_pair_score(None, None, None, 4, 4, 128, 1, 0, 0, 0, 0, None, None, None, 0, None)
    assert rc == 1 and b"NULL" in handle.mdg_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = handle.mdg_pair_score(p, p, p, 4, 4, 100, 1, 0, 0, 0, 0, None, p, None, 0, None)
    assert rc == 2 and b"D=100" in handle.mdg_last_error()
    rc = handle.mdg_pair_score(p, p, p, 4, 4, 128, 1, 0, 2, 0, 0, None, p, None, 0, None)
    assert rc == 1 and b"rank table" in handle.mdg_last_error()
    rc = handle.mdg_rank_table_build(p, 1, 70000, p + 16, p, p, None)
    assert rc != 0


def test_workspace_size_is_monotone(handle):
    a = handle.mdg_pair_score_workspace_bytes(1024, 1024, 128, 86, 0)
    b = handle.mdg_pair_score_workspace_bytes(1024, 1024, 128, 86, 1)
    c = handle.mdg_pair_score_workspace_bytes(4096, 4096, 256, 86, 0)
    assert 0 < a < b and a < c


def test_python_ops_refuse_cpu_tensors():
    import torch
    import madrigal_b200 as mb
    z = torch.zeros(4, 128)
    W = torch.zeros(1, 128, 128)
    with pytest.raises(RuntimeError, match="CUDA"):
        mb.pair_score(z, z, W)


def test_chemcpa_dropin_has_the_reference_state_dict_keys():
    """madrigal_b200.chemcpa.TxAdaptingComPert registers exactly the reference module's parameters / buffers for the
    parts on the path (keys recorded from the reference by tests/golden/make_golden_chemcpa.py; its decoder.* and
    adversary_* entries have no counterpart).  Module construction only: no GPU call."""
    import numpy as np
    import torch
    import synth
    from madrigal_b200 import chemcpa
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_chemcpa.npz"))
    for case in synth.CHEMCPA_CASES:
        _, table, _ = synth.chemcpa_case(case)
        emb = torch.nn.Embedding.from_pretrained(torch.from_numpy(table), freeze=True)
        mod = chemcpa.TxAdaptingComPert(num_genes=case["num_genes"], num_drugs=case["num_drugs"],
                                        covariate_names_unique={"cell_iname": [f"C{i}" for i in range(case["n_cell"])]},
                                        doser_type=case["doser_type"], hparams=dict(case["hparams"]),
                                        drug_embeddings=emb, use_drugs=case["use_drugs"], disable_adv=True)
        assert list(mod.state_dict().keys()) == [str(k) for k in g[f"{case['name']}.keys"]], case["name"]


def test_more_entry_points_validate_arguments_first(handle):
    """Every entry point added after the first ABI draft rejects NULL / out-of-range arguments with a message before it
    touches the device (so these calls are safe on a machine without a GPU)."""
    buf = ctypes.create_string_buffer(256)
    p = ctypes.addressof(buf)
    assert handle.mdg_tx_latent_combine(None, None, None, None, None, None, 0, None, None, 4, 8, None, None) == 1
    assert b"NULL" in handle.mdg_last_error()
    assert handle.mdg_tx_latent_combine(p, None, None, None, None, None, 7, None, None, 4, 8, p, None) != 0
    assert b"doser" in handle.mdg_last_error()
    assert handle.mdg_tx_latent_combine(p, p, p, None, None, None, 2, None, None, 4, 8, p, None) == 1   # logsigm needs idx/beta/bias
    assert handle.mdg_tx_latent_combine(p, None, None, None, None, None, 0, p, None, 4, 8, p, None) == 1  # table without index
    assert handle.mdg_tx_latent_combine(p, None, None, None, None, None, 0, None, None, 0, 8, p, None) == 0  # empty batch: no-op
    assert handle.mdg_masked_pool(None, None, 4, 4, 8, 0, None, None) == 1
    assert handle.mdg_masked_pool(p, p, 4, 4, 8, 5, p, None) == 1 and b"bad arguments" in handle.mdg_last_error()
    assert handle.mdg_exact_rank(None, 1, 8, None, None, 0, None) == 1
    assert handle.mdg_exact_rank(p, 1, 100000, p, p, 256, None) == 2 and b"32 bits" in handle.mdg_last_error()
    assert handle.mdg_exact_rank_workspace_bytes(100000) == 0 and handle.mdg_exact_rank_workspace_bytes(4096) > 0
    assert handle.mdg_exact_rank(p, 0, 8, p, None, 0, None) == 0   # no outcomes: no-op
    # prepared decoder / row normalisation (ABI 3)
    assert handle.mdg_pair_prepared_bytes(256, 86, 0) == 86 * 256 * 256 * 2
    assert handle.mdg_pair_prepared_bytes(256, 86, 1) == 2 * 86 * 256 * 256 * 2
    assert handle.mdg_pair_prepare(None, 256, 86, 0, None, 0, None) == 1 and b"NULL" in handle.mdg_last_error()
    assert handle.mdg_pair_prepare(p, 100, 1, 0, p, 256, None) == 2 and b"D=100" in handle.mdg_last_error()
    assert handle.mdg_pair_prepare(p, 128, 1, 0, p, 256, None) == 3   # buffer too small
    assert handle.mdg_pair_score_prepared(p, p, None, 0, 4, 4, 128, 1, 0, 0, 0, 0, None, p, None, 0, None) == 1
    assert handle.mdg_pair_score_prepared(p, p, p, -1, 4, 4, 128, 1, 0, 0, 0, 0, None, p, None, 0, None) == 1
    assert handle.mdg_l2_normalize_rows(None, 4, 8, None, None) == 1
    assert handle.mdg_l2_normalize_rows(p, 4, 0, p, None) == 1
    assert handle.mdg_l2_normalize_rows(p, 0, 8, p, None) == 0     # no rows: no-op
    # per-drug 'mlp' dosers (ABI 4)
    assert handle.mdg_doser_mlp(p, p, 4, 0, 8, 2, p, p, p, p, p, p, p, None) == 1 and b"bad sizes" in handle.mdg_last_error()
    assert handle.mdg_doser_mlp(p, p, 4, 3, 512, 2, p, p, p, p, p, p, p, None) == 2 and b"width" in handle.mdg_last_error()
    assert handle.mdg_doser_mlp(p, p, 4, 3, 8, 2, p, p, None, None, p, p, p, None) == 1   # depth 2 needs the hidden layer
    assert handle.mdg_doser_mlp(None, None, 0, 3, 8, 1, None, None, None, None, None, None, None, None) == 0  # empty batch
    # host half of the packed transfer (ABI 5)
    assert handle.mdg_host_mirror_tiles(None, 1, 32, None, 1) == 1 and b"NULL" in handle.mdg_last_error()
    assert handle.mdg_host_mirror_tiles(p, -1, 32, p, 1) == 1
    assert handle.mdg_host_mirror_tiles(p, 0, 32, p, 1) == 0      # no outcomes: no-op


def test_stale_library_abi_is_rejected(monkeypatch):
    """_lib.lib() refuses a library whose mdg_abi_version() differs from the binding's (structs go by pointer)."""
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setattr(_lib, "EXPECTED_ABI", 999)
    with pytest.raises(RuntimeError, match="ABI version"):
        _lib.lib()
    monkeypatch.setattr(_lib, "EXPECTED_ABI", 5)
    monkeypatch.setattr(_lib, "_LIB", None)
    assert _lib.lib() is not None


def test_non_tx_modalities_is_a_constructor_argument():
    """The reference reads NON_TX_MODALITIES from the environment at import (utils.py:30-37); here it is also a
    constructor argument: token count, pooling key mask, positional-encoding length and src_mask follow it."""
    import madrigal_b200 as mb
    hp = dict(transformer_num_layers=1, transformer_att_heads=2, transformer_head_dim=32, transformer_ffn_dim=64,
              transformer_dropout=0.0, transformer_actn='gelu', transformer_norm_first=True,
              transformer_batch_first=False, transformer_agg='x-attn')
    proj = dict(proj_hidden_dims=[32], proj_dropout=0.0, proj_norm='ln', proj_actn='relu', proj_order='nd')
    for mods, n in ((None, 3), (4, 4), (["str", "kg", "cv", "bs"], 4)):
        enc = mb.FusionEncoder(64, 2, 0.0, hp, proj, fusion='transformer', pos_emb_type='sinusoidal',
                               non_tx_modalities=mods)
        T = n + 16 + 2
        assert enc.transformer.x_attn_key_padding_mask.shape == (1, T)
        assert enc.transformer.x_attn_key_padding_mask[0].tolist() == [True] * n + [False] * 2 + [True] * 16
        assert enc.pos_encoder.pe.shape == (1, T, 64) and enc.pos_emb_max_len == n
        sm = enc._src_mask("cpu")
        assert sm.shape == (T, T) and bool(sm[0, T - 1]) and bool(sm[T - 1, 0]) and not bool(sm[n, 0])
    with pytest.raises(ValueError):
        mb.FusionEncoder(64, 0, 0.0, hp, proj, non_tx_modalities=["kg", "str", "cv"])


@pytest.mark.parametrize("N", [1, 31, 32, 33, 64, 100, 257, 1000])
@pytest.mark.parametrize("threads", [1, 3])
def test_host_mirror_of_packed_tiles_equals_the_layout_definition(N, threads):
    """mdg_host_mirror_tiles (the host half of score_all_pairs_to_host(host_mirror=True): packed lower-triangular rank
    tiles -> the normaliser's [L, N, N] layout, normalize_scores.py:67-70) against `unpack_packed_tiles`, the plain
    statement of MDG_PAIRS_PACKED_TILES: ragged N, garbage beyond N inside the edge tiles, zero diagonal, both the
    streaming-store path (N % 32 == 0) and the plain one, guard bands around the output."""
    import numpy as np
    import torch
    from madrigal_b200 import decoder
    L = 3
    nb = (N + 31) // 32
    T = nb * (nb + 1) // 2
    rng = np.random.default_rng(N)
    full = rng.integers(1, 65535, size=(L, nb * 32, nb * 32), dtype=np.uint16)   # incl. garbage in rows/cols >= N
    packed = np.zeros((L, T, 32, 32), dtype=np.uint16)
    t = 0
    for bi in range(nb):
        for bj in range(bi + 1):
            tile = full[:, 32 * bi:32 * bi + 32, 32 * bj:32 * bj + 32].copy()
            packed[:, t] = np.tril(tile, -1) if bi == bj else tile
            t += 1
    want = decoder.unpack_packed_tiles(packed, N)
    assert np.array_equal(want, np.swapaxes(want, 1, 2)) and not want[:, np.arange(N), np.arange(N)].any()
    guard = 512
    flat = torch.full((L * N * N + 2 * guard,), 0x5A5A, dtype=torch.int16)
    out = flat[guard:guard + L * N * N].view(torch.uint16).view(L, N, N)
    got = decoder.mirror_packed_tiles_host(torch.from_numpy(packed.view(np.int16)).view(torch.uint16), N, out=out,
                                           threads=threads)
    assert got.data_ptr() == out.data_ptr()
    assert np.array_equal(got.view(torch.int16).numpy().view(np.uint16), want)
    assert bool((flat[:guard] == 0x5A5A).all()) and bool((flat[guard + L * N * N:] == 0x5A5A).all())
    with pytest.raises(ValueError):
        decoder.mirror_packed_tiles_host(torch.zeros((L, T + 1, 32, 32), dtype=torch.int16).view(torch.uint16), N)
