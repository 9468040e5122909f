"""CPU (gloo, world_size 2): the multi-GPU partition logic of SURVEY §8e — outcome/row shards and the single
all-gather of the fused-embedding table.  No CUDA compute: the decoder step is replaced by the oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from madrigal_b200 import scoring


def test_outcome_shards_partition_exactly():
    for L in (1, 7, 86, 953, 963):
        for world in (1, 2, 3, 4, 8):
            spans = [scoring.outcome_shard(L, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == L
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, N, D, L, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synth
    from oracle import oracle
    z_full, W = synth.decoder_inputs(N, D, L, seed=1)
    # 1. each rank "encodes" its row shard, 2. ONE all-gather replicates z, 3. each rank scores its outcome shard
    r0, r1 = scoring.row_shard(N, rank, world)
    z = scoring.all_gather_embeddings(torch.from_numpy(z_full[r0:r1].copy()), N)
    assert torch.equal(z, torch.from_numpy(z_full))
    # the exchange-step object the scoring drivers use: on CPU tensors it reports the collective fallback and gives the
    # same table (its peer-push mode needs CUDA peer mappings and is covered on the GPUs); equal shards reuse `out`
    g = scoring.PeerAllGather(N, D, torch.device("cpu"))
    assert g.mode == "collective" and (g.r0, g.r1) == (r0, r1)
    assert torch.equal(g.gather(torch.from_numpy(z_full[r0:r1].copy())), torch.from_numpy(z_full))
    if N % world == 0:
        buf = torch.empty((N, D))
        z2 = scoring.all_gather_embeddings(torch.from_numpy(z_full[r0:r1].copy()), N, out=buf)
        assert z2.data_ptr() == buf.data_ptr() and torch.equal(z2, torch.from_numpy(z_full))
    l0, l1 = scoring.outcome_shard(L, rank, world)
    part = oracle.bilinear_scores(z.numpy(), z.numpy(), W, (l0, l1))
    np.save(os.path.join(tmp, f"part{rank}.npy"), part)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_scoring_matches_single_rank(tmp_path):
    N, D, L, world = 37, 16, 5, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, N, D, L, str(tmp_path)), nprocs=world, join=True)
    import synth
    from oracle import oracle
    z, W = synth.decoder_inputs(N, D, L, seed=1)
    full = oracle.bilinear_scores(z, z, W)
    got = np.concatenate([np.load(tmp_path / f"part{r}.npy") for r in range(world)], axis=0)
    assert np.array_equal(got, full)  # sharding does not change any arithmetic


def test_two_rank_equal_shards_gather_in_place(tmp_path):
    N, D, L, world = 40, 16, 4, 2   # N divisible by the world size: all_gather_into_tensor straight into the table
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, N, D, L, str(tmp_path)), nprocs=world, join=True)


def test_novel_ddi_encoder_dropin_plumbing_cpu():
    """Host logic of madrigal_b200.NovelDDIEncoder.encode (models.py:717-775, 889-893) with stub modality encoders and
    a stub fusion section: token order [str, kg, cv, tx...], KG rows behind a permuted drug_index_map, per-cell-line
    encoders vs the chemCPA call (one predict over all cell lines, split back), raw_encoder_output."""
    import numpy as np
    import torch
    import torch.nn as nn
    import madrigal_b200 as mb
    from madrigal_b200.constants import CELL_LINES

    class FakeFusion(nn.Module):
        embed_dim, normalize = 8, False  # normalisation is a CUDA kernel (mdg_l2_normalize_rows): GPU tests cover it

        def __init__(self):
            super().__init__()
            self.uni_projector = nn.Identity()

        def forward(self, e, m):
            return (e * (~m)[:, :, None]).sum(1)

    B, E = 5, 8
    rng = np.random.default_rng(0)
    embeds = torch.from_numpy(rng.standard_normal((B, 19, E)).astype(np.float32))
    masks = torch.from_numpy(rng.random((B, 19)) < 0.5)
    masks[:, 0] = False
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))

    class Mols:
        node_feature = torch.zeros(1)

    class KG:
        x_dict = None
        edge_index_dict = None

    common = dict(str_encoder=lambda m, f: {"graph_feature": embeds[:, 0]},
                  kg_encoder=lambda x, e: {"drug": embeds[:, 1][perm]}, cv_encoder=lambda cv: embeds[:, 2])
    kg = {"data": KG(), "drug_index_map": perm}
    enc = mb.NovelDDIEncoder(FakeFusion(), tx_encoder_dict={c: (lambda s: s) for c in CELL_LINES}, **common)
    tx = {c: {"sigs": embeds[:, 3 + i]} for i, c in enumerate(CELL_LINES)}
    z = enc(torch.arange(B), masks, Mols(), kg, None, tx)
    assert torch.allclose(z, FakeFusion()(embeds, masks))
    raw = enc.encode(torch.arange(B), masks, Mols(), kg, None, tx, raw_encoder_output=True)
    assert raw.shape == (int((~masks).sum()), E)

    class StubTx(nn.Module):
        def predict(self, genes, drugs_idx, dosages, covariates, return_latent_basal, return_latent_treated):
            assert covariates[0].shape == (genes.shape[0], len(CELL_LINES))
            assert (return_latent_basal, return_latent_treated) == (False, True)
            return None, None, genes + covariates[0].argmax(1)[:, None].float()

    class OneHot:
        def transform(self, a):
            return np.eye(len(CELL_LINES))[[CELL_LINES.index(x) for x in a[:, 0]]]

    tx2 = {c: {"sigs": embeds[:, 3 + i], "drugs": torch.arange(B), "dosages": torch.ones(B),
               "cell_lines": np.array([c] * B)} for i, c in enumerate(CELL_LINES)}
    enc2 = mb.NovelDDIEncoder(FakeFusion(), tx_encoder=StubTx(), tx_cell_line_onehot_encoder=OneHot(), **common)
    want = embeds.clone()
    for i in range(len(CELL_LINES)):
        want[:, 3 + i] += i
    assert torch.allclose(enc2(torch.arange(B), masks, Mols(), kg, None, tx2), FakeFusion()(want, masks))
    import pytest
    with pytest.raises(ValueError):
        mb.NovelDDIEncoder(FakeFusion(), **common)


def test_bind_host_thread_to_gpu_is_a_noop_without_nvml_affinity():
    """scoring.bind_host_thread_to_gpu: on a host without NVML / NUMA affinity information nothing changes."""
    import os
    from madrigal_b200 import scoring
    before = set(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None
    res = scoring.bind_host_thread_to_gpu(0)
    if res is None and before is not None:
        assert set(os.sched_getaffinity(0)) == before
    elif res is not None:  # a host with affinity information: the new set is a subset of the old one; restore
        assert res[1] <= res[0]
        os.sched_setaffinity(0, res[0])


def test_encoder_plan_replicates_single_wave_catalogues():
    """scoring.encoder_is_replicated: one wave of 128-row encoder tiles (<= 148 SMs) -> every rank encodes the whole
    catalogue and the step has no exchange; more tiles -> row shards + the one all-gather of z."""
    from madrigal_b200 import scoring
    assert scoring.encoder_is_replicated(4096, 4, 8)          # BASELINE configs[2]: 128 tiles
    assert scoring.encoder_is_replicated(4736, 4, 2)          # exactly 148 tiles
    assert not scoring.encoder_is_replicated(4737, 4, 2)
    assert not scoring.encoder_is_replicated(20000, 4, 8)     # configs[3]: 625 tiles, sharded
    assert not scoring.encoder_is_replicated(4096, 23, 4)     # production T = 23: 5 drugs per tile
    assert scoring.encoder_is_replicated(10 ** 6, 4, 1)       # one rank: nothing to exchange
