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
