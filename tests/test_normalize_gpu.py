"""GPU parity: exact in-sample normalised rank (mdg_exact_rank) and the reference-quantile builder vs the oracle and
the committed outputs of the reference's own run_slice (notebooks/normalize_scores.py:62-74).  Integer rank work:
bit-exact float32 outputs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
META = json.load(open(os.path.join(HERE, "golden", "golden_meta.json")))


@pytest.fixture(scope="module")
def norm(cuda_device):
    from madrigal_b200 import normalize
    return normalize


@pytest.mark.parametrize("case", META["normalizer"], ids=lambda c: c["name"])
def test_exact_rank_vs_reference_golden(norm, cuda_device, case):
    g = np.load(os.path.join(HERE, "golden", "golden_normalizer.npz"))
    rng = np.random.default_rng(case["seed"])
    raw = rng.standard_normal((case["L"], case["N"], case["N"])).astype(np.float32)
    if case["ties"]:
        raw = np.round(raw * 4) / 4
    got = norm.exact_normalized_ranks(torch.from_numpy(raw).to(cuda_device)).cpu().numpy()
    ref = g[f"{case['name']}.out"]
    if not case["ties"]:
        assert np.array_equal(got, ref)  # bit-identical to the reference's float32 memmap contents
    else:
        assert np.array_equal(got, oracle.normalize_scores(raw, kind="stable"))
        N, M = case["N"], case["N"] * (case["N"] - 1) // 2
        i, j = np.tril_indices(N, -1)
        for l in range(case["L"]):
            v = np.sort(oracle.lower_triangle_values(raw[l]))
            r = np.rint(got[l][i, j].astype(np.float64) * M).astype(np.int64)
            assert (r >= np.searchsorted(v, raw[l][i, j], "left") + 1).all()
            assert (r <= np.searchsorted(v, raw[l][i, j], "right")).all()


@pytest.mark.parametrize("L,N", [(2, 257), (1, 2), (3, 1), (1, 700), (2, 3001)])  # 3001: grouped-placement path
def test_exact_rank_vs_oracle(norm, cuda_device, L, N):
    rng = np.random.default_rng(N)
    raw = rng.standard_normal((L, N, N)).astype(np.float32)
    raw[:, 0, :] = 0.0  # some exact ties incl. -0.0
    raw[:, :, 0] *= -0.0 if N > 2 else 1.0
    got = norm.exact_normalized_ranks(torch.from_numpy(raw).to(cuda_device)).cpu().numpy()
    ref = oracle.normalize_scores(raw, kind="stable")
    assert got.shape == ref.shape and np.array_equal(got, ref)
    assert np.array_equal(got, got.swapaxes(1, 2)) and (np.diagonal(got, axis1=1, axis2=2) == 0).all()


def test_quantile_builder_vs_oracle(norm, cuda_device):
    rng = np.random.default_rng(4)
    raw = rng.standard_normal((3, 200, 200)).astype(np.float32)
    for Q in (1, 17, 4096, 19900):
        got = norm.lower_triangle_quantiles(torch.from_numpy(raw).to(cuda_device), Q).cpu().numpy()
        assert np.array_equal(got, oracle.reference_quantiles(raw, Q))


def test_fused_quantile_rank_tracks_exact_rank_within_one_over_q(norm, cuda_device):
    """End-to-end link between the two formulations: |fused rank / Q - reference normalised rank| <= 1/Q + snapping."""
    import madrigal_b200 as mb
    import synth
    N, D, L, Q = 300, 128, 2, 4096
    z, W = synth.decoder_inputs(N, D, L, seed=21)
    zt, Wt = torch.from_numpy(z).to(cuda_device), torch.from_numpy(W).to(cuda_device)
    logits = mb.pair_score(zt, zt, Wt, precision="fp32")
    exact = norm.exact_normalized_ranks(logits).cpu().numpy()
    table = norm.build_rank_table(zt, Wt, Q, precision="fp32")
    fused = mb.pair_score(zt, zt, Wt, precision="fp32", out="rank", table=table).cpu().numpy().astype(np.float64) / Q
    i, j = np.tril_indices(N, -1)
    for l in range(L):
        assert np.abs(fused[l][i, j] - exact[l][i, j]).max() <= 1.0 / Q + 4.0 / 131072 + 1e-7


def test_gmean_ensemble_and_rerank_vs_oracle(norm, cuda_device):
    """generate_embeddings.ipynb cells 18 + 20: geometric mean of 5 checkpoints' normalised ranks, then run_slice again."""
    rng = np.random.default_rng(9)
    K, L, N = 5, 2, 120
    members = [oracle.normalize_scores(rng.standard_normal((L, N, N)).astype(np.float32)) for _ in range(K)]
    g_ref = oracle.gmean_normalized_ranks(members)
    dev_members = [torch.from_numpy(m).to(cuda_device) for m in members]
    g = norm.gmean_normalized_ranks(dev_members).cpu().numpy()
    assert (np.diagonal(g, axis1=1, axis2=2) == 0).all()                       # log(0) = -inf -> exp = 0, as scipy
    assert np.abs(g - g_ref).max() <= 4e-6 * g_ref.max()                        # logf/expf vs numpy's float32 log/exp
    # re-normalisation: bit-identical to the reference normaliser applied to the SAME gmean values (stable ties)
    final = norm.ensemble_normalized_ranks(dev_members).cpu().numpy()
    assert np.array_equal(final, oracle.normalize_scores(g, kind="stable"))
    # and within one rank step of the all-numpy chain wherever its gmean values are not nearly tied
    ref_final = oracle.ensemble_normalized_ranks(members, kind="stable")
    M = N * (N - 1) // 2
    assert np.abs(final - ref_final).max() <= 3.0 / M
    # uint16 fused-rank members: gmean of rank / Q
    Q = 4096
    ranks = [torch.from_numpy(rng.integers(0, Q + 1, size=(L, N, N)).astype(np.uint16)).to(cuda_device) for _ in range(3)]
    gu = norm.gmean_normalized_ranks(ranks, Q).cpu().numpy()
    ref_u = oracle.gmean_normalized_ranks([r.cpu().numpy().astype(np.float32) * np.float32(1.0 / Q) for r in ranks])
    assert np.abs(gu - ref_u).max() <= 4e-6


def test_mean_sigmoid_ensemble_vs_oracle(norm, cuda_device):
    import madrigal_b200 as mb
    from madrigal_b200 import scoring
    import synth
    zs, Ws, lgs = [], [], []
    for k in range(3):
        z, W = synth.decoder_inputs(96, 64, 4, seed=40 + k)
        zs.append(torch.from_numpy(z).to(cuda_device)); Ws.append(torch.from_numpy(W).to(cuda_device))
        lgs.append(oracle.bilinear_scores(z, z, W, (1, 3)))
    got = scoring.ensemble_mean_sigmoid(zs, Ws, precision="fp32", label_range=(1, 3)).cpu().numpy()
    assert np.abs(got - oracle.ensemble_mean_sigmoid(lgs)).max() <= 2e-4


def test_streamed_ensemble_normalisation_vs_oracle_chain(norm, cuda_device):
    """scoring.ensemble_normalized_ranks_chunks == per-checkpoint scores -> normaliser -> gmean -> normaliser, against
    the all-numpy chain (within rank steps where logits / gmeans are nearly tied) and bit-identical to the library's
    own materialised chain."""
    from madrigal_b200 import scoring
    import synth
    K, N, D, L = 3, 150, 64, 5
    zs, Ws, members = [], [], []
    for k in range(K):
        z, W = synth.decoder_inputs(N, D, L, seed=70 + k)
        zs.append(torch.from_numpy(z).to(cuda_device)); Ws.append(torch.from_numpy(W).to(cuda_device))
        members.append(oracle.normalize_scores(oracle.bilinear_scores(z, z, W), kind="stable"))
    ref = oracle.ensemble_normalized_ranks(members, kind="stable")
    got = np.zeros_like(ref)
    seen = []
    for l0, l1, r in scoring.ensemble_normalized_ranks_chunks(zs, Ws, precision="fp32", chunk=2):
        got[l0:l1] = r.cpu().numpy()
        seen.append((l0, l1))
    assert seen == [(0, 2), (2, 4), (4, 5)]
    M = N * (N - 1) // 2
    assert np.array_equal(got, got.transpose(0, 2, 1)) and (np.diagonal(got, axis1=1, axis2=2) == 0).all()
    # 1e-3-accurate logits can swap near-tied neighbours in a member's ranking: a few rank steps after the gmean
    assert np.abs(got - ref).max() <= 12.0 / M and np.abs(got - ref).mean() <= 0.5 / M
    import madrigal_b200 as mb
    mats = [norm.exact_normalized_ranks(mb.pair_score(z, z, W, precision="fp32", out="logit")) for z, W in zip(zs, Ws)]
    assert np.array_equal(got, norm.ensemble_normalized_ranks(mats).cpu().numpy())


def test_config1_end_to_end_vs_oracle_chain(norm, cuda_device):
    """SURVEY 4, integration: BASELINE config 1 (1,024 drugs, 4 modality tokens, 86 outcomes) through the whole path —
    fusion encoder -> bilinear decoder -> normalised ranks — against the CPU oracle chain on the same seeded weights.
    Outcomes are sub-sampled (the CPU normaliser takes ~0.2 s per outcome at this size)."""
    import madrigal_b200 as mb
    import synth
    N, T, E, L = 1024, 4, 128, 86
    cfg = dict(embed_dim=E, num_layers=2, num_heads=8, head_dim=32, ffn_dim=512, actn="gelu", norm_first=True,
               agg="mean", nb=0)
    sd = synth.fusion_state_dict(cfg, seed=11)
    tokens, masks = synth.fusion_inputs(N, T, E, seed=11)
    _, W = synth.decoder_inputs(1, E, L, seed=12)
    pick = [0, 17, 42, 85]
    # CPU oracle chain (float64 encoder/decoder so that only the GPU's own error is measured, fp32 normaliser as the reference)
    z_ref = oracle.fusion_forward(sd, cfg, tokens, masks, dtype=np.float64)
    lg_ref = np.stack([oracle.bilinear_scores(z_ref, z_ref, W, (l, l + 1), dtype=np.float64)[0] for l in pick]).astype(np.float32)
    rk_ref = oracle.normalize_scores(lg_ref, kind="stable")
    M = N * (N - 1) // 2
    i, j = np.tril_indices(N, -1)
    for precision, tol_logit, tol_rank in (("fp32", 1e-3, 2e-3), ("bf16", 3e-2, 3e-2)):
        enc = mb.TransformerFusion(E, 0, 2, 8, 32, 512, transformer_actn="gelu", transformer_norm_first=True,
                                   transformer_batch_first=False, transformer_agg="mean", precision=precision)
        enc.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        enc = enc.to(cuda_device).eval()
        with torch.no_grad():
            z = enc(torch.from_numpy(tokens).to(cuda_device), torch.from_numpy(masks).to(cuda_device))
        Wt = torch.from_numpy(W[pick]).to(cuda_device)
        logits = mb.pair_score(z, z, Wt, precision=precision, out="logit")
        ranks = norm.exact_normalized_ranks(logits).cpu().numpy()
        lg = logits.cpu().numpy()
        rms = np.sqrt(np.mean(lg_ref.astype(np.float64) ** 2))
        if precision == "fp32":   # 1e-3 bar against max(|ref|, rms) element-wise, loosened to the scale of the tensor
            assert np.abs(lg - lg_ref).max() <= tol_logit * max(rms, np.abs(lg_ref).max() * 0.1), precision
        else:                     # bf16 operands through 2 transformer layers AND the quadratic form: 1e-2 of max|ref|
            assert np.abs(lg - lg_ref).max() <= 1e-2 * np.abs(lg_ref).max(), precision
            assert np.abs(lg - lg_ref).mean() <= 1e-2 * rms, precision
        # the normaliser output has the reference's structure exactly ...
        assert np.array_equal(ranks, ranks.swapaxes(1, 2)) and (np.diagonal(ranks, axis1=1, axis2=2) == 0).all()
        for a in range(len(pick)):
            assert np.array_equal(np.sort(ranks[a][i, j]), (np.arange(1, M + 1) / M).astype(np.float32))  # a permutation of 1..M
            # ... and its values track the reference chain: a logit error e moves a rank by ~ density * e
            d = np.abs(ranks[a][i, j].astype(np.float64) - rk_ref[a][i, j])
            assert d.max() <= tol_rank and d.mean() <= tol_rank / 10, (precision, pick[a], d.max(), d.mean())
        # fused quantile ranks from the same logits: within (1 + snapping)/Q of the exact in-sample ranks
        if precision == "bf16":
            table = norm.build_rank_table(z, Wt, 16384, precision="bf16")
            fused = mb.pair_score(z, z, Wt, precision="bf16", out="rank", table=table, symmetric=True).cpu().numpy()
            for a in range(len(pick)):
                assert np.abs(fused[a][i, j] / 16384.0 - ranks[a][i, j]).max() <= 1.0 / 16384 + 4.0 / 131072 + 1e-7


def test_quantile_formulation_ensemble_is_bit_exact_and_tracks_the_reference(norm, cuda_device):
    """mdg_ensemble_rank_u16 (fixed-point log-sum of the members' fused ranks + ensemble-table lookup):
    (1) bit-exact against oracle.ensemble_quantile_ranks on the same member ranks / ilog table / thresholds;
    (2) the builder mode's float32 log-sums equal the integer sums;
    (3) with full-sample tables it tracks the reference's gmean + re-rank (oracle.ensemble_normalized_ranks on the exact
        in-sample ranks of the same logits) within 2e-3 of normalised rank: K * 0.5 / 2048 of log2 resolution moves a
        gmean by < 0.06 %, the quantile tables add (1 + snapping) / Q per lookup."""
    import madrigal_b200 as mb
    import synth
    from madrigal_b200 import scoring
    N, D, L, K = 256, 128, 3, 4
    M = N * (N - 1) // 2
    zs, Ws, tables, members, logits = [], [], [], [], []
    for k in range(K):
        z, W = synth.decoder_inputs(N, D, L, seed=50 + k)
        zt, Wt = torch.from_numpy(z).to(cuda_device), torch.from_numpy(W).to(cuda_device)
        tbl = norm.build_rank_table(zt, Wt, M, precision="bf16")          # Q = M: every pair of the sample is a threshold
        zs.append(zt), Ws.append(Wt), tables.append(tbl)
        members.append(mb.pair_score(zt, zt, Wt, precision="bf16", out="rank", table=tbl, symmetric=True))
        logits.append(mb.pair_score(zt, zt, Wt, precision="bf16", out="logit").cpu().numpy())
    Qm = tables[0].Q
    ens = norm.build_ensemble_rank_table(members, Qm, Q=M)
    got = norm.ensemble_fused_ranks(members, ens).cpu().numpy()
    mem_np = [m.cpu().numpy() for m in members]
    # (1)
    exp = oracle.ensemble_quantile_ranks(mem_np, oracle.ilog_table(Qm), ens.table.thresholds.cpu().numpy())
    assert np.array_equal(got, exp)
    assert np.array_equal(ens.ilog_host, oracle.ilog_table(Qm))
    # (2)
    g = norm.ensemble_logsum(members, Qm).cpu().numpy()
    gi = sum(oracle.ilog_table(Qm)[m.astype(np.int64)].astype(np.int64) for m in mem_np)
    assert np.array_equal(g, gi.astype(np.float32))
    # (3)
    ref_members = [oracle.normalize_scores(lg, kind="stable") for lg in logits]
    ref = oracle.ensemble_normalized_ranks(ref_members, kind="stable")
    i, j = np.tril_indices(N, -1)
    d = np.abs(got[:, i, j] / float(ens.Q) - ref[:, i, j])
    assert d.max() <= 2e-3 and d.mean() <= 3e-4, (d.max(), d.mean())
    assert (np.diagonal(got, axis1=1, axis2=2) == 0).all() and np.array_equal(got, got.swapaxes(1, 2))
    # streamed driver == one-shot
    outs = np.concatenate([r.cpu().numpy().copy() for _, _, r in
                           scoring.ensemble_fused_ranks_chunks(zs, Ws, tables, ens, precision="bf16", chunk=2)])
    assert np.array_equal(outs, got)


def test_ensemble_kernel_writes_inside_its_output_only(norm, cuda_device):
    """Guard bands around mdg_ensemble_rank_u16's output (n not a multiple of 8 for L = 1: the scalar tail path)."""
    import madrigal_b200 as mb
    import synth
    N, D, K, Q = 123, 64, 3, 256
    members, tables = [], []
    for k in range(K):
        z, W = synth.decoder_inputs(N, D, 1, seed=70 + k)
        zt, Wt = torch.from_numpy(z).to(cuda_device), torch.from_numpy(W).to(cuda_device)
        tbl = norm.build_rank_table(zt, Wt, Q, precision="bf16")
        members.append(mb.pair_score(zt, zt, Wt, precision="bf16", out="rank", table=tbl, symmetric=True))
    ens = norm.build_ensemble_rank_table(members, Q, Q=Q)
    n = N * N
    assert n % 8 != 0
    GUARD = 1024
    flat = torch.full((n + 2 * GUARD,), -12345, dtype=torch.int16, device=cuda_device)
    out = flat[GUARD:GUARD + n].view(torch.uint16).view(1, N, N)
    got = norm.ensemble_fused_ranks(members, ens, out=out)
    torch.cuda.synchronize()
    assert bool((flat[:GUARD] == -12345).all()) and bool((flat[-GUARD:] == -12345).all())
    exp = oracle.ensemble_quantile_ranks([m.cpu().numpy() for m in members], oracle.ilog_table(Q),
                                         ens.table.thresholds.cpu().numpy())
    assert np.array_equal(got.cpu().numpy(), exp)
