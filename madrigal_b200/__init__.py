"""madrigal_b200 — B200-native (sm_100a) drug-pair scoring path of Madrigal behind the reference's call signatures.

Only the path is here: fusion encoder -> bilinear decoder -> rank normalisation.  Every operator
calls the C ABI in include/madrigal_b200.h through ctypes; nothing falls back to PyTorch or the CPU.
"""
from .decoder import (BilinearDDIScorer, PreparedDecoder, RankTable, Symmetric, ensemble_reduce, pair_score,  # noqa: F401
                      pair_score_gather, pair_topk)
from .fusion import (FusionEncoder, MLPAdaptor, MLPEncoder, PositionEncodingLearnable, PositionEncodingSinusoidal,  # noqa: F401
                     TransformerFusion, masked_pool)
from .model import NovelDDIEncoder, NovelDDIMultilabel, PrecomputedEmbeddingEncoder  # noqa: F401
from . import chemcpa  # noqa: F401  (tx modality encoder: chemcpa.TxAdaptingComPert, chemcpa.MLP)
from . import ops  # noqa: F401  (registers torch.ops.madrigal_b200.*)

__all__ = ["BilinearDDIScorer", "PreparedDecoder", "RankTable", "Symmetric", "pair_score", "pair_topk", "pair_score_gather", "ensemble_reduce", "TransformerFusion", "MLPAdaptor", "MLPEncoder",
           "FusionEncoder", "PositionEncodingSinusoidal", "PositionEncodingLearnable", "masked_pool",
           "NovelDDIEncoder", "NovelDDIMultilabel", "PrecomputedEmbeddingEncoder", "chemcpa"]
