"""Model wrapper drop-in (reference: NovelDDIMultilabel, madrigal/models/models.py:914-953)."""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .decoder import BilinearDDIScorer, Symmetric, pair_score, pair_score_gather


class NovelDDIMultilabel(nn.Module):
    """Same constructor and `forward` signature as the reference.  `encoder` is any module with the reference
    encoder's call signature `(drugs, masks, mols, kg, cv, tx_dict, **tabular_mods) -> z [N, feat_dim]`
    (models.py:945-946) — e.g. the reference's own NovelDDIEncoder with its `encode` fusion section replaced by
    `madrigal_b200.FusionEncoder`, or `PrecomputedEmbeddingEncoder` below.  The decoder weight is registered through
    the `Symmetric` parametrisation exactly as models.py:921-922, so the state_dict keys are
    `decoder.parametrizations.weight.original` / `decoder.bias`.
    """

    def __init__(self, encoder, feat_dim, prediction_dim, prediction_dim_single_drug=None, normalize=False,
                 use_single_drug=False, precision: str = "fp32"):
        super().__init__()
        self.encoder = encoder
        self.embed_dim = feat_dim
        self.normalize = normalize
        self.use_single_drug = use_single_drug
        self.decoder = BilinearDDIScorer(feat_dim, feat_dim, prediction_dim, precision=precision)
        nn.utils.parametrize.register_parametrization(self.decoder, 'weight', Symmetric())

    def forward(self, batch_head, batch_tail, batch_head_mod_masks, batch_tail_mod_masks, batch_kg,
                label_range: Optional[Tuple[int, int]] = None, single_drug=False):
        z_head = self.encoder(batch_head['drugs'], batch_head_mod_masks, batch_head['strs'], batch_kg,
                              batch_head['cv'], batch_head['tx'])
        z_tail = self.encoder(batch_tail['drugs'], batch_tail_mod_masks, batch_tail['strs'], batch_kg,
                              batch_tail['cv'], batch_tail['tx'])
        weight = self.decoder.weight
        if label_range is not None:
            assert len(label_range) == 2
            weight = weight[label_range[0]:label_range[1]]
        # F.normalize of both embedding tables (models.py:947-949) is fused into the decoder's operand preparation
        return pair_score(z_head, z_tail, weight, precision=self.decoder.precision, out="logit",
                          normalize=bool(self.normalize))


    def forward_triples(self, batch_head, batch_tail, batch_head_mod_masks, batch_tail_mod_masks, batch_kg, ddi_labels,
                        head_idx, tail_idx, sigmoid: bool = True):
        """`sigmoid(model(...))[ddi_labels, head_idx, tail_idx]` (train_ddi_batch.py:285-286, evaluate.py:191-195)
        without the dense [L, Nh, Nt] tensor the reference materialises first."""
        z_head = self.encoder(batch_head['drugs'], batch_head_mod_masks, batch_head['strs'], batch_kg,
                              batch_head['cv'], batch_head['tx'])
        z_tail = self.encoder(batch_tail['drugs'], batch_tail_mod_masks, batch_tail['strs'], batch_kg,
                              batch_tail['cv'], batch_tail['tx'])
        return pair_score_gather(z_head, z_tail, self.decoder.weight, ddi_labels, head_idx, tail_idx,
                                 precision=self.decoder.precision, out="sigmoid" if sigmoid else "logit",
                                 normalize=bool(self.normalize))


class PrecomputedEmbeddingEncoder(nn.Module):
    """Adapter with the reference encoder's call signature around `FusionEncoder`, for callers that already hold the
    stacked modality embeddings: `batch_tx_dict` carries them under the key 'all_embeds' ([B, 19, E])."""

    def __init__(self, fusion_encoder):
        super().__init__()
        self.fusion_encoder = fusion_encoder

    def forward(self, batch_drugs, batch_masks, batch_mols, batch_kg, batch_cv, batch_tx_dict, **kwargs):
        return self.fusion_encoder(batch_tx_dict['all_embeds'], batch_masks)
