"""Model wrapper drop-in (reference: NovelDDIMultilabel, madrigal/models/models.py:914-953)."""
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .decoder import BilinearDDIScorer, Symmetric, l2_normalize_rows, pair_score, pair_score_gather


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

    def _tabular_mods(self, batch_head, batch_tail):
        """Extra tabular modalities forwarded to the encoder as keyword arguments when the token layout has more than
        str / kg / cv (models.py:936-943).  The count comes from the encoder's own layout when it exposes one
        (`non_tx_modalities`), else from the module default (environment variable, as in the reference)."""
        from .constants import NUM_NON_TX_MODALITIES
        enc = self.encoder
        mods = getattr(getattr(enc, 'fusion_encoder', enc), 'non_tx_modalities', None)
        n_non_tx = len(mods) if mods is not None else NUM_NON_TX_MODALITIES
        head, tail = {}, {}
        if n_non_tx > 3:
            for mod in batch_head.keys():
                if mod not in ('drugs', 'strs', 'masks', 'cv', 'tx'):
                    head[mod] = batch_head[mod]
                    tail[mod] = batch_tail[mod]
        return head, tail

    def forward(self, batch_head, batch_tail, batch_head_mod_masks, batch_tail_mod_masks, batch_kg,
                label_range: Optional[Tuple[int, int]] = None, single_drug=False):
        head_extra, tail_extra = self._tabular_mods(batch_head, batch_tail)
        z_head = self.encoder(batch_head['drugs'], batch_head_mod_masks, batch_head['strs'], batch_kg,
                              batch_head['cv'], batch_head['tx'], **head_extra)
        z_tail = self.encoder(batch_tail['drugs'], batch_tail_mod_masks, batch_tail['strs'], batch_kg,
                              batch_tail['cv'], batch_tail['tx'], **tail_extra)
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
        head_extra, tail_extra = self._tabular_mods(batch_head, batch_tail)
        z_head = self.encoder(batch_head['drugs'], batch_head_mod_masks, batch_head['strs'], batch_kg,
                              batch_head['cv'], batch_head['tx'], **head_extra)
        z_tail = self.encoder(batch_tail['drugs'], batch_tail_mod_masks, batch_tail['strs'], batch_kg,
                              batch_tail['cv'], batch_tail['tx'], **tail_extra)
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


class NovelDDIEncoder(nn.Module):
    """The reference encoder's call signature (`NovelDDIEncoder.encode`, models.py:717-900) around `FusionEncoder`.

    The modality encoders are passed in as modules and called exactly as the reference calls them (models.py:720-769):
    the structure / knowledge-graph encoders are the reference's own GNNs (out of scope here), `cv_encoder` and
    `tx_encoder_dict` may be `madrigal_b200.MLPEncoder`s and `tx_encoder` a `madrigal_b200.chemcpa.TxAdaptingComPert`.
    Everything from `all_embeds` on (:772-896: stacking, bottleneck tokens, masks, positional encoding, fusion
    transformer, unimodal bypass) is `fusion_encoder`.  Attribute names follow the reference so that its checkpoints'
    encoder entries line up (`str_encoder.*`, `kg_encoder.*`, `cv_encoder.*`, `tx_encoder.*`; the fusion section's
    parameters live under `fusion_encoder.*`).
    """

    def __init__(self, fusion_encoder, str_encoder, kg_encoder, cv_encoder, tx_encoder=None, tx_encoder_dict=None,
                 tabular_mod_encoders=None, kg_encoder_name: str = 'hgt', use_tx_basal: bool = False,
                 tx_cell_line_onehot_encoder=None):
        super().__init__()
        from .constants import CELL_LINES
        if (tx_encoder is None) == (tx_encoder_dict is None):
            raise ValueError("give exactly one of tx_encoder (chemCPA) and tx_encoder_dict (one encoder per cell line)")
        self.fusion_encoder = fusion_encoder
        self.embed_dim = fusion_encoder.embed_dim
        self.str_encoder, self.kg_encoder, self.cv_encoder = str_encoder, kg_encoder, cv_encoder
        self.kg_encoder_name = kg_encoder_name
        self.tx_encoder = tx_encoder
        self.tx_encoder_dict = tx_encoder_dict if tx_encoder_dict is None or isinstance(tx_encoder_dict, nn.Module) \
            else _as_module_dict(tx_encoder_dict)
        mods = tabular_mod_encoders or {}
        self.tabular_mod_encoders = mods if isinstance(mods, nn.Module) else _as_module_dict(mods)
        self.use_tx_basal = use_tx_basal
        self.tx_cell_line_onehot_encoder = tx_cell_line_onehot_encoder
        self._cell_lines = list(CELL_LINES)
        from .constants import resolve_non_tx
        mods = resolve_non_tx(getattr(fusion_encoder, 'non_tx_modalities', None))
        self._extra_mods = list(mods[3:])  # reference: NON_TX_MODALITIES[3:] (models.py:747)
        if len(self.tabular_mod_encoders) > 0 and sorted(self.tabular_mod_encoders.keys()) != sorted(self._extra_mods):
            raise ValueError(f"tabular_mod_encoders {sorted(self.tabular_mod_encoders.keys())} do not match the fusion "
                             f"encoder's extra non-TX modalities {self._extra_mods}")

    def encode(self, batch_drugs, batch_masks, batch_mols, batch_kg, batch_cv, batch_tx_dict, raw_encoder_output=False,
               **kwargs):
        # structure (models.py:720-721)
        str_out = self.str_encoder(batch_mols, batch_mols.node_feature.float())["graph_feature"]
        # knowledge graph (:724-736)
        kg_data, kg_map = batch_kg['data'], batch_kg['drug_index_map']
        if 'han' in self.kg_encoder_name or 'hgt' in self.kg_encoder_name:
            kg_valid = self.kg_encoder(kg_data.x_dict, kg_data.edge_index_dict)['drug']
        elif 'rgcn' in self.kg_encoder_name:
            kg_valid = self.kg_encoder(kg_data.node_embeddings, kg_data.edge_index, kg_data.node_type,
                                       kg_data.edge_type)[:(kg_data.node_type == 0).sum().item()]
        else:
            raise NotImplementedError(self.kg_encoder_name)
        # drugs outside the KG get a filler row: the reference draws randn (:733), the slot is masked either way
        n_rows = max(int(batch_drugs.max().item()) + 1, int(kg_map.max().item()) + 1)
        kg_out = torch.zeros((n_rows, self.embed_dim), dtype=kg_valid.dtype, device=kg_valid.device)
        kg_out[kg_map.to(kg_valid.device)] = kg_valid
        kg_out = kg_out[batch_drugs.to(kg_valid.device)]
        # every drug whose KG view is marked present must be in the KG (models.py:738)
        bm = batch_masks.to(torch.bool)
        assert torch.isin(batch_drugs[~bm[:, 1]].to(kg_map.device), kg_map).all()
        cv_out = self.cv_encoder(batch_cv)  # :741
        other = []
        if len(self.tabular_mod_encoders) > 0:  # :746-750
            for mod in self._extra_mods:
                if kwargs.get(mod, None) is None:
                    raise AssertionError(f"Missing {mod} in the input batch")
                other.append(self.tabular_mod_encoders[mod](kwargs[mod]))
        if self.tx_encoder_dict is not None:  # :753-754
            tx_out = [self.tx_encoder_dict[c](batch_tx_dict[c]['sigs']) for c in self._cell_lines]
        else:  # chemCPA over all cell lines at once (:756-769)
            import numpy as np
            sigs = torch.cat([batch_tx_dict[c]['sigs'] for c in self._cell_lines], dim=0)
            drugs = torch.cat([batch_tx_dict[c]['drugs'] for c in self._cell_lines], dim=0)
            dosages = torch.cat([batch_tx_dict[c]['dosages'] for c in self._cell_lines], dim=0)
            cells = np.concatenate([batch_tx_dict[c]['cell_lines'] for c in self._cell_lines], axis=0)
            onehot = torch.from_numpy(self.tx_cell_line_onehot_encoder.transform(cells.reshape(-1, 1))).long().to(sigs.device)
            lat = self.tx_encoder.predict(genes=sigs, drugs_idx=drugs, dosages=dosages, covariates=[onehot],
                                          return_latent_basal=self.use_tx_basal,
                                          return_latent_treated=(not self.use_tx_basal))[2]
            tx_out = list(torch.split(lat, lat.shape[0] // len(self._cell_lines), dim=0))
        all_embeds = torch.stack([str_out, kg_out, cv_out] + other + tx_out, dim=1).float().contiguous()  # :772-775
        if raw_encoder_output:  # :889-893
            uni = all_embeds[~batch_masks.to(torch.bool), :]
            if self.fusion_encoder.normalize:
                uni = l2_normalize_rows(uni.contiguous())
            return self.fusion_encoder.uni_projector(uni.contiguous())
        return self.fusion_encoder(all_embeds, batch_masks)

    def forward(self, batch_drugs, batch_masks, batch_mols, batch_kg, batch_cv, batch_tx_dict, raw_encoder_output=False,
                **kwargs):
        return self.encode(batch_drugs, batch_masks, batch_mols, batch_kg, batch_cv, batch_tx_dict, raw_encoder_output,
                           **kwargs)


class _Callable(nn.Module):
    """Wraps a plain callable so that it can sit in an nn.ModuleDict next to real modules."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, *a, **k):
        return self.fn(*a, **k)


def _as_module_dict(d):
    return nn.ModuleDict({k: (v if isinstance(v, nn.Module) else _Callable(v)) for k, v in d.items()})
