"""Fusion-encoder drop-ins (reference: madrigal/models/models.py:351-518, 551-603, 772-865) over the C ABI.

Every class keeps the reference's constructor arguments and parameter names, so a reference `state_dict` loads
unchanged.  torch.nn modules are used ONLY as parameter containers with the reference's key layout —
their `forward` is never called: all arithmetic runs in mdg_fusion_encode / mdg_mlp_forward / mdg_assemble_tokens
(tcgen05 GEMMs + fused glue kernels).  There is no PyTorch fallback; CPU tensors raise.
"""
import ctypes
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import MdgFusionCfg, MdgFusionWeights, MdgMlp
from .constants import CELL_LINES, resolve_non_tx
from .decoder import _PRECISION, _require_cuda_f32, _stream_ptr, _workspace, l2_normalize_rows


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _mask_u8(mask: torch.Tensor, name: str) -> torch.Tensor:
    if not mask.is_cuda:
        raise RuntimeError(f"madrigal_b200: `{name}` must be a CUDA tensor (no CPU path exists)")
    return mask.to(torch.uint8).contiguous()


class TransformerFusion(nn.Module):
    """Drop-in for the reference `TransformerFusion` (models.py:352-455): same ctor, state_dict keys and
    `forward(fusion_sequence [B,T,E], fusion_mask [B,T] bool True=missing, src_mask [T,T] bool True=blocked) -> [B,E]`.

    `transformer_batch_first` is accepted and ignored: the input is always [B,T,E] and the result is what the reference
    computes with batch_first=False (its batch_first=True x-attn path returns only the first drug, models.py:443).
    """

    def __init__(self, embed_dim, num_tx_bottlenecks, transformer_num_layers, transformer_att_heads,
                 transformer_head_dim, transformer_ffn_dim, transformer_dropout=0.1, transformer_actn='relu',
                 transformer_norm_first=False, transformer_batch_first=True, transformer_agg='mean',
                 precision: str = "fp32", non_tx_modalities=None):
        super().__init__()
        n_non_tx = len(resolve_non_tx(non_tx_modalities))  # reference: module-level NUM_NON_TX_MODALITIES (utils.py:36)
        if transformer_actn not in _lib.MDG_ACTN:
            raise NotImplementedError(f"transformer_actn={transformer_actn!r} (supported: relu, gelu)")
        if transformer_agg not in _lib.MDG_AGG:
            raise NotImplementedError(transformer_agg)  # models.py:453
        self.embed_dim = embed_dim
        self.batch_first = transformer_batch_first
        self.norm_first = transformer_norm_first
        self.num_heads = transformer_att_heads
        self.head_dim = transformer_head_dim
        self.ffn_dim = transformer_ffn_dim
        self.actn = transformer_actn
        self.latent_dim = transformer_head_dim * transformer_att_heads
        self.transformer_agg = transformer_agg
        self.precision = precision
        # parameter containers with the reference's names (never called)
        self.embed2latent = nn.Linear(embed_dim, self.latent_dim)
        layer = nn.TransformerEncoderLayer(d_model=self.latent_dim, nhead=transformer_att_heads,
                                           dim_feedforward=transformer_ffn_dim, dropout=transformer_dropout,
                                           activation=transformer_actn, norm_first=transformer_norm_first,
                                           batch_first=False)
        self.transformer_encoder = nn.TransformerEncoder(layer, num_layers=transformer_num_layers,
                                                         enable_nested_tensor=False)
        self.latent2embed = nn.Linear(self.latent_dim, embed_dim)
        if transformer_agg == 'x-attn':
            self.x_attn_kv_norm = nn.LayerNorm(self.latent_dim)
            self.x_attn_query_norm = nn.LayerNorm(self.latent_dim)
            self.x_attn_mha_layer = nn.MultiheadAttention(embed_dim=self.latent_dim, num_heads=transformer_att_heads,
                                                          dropout=transformer_dropout, batch_first=False)
            self.x_attn_query = nn.Parameter(torch.randn(1, self.latent_dim))
            # constant pooling key mask (models.py:382-385): with bottlenecks only they are visible
            m = torch.zeros(1, n_non_tx + len(CELL_LINES) + num_tx_bottlenecks, dtype=torch.bool)
            if num_tx_bottlenecks > 0:
                m[:, :n_non_tx] = True
                m[:, -len(CELL_LINES):] = True
            self.x_attn_key_padding_mask = m

    def _weights(self) -> MdgFusionWeights:
        w = MdgFusionWeights()
        w.embed2latent_weight, w.embed2latent_bias = _ptr(self.embed2latent.weight), _ptr(self.embed2latent.bias)
        w.latent2embed_weight, w.latent2embed_bias = _ptr(self.latent2embed.weight), _ptr(self.latent2embed.bias)
        for i, layer in enumerate(self.transformer_encoder.layers):
            L = w.layers[i]
            L.in_proj_weight, L.in_proj_bias = _ptr(layer.self_attn.in_proj_weight), _ptr(layer.self_attn.in_proj_bias)
            L.out_proj_weight = _ptr(layer.self_attn.out_proj.weight)
            L.out_proj_bias = _ptr(layer.self_attn.out_proj.bias)
            L.linear1_weight, L.linear1_bias = _ptr(layer.linear1.weight), _ptr(layer.linear1.bias)
            L.linear2_weight, L.linear2_bias = _ptr(layer.linear2.weight), _ptr(layer.linear2.bias)
            L.norm1_weight, L.norm1_bias = _ptr(layer.norm1.weight), _ptr(layer.norm1.bias)
            L.norm2_weight, L.norm2_bias = _ptr(layer.norm2.weight), _ptr(layer.norm2.bias)
        if self.transformer_agg == 'x-attn':
            w.x_attn_query = _ptr(self.x_attn_query)
            w.x_attn_kv_norm_weight, w.x_attn_kv_norm_bias = _ptr(self.x_attn_kv_norm.weight), _ptr(self.x_attn_kv_norm.bias)
            w.x_attn_query_norm_weight = _ptr(self.x_attn_query_norm.weight)
            w.x_attn_query_norm_bias = _ptr(self.x_attn_query_norm.bias)
            w.x_attn_in_proj_weight = _ptr(self.x_attn_mha_layer.in_proj_weight)
            w.x_attn_in_proj_bias = _ptr(self.x_attn_mha_layer.in_proj_bias)
            w.x_attn_out_proj_weight = _ptr(self.x_attn_mha_layer.out_proj.weight)
            w.x_attn_out_proj_bias = _ptr(self.x_attn_mha_layer.out_proj.bias)
        return w

    def _prepared_weights(self, device, T):
        """Weights struct + bf16 operand copies, cached until a parameter changes (version counter / storage)."""
        params = list(self.parameters())
        key = (self.precision, str(device), tuple((p.data_ptr(), p._version) for p in params))
        cache = getattr(self, "_mdg_cache", None)
        if cache is None or cache[0] != key:
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("madrigal_b200: module parameters must be contiguous float32 CUDA tensors")
            w = self._weights()
            cfg = MdgFusionCfg(self.embed_dim, len(self.transformer_encoder.layers), self.num_heads, self.head_dim,
                               self.ffn_dim, _lib.MDG_ACTN[self.actn], int(self.norm_first),
                               _lib.MDG_AGG[self.transformer_agg], 1)
            prec = _PRECISION[self.precision]
            fn = _lib.lib()
            nbytes = fn.mdg_fusion_prepared_bytes(ctypes.byref(cfg), prec)
            buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
            with torch.cuda.device(device):
                _lib.check(fn.mdg_fusion_prepare(ctypes.byref(w), ctypes.byref(cfg), prec, buf.data_ptr(), buf.numel(),
                                                 _stream_ptr(device)), "mdg_fusion_prepare")
            pm = None
            if self.transformer_agg == 'x-attn':
                pm = self.x_attn_key_padding_mask.reshape(-1).to(device=device, dtype=torch.uint8).contiguous()
            cache = (key, w, buf, pm, self.x_attn_key_padding_mask if pm is not None else None)
            self._mdg_cache = cache
        if cache[3] is not None and cache[4] is not self.x_attn_key_padding_mask:  # attribute was replaced
            pm = self.x_attn_key_padding_mask.reshape(-1).to(device=device, dtype=torch.uint8).contiguous()
            cache = (cache[0], cache[1], cache[2], pm, self.x_attn_key_padding_mask)
            self._mdg_cache = cache
        return cache[1], cache[2], cache[3]

    def forward(self, fusion_sequence: torch.Tensor, fusion_mask: torch.Tensor, src_mask: Optional[torch.Tensor] = None):
        x = _require_cuda_f32(fusion_sequence, "fusion_sequence")
        if x.dim() != 3 or x.shape[2] != self.embed_dim:
            raise ValueError(f"fusion_sequence must be [B, T, {self.embed_dim}]")
        B, T, E = x.shape
        km = _mask_u8(fusion_mask, "fusion_mask")
        if tuple(km.shape) != (B, T):
            raise ValueError("fusion_mask must be [B, T]")
        sm = None
        if src_mask is not None:
            sm = _mask_u8(src_mask, "src_mask")
            if tuple(sm.shape) != (T, T):
                raise ValueError("src_mask must be [T, T]")
        w, prepared, pm = self._prepared_weights(x.device, T)
        if pm is not None and pm.numel() != T:
            raise ValueError(f"x_attn_key_padding_mask has {pm.numel()} keys but the sequence has {T} tokens")
        cfg = MdgFusionCfg(E, len(self.transformer_encoder.layers), self.num_heads, self.head_dim, self.ffn_dim,
                           _lib.MDG_ACTN[self.actn], int(self.norm_first), _lib.MDG_AGG[self.transformer_agg], T)
        prec = _PRECISION[self.precision]
        fn = _lib.lib()
        z = torch.empty((B, E), dtype=torch.float32, device=x.device)
        if B == 0:
            return z
        nbytes = fn.mdg_fusion_workspace_bytes(ctypes.byref(cfg), B, prec)
        ws = _workspace(x.device, max(nbytes, 256))  # nbytes == 0: unsupported config, the C side reports why
        with torch.cuda.device(x.device):
            _lib.check(fn.mdg_fusion_encode(ctypes.byref(w), ctypes.byref(cfg), prepared.data_ptr(), x.data_ptr(),
                                            km.data_ptr(), _ptr(sm), _ptr(pm), z.data_ptr(), B, prec, ws.data_ptr(),
                                            ws.numel(), _stream_ptr(x.device)), "mdg_fusion_encode")
        self.last_launch_count = fn.mdg_last_launch_count()
        return z


class MLPAdaptor(nn.Module):
    """Drop-in for the reference `MLPAdaptor` (models.py:459-518): same ctor and `fc.*` state_dict keys.
    Eval-mode arithmetic only (Dropout = identity); norm 'ln', 'bn' or None; activation relu or gelu.  norm='bn'
    (nn.BatchNorm1d, models.py:154,492) is an inference-only path: in eval mode it is a per-feature affine map of the
    running statistics, folded once into the Linear that follows it (W' = W diag(s), b' = b + W t) on the device."""

    def __init__(self, in_dim: int, hidden_dims: list, output_dim: int, p: float, norm: str, actn: str,
                 order: str = 'nd', precision: str = "fp32"):
        super().__init__()
        if actn not in _lib.MDG_ACTN:
            raise NotImplementedError(actn)
        if norm not in ('ln', 'bn', None, 'None'):
            raise NotImplementedError(norm)
        if order not in ('nd', 'dn'):
            raise NotImplementedError(order)
        self.actn, self.precision = actn, precision
        act = lambda: nn.ReLU() if actn == 'relu' else nn.GELU()
        layers = [nn.Linear(in_dim, hidden_dims[0]), act()]
        for i in range(len(hidden_dims) - 1):
            block = []
            nrm = nn.LayerNorm(hidden_dims[i]) if norm == 'ln' else (nn.BatchNorm1d(hidden_dims[i]) if norm == 'bn' else None)
            drop = nn.Dropout(p) if p != 0 else None
            for m in ((nrm, drop) if order == 'nd' else (drop, nrm)):
                if m is not None:
                    block.append(m)
            layers += block + [nn.Linear(hidden_dims[i], hidden_dims[i + 1]), act()]
        layers.append(nn.Linear(hidden_dims[-1], output_dim))
        self.fc = nn.Sequential(*layers)  # container only
        self._bn_folded = None

    def _fold_batchnorm(self):
        """{index of the Linear: (W', b')} with the preceding eval-mode BatchNorm1d folded in; cached until a parameter
        or running statistic changes.  Weight preprocessing (like mdg_fusion_prepare's LayerNorm fold), not path
        arithmetic: the activations only ever see mdg_mlp_forward."""
        mods = list(self.fc)
        tensors = [t for m in mods for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False))]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._bn_folded is not None and self._bn_folded[0] == key:
            return self._bn_folded[1]
        out, pending = {}, None
        with torch.no_grad():
            for i, m in enumerate(mods):
                if isinstance(m, nn.BatchNorm1d):
                    s = m.weight / torch.sqrt(m.running_var + m.eps)
                    pending = (s, m.bias - m.running_mean * s)
                elif isinstance(m, nn.Linear) and pending is not None:
                    s, t = pending
                    out[i] = ((m.weight * s[None, :]).contiguous(), (m.bias + m.weight @ t).contiguous())
                    pending = None
        self._bn_folded = (key, out)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        has_bn = any(isinstance(m, nn.BatchNorm1d) for m in self.fc)
        if has_bn and self.training:
            raise RuntimeError("madrigal_b200.MLPAdaptor(norm='bn') is an inference path (BatchNorm1d uses its running "
                               "statistics): call .eval() first")
        folded = self._fold_batchnorm() if has_bn else {}
        x2 = _require_cuda_f32(x, "x")
        lead = x2.shape[:-1]
        x2 = x2.reshape(-1, x2.shape[-1]).contiguous()
        B = x2.shape[0]
        if B == 0:
            return torch.empty((*lead, [l for l in self.fc if isinstance(l, nn.Linear)][-1].out_features),
                               dtype=torch.float32, device=x2.device)
        m = MdgMlp()
        linears = [l for l in self.fc if isinstance(l, nn.Linear)]
        m.n_linear = len(linears)
        m.actn = _lib.MDG_ACTN[self.actn]
        pending_ln = None
        i = 0
        for pos, layer in enumerate(self.fc):
            if isinstance(layer, nn.LayerNorm):
                pending_ln = layer
            elif isinstance(layer, nn.Linear):
                m.dims[i] = layer.in_features
                m.dims[i + 1] = layer.out_features
                wt, bs = folded.get(pos, (layer.weight, layer.bias))
                m.weight[i], m.bias[i] = wt.data_ptr(), bs.data_ptr()
                if pending_ln is not None:
                    m.ln_weight[i], m.ln_bias[i] = pending_ln.weight.data_ptr(), pending_ln.bias.data_ptr()
                    pending_ln = None
                i += 1
        y = torch.empty((B, linears[-1].out_features), dtype=torch.float32, device=x2.device)
        prec = _PRECISION[self.precision]
        fn = _lib.lib()
        ws = _workspace(x2.device, fn.mdg_mlp_workspace_bytes(ctypes.byref(m), B, prec))
        with torch.cuda.device(x2.device):
            _lib.check(fn.mdg_mlp_forward(ctypes.byref(m), x2.data_ptr(), y.data_ptr(), B, prec, ws.data_ptr(),
                                          ws.numel(), _stream_ptr(x2.device)), "mdg_mlp_forward")
        return y.reshape(*lead, y.shape[-1])


class PositionEncodingSinusoidal(nn.Module):
    """Buffer `pe` built exactly as models.py:551-579 (zero-padded beyond max_len when bottlenecks are used)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 19, num_tx_bottlenecks: int = 0,
                 transformer_agg: str = 'cls', non_tx_modalities=None):
        super().__init__()
        import math
        NUM_MODALITIES = len(resolve_non_tx(non_tx_modalities)) + len(CELL_LINES)
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(1, max_len, d_model)
        pe[0, :, 0::2] = torch.sin(position * div_term)
        pe[0, :, 1::2] = torch.cos(position * div_term)
        if num_tx_bottlenecks > 0:
            seq_len = NUM_MODALITIES + num_tx_bottlenecks + (1 if transformer_agg == 'cls' else 0)
            padded = torch.zeros(1, seq_len, d_model)
            padded[:, :max_len] = pe
            pe = padded
        self.register_buffer('pe', pe)


class PositionEncodingLearnable(nn.Module):
    """Parameter `pe` [1, max_len, d_model] added to the first max_len tokens (models.py:590-603)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 19, num_tx_bottlenecks: int = 0,
                 transformer_agg: str = 'cls'):
        super().__init__()
        self.max_len = max_len
        self.pe = nn.Parameter(torch.randn(1, max_len, d_model))


class MLPEncoder(MLPAdaptor):
    """Drop-in for the reference `MLPEncoder` (models.py:121-180), the tabular modality encoder used for the
    cell-viability (`cv_encoder: mlp`, get_tabular_mod_encoder :250-259) and `tx_encoder: mlp` modalities.  The
    reference class is line-for-line the same module as `MLPAdaptor` (same ctor, same `fc.*` keys), so it runs on the
    same kernel chain (mdg_mlp_forward); pretrained `cv_model_ae.pt` state_dicts load unchanged."""


def masked_pool(tokens: torch.Tensor, masks: torch.Tensor, mode: str) -> torch.Tensor:
    """fusion='mean' / 'add' (models.py:870-878): masked mean / sum over visible modality tokens."""
    x = _require_cuda_f32(tokens, "tokens")
    km = _mask_u8(masks, "masks")
    B, T, E = x.shape
    z = torch.empty((B, E), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().mdg_masked_pool(x.data_ptr(), km.data_ptr(), B, T, E, {"mean": 0, "max": 1, "add": 2}[mode],
                                              z.data_ptr(), _stream_ptr(x.device)), "mdg_masked_pool")
    return z


class FusionEncoder(nn.Module):
    """The fusion section of the reference `NovelDDIEncoder` (models.py:653-714 construction, :772-896 arithmetic),
    starting from the stacked modality embeddings.  Parameter names match the reference encoder's
    (`tx_bottleneck_tokens`, `cls`, `pos_encoder.pe`, `transformer.*`, `uni_projector.fc.*`, `uni_fuser.fc.*`), so
    `load_state_dict(reference_encoder_state_dict, strict=False)` picks them up; the modality encoders (GNNs, chemCPA)
    are out of scope and stay the reference's.

    forward(all_embeds [B, 19, E] in the order [str, kg, cv, (extra tabular modalities), tx_a375 ... tx_yapc]
            (models.py:772, utils.py:28), batch_masks [B, 19] bool True = modality missing) -> z [B, E]
    `non_tx_modalities`: the reference's NON_TX_MODALITIES knob (utils.py:30-37) as a constructor argument — None
    (environment / default 3), a count, or the list of names; 19 becomes len(non_tx) + 16.
    """

    def __init__(self, feat_dim, num_tx_bottlenecks, pos_emb_dropout, transformer_fusion_hparams, proj_hparams,
                 fusion='transformer_uni_proj', normalize=False, pos_emb_type='learnable', adapt_before_fusion=False,
                 precision: str = "fp32", non_tx_modalities=None, **kwargs):
        super().__init__()
        self.non_tx_modalities = resolve_non_tx(non_tx_modalities)
        NUM_NON_TX_MODALITIES = len(self.non_tx_modalities)
        NUM_MODALITIES = NUM_NON_TX_MODALITIES + len(CELL_LINES)
        self.embed_dim, self.fusion, self.normalize = feat_dim, fusion, normalize
        self.adapt_before_fusion = adapt_before_fusion
        self.num_tx_bottlenecks = num_tx_bottlenecks
        self.transformer_agg = transformer_fusion_hparams['transformer_agg']
        self.pos_emb_type = pos_emb_type
        max_len = NUM_MODALITIES if num_tx_bottlenecks == 0 else NUM_NON_TX_MODALITIES  # models.py:668-676
        if self.transformer_agg == 'cls':
            max_len += 1
        self.pos_emb_max_len = max_len
        if num_tx_bottlenecks > 0:
            self.tx_bottleneck_tokens = nn.Parameter(torch.randn(num_tx_bottlenecks, feat_dim))
        if pos_emb_type == 'learnable':
            self.pos_encoder = PositionEncodingLearnable(feat_dim, pos_emb_dropout, max_len, num_tx_bottlenecks,
                                                         self.transformer_agg)
        elif pos_emb_type == 'sinusoidal':
            self.pos_encoder = PositionEncodingSinusoidal(feat_dim, pos_emb_dropout, max_len, num_tx_bottlenecks,
                                                          self.transformer_agg, self.non_tx_modalities)
        else:
            raise NotImplementedError(pos_emb_type)
        self.transformer = TransformerFusion(feat_dim, num_tx_bottlenecks, precision=precision,
                                             non_tx_modalities=self.non_tx_modalities, **transformer_fusion_hparams)
        if self.transformer_agg == 'cls':
            self.cls = nn.Parameter(torch.randn(1, feat_dim))
        mk = lambda: MLPAdaptor(feat_dim, proj_hparams['proj_hidden_dims'], feat_dim, proj_hparams['proj_dropout'],
                                proj_hparams['proj_norm'], proj_hparams['proj_actn'], proj_hparams['proj_order'],
                                precision=precision)
        self.uni_projector = mk()
        if fusion == 'transformer_uni_proj':
            self.uni_fuser = mk()

    def _src_mask(self, device) -> Optional[torch.Tensor]:
        nb = self.num_tx_bottlenecks
        if nb == 0:
            return None
        n_tx, n_non = len(CELL_LINES), len(self.non_tx_modalities)
        T = n_non + n_tx + nb
        sm = torch.zeros((T, T), dtype=torch.bool)
        sm[:n_non, T - n_tx:] = True  # non-TX queries never see TX keys ...
        sm[T - n_tx:, :n_non] = True  # ... and vice versa; bottleneck tokens see everything (models.py:813-816)
        if self.transformer_agg == 'cls':  # CLS row/column attends to / is attended by all (models.py:828-842)
            full = torch.zeros((T + 1, T + 1), dtype=torch.bool)
            full[1:, 1:] = sm
            sm = full
        return sm.to(device)

    def assemble(self, embeds: torch.Tensor, masks: torch.Tensor):
        """mdg_assemble_tokens: [B, 19, E] + masks -> position-encoded sequence [B, T, E] and its key mask [B, T]."""
        x = _require_cuda_f32(embeds, "all_embeds")
        km = _mask_u8(masks, "batch_masks")
        B, M, E = x.shape
        nb = self.num_tx_bottlenecks
        has_cls = self.transformer_agg == 'cls'
        T = M + nb + int(has_cls)
        pe = self.pos_encoder.pe
        pe2 = pe.reshape(pe.shape[1], E).contiguous()
        seq = torch.empty((B, T, E), dtype=torch.float32, device=x.device)
        smask = torch.empty((B, T), dtype=torch.uint8, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().mdg_assemble_tokens(
                x.data_ptr(), km.data_ptr(), B, M, E, len(self.non_tx_modalities), nb,
                _ptr(self.tx_bottleneck_tokens) if nb > 0 else None, _ptr(self.cls) if has_cls else None,
                pe2.data_ptr(), min(pe2.shape[0], T), int(bool(self.normalize)), seq.data_ptr(), smask.data_ptr(),
                _stream_ptr(x.device)), "mdg_assemble_tokens")
        return seq, smask

    def forward(self, all_embeds: torch.Tensor, batch_masks: torch.Tensor) -> torch.Tensor:
        embeds = _require_cuda_f32(all_embeds, "all_embeds")
        masks = batch_masks.to(torch.bool)
        if self.adapt_before_fusion:
            embeds = self.uni_projector(embeds)  # models.py:776-777
        if self.fusion in ('mean', 'add'):  # models.py:870-878
            x = l2_normalize_rows(embeds) if self.normalize else embeds
            return masked_pool(x, masks, self.fusion)
        if self.fusion not in ('transformer', 'transformer_uni_proj'):
            raise NotImplementedError(self.fusion)
        if self.fusion == 'transformer':
            seq, smask = self.assemble(embeds, masks)
            return self.transformer(seq, smask, self._src_mask(embeds.device))
        # transformer_uni_proj (models.py:781-790, 855-865): single-modality drugs bypass the transformer
        visible = (~masks).sum(dim=1)
        assert torch.all(visible > 0)  # models.py:783
        multi = visible > 1
        z = torch.empty((embeds.shape[0], self.embed_dim), dtype=torch.float32, device=embeds.device)
        if bool(multi.any()):
            seq, smask = self.assemble(embeds[multi].contiguous(), masks[multi].contiguous())
            z[multi] = self.transformer(seq, smask, self._src_mask(embeds.device))
        if bool((~multi).any()):
            uni_rows = (~multi).nonzero(as_tuple=True)[0]
            uni_mod = (~masks[uni_rows]).to(torch.uint8).argmax(dim=1)
            uni = embeds[uni_rows, uni_mod].contiguous()
            if self.normalize:
                uni = l2_normalize_rows(uni)
            z[uni_rows] = self.uni_fuser(uni)
        return z
