"""Transcriptomic (`tx: chemcpa`) modality encoder — the token chemCPA contributes to the fusion sequence.

Reference: `TxAdaptingComPert` (madrigal/chemcpa/chemCPA/model.py:290-712) as NovelDDIEncoder.encode calls it
(models.py:756-769): `_, _, tx = tx_encoder.predict(genes, drugs_idx, dosages, covariates, return_latent_basal=...,
return_latent_treated=...)`.  Only the latent path is on the scoring path:

    latent_basal   = encoder(genes)                                             (model.py:678)
    latent_treated = latent_basal + dose_scale * drug_embedding_encoder(E[idx])  (:683-689, use_drugs only)
                                  + covariates_embeddings[c][argmax(onehot_c)]   (:693-697)

The gene decoder, the adversaries and the training losses are not built (Madrigal discards the reconstruction and
instantiates with disable_adv=True, models.py:278-288); `predict` returns None in the first two tuple slots.

`MLP` keeps the reference's constructor and `network.*` state_dict keys (Linear -> BatchNorm1d -> ReLU chains).  In
eval mode BatchNorm1d is a per-feature affine map, folded here into the preceding Linear at first use, so the chain
is `mdg_mlp_forward` (tcgen05 GEMMs, ReLU in the epilogue); `mdg_tx_latent_combine` finishes the token.  CUDA only.
"""
import ctypes
import json
from collections import OrderedDict
from typing import Dict, List, Union

import torch
import torch.nn as nn

from . import _lib
from ._lib import MdgMlp
from .decoder import _require_cuda_f32, _stream_ptr, _workspace
from .fusion import _PRECISION

_DOSER = {None: 0, "sigm": 1, "logsigm": 2, "amortized": 3, "mlp": 3}  # 3 = the scale is computed before the combine


class MLP(nn.Module):
    """Drop-in for chemCPA's `MLP` (model.py:161-231), eval-mode arithmetic, last_layer_act='linear' only (the 'ReLU'
    variant is the gene decoder's)."""

    def __init__(self, sizes, batch_norm=True, last_layer_act="linear", append_layer_width=None,
                 append_layer_position=None, precision: str = "fp32"):
        super().__init__()
        if last_layer_act != "linear":
            raise NotImplementedError("last_layer_act='ReLU' is the gene decoder's variant; not on the scoring path")
        layers = []
        for s in range(len(sizes) - 1):
            layers.append(nn.Linear(sizes[s], sizes[s + 1]))
            if batch_norm and s < len(sizes) - 2:
                layers.append(nn.BatchNorm1d(sizes[s + 1]))
            layers.append(nn.ReLU())
        layers = layers[:-1]
        self.activation = last_layer_act
        named = OrderedDict()
        if append_layer_width:
            assert append_layer_position in ("first", "last")
            if append_layer_position == "first":
                named["append_linear"] = nn.Linear(append_layer_width, sizes[0])
                named["append_bn1d"] = nn.BatchNorm1d(sizes[0])
                named["append_relu"] = nn.ReLU()
                for i, mod in enumerate(layers):
                    named[str(i)] = mod
            else:
                for i, mod in enumerate(layers):
                    named[str(i)] = mod
                named["append_bn1d"] = nn.BatchNorm1d(sizes[-1])
                named["append_relu"] = nn.ReLU()
                named["append_linear"] = nn.Linear(sizes[-1], append_layer_width)
        else:
            for i, mod in enumerate(layers):
                named[str(i)] = mod
        self.network = nn.Sequential(named)  # container only
        self.precision = precision
        self._folded = None
        self._folded_key = None

    def _fold(self):
        """[(W', b')] per Linear with the following BatchNorm1d's eval statistics folded in:
        W' = diag(g) W, b' = (b - mean) g + beta, g = gamma / sqrt(var + eps)."""
        mods = list(self.network)
        tensors = [t for m in mods for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False))]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._folded is not None and key == self._folded_key:
            return self._folded
        out = []
        with torch.no_grad():
            for i, m in enumerate(mods):
                if not isinstance(m, nn.Linear):
                    continue
                w, b = m.weight.detach().float(), m.bias.detach().float()
                bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) else None
                if bn is not None:
                    g = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
                    w = w * g[:, None]
                    b = (b - bn.running_mean.float()) * g + bn.bias.detach().float()
                out.append((w.contiguous(), b.contiguous()))
        self._folded, self._folded_key = out, key
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise RuntimeError("madrigal_b200.chemcpa.MLP is an inference path (BatchNorm1d uses running statistics): "
                               "call .eval()")
        x2 = _require_cuda_f32(x, "x")
        if x2.dim() != 2:
            raise ValueError("x must be [B, in_features]")
        x2 = x2.contiguous()
        folded = self._fold()
        B = x2.shape[0]
        n_out = folded[-1][0].shape[0]
        y = torch.empty((B, n_out), dtype=torch.float32, device=x2.device)
        if B == 0:
            return y
        if len(folded) > _lib.MDG_MAX_MLP_LINEAR:
            raise NotImplementedError(f"more than {_lib.MDG_MAX_MLP_LINEAR} linear layers")
        m = MdgMlp()
        m.n_linear = len(folded)
        m.actn = _lib.MDG_ACTN["relu"]
        for i, (w, b) in enumerate(folded):
            if w.device != x2.device:
                raise ValueError("module parameters and input are on different devices")
            m.dims[i], m.dims[i + 1] = w.shape[1], w.shape[0]
            m.weight[i], m.bias[i] = w.data_ptr(), b.data_ptr()
        prec = _PRECISION[self.precision]
        fn = _lib.lib()
        ws = _workspace(x2.device, fn.mdg_mlp_workspace_bytes(ctypes.byref(m), B, prec))
        with torch.cuda.device(x2.device):
            _lib.check(fn.mdg_mlp_forward(ctypes.byref(m), x2.data_ptr(), y.data_ptr(), B, prec, ws.data_ptr(),
                                          ws.numel(), _stream_ptr(x2.device)), "mdg_mlp_forward")
        return y


class GeneralizedSigmoid(nn.Module):
    """Parameter holder for chemCPA's dose-response curve (model.py:234-287): `beta`, `bias` [1, num_drugs].  The
    curve itself is evaluated inside mdg_tx_latent_combine."""

    def __init__(self, dim, device=None, nonlin="sigm"):
        super().__init__()
        assert nonlin in ("sigm", "logsigm", None)
        self.nonlin = nonlin
        self.beta = nn.Parameter(torch.ones(1, dim))
        self.bias = nn.Parameter(torch.zeros(1, dim))


class TxAdaptingComPert(nn.Module):
    """Drop-in for the latent path of chemCPA's `TxAdaptingComPert` (model.py:290-712): same constructor arguments and
    the same state_dict keys for `encoder.*`, `drug_embedding_encoder.*`, `dosers.*`, `drug_embeddings.weight`,
    `covariates_embeddings.N.weight` (load reference checkpoints with strict=False, as models.py:329 does: their
    `decoder.*` / `adversary_*` entries have no counterpart here)."""

    def __init__(self, num_genes: int, num_drugs: int, covariate_names_unique: Dict[str, List[str]], seed=0,
                 patience=5, doser_type="logsigm", decoder_activation="linear", hparams: Union[str, dict] = "",
                 drug_embeddings: Union[None, nn.Embedding] = None, append_layer_width=None, use_drugs=True,
                 disable_adv=False, precision: str = "fp32", **kwargs):
        super().__init__()
        self.num_genes, self.num_drugs = num_genes, num_drugs
        self.covariate_names_unique = covariate_names_unique
        self.num_covariates = [len(names) for names in covariate_names_unique.values()]
        assert 0 not in self.num_covariates
        self.use_drugs, self.use_drugs_idx = use_drugs, True
        if isinstance(hparams, str):
            if not hparams:
                raise ValueError("hparams must be given (dict or JSON): the reference's random defaults "
                                 "(model.py:541-565) are a training-time sweep, not an inference configuration")
            hparams = json.loads(hparams)
        self.hparams = dict(hparams)
        hp = self.hparams
        self.encoder = MLP([num_genes] + [hp["autoencoder_width"]] * hp["autoencoder_depth"] + [hp["dim"]],
                           append_layer_width=append_layer_width, append_layer_position="first", precision=precision)
        if append_layer_width:
            self.num_genes = append_layer_width
        if use_drugs:
            if doser_type not in _DOSER:
                raise NotImplementedError(f"doser_type={doser_type!r}")
            self.drug_embeddings = drug_embeddings if drug_embeddings is not None else nn.Embedding(num_drugs, hp["dim"])
            self.drug_embedding_encoder = MLP(
                [self.drug_embeddings.embedding_dim]
                + [hp.get("embedding_encoder_width", 512)] * hp.get("embedding_encoder_depth", 0) + [hp["dim"]],
                last_layer_act="linear", precision=precision)
            if doser_type == "mlp":  # one small MLP per drug on the scalar dosage (model.py:405-416)
                if hp["dosers_depth"] < 1:
                    raise NotImplementedError("'mlp' dosers need dosers_depth >= 1")
                self.dosers = nn.ModuleList([MLP([1] + [hp["dosers_width"]] * hp["dosers_depth"] + [1], batch_norm=False)
                                             for _ in range(num_drugs)])
                self._doser_stack, self._doser_key = None, None
            elif doser_type == "amortized":
                self.dosers = MLP([self.drug_embeddings.embedding_dim + 1]
                                  + [hp["dosers_width"]] * hp["dosers_depth"] + [1], precision=precision)
            else:
                self.dosers = GeneralizedSigmoid(num_drugs, None, nonlin=doser_type)
            self.doser_type = doser_type
        else:
            self.drug_embeddings = self.drug_embedding_encoder = self.dosers = None
        self.covariates_embeddings = nn.ModuleList([nn.Embedding(n, hp["dim"]) for n in self.num_covariates])

    def _mlp_doser_scale(self, dosage, idx):
        """sigmoid(dosers[idx[b]](dosage[b])) for the per-drug 'mlp' dosers (model.py:609-621) through mdg_doser_mlp; the
        drugs' parameters are stacked once ([num_drugs, ...]) and re-stacked when any of them changes."""
        lin = [[m for m in d.network if isinstance(m, nn.Linear)] for d in self.dosers]
        tensors = [t for ls in lin for m in ls for t in (m.weight, m.bias)]
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if self._doser_stack is None or key != self._doser_key:
            f = lambda t: t.detach().float()
            w = lin[0][0].out_features
            st = dict(
                w_in=torch.stack([f(ls[0].weight).reshape(w) for ls in lin]).contiguous(),
                b_in=torch.stack([f(ls[0].bias) for ls in lin]).contiguous(),
                w_out=torch.stack([f(ls[-1].weight).reshape(w) for ls in lin]).contiguous(),
                b_out=torch.stack([f(ls[-1].bias).reshape(()) for ls in lin]).contiguous(),
                w_hid=None, b_hid=None, width=w, depth=len(lin[0]) - 1)
            if st["depth"] > 1:
                st["w_hid"] = torch.stack([torch.stack([f(m.weight) for m in ls[1:-1]]) for ls in lin]).contiguous()
                st["b_hid"] = torch.stack([torch.stack([f(m.bias) for m in ls[1:-1]]) for ls in lin]).contiguous()
            self._doser_stack, self._doser_key = st, key
        st = self._doser_stack
        if st["w_in"].device != dosage.device:
            raise ValueError("module parameters and input are on different devices")
        scale = torch.empty_like(dosage)
        ptr = lambda t: 0 if t is None else t.data_ptr()
        with torch.cuda.device(dosage.device):
            _lib.check(_lib.lib().mdg_doser_mlp(
                dosage.data_ptr(), idx.data_ptr(), dosage.shape[0], len(self.dosers), st["width"], st["depth"],
                ptr(st["w_in"]), ptr(st["b_in"]), ptr(st["w_hid"]), ptr(st["b_hid"]), ptr(st["w_out"]), ptr(st["b_out"]),
                scale.data_ptr(), _stream_ptr(dosage.device)), "mdg_doser_mlp")
        return scale

    def _combine(self, basal, drug_latent, dosage, drugs_idx, cov_table, cov_idx, doser):
        B, dim = basal.shape
        out = torch.empty_like(basal)
        if B == 0:
            return out
        ptr = lambda t: 0 if t is None else t.data_ptr()
        beta = bias = None
        if doser in (1, 2):
            beta = self.dosers.beta.detach().reshape(-1).contiguous()
            bias = self.dosers.bias.detach().reshape(-1).contiguous()
        with torch.cuda.device(basal.device):
            _lib.check(_lib.lib().mdg_tx_latent_combine(
                basal.data_ptr(), ptr(drug_latent), ptr(dosage), ptr(drugs_idx), ptr(beta), ptr(bias), doser,
                ptr(cov_table), ptr(cov_idx), B, dim, out.data_ptr(), _stream_ptr(basal.device)),
                "mdg_tx_latent_combine")
        return out

    @torch.no_grad()
    def predict(self, genes, drugs=None, drugs_idx=None, dosages=None, covariates=None, return_latent_basal=False,
                return_latent_treated=False):
        """-> (None, None[, latent_basal][, latent_treated]); the reference's first two slots (gene reconstruction,
        cell/drug embedding) are not on the scoring path."""
        if drugs is not None:
            raise NotImplementedError("one-hot `drugs` dose matrices are a training-time input; pass drugs_idx/dosages")
        genes = _require_cuda_f32(genes, "genes")
        dev = genes.device
        latent_basal = self.encoder(genes)
        output = (None, None)
        if return_latent_basal:
            output += (latent_basal,)
        if not return_latent_treated:
            return output
        drug_latent = dosage = idx = None
        doser = 0
        if self.num_drugs > 0 and self.use_drugs:
            assert drugs_idx is not None and dosages is not None
            idx = drugs_idx.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
            dosage = dosages.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
            assert idx.shape == dosage.shape and idx.shape[0] == genes.shape[0]
            emb = self.drug_embeddings.weight.detach().float()[idx]  # gather (plumbing), model.py:601
            doser = _DOSER[self.doser_type]
            if self.doser_type == "mlp":
                dosage = self._mlp_doser_scale(dosage, idx)
            elif doser == 3:  # amortized: MLP over [embedding | dosage] (model.py:622-627)
                dosage = self.dosers(torch.cat([emb, dosage[:, None]], dim=1)).reshape(-1).contiguous()
            drug_latent = self.drug_embedding_encoder(emb.contiguous())
        latent = latent_basal
        if self.num_covariates[0] > 0:
            for c, table in enumerate(self.covariates_embeddings):
                cov_idx = covariates[c].to(dev).argmax(1).to(torch.int64).contiguous()
                latent = self._combine(latent, drug_latent if c == 0 else None, dosage, idx,
                                       table.weight.detach().float().contiguous(), cov_idx, doser)
        elif drug_latent is not None:
            latent = self._combine(latent, drug_latent, dosage, idx, None, None, doser)
        return output + (latent,)
