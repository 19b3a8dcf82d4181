"""Deterministic synthetic weights and loader-shaped inputs for the Sub-GC hot path.

There is no dataset or checkpoint offline, so tests, `bench.py` and `smoke()` all draw from here.
Layouts follow the reference loaders exactly (dataloaders/dataloader.py:194-206,224-304 for training,
dataloaders/dataloader_test.py:191-203,221-273 for inference): dummy node 36 / dummy edge 64,
pad indices, diagonal pooling matrices, ×5 sentence copies.  Every tensor comes from a CPU
`torch.Generator` seeded from (seed, crc32(name)), so the same call gives the same bits on any box with the
same torch build; `fingerprint()` lets a fixture record what it was generated from.
"""
from __future__ import annotations

import zlib
from collections import OrderedDict

import torch

from .config import Dims


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
    return g


def param_shapes(d: Dims) -> "OrderedDict[str, tuple]":
    """state_dict contract of the reference TopDownModel (models/AttModel.py:72-120,393-443,
    models/lib/graph_conv_unit.py:9-15, models/lib/gpn.py:25-39)."""
    s = OrderedDict()
    s["obj_v_proj.weight"] = (d.gcn, d.att_feat); s["obj_v_proj.bias"] = (d.gcn,)
    s["sg_obj_embed.weight"] = (d.obj_classes, d.embed)
    s["obj_emb_proj.weight"] = (d.gcn, d.embed); s["obj_emb_proj.bias"] = (d.gcn,)
    s["sg_pred_embed.weight"] = (d.pred_classes, d.embed)
    s["pred_emb_prj.weight"] = (d.gcn, d.embed); s["pred_emb_prj.bias"] = (d.gcn,)
    for l in range(d.gcn_layers):
        for u in range(4):
            p = f"gcn_backbone.gcn.{l}.gcn_collect.collect_units.{u}."
            s[p + "fc_lft.weight"] = (d.low_rank, d.gcn); s[p + "fc_lft.bias"] = (d.low_rank,)
            s[p + "fc_rgt.weight"] = (d.gcn, d.low_rank); s[p + "fc_rgt.bias"] = (d.gcn,)
    s["gpn_layer.gpn_fc.0.weight"] = (d.att_hid, 2 * d.gcn); s["gpn_layer.gpn_fc.0.bias"] = (d.att_hid,)
    s["gpn_layer.gpn_fc.3.weight"] = (1, d.att_hid); s["gpn_layer.gpn_fc.3.bias"] = (1,)
    s["gpn_layer.read_out_proj.0.weight"] = (d.att_hid, 2 * d.gcn); s["gpn_layer.read_out_proj.0.bias"] = (d.att_hid,)
    s["gpn_layer.read_out_proj.1.weight"] = (2 * d.gcn, d.att_hid); s["gpn_layer.read_out_proj.1.bias"] = (2 * d.gcn,)
    s["logit.weight"] = (d.v1, d.rnn); s["logit.bias"] = (d.v1,)
    s["embed.0.weight"] = (d.v1, d.enc)
    s["fc_embed.0.weight"] = (d.fc_feat, d.att_feat); s["fc_embed.0.bias"] = (d.fc_feat,)
    s["fc_embed.2.weight"] = (d.rnn, d.fc_feat); s["fc_embed.2.bias"] = (d.rnn,)
    s["att_embed.0.weight"] = (d.rnn, d.gcn); s["att_embed.0.bias"] = (d.rnn,)
    s["ctx2att.weight"] = (d.att_hid, d.rnn); s["ctx2att.bias"] = (d.att_hid,)
    s["core.attention.h2att.weight"] = (d.att_hid, d.rnn); s["core.attention.h2att.bias"] = (d.att_hid,)
    s["core.attention.alpha_net.weight"] = (1, d.att_hid); s["core.attention.alpha_net.bias"] = (1,)
    s["core.att_lstm.weight_ih"] = (4 * d.rnn, d.enc + 2 * d.rnn); s["core.att_lstm.weight_hh"] = (4 * d.rnn, d.rnn)
    s["core.att_lstm.bias_ih"] = (4 * d.rnn,); s["core.att_lstm.bias_hh"] = (4 * d.rnn,)
    s["core.lang_lstm.weight_ih"] = (4 * d.rnn, 2 * d.rnn); s["core.lang_lstm.weight_hh"] = (4 * d.rnn, d.rnn)
    s["core.lang_lstm.bias_ih"] = (4 * d.rnn,); s["core.lang_lstm.bias_hh"] = (4 * d.rnn,)
    return s


def make_state_dict(d: Dims, seed: int = 0, gcn_std: float | None = None, logit_gain: float = 1.0,
                    lstm_gain: float = 1.0, eos_bias: float = 0.0):
    """Random fp32 state_dict with the reference's key names.

    Linear-like tensors are U(±1/sqrt(fan_in)) (nn.Linear/LSTMCell default), embeddings N(0,1).
    `gcn_std=None` draws the GCN units like every other Linear so the message-passing path contributes O(1)
    to `x_obj` (a meaningful parity test); `gcn_std=1e-3` reproduces the reference initialiser
    (models/lib/graph_conv_unit.py:5-20: normal(0, 0.001), zero bias).  `logit_gain>1` sharpens the 9488-way
    distribution ("peaked-logit" variant, SURVEY §7 hard part 1); `lstm_gain>1` makes the recurrent state
    (and therefore the token sequence) depend more strongly on the fed-back token; `eos_bias` is added to the
    logit bias of token 0 so that some rows finish before the maximum length.
    """
    sd = OrderedDict()
    for name, shape in param_shapes(d).items():
        g = _gen(seed, name)
        if name in ("sg_obj_embed.weight", "sg_pred_embed.weight", "embed.0.weight"):
            t = torch.randn(shape, generator=g)
        elif "collect_units" in name and gcn_std is not None:
            t = torch.randn(shape, generator=g) * gcn_std if name.endswith("weight") else torch.zeros(shape)
        else:
            if name.startswith("core.") and "lstm" in name:
                fan_in = d.rnn
            elif len(shape) == 2:
                fan_in = shape[1]
            else:  # bias: fan_in of its weight
                wname = name[:-4] + "weight"
                fan_in = param_shapes(d)[wname][1]
            bound = 1.0 / (fan_in ** 0.5)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        if name == "logit.weight":
            t = t * logit_gain
        if name == "logit.bias" and eos_bias:
            t[0] += eos_bias
        if name.startswith("core.") and name.endswith(("weight_ih", "weight_hh")):
            t = t * lstm_gain
        sd[name] = t.contiguous()
    return sd


def _subgraph(g, d: Dims, ragged: bool):
    """(sorted node ids, count) of one sampled sub-graph."""
    real = d.obj_num - 1
    if not ragged:
        return torch.arange(real), real
    cnt = int(torch.randint(3, real + 1, (1,), generator=g))
    ids = torch.randperm(real, generator=g)[:cnt].sort().values
    return ids, cnt


def _fill_subgraph_tensors(g, d: Dims, rows, halves, per_half, ragged, identical_rows):
    N, K = d.obj_num, d.rel_num
    obj_ind = torch.full((rows, halves, per_half, N), N - 1, dtype=torch.int64)
    mask = torch.zeros(rows, halves, per_half, N)
    pool = torch.zeros(rows, halves, per_half, N, N)
    pred_ind = torch.full((rows, halves, per_half, K), K - 1, dtype=torch.int64)
    nrel_ind = torch.full((rows, halves, per_half, K, 2), N - 1, dtype=torch.int64)
    for r in range(rows):
        if identical_rows and r % identical_rows != 0:
            src = r - r % identical_rows
            obj_ind[r], mask[r], pool[r], pred_ind[r], nrel_ind[r] = obj_ind[src], mask[src], pool[src], pred_ind[src], nrel_ind[src]
            continue
        for h in range(halves):
            for m in range(per_half):
                ids, cnt = _subgraph(g, d, ragged)
                obj_ind[r, h, m, :cnt] = ids
                mask[r, h, m, :cnt] = 1
                ar = torch.arange(cnt)
                pool[r, h, m, ar, ar] = 1
                ne = int(torch.randint(0, K - 1, (1,), generator=g)) if ragged else min(cnt, K - 1)
                pred_ind[r, h, m, :ne] = torch.randperm(K - 1, generator=g)[:ne].sort().values
                nrel_ind[r, h, m, :ne] = torch.randint(0, max(cnt, 1), (ne, 2), generator=g)
    return obj_ind, mask, pool, pred_ind, nrel_ind


def make_graph_inputs(d: Dims, seed: int, n_images: int, ragged_edges: bool = False):
    """Per-image scene-graph tensors (identical layout for train and test loaders)."""
    g = _gen(seed, "graph")
    B, N, K = n_images, d.obj_num, d.rel_num
    att = torch.randn(B, N, d.att_feat, generator=g).abs_()
    att[:, N - 1] = 0
    obj_dist = torch.rand(B, N, d.obj_classes, generator=g)
    obj_dist[:, N - 1] = 0
    obj_dist[:, N - 1, 0] = 1
    pred_dist = torch.rand(B, K, d.pred_classes, generator=g)
    rel_ind = torch.full((B, K, 2), N - 1, dtype=torch.int64)
    for b in range(B):
        ne = int(torch.randint(2, K, (1,), generator=g)) if ragged_edges else K - 1
        rel_ind[b, :ne] = torch.randint(0, N - 1, (ne, 2), generator=g)
        pred_dist[b, ne:] = 0
        pred_dist[b, ne:, 0] = 1
    return dict(fc_feats=torch.zeros(B, d.att_feat), att_feats=att, obj_dist=obj_dist, rel_ind=rel_ind,
                pred_dist=pred_dist, trip_pred=None, obj_box=None, pred_fmap=None)


def make_test_inputs(d: Dims, seed: int = 0, n_images: int = 1, per_half: int = 1, ragged: bool = False,
                     ragged_edges: bool = False, seq_per_img: int = 5):
    """Inference batch: `n_images` scene graphs, each with 2*per_half candidate sub-graphs, the five sentence
    copies identical (dataloaders/dataloader_test.py:226-273).  n_images=1 is the reference's own call shape."""
    data = make_graph_inputs(d, seed, n_images, ragged_edges)
    g = _gen(seed, "subgraphs")
    rows = n_images * seq_per_img
    obj_ind, mask, pool, pred_ind, nrel_ind = _fill_subgraph_tensors(g, d, rows, 2, per_half, ragged, seq_per_img)
    data.update(att_masks=mask, gpn_obj_ind=obj_ind, gpn_pred_ind=pred_ind, gpn_nrel_ind=nrel_ind, gpn_pool_mtx=pool)
    return data


def make_train_inputs(d: Dims, seed: int = 0, n_images: int = 2, gpn_batch: int = 2, ragged: bool = True,
                      ragged_edges: bool = True, seq_per_img: int = 5, label_len: int = 18):
    """Training batch: 5 sentences per image, each with `gpn_batch` positive and negative sub-graphs
    (dataloaders/dataloader.py:224-304), labels [5B, label_len] = BOS 0, words, EOS 0 padding and masks covering
    BOS+words+EOS (dataloaders/dataloader.py:359-364)."""
    data = make_graph_inputs(d, seed, n_images, ragged_edges)
    g = _gen(seed, "train")
    rows = n_images * seq_per_img
    obj_ind, mask, pool, pred_ind, nrel_ind = _fill_subgraph_tensors(g, d, rows, 2, gpn_batch, ragged, 0)
    labels = torch.zeros(rows, label_len, dtype=torch.int64)
    masks = torch.zeros(rows, label_len)
    for r in range(rows):
        nw = int(torch.randint(3, label_len - 1, (1,), generator=g))
        labels[r, 1:1 + nw] = torch.randint(1, d.vocab + 1, (nw,), generator=g)
        masks[r, :nw + 2] = 1
    data.update(att_masks=mask, gpn_obj_ind=obj_ind, gpn_pred_ind=pred_ind, gpn_nrel_ind=nrel_ind, gpn_pool_mtx=pool,
                labels=labels, masks=masks)
    return data


SAMPLE_ARG_ORDER = ("fc_feats", "att_feats", "att_masks", "trip_pred", "obj_dist", "obj_box", "rel_ind", "pred_fmap",
                    "pred_dist", "gpn_obj_ind", "gpn_pred_ind", "gpn_nrel_ind", "gpn_pool_mtx")
"""Positional order of `model(..., mode='sample')` (misc/eval_utils.py:102-104)."""

FORWARD_ARG_ORDER = ("fc_feats", "att_feats", "labels", "att_masks", "trip_pred", "obj_dist", "obj_box", "rel_ind",
                     "pred_fmap", "pred_dist", "gpn_obj_ind", "gpn_pred_ind", "gpn_nrel_ind", "gpn_pool_mtx")
"""Positional order of `model(...)` in forward mode (misc/eval_utils.py:82-83, models/loss_wrapper.py:18-19)."""


def sample_args(data):
    return [data[k] for k in SAMPLE_ARG_ORDER]


def forward_args(data):
    return [data[k] for k in FORWARD_ARG_ORDER]


def fingerprint(tensors) -> float:
    """Order-dependent float64 checksum of a dict/list of tensors (recorded in fixtures to detect RNG drift)."""
    items = tensors.items() if hasattr(tensors, "items") else enumerate(tensors)
    acc = 0.0
    for i, (_, t) in enumerate(items):
        if t is None:
            continue
        acc += (i + 1) * float(t.double().sum()) + 0.5 * float(t.double().abs().max())
    return acc
