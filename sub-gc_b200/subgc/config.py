"""Model dimensions for the Sub-GC hot path, read from the reference's `opt` namespace.

The attribute names are the ones `AttModel.__init__` reads (reference models/AttModel.py:44-69,94-98);
`make_opt()` builds a namespace with the Sub_GC_Kar values of train.sh:20-24 / opts.py defaults so that
tests, bench and smoke can construct the model without argparse.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, asdict
from types import SimpleNamespace

# class-name tables bundled with the reference (data/object_names_1600-0-20.npy, predicate_names_...):
# only their lengths are used by the model once a checkpoint is loaded.
DEFAULT_OBJ_CLASSES = 1599
DEFAULT_PRED_CLASSES = 21
GCN_LOW_RANK = 512  # models/lib/graph_conv.py:11 (dim_lr default, never overridden)


@dataclass(frozen=True)
class Dims:
    vocab: int = 9487          # V; logits have V+1 columns
    enc: int = 1000            # input_encoding_size X
    rnn: int = 1000            # rnn_size H
    att_hid: int = 512         # att_hid_size AH (also sGPN hidden)
    fc_feat: int = 2048        # fc_feat_size
    att_feat: int = 2048       # att_feat_size A
    gcn: int = 1024            # gcn_dim L
    low_rank: int = GCN_LOW_RANK
    embed: int = 300           # embed_dim E (GloVe)
    obj_classes: int = DEFAULT_OBJ_CLASSES
    pred_classes: int = DEFAULT_PRED_CLASSES
    gcn_layers: int = 2
    gcn_residual: int = 2
    pred_emb_type: int = 1
    seq_length: int = 20
    obj_num: int = 37          # N (36 boxes + dummy)
    rel_num: int = 65          # K (64 edges + dummy)

    @property
    def v1(self) -> int:
        return self.vocab + 1

    def as_dict(self):
        return asdict(self)


def _count_names(path, default):
    if path and os.path.isfile(path):
        import numpy as np
        return int(np.load(path, encoding="latin1").shape[0])
    return default


def dims_from_opt(opt) -> Dims:
    """Translate the reference's `opt` into Dims; rejects the variants this path does not implement."""
    if getattr(opt, "use_gpn", 1) != 1 or getattr(opt, "noun_fuse", 1) != 1:
        raise NotImplementedError("only the Sub-GC configuration (use_gpn=1, noun_fuse=1) is implemented; "
                                  "Full-GC is SURVEY §8f 'next'")
    if getattr(opt, "gcn_bn", 0) != 0 or getattr(opt, "use_bn", 0) != 0:
        raise NotImplementedError("BatchNorm variants (gcn_bn/use_bn) are not part of the Sub-GC hot path")
    obj_classes = getattr(opt, "sg_obj_cnt", None) or _count_names(getattr(opt, "obj_name_path", None),
                                                                    DEFAULT_OBJ_CLASSES)
    pred_classes = getattr(opt, "sg_pred_cnt", None) or _count_names(getattr(opt, "rel_name_path", None),
                                                                      DEFAULT_PRED_CLASSES)
    return Dims(
        vocab=opt.vocab_size, enc=opt.input_encoding_size, rnn=opt.rnn_size, att_hid=opt.att_hid_size,
        fc_feat=opt.fc_feat_size, att_feat=opt.att_feat_size, gcn=opt.gcn_dim, embed=opt.embed_dim,
        obj_classes=obj_classes, pred_classes=pred_classes,
        gcn_layers=opt.gcn_layers, gcn_residual=opt.gcn_residual, pred_emb_type=opt.pred_emb_type,
        seq_length=(getattr(opt, "max_length", 0) or opt.seq_length),
        obj_num=getattr(opt, "obj_num", 37), rel_num=getattr(opt, "rel_num", 65),
        low_rank=getattr(opt, "gcn_low_rank", GCN_LOW_RANK),
    )


def make_opt(dims: Dims | None = None, **overrides) -> SimpleNamespace:
    """An `opt` namespace as train.py/test.py would hand to `models.setup` (Sub_GC_Kar values)."""
    d = dims or Dims()
    opt = SimpleNamespace(
        caption_model="topdown", vocab_size=d.vocab, input_encoding_size=d.enc, rnn_size=d.rnn, num_layers=1,
        drop_prob_lm=0.5, max_length=d.seq_length, seq_length=d.seq_length, fc_feat_size=d.fc_feat,
        att_feat_size=d.att_feat, att_hid_size=d.att_hid, use_bn=0, sampling_prob=0.0, use_gpn=1,
        embed_dim=d.embed, gcn_dim=d.gcn, noun_fuse=1, pred_emb_type=d.pred_emb_type, gcn_layers=d.gcn_layers,
        gcn_residual=d.gcn_residual, gcn_bn=0, obj_name_path=None, rel_name_path=None,
        sg_obj_cnt=d.obj_classes, sg_pred_cnt=d.pred_classes, obj_num=d.obj_num, rel_num=d.rel_num,
        gcn_low_rank=d.low_rank,
        # test-time flags (test.py:30-169)
        test_LSTM=0, use_topk_sampling=0, topk_temp=0.6, the_k=3, sct=0, gpn_nms_thres=0.75, gpn_max_subg=1,
        use_gt_subg=0, start_from=None, id="topdown",
    )
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt


SMALL = Dims(vocab=61, enc=24, rnn=40, att_hid=16, fc_feat=48, att_feat=48, gcn=24, low_rank=512, embed=12,
             obj_classes=23, pred_classes=7, seq_length=8, obj_num=37, rel_num=65)
"""Tiny shape set used by golden fixtures (every dimension deliberately not a multiple of a tile size).
low_rank stays 512 because the reference hard-codes it (models/lib/graph_conv.py:11); obj_num stays 37 because the
reference's NMS asserts on the literal pad index 36 (models/lib/gpn.py:117-118)."""
