"""Full-GC variant of the reference's AttModel (train.sh:27-36: use_gpn=0, noun_fuse=0, pred_emb_type=2, gcn_layers=4,
gcn_residual=1, gcn_bn=1), inference: `FullGCModel(opt)(..., mode='sample')` = models/AttModel.py:236-326 on its `else` branches
(:196-206 / :261-271: no sGPN, mean-pool read-out over the full scene graph, one caption per image).

Everything runs through the same C ABI as the Sub-GC model (no ATen arithmetic on the path):
  feat_fusion  (AttModel.py:370-387, noun_fuse=0): x0 = obj_v_proj(att_feats) (no ReLU, no class embedding); predicates embedded from
               the arg-max class INCLUDING background (pred_emb_type=2)            subgc_gemm_nt / subgc_class_argmax / subgc_gather_rows
  gcn_backbone (gcn_backbone.py:29-53): every layer is a residual boundary; BatchNorm1d of a unit (graph_conv_unit.py:23-32) is an affine
               map in eval mode and folds, with fc_rgt(fc_lft(.)), into ONE [2L, L] weight per direction        subgc_gcn_forward (gcn_fold)
  read-out     torch.mean(att_feats, 1) -> read_out_proj                                      subgc_mean_nodes / subgc_gemm_nt
  _prepare_feature + decode loop                                                              subgc_gemm_nt / subgc_decode_sample

The reference decodes `att_feats[0:1]` only (its test loader feeds one image per call); a batch of B images here is B independent rows,
i.e. what B reference calls return.  Training (BatchNorm batch statistics) and beam search are not part of this variant.
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, lib, ptr
from .config import DEFAULT_PRED_CLASSES, GCN_LOW_RANK, Dims, _count_names

BN_EPS = 1e-5   # nn.BatchNorm1d default (graph_conv_unit.py:24)


def _holder(**mods):
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(k, v)
    return m


class _Unit(nn.Module):
    """Parameter holder of _Collection_Unit (models/lib/graph_conv_unit.py:12-26)."""

    def __init__(self, dim, low_rank, use_bn):
        super().__init__()
        self.fc_lft = nn.Linear(dim, low_rank)
        self.fc_rgt = nn.Linear(low_rank, dim)
        for lin in (self.fc_lft, self.fc_rgt):
            nn.init.normal_(lin.weight, 0.0, 0.001)
            nn.init.zeros_(lin.bias)
        if use_bn:
            self.bn = nn.BatchNorm1d(dim)


def fullgc_dims(opt) -> Dims:
    pred_classes = getattr(opt, "sg_pred_cnt", None) or _count_names(getattr(opt, "rel_name_path", None), DEFAULT_PRED_CLASSES)
    obj_classes = getattr(opt, "sg_obj_cnt", None) or 2
    return Dims(vocab=opt.vocab_size, enc=opt.input_encoding_size, rnn=opt.rnn_size, att_hid=opt.att_hid_size, fc_feat=opt.fc_feat_size,
                att_feat=opt.att_feat_size, gcn=opt.gcn_dim, embed=opt.embed_dim, obj_classes=obj_classes, pred_classes=pred_classes,
                gcn_layers=opt.gcn_layers, gcn_residual=opt.gcn_residual, pred_emb_type=opt.pred_emb_type,
                seq_length=(getattr(opt, "max_length", 0) or opt.seq_length), obj_num=getattr(opt, "obj_num", 37),
                rel_num=getattr(opt, "rel_num", 65), low_rank=getattr(opt, "gcn_low_rank", GCN_LOW_RANK))


def fold_unit_pair(sd, prefix_of, units, use_bn):
    """[2L, L] weight and [2L] bias of two collection units sharing their input: bn(fc_rgt(fc_lft(x))) in eval mode, fp64."""
    ws, bs = [], []
    for u in units:
        pre = prefix_of(u)
        wl, bl = sd[pre + "fc_lft.weight"].double(), sd[pre + "fc_lft.bias"].double()
        wr, br = sd[pre + "fc_rgt.weight"].double(), sd[pre + "fc_rgt.bias"].double()
        wf, bf = wr @ wl, wr @ bl + br
        if use_bn:
            g, b = sd[pre + "bn.weight"].double(), sd[pre + "bn.bias"].double()
            mu, var = sd[pre + "bn.running_mean"].double(), sd[pre + "bn.running_var"].double()
            k = g / torch.sqrt(var + BN_EPS)
            wf, bf = wf * k[:, None], (bf - mu) * k + b
        ws.append(wf)
        bs.append(bf)
    wf, bf = torch.cat(ws, 0), torch.cat(bs, 0)
    mx = float(wf.abs().max())
    e = 0 if not (mx > 0.0 and math.isfinite(mx)) else max(-40, min(40, int(math.floor(math.log2(1024.0 / mx)))))
    scale = float(2.0 ** e)
    return (wf * scale).float().contiguous(), (bf * scale).float().contiguous(), scale


class FullGCModel(nn.Module):
    """Drop-in for the reference's TopDownModel built with use_gpn=0 (Full-GC), inference (`mode='sample'`, beam_size 1)."""

    def __init__(self, opt):
        super().__init__()
        if getattr(opt, "use_gpn", 1) != 0 or getattr(opt, "noun_fuse", 1) != 0:
            raise NotImplementedError("FullGCModel is the use_gpn=0, noun_fuse=0 configuration (train.sh:27-36)")
        if getattr(opt, "use_bn", 0) != 0:
            raise NotImplementedError("use_bn (BatchNorm inside att_embed) is not used by any train.sh configuration")
        d = self.dims = fullgc_dims(opt)
        self.use_bn = getattr(opt, "gcn_bn", 0) != 0
        self.vocab_size, self.rnn_size, self.seq_length, self.num_layers = d.vocab, d.rnn, d.seq_length, 2
        self.drop_prob_lm = getattr(opt, "drop_prob_lm", 0.5)
        self.gpn = False
        self.topk_sampling = getattr(opt, "use_topk_sampling", 0) != 0
        self.topk_temp = getattr(opt, "topk_temp", 0.6)
        self.the_k = getattr(opt, "the_k", 3)
        # parameter holders named like the reference modules (state_dict contract): models/AttModel.py:70-120 with self.gpn False
        self.obj_v_proj = nn.Linear(d.att_feat, d.gcn)
        self.sg_pred_embed = nn.Embedding(d.pred_classes, d.embed)
        self.pred_emb_prj = nn.Linear(d.embed, d.gcn)
        gcn = nn.ModuleList()
        for _ in range(d.gcn_layers):
            gcn.append(_holder(gcn_collect=_holder(collect_units=nn.ModuleList([_Unit(d.gcn, d.low_rank, self.use_bn) for _ in range(4)]))))
        self.gcn_backbone = _holder(gcn=gcn)
        self.read_out_proj = nn.Sequential(nn.Linear(d.gcn, d.att_hid), nn.Linear(d.att_hid, 2 * d.gcn))
        self.logit = nn.Linear(d.rnn, d.v1)
        self.embed = nn.Sequential(nn.Embedding(d.v1, d.enc), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.fc_embed = nn.Sequential(nn.Linear(d.att_feat, d.fc_feat), nn.ReLU(), nn.Linear(d.fc_feat, d.rnn), nn.ReLU(),
                                      nn.Dropout(self.drop_prob_lm))
        self.att_embed = nn.Sequential(nn.Linear(d.gcn, d.rnn), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.ctx2att = nn.Linear(d.rnn, d.att_hid)
        self.core = _holder(attention=_holder(h2att=nn.Linear(d.rnn, d.att_hid), alpha_net=nn.Linear(d.att_hid, 1)),
                            att_lstm=nn.LSTMCell(d.enc + 2 * d.rnn, d.rnn), lang_lstm=nn.LSTMCell(2 * d.rnn, d.rnn))
        self._cdims = _lib.Dims(d.v1, d.enc, d.rnn, d.att_hid, d.fc_feat, d.att_feat, d.gcn, d.low_rank, d.embed, d.obj_classes,
                                d.pred_classes, d.gcn_layers, d.gcn_residual, d.pred_emb_type, d.seq_length, d.obj_num, d.rel_num)
        self._wkey = self._w = self._fold = self._ws = None
        self.last_steps = None

    def forward(self, *args, **kwargs):
        mode = kwargs.pop("mode", "forward")
        if mode != "sample":
            raise NotImplementedError("FullGCModel implements mode='sample' (inference); training needs BatchNorm batch statistics")
        return self._sample(*args, **kwargs)

    # ---- plumbing ---------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def _scratch(self, nbytes, dev):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(max(int(nbytes) + 256, 1 << 22), dtype=torch.uint8, device=dev)
        return self._ws

    def _linear(self, x, w, b, relu=False):
        L = lib()
        M, (N, K) = x.shape[0], w.shape
        out = torch.empty(M, N, device=x.device)
        ws = self._scratch(L.subgc_gemm_nt_workspace_bytes(M, N, K), x.device)
        check(L.subgc_gemm_nt(M, N, K, ptr(x), x.stride(0), None, ptr(w), w.stride(0), ptr(b), int(relu), 0, ptr(out), N, ptr(ws), ws.numel(),
                              self._st()), "subgc_gemm_nt")
        return out

    def _weights(self):
        """subgc_weights over the live parameters (+ the folded GCN weights), rebuilt when a tensor moved or changed."""
        sd = OrderedDict((n, t) for n, t in list(self.named_parameters()) + list(self.named_buffers()))
        key = tuple((t.data_ptr(), t._version) for t in sd.values())
        if self._wkey == key:
            return self._w
        for n, t in sd.items():
            if t.is_floating_point() and not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise _lib.SubgcError(f"parameter {n} must be a contiguous fp32 CUDA tensor (the path has no CPU implementation)")
        g = lambda n: sd[n].data_ptr()
        lin = lambda n: _lib.Linear(g(n + ".weight"), g(n + ".bias"))
        w = _lib.Weights()
        w.obj_v_proj = lin("obj_v_proj"); w.sg_pred_embed = g("sg_pred_embed.weight"); w.pred_emb_prj = lin("pred_emb_prj")
        w.logit = lin("logit"); w.embed = g("embed.0.weight")
        w.fc_embed0 = lin("fc_embed.0"); w.fc_embed2 = lin("fc_embed.2"); w.att_embed = lin("att_embed.0"); w.ctx2att = lin("ctx2att")
        w.h2att = lin("core.attention.h2att"); w.alpha_net = lin("core.attention.alpha_net")
        w.att_w_ih, w.att_w_hh = g("core.att_lstm.weight_ih"), g("core.att_lstm.weight_hh")
        w.att_b_ih, w.att_b_hh = g("core.att_lstm.bias_ih"), g("core.att_lstm.bias_hh")
        w.lang_w_ih, w.lang_w_hh = g("core.lang_lstm.weight_ih"), g("core.lang_lstm.weight_hh")
        w.lang_b_ih, w.lang_b_hh = g("core.lang_lstm.bias_ih"), g("core.lang_lstm.bias_hh")
        self._fold = {}
        with torch.no_grad():
            for l in range(self.dims.gcn_layers):
                pre = lambda u, l=l: f"gcn_backbone.gcn.{l}.gcn_collect.collect_units.{u}."
                for u in range(4):
                    w.gcn_lft[l][u] = lin(pre(u) + "fc_lft")
                    w.gcn_rgt[l][u] = lin(pre(u) + "fc_rgt")
                for dr, units in enumerate(((0, 1), (2, 3))):
                    wf, bf, scale = fold_unit_pair(sd, pre, units, self.use_bn)
                    self._fold[(l, dr)] = (wf, bf)
                    w.gcn_fold[l][dr] = _lib.Linear(wf.data_ptr(), bf.data_ptr())
                    w.gcn_fold_scale[l][dr] = scale
        self._wkey, self._w = key, w
        return w

    # ---- inference --------------------------------------------------------------------------------------------------------------
    def encode(self, att_feats, pred_dist, rel_ind):
        """feat_fusion (noun_fuse=0) + gcn_backbone: x_obj [B, N, L]."""
        L, d, w, cd = lib(), self.dims, self._weights(), self._cdims
        dev = att_feats.device
        B, N, K = att_feats.shape[0], d.obj_num, d.rel_num
        st = self._st()
        x0 = self._linear(att_feats.reshape(B * N, -1), self.obj_v_proj.weight, self.obj_v_proj.bias)
        cls = torch.empty(B * K, dtype=torch.int64, device=dev)
        check(L.subgc_class_argmax(B * K, d.pred_classes, 0 if d.pred_emb_type == 2 else 1, ptr(pred_dist), ptr(cls), st), "subgc_class_argmax")
        emb = torch.empty(B * K, d.embed, device=dev)
        check(L.subgc_gather_rows(B * K, d.embed, ptr(self.sg_pred_embed.weight), d.embed, ptr(cls), ptr(emb), 0, st), "subgc_gather_rows")
        p0 = self._linear(emb, self.pred_emb_prj.weight, self.pred_emb_prj.bias)
        x_obj = torch.empty(B, N, d.gcn, device=dev)
        ws = self._scratch(L.subgc_encoder_workspace_bytes(C.byref(cd), B), dev)
        check(L.subgc_gcn_forward(C.byref(cd), C.byref(w), B, ptr(x0), ptr(p0), ptr(rel_ind), ptr(x_obj), None, ptr(ws), ws.numel(), st),
              "subgc_gcn_forward")
        self._x0, self._p0 = x0.view(B, N, -1), p0.view(B, K, -1)
        return x_obj

    def _sample(self, fc_feats, att_feats, att_masks=None, trip_pred=None, obj_dist=None, obj_box=None, rel_ind=None, pred_fmap=None,
                pred_dist=None, gpn_obj_ind=None, gpn_pred_ind=None, gpn_nrel_ind=None, gpn_pool_mtx=None, opt={}):
        """models/AttModel.py:236-326 with self.gpn False.  Returns (seq [B, T], seqLogprobs [B, T], subgraph_score [B], keep_ind [B])."""
        if opt.get("beam_size", 1) != 1:
            raise NotImplementedError("FullGCModel decodes greedily or with top-k sampling (beam_size 1)")
        if not att_feats.is_cuda:
            raise _lib.SubgcError("FullGCModel runs on a CUDA device only (no CPU implementation)")
        L, d, cd = lib(), self.dims, self._cdims
        dev = att_feats.device
        att_feats = att_feats.float().contiguous()
        pred_dist, rel_ind = pred_dist.float().contiguous(), rel_ind.long().contiguous()
        B, N, T, H = att_feats.shape[0], d.obj_num, d.seq_length, d.rnn
        w = self._weights()
        x_obj = self.encode(att_feats, pred_dist.reshape(-1, d.pred_classes), rel_ind)
        st = self._st()
        read_out = torch.empty(B, d.gcn, device=dev)
        check(L.subgc_mean_nodes(B, N, d.gcn, ptr(x_obj), ptr(read_out), st), "subgc_mean_nodes")
        g_fc = self._linear(self._linear(read_out, self.read_out_proj[0].weight, self.read_out_proj[0].bias), self.read_out_proj[1].weight,
                            self.read_out_proj[1].bias)
        # att_masks[0:1, 0, 0] with the first 36 entries forced to 1 (AttModel.py:268-269); every image of the batch gets its own row
        m = att_masks.float()
        rows5 = m.shape[0] // B if m.shape[0] % B == 0 and m.shape[0] >= B else 1
        masks = m[::rows5][:B, 0, 0].clone().contiguous() if m.dim() == 4 else m[:B].clone().contiguous()
        masks[:, :36] = 1.0
        ln = int(masks.long().sum(1).max())            # clip_att (AttModel.py:348-354): the one host read of the call
        # _prepare_feature (AttModel.py:356-368): rows past a sequence's length are exact zeros (pack_wrapper)
        fc = self._linear(self._linear(g_fc, self.fc_embed[0].weight, self.fc_embed[0].bias, relu=True), self.fc_embed[2].weight,
                          self.fc_embed[2].bias, relu=True)
        att = self._linear(x_obj[:, :ln].reshape(B * ln, d.gcn), self.att_embed[0].weight, self.att_embed[0].bias, relu=True)
        masks = masks[:, :ln].contiguous()
        valid = (torch.arange(ln, device=dev).view(1, -1) < masks.long().sum(1).view(-1, 1)).float().view(B * ln, 1).expand(-1, H).contiguous()
        check(L.subgc_ew(0, att.numel(), ptr(att), ptr(valid), ptr(att), 0.0, st), "subgc_ew")
        p_att = self._linear(att, self.ctx2att.weight, self.ctx2att.bias)
        seq = torch.empty(B, T, dtype=torch.int64, device=dev)
        lps = torch.empty(B, T, device=dev)
        steps = torch.empty(1, dtype=torch.int32, device=dev)
        uniforms = None
        if self.topk_sampling:
            uniforms = opt.get("topk_uniforms")
            uniforms = uniforms.to(dev).float().contiguous() if uniforms is not None else torch.rand(T, B, device=dev)
        ws = self._scratch(L.subgc_decode_workspace_bytes(C.byref(cd), B, ln), dev)
        check(L.subgc_decode_sample(C.byref(cd), C.byref(w), B, ln, 1 if self.topk_sampling else 0, float(self.topk_temp), int(self.the_k), 0, 0,
                                    ptr(uniforms), ptr(fc), ptr(att), ptr(p_att), ptr(masks), ptr(seq), ptr(lps), None, ptr(steps), ptr(ws),
                                    ws.numel(), st), "subgc_decode_sample")
        self.last_steps = steps
        self.last_g_fc, self.last_fc, self.last_att, self.last_p_att, self.last_x_obj = g_fc, fc, att.view(B, ln, H), p_att.view(B, ln, -1), x_obj
        keep_ind = torch.arange(B, device=dev).type_as(gpn_obj_ind) if gpn_obj_ind is not None else torch.arange(B, device=dev)
        return seq, lps, torch.ones(B, device=dev), keep_ind


def make_fullgc_state_dict(model_or_shapes, seed):
    """Seeded random state_dict for a Full-GC model (tests / fixtures): Linear-like tensors U(+-1/sqrt(fan_in)), embeddings N(0,1),
    BatchNorm with non-trivial affine parameters and running statistics.  `model_or_shapes`: a module or {name: shape}."""
    shapes = {n: tuple(t.shape) for n, t in model_or_shapes.state_dict().items()} if isinstance(model_or_shapes, nn.Module) else model_or_shapes
    sd = OrderedDict()
    for i, (name, shape) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if name.endswith("num_batches_tracked"):
            t = torch.tensor(100, dtype=torch.int64)
        elif name.endswith("running_var"):
            t = torch.rand(shape, generator=g) * 1.5 + 0.25
        elif name.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.2
        elif ".bn.weight" in name:
            t = torch.rand(shape, generator=g) + 0.5
        elif ".bn.bias" in name:
            t = torch.randn(shape, generator=g) * 0.1
        elif name in ("sg_pred_embed.weight", "embed.0.weight"):
            t = torch.randn(shape, generator=g)
        else:
            fan_in = shape[1] if len(shape) == 2 else shapes[name[:-4] + "weight"][1] if name.endswith("bias") and name[:-4] + "weight" in shapes \
                else shapes[name.replace("bias", "weight")][1]
            t = (torch.rand(shape, generator=g) * 2 - 1) / (fan_in ** 0.5)
        sd[name] = t.contiguous()
    return sd
