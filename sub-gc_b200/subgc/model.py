"""Host-side mirror of the reference model API for the Sub-GC hot path.

`setup(opt)` -> `TopDownModel` keeps the surface of reference models/__init__.py:43-59 / models/AttModel.py /
models/CaptionModel.py: constructor from `opt`, identical `state_dict` keys and shapes (a reference checkpoint loads
with `load_state_dict`), `model(*tensors, mode='forward'|'sample', opt={...})` with the reference's positional
signatures and return tuples, `get_logprobs_state`, `init_hidden`, and the attributes the drivers touch
(`gpn`, `ss_prob`, `done_beams`, `seq_length`, `vocab_size`).  The modules below only *hold* parameters under the
reference's names — all arithmetic runs in the CUDA kernels behind the C ABI (include/subgc_b200.h); there is no
torch / CPU fallback and calls fail loudly when the library or a CUDA device is missing.

One extension over the reference: `mode='sample'` accepts B >= 1 images per call (the reference asserts B == 1,
models/lib/gpn.py:84); images are encoded / scored / NMS-ed independently and decoded as one batch, and the image
of every returned row is in `model.last_image_of_row`.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.nn as nn

from . import _lib, packing
from ._lib import check, lib, ptr
from .config import Dims, dims_from_opt


def _holder(**mods):
    m = nn.Module()
    for k, v in mods.items():
        m.add_module(k, v)
    return m


class _CollectionUnit(nn.Module):
    """Parameter holder for _Collection_Unit (reference models/lib/graph_conv_unit.py:5-26)."""

    def __init__(self, dim, low_rank):
        super().__init__()
        self.fc_lft = nn.Linear(dim, low_rank)
        self.fc_rgt = nn.Linear(low_rank, dim)
        for lin in (self.fc_lft, self.fc_rgt):  # normal_init(m, 0, 0.001) with zero bias
            nn.init.normal_(lin.weight, 0.0, 0.001)
            nn.init.zeros_(lin.bias)


def _gcn_backbone(layers, dim, low_rank):
    gcn = nn.ModuleList()
    for _ in range(layers):
        gcn.append(_holder(gcn_collect=_holder(collect_units=nn.ModuleList([_CollectionUnit(dim, low_rank) for _ in range(4)]))))
    return _holder(gcn=gcn)


class Workspace:
    """Grow-only device scratch buffer handed to the C ABI (which never allocates)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        nbytes = int(nbytes) + 256
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return self.buf


class _DecodePlan:
    """Static buffers + captured CUDA graph of one decode shape.  The decode loops issue ~250 dependent launches;
    replaying them as a graph removes the per-launch CPU cost and most of the inter-kernel gaps.  The C ABI never
    allocates or synchronises, so a call is capturable as is; every pointer it receives lives in this object."""

    def __init__(self, dev, d, n_rows, len_max):
        self.fc = torch.empty(n_rows, d.rnn, device=dev)
        self.g_fc = torch.empty(n_rows, 2 * d.gcn, device=dev)
        self.att = torch.empty(n_rows, len_max, d.rnn, device=dev)
        self.p_att = torch.empty(n_rows, len_max, d.att_hid, device=dev)
        self.masks = torch.empty(n_rows, len_max, device=dev)
        self.graph = None
        self.calls = 0
        self.launches = 0
        self.out = {}

    replayed_launches = 0   # kernels launched through graph replays (the C ABI's own counter only sees direct launches)

    def run(self, launch, use_graph):
        """launch() enqueues the C call on the current stream.  First call: eager (also sets kernel attributes);
        second call: capture; afterwards: replay."""
        self.calls += 1
        if not use_graph:
            launch()
        elif self.graph is not None:
            self.graph.replay()
            _DecodePlan.replayed_launches += self.launches
        elif self.calls == 1:
            launch()
        else:
            c0 = lib().subgc_launch_count()
            g = torch.cuda.CUDAGraph()
            # torch.cuda.graph caches ONE default capture stream per process (on the device that was current first): pass a stream
            # of the device this plan lives on
            with torch.cuda.graph(g, stream=_capture_stream(self.fc.device)):
                launch()
            self.launches = int(lib().subgc_launch_count() - c0)   # kernels per replay (counted while capturing; capture does not run them)
            self.graph = g
            g.replay()


_capture_streams = {}


def _capture_stream(dev):
    st = _capture_streams.get(dev.index)
    if st is None:
        st = _capture_streams[dev.index] = torch.cuda.Stream(device=dev)
    return st


class _TrainStep(torch.autograd.Function):
    """Autograd hook-up of the CUDA training step: forward = subgc.train.forward (saves activations), backward =
    subgc.train.backward (hand-written gradients of all parameters).  Inputs after `names` are the parameters, in order."""

    @staticmethod
    def forward(ctx, model, data, drop, names, *params):
        from . import train
        P = dict(zip(names, (p.detach() for p in params)))
        ops = model._train_ops()
        outputs, gpn_loss, score, saved = train.forward(ops, P, model._weights(), model.dims, data, drop, model.seq_per_img, ss=data.get("ss"),
                                                        loss=data.get("loss"))   # with `loss`, `outputs` is the scalar lang_loss
        ctx.pack = (model, ops, P, saved, names)
        ctx.mark_non_differentiable(score)
        return outputs, gpn_loss[0], score

    @staticmethod
    def backward(ctx, d_outputs, d_gpn_loss, _d_score):
        from . import train
        model, ops, P, saved, names = ctx.pack
        ctx.pack = None   # the saved activations (~1.5 GB per step at 160 sentences) go back to the allocator when this call returns
        if d_outputs is None:
            d_outputs = torch.zeros_like(saved["outputs"]) if saved.get("outputs") is not None else torch.zeros((), device=saved["x0"].device)
        dg = 0.0 if d_gpn_loss is None else float(d_gpn_loss)
        G = train.backward(ops, P, model.dims, saved, d_outputs.contiguous(), dg, reducer=getattr(model, "grad_reducer", None))
        return (None, None, None, None) + tuple(G.get(n) for n in names)


class _DeviceState:
    """Everything the model caches per device: scratch, weight tables, packed copies, plans / captured graphs, the fp16-range flag.
    nn.DataParallel (reference train.py:96-98) runs replicas of ONE module object's __dict__ in one thread per device: state kept on
    the module itself would be shared by those threads, state keyed by the device of the calling replica's parameters is not."""

    def __init__(self):
        self.ws = Workspace()
        self.wcache = None
        self.packs = packing.PackCache()
        self.pack_key = None
        self.ovf_dev = self.ovf_host = self.ovf_event = None
        self.early = self.early_key = None
        self.fold = self.fold_key = None
        self.pfold = self.pfold_key = None
        self.mega = self.mega_key = None
        self.plans = {}
        self.tops = None
        self.cur_plan = None
        self.params = None
        self.wfast = None


def _state_property(name):
    return property(lambda self: getattr(self._state(), name), lambda self, v: setattr(self._state(), name, v))


class TopDownModel(nn.Module):
    """Drop-in for reference `TopDownModel(AttModel(CaptionModel))`, Sub-GC configuration."""
    _ws = _state_property("ws")
    _wcache = _state_property("wcache")
    _packs = _state_property("packs")
    _pack_key = _state_property("pack_key")
    _ovf_dev = _state_property("ovf_dev")
    _ovf_host = _state_property("ovf_host")
    _ovf_event = _state_property("ovf_event")
    _early = _state_property("early")
    _early_key = _state_property("early_key")
    _fold = _state_property("fold")
    _fold_key = _state_property("fold_key")
    _pfold = _state_property("pfold")
    _pfold_key = _state_property("pfold_key")
    _mega = _state_property("mega")
    _mega_key = _state_property("mega_key")
    _plans = _state_property("plans")
    _tops = _state_property("tops")
    _cur_plan = _state_property("cur_plan")

    def _state(self):
        dev = self.logit.weight.device
        key = dev.index if dev.type == "cuda" else -1
        st = self._states.get(key)
        if st is None:
            st = self._states.setdefault(key, _DeviceState())
        return st

    def _named_params(self):
        """name -> tensor of every parameter of THIS module object.  A DataParallel replica has no registered Parameters (its weights
        are plain attributes, torch/nn/parallel/replicate.py), so the tensors are fetched by attribute path, not by named_parameters()."""
        st = self._state()
        cached = st.params
        if cached is not None and cached[0] is self.logit.weight and cached[1] is self.obj_v_proj.weight:
            return cached[2]
        out = {}
        for name in self._param_names:
            obj = self
            for part in name.split("."):
                obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
            out[name] = obj
        st.params = (self.logit.weight, self.obj_v_proj.weight, out)
        return out

    def __init__(self, opt):
        super().__init__()
        self.dims: Dims = dims_from_opt(opt)
        d = self.dims
        self.vocab_size = d.vocab
        self.input_encoding_size = d.enc
        self.rnn_size = d.rnn
        self.num_layers = 2
        self.drop_prob_lm = getattr(opt, "drop_prob_lm", 0.5)
        self.seq_length = d.seq_length
        self.fc_feat_size = d.fc_feat
        self.att_feat_size = d.att_feat
        self.att_hid_size = d.att_hid
        self.ss_prob = getattr(opt, "sampling_prob", 0.0)
        self.gpn = True
        self.GCN_dim = d.gcn
        self.test_LSTM = getattr(opt, "test_LSTM", 0) != 0
        self.topk_sampling = getattr(opt, "use_topk_sampling", 0) != 0
        self.topk_temp = getattr(opt, "topk_temp", 0.6)
        self.the_k = getattr(opt, "the_k", 3)
        self.sct = getattr(opt, "sct", 0) != 0
        self.seq_per_img = getattr(opt, "seq_per_img", 5)
        if getattr(opt, "use_gt_subg", 0) != 0:
            raise NotImplementedError("use_gt_subg (Sup. SCT model) is outside the Sub-GC hot path")

        # parameter holders, named exactly like the reference modules (state_dict contract)
        self.obj_v_proj = nn.Linear(d.att_feat, d.gcn)
        self.sg_obj_embed = nn.Embedding(d.obj_classes, d.embed)   # GloVe-initialised in the reference; a checkpoint overwrites it
        self.obj_emb_proj = nn.Linear(d.embed, d.gcn)
        self.sg_pred_embed = nn.Embedding(d.pred_classes, d.embed)
        self.pred_emb_prj = nn.Linear(d.embed, d.gcn)
        self.gcn_backbone = _gcn_backbone(d.gcn_layers, d.gcn, d.low_rank)
        gpn = nn.Module()
        gpn.gpn_fc = nn.Sequential(nn.Linear(2 * d.gcn, d.att_hid), nn.ReLU(inplace=True), nn.Dropout(0.5), nn.Linear(d.att_hid, 1))
        gpn.read_out_proj = nn.Sequential(nn.Linear(2 * d.gcn, d.att_hid), nn.Linear(d.att_hid, 2 * d.gcn))
        for lin in (gpn.gpn_fc[0], gpn.gpn_fc[3], gpn.read_out_proj[0], gpn.read_out_proj[1]):
            nn.init.zeros_(lin.bias)
        gpn.use_nms = not self.sct
        gpn.iou_thres = getattr(opt, "gpn_nms_thres", 0.75)
        gpn.max_subgraphs = getattr(opt, "gpn_max_subg", 1)
        gpn.test_LSTM = self.test_LSTM
        self.gpn_layer = gpn
        self.logit = nn.Linear(d.rnn, d.v1)
        self.embed = nn.Sequential(nn.Embedding(d.v1, d.enc), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.fc_embed = nn.Sequential(nn.Linear(d.att_feat, d.fc_feat), nn.ReLU(), nn.Linear(d.fc_feat, d.rnn), nn.ReLU(),
                                      nn.Dropout(self.drop_prob_lm))
        self.att_embed = nn.Sequential(nn.Linear(d.gcn, d.rnn), nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.ctx2att = nn.Linear(d.rnn, d.att_hid)
        self.core = _holder(attention=_holder(h2att=nn.Linear(d.rnn, d.att_hid), alpha_net=nn.Linear(d.att_hid, 1)),
                            att_lstm=nn.LSTMCell(d.enc + 2 * d.rnn, d.rnn), lang_lstm=nn.LSTMCell(2 * d.rnn, d.rnn))

        self.done_beams = []
        self.last_gpn_loss = None
        self.last_image_of_row = None
        self.last_steps = None
        self.last_host_results = None   # host copies of the most recent whole-call-graph results (see _sample_dyn)
        self._states = {}   # device index -> _DeviceState (shared by DataParallel replicas, which live on different devices)
        self.use_packed = True   # inference contractions read split-fp16 copies of the weights (subgc.packing)
        self.fused_loss = True   # LossWrapper in train mode: log-softmax + LanguageModelCriterion inside the decoder stage
        self.use_mega = True     # greedy / top-k loops of <= 128 rows run as one persistent kernel (csrc/mega_decode.cu)
        self.grad_reducer = None     # subgc.parallel.GradReducer: gradient all-reduce overlapped with the hand-written backward (one process per GPU)
        self.use_step_graph = True   # ... and then the whole call (encoder .. decode) replays as ONE CUDA graph without host round trips
        self.stage_events = None  # set to [] to collect (name, start_event, end_event) per stage (bench / profiling)
        self.dropout_enabled = True   # tests switch it off: Philox masks cannot match torch's RNG stream (SURVEY §7 hard part 5)
        self.force_train_path = False  # run the autograd-capable path in eval mode too (gradient parity tests)
        self.use_graphs = os.environ.get("SUBGC_NO_GRAPH", "0") != "1"
        self._cdims = _lib.Dims(d.v1, d.enc, d.rnn, d.att_hid, d.fc_feat, d.att_feat, d.gcn, d.low_rank, d.embed, d.obj_classes,
                                d.pred_classes, d.gcn_layers, d.gcn_residual, d.pred_emb_type, d.seq_length, d.obj_num, d.rel_num)
        self._param_names = tuple(n for n, _ in self.named_parameters())

    # ------------------------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------------------------
    def forward(self, *args, **kwargs):
        """Reference models/CaptionModel.py:21-26."""
        mode = kwargs.get("mode", "forward")
        if "mode" in kwargs:
            del kwargs["mode"]
        return getattr(self, "_" + mode)(*args, **kwargs)

    def init_hidden(self, bsz):
        """Reference models/AttModel.py:343-346."""
        w = self.logit.weight
        return (w.new_zeros(self.num_layers, bsz, self.rnn_size), w.new_zeros(self.num_layers, bsz, self.rnn_size))

    def _weights(self):
        """subgc_weights over the live parameter storage (rebuilt when a tensor moved, e.g. after .cuda() / load)."""
        params = self._named_params()
        st = self._state()
        # fast path of every call: nothing moved, nothing was updated in place, no switch was flipped since the last validation
        fast = (tuple([(p.data_ptr(), p._version) for p in params.values()]), self.training, self.use_packed, self.use_mega)
        if st.wfast == fast and st.wcache is not None:
            return st.wcache[1]
        key = tuple(k[0] for k in fast[0])
        if self._wcache is not None and self._wcache[0] == key:
            w = self._wcache[1]
            self._attach_packs(w, params)
            st.wfast = fast
            return w
        for n, p in params.items():
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise _lib.SubgcError(f"parameter {n} must be a contiguous fp32 CUDA tensor (got {p.dtype} on {p.device}); "
                                      "the Sub-GC path has no CPU implementation")
        g = lambda n: params[n].data_ptr()
        lin = lambda n: _lib.Linear(g(n + ".weight"), g(n + ".bias"))
        w = _lib.Weights()
        w.obj_v_proj = lin("obj_v_proj"); w.sg_obj_embed = g("sg_obj_embed.weight"); w.obj_emb_proj = lin("obj_emb_proj")
        w.sg_pred_embed = g("sg_pred_embed.weight"); w.pred_emb_prj = lin("pred_emb_prj")
        for l in range(self.dims.gcn_layers):
            for u in range(4):
                pre = f"gcn_backbone.gcn.{l}.gcn_collect.collect_units.{u}."
                w.gcn_lft[l][u] = lin(pre + "fc_lft")
                w.gcn_rgt[l][u] = lin(pre + "fc_rgt")
        w.gpn_fc0 = lin("gpn_layer.gpn_fc.0"); w.gpn_fc3 = lin("gpn_layer.gpn_fc.3")
        w.read_out0 = lin("gpn_layer.read_out_proj.0"); w.read_out1 = lin("gpn_layer.read_out_proj.1")
        w.logit = lin("logit"); w.embed = g("embed.0.weight")
        w.fc_embed0 = lin("fc_embed.0"); w.fc_embed2 = lin("fc_embed.2"); w.att_embed = lin("att_embed.0"); w.ctx2att = lin("ctx2att")
        w.h2att = lin("core.attention.h2att"); w.alpha_net = lin("core.attention.alpha_net")
        w.att_w_ih = g("core.att_lstm.weight_ih"); w.att_w_hh = g("core.att_lstm.weight_hh")
        w.att_b_ih = g("core.att_lstm.bias_ih"); w.att_b_hh = g("core.att_lstm.bias_hh")
        w.lang_w_ih = g("core.lang_lstm.weight_ih"); w.lang_w_hh = g("core.lang_lstm.weight_hh")
        w.lang_b_ih = g("core.lang_lstm.bias_ih"); w.lang_b_hh = g("core.lang_lstm.bias_hh")
        self._wcache = (key, w)
        self._plans.clear()  # captured graphs hold the old parameter addresses
        self._pack_key = None
        self._attach_packs(w, params)
        st.wfast = fast
        return w

    # weight matrices that are the W operand of a contraction on the inference path (embedding tables are A operands)
    _PACK_SKIP = ("sg_obj_embed.weight", "sg_pred_embed.weight", "embed.0.weight")

    def _attach_packs(self, w, params):
        """Inference only: split-fp16 copies of the weight matrices for the tensor cores (subgc.packing), refreshed when a
        parameter was updated in place; training steps read the fp32 parameters directly."""
        use = (not self.training) and self.use_packed and os.environ.get("SUBGC_H3", "1") != "0"
        if not use:
            if w.n_packs:
                w.packs, w.n_packs = None, 0
                w.h3_overflow = None
                w.lang_early_w = None
                self._plans.clear()
            w.mega, w.mega_bytes, w.mega_ctas = None, 0, 0
            w.prep_fold = _lib.Linear(None, None)
            for l in range(self.dims.gcn_layers):   # folded GCN weights are an inference-only derived copy
                for dr in range(2):
                    w.gcn_fold[l][dr] = _lib.Linear(None, None)
                    w.gcn_fold_scale[l][dr] = 0.0
            return
        dev = next(iter(params.values())).device
        if self._ovf_dev is None or self._ovf_dev.device != dev:
            self._ovf_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._ovf_host = torch.zeros(1, dtype=torch.int32).pin_memory()
            self._plans.clear()
        w.h3_overflow = self._ovf_dev.data_ptr()
        named = {n: p for n, p in params.items() if p.dim() == 2 and p.shape[0] >= 64 and n not in self._PACK_SKIP}
        H = self.rnn_size   # K segments of the un-concatenated LSTM inputs: [h_lang | fc | x_t] and [ctx | h_att]
        # derived tensor of the merged pre-attention contraction (include/subgc_b200.h: subgc_weights.lang_early_w):
        # [[h2att.weight, 0], [lang weight_ih[:, H:2H], lang weight_hh]], rebuilt when one of its sources changed
        src = (params["core.attention.h2att.weight"], params["core.lang_lstm.weight_ih"], params["core.lang_lstm.weight_hh"])
        ekey = tuple((t.data_ptr(), t._version) for t in src)
        if self._early_key != ekey:
            with torch.no_grad():
                top = torch.cat([src[0], src[0].new_zeros(src[0].shape[0], H)], 1)
                bot = torch.cat([src[1][:, H:2 * H], src[2]], 1)
                self._early = torch.cat([top, bot], 0).contiguous()
            self._early_key = ekey
            self._plans.clear()
        named["__lang_early"] = self._early
        w.lang_early_w = self._early.data_ptr()
        self._attach_gcn_fold(w, params, named)
        self._attach_prep_fold(w, params, named)
        arr, cnt = self._packs.build(named, {"core.att_lstm.weight_ih": [H, 2 * H], "core.lang_lstm.weight_ih": [H], "__lang_early": [H]})
        if self._pack_key != self._packs.array_key or not w.n_packs:
            w.packs, w.n_packs = arr, cnt
            self._pack_key = self._packs.array_key
            self._plans.clear()  # captured graphs hold the old packed-copy addresses
        self._attach_mega(w, params, dev)

    def _attach_prep_fold(self, w, params, named):
        """read_out_proj.0 -> read_out_proj.1 -> fc_embed.0 are three Linears in a row (reference models/lib/gpn.py:35-36,
        models/AttModel.py:109): folded in fp64 into one [FC, 2L] matrix for calls that do not need the intermediate `fc_feats`
        (include/subgc_b200.h: subgc_weights.prep_fold)."""
        if os.environ.get("SUBGC_PREP_FOLD", "1") == "0":
            w.prep_fold = _lib.Linear(None, None)
            return
        names = ("gpn_layer.read_out_proj.0", "gpn_layer.read_out_proj.1", "fc_embed.0")
        src = [params[n + k] for n in names for k in (".weight", ".bias")]
        key = tuple((t.data_ptr(), t._version) for t in src)
        if self._pfold_key != key:
            with torch.no_grad():
                w0, b0, w1, b1, w2, b2 = (t.double() for t in src)
                wf = w2 @ w1 @ w0
                bf = w2 @ (w1 @ b0 + b1) + b2
                self._pfold = (wf.float().contiguous(), bf.float().contiguous())
            self._pfold_key = key
            self._plans.clear()
        named["__prep_fold"] = self._pfold[0]
        w.prep_fold = _lib.Linear(self._pfold[0].data_ptr(), self._pfold[1].data_ptr())

    def _attach_gcn_fold(self, w, params, named):
        """Folded GCN weights (include/subgc_b200.h: subgc_weights.gcn_fold): fc_rgt(fc_lft(.)) of a _Collection_Unit is linear
        (reference models/lib/graph_conv_unit.py:29-31), so both units of a direction become one [2L, L] matrix, computed in fp64 and
        scaled by a power of two (undone exactly by the consumer).  Rebuilt when a source parameter changed."""
        Ln = self.dims.gcn_layers
        if os.environ.get("SUBGC_GCN_FOLD", "1") == "0" or Ln == 0:
            for l in range(Ln):
                for dr in range(2):
                    w.gcn_fold[l][dr] = _lib.Linear(None, None)
                    w.gcn_fold_scale[l][dr] = 0.0
            return
        pre = lambda l, u: f"gcn_backbone.gcn.{l}.gcn_collect.collect_units.{u}."
        src = [params[pre(l, u) + k] for l in range(Ln) for u in range(4) for k in ("fc_lft.weight", "fc_lft.bias", "fc_rgt.weight", "fc_rgt.bias")]
        fkey = tuple((t.data_ptr(), t._version) for t in src)
        if self._fold_key != fkey:
            fold = {}
            with torch.no_grad():
                for l in range(Ln):
                    for dr, units in enumerate(((0, 1), (2, 3))):
                        ws, bs = [], []
                        for u in units:
                            wl, bl = params[pre(l, u) + "fc_lft.weight"].double(), params[pre(l, u) + "fc_lft.bias"].double()
                            wr, br = params[pre(l, u) + "fc_rgt.weight"].double(), params[pre(l, u) + "fc_rgt.bias"].double()
                            ws.append(wr @ wl)
                            bs.append(wr @ bl + br)
                        wf, bf = torch.cat(ws, 0), torch.cat(bs, 0)
                        mx = float(wf.abs().max())
                        e = 0 if not (mx > 0.0 and math.isfinite(mx)) else max(0, min(40, int(math.floor(math.log2(1024.0 / mx)))))
                        scale = float(2.0 ** e)
                        fold[(l, dr)] = ((wf * scale).float().contiguous(), (bf * scale).float().contiguous(), scale)
            self._fold, self._fold_key = fold, fkey
            self._plans.clear()
        for (l, dr), (wf, bf, scale) in self._fold.items():
            named[f"__gcn_fold_{l}_{dr}"] = wf
            w.gcn_fold[l][dr] = _lib.Linear(wf.data_ptr(), bf.data_ptr())
            w.gcn_fold_scale[l][dr] = scale

    _MEGA_SOURCES = ("core.att_lstm.weight_ih", "core.att_lstm.weight_hh", "core.lang_lstm.weight_ih", "core.lang_lstm.weight_hh",
                     "core.attention.h2att.weight", "logit.weight")

    def _attach_mega(self, w, params, dev):
        """Schedule tables + stream pack of the persistent decode kernel (include/subgc_b200.h: subgc_mega_pack), rebuilt when one of
        the six decoder weight matrices changed.  Dimensions the kernel does not take (or weights beyond the fp16 range) leave
        w.mega NULL: the decode loop then runs one launch per stage."""
        if os.environ.get("SUBGC_MEGA", "1") == "0" or not self.use_mega:
            if w.mega:
                self._plans.clear()
            w.mega, w.mega_bytes, w.mega_ctas = None, 0, 0
            return
        key = tuple((params[n].data_ptr(), params[n]._version) for n in self._MEGA_SOURCES) + (dev.index,)
        if self._mega_key != key:
            L, cd = lib(), self._cdims
            n_cta = torch.cuda.get_device_properties(dev).multi_processor_count
            nbytes = int(L.subgc_mega_pack_bytes(C.byref(cd), n_cta))
            self._mega = None
            if nbytes:
                buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
                base = (buf.data_ptr() + 1023) // 1024 * 1024
                flag = torch.zeros(1, dtype=torch.int32, device=dev)
                check(L.subgc_mega_pack(C.byref(cd), C.byref(w), n_cta, base, nbytes, ptr(flag), self._stream()), "subgc_mega_pack")
                if int(flag.item()) == 0:   # one host read at parameter-load time, never inside a decode loop
                    self._mega = (buf, base, nbytes, n_cta)
            self._mega_key = key
            self._plans.clear()
        if self._mega is None:
            w.mega, w.mega_bytes, w.mega_ctas = None, 0, 0
        else:
            if not w.mega:
                self._plans.clear()
            w.mega, w.mega_bytes, w.mega_ctas = self._mega[1], self._mega[2], self._mega[3]

    def _arm_overflow_check(self):
        """Queue an asynchronous read-back of the fp16-range flag the split-fp16 kernels raise (no synchronisation here)."""
        if self._ovf_dev is not None and self._wcache is not None and self._wcache[1].n_packs:
            self._ovf_host.copy_(self._ovf_dev, non_blocking=True)
            self._ovf_event = torch.cuda.Event()
            self._ovf_event.record()

    def check_numerics(self, block=True):
        """The split-fp16 tensor-core path saturates activations beyond +-65504 and raises a device flag.  Polled at the start of
        every call (non-blocking) and on demand (block=True): if the flag is up, the packed path is switched off for this model
        and the caller is told that the previous results are invalid."""
        ev = self._ovf_event
        if ev is None or not (block or ev.query()):
            return
        ev.synchronize()
        self._ovf_event = None
        if int(self._ovf_host[0]) != 0:
            self._ovf_dev.zero_()
            self.use_packed = False
            raise _lib.SubgcError("an activation exceeded the fp16 range of the split-fp16 tensor-core path (|x| > 65504): the results of "
                                  "the previous call are invalid.  Packed weights are now disabled for this model (fp32 split-TF32 "
                                  "path); run the call again.")

    def _check_call(self, steps_dev):
        """Status of the decode call that was just enqueued, read back with ONE small blocking copy: (executed steps, fp16-range
        flag).  Results of a call whose activations saturated the split-fp16 path are never handed out: the packed path is switched
        off and None is returned so that the caller repeats the call on the fp32 path (the reference has no such failure mode)."""
        have_flag = self._ovf_dev is not None and self._wcache is not None and self._wcache[1].n_packs
        st = torch.cat([steps_dev.view(1), self._ovf_dev.view(1)]) if have_flag else steps_dev.view(1)
        st_h = st.cpu()
        steps = int(st_h[0])
        if steps < 0:
            raise _lib.SubgcError(f"persistent decode kernel timed out (wait site {(-steps) // 1000}, CTA {(-steps) % 1000 - 1}); "
                                  "set SUBGC_MEGA=0 to decode with one launch per stage")
        if have_flag and int(st_h[1]) != 0:
            import warnings
            self._ovf_dev.zero_()
            self.use_packed = False
            warnings.warn("subgc: an activation exceeded the fp16 range of the split-fp16 tensor-core path (|x| > 65504); "
                          "repeating the call on the fp32 (split-TF32) path, which this model keeps using from now on")
            return None
        return steps

    def _train_ops(self):
        from .train import CudaOps
        if self._tops is None:
            self._tops = CudaOps(self._cdims, self.seq_per_img)
        return self._tops

    @staticmethod
    def _stream():
        return torch.cuda.current_stream().cuda_stream

    def _mark(self, name=None):
        """Stage timing hook: _mark() opens a region on the current stream, _mark(name) closes it."""
        if self.stage_events is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if name is None:
            self._open_ev = ev
        else:
            self.stage_events.append((name, self._open_ev, ev))
            self._open_ev = ev

    def _f32(self, t):
        return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t

    def _i64(self, t):
        return t.contiguous().long() if (t.dtype != torch.int64 or not t.is_contiguous()) else t

    def _check_device(self, *tensors):
        dev = self.logit.weight.device
        if dev.type != "cuda":
            raise _lib.SubgcError("the model must live on a CUDA device (model.cuda()); there is no CPU path")
        for t in tensors:
            if t is not None and t.device != dev:
                raise _lib.SubgcError(f"input tensor on {t.device}, model on {dev}")
        return dev

    # ------------------------------------------------------------------------------------------------------------
    # stages (each one is one C-ABI call)
    # ------------------------------------------------------------------------------------------------------------
    def encode(self, att_feats, obj_dist, pred_dist, rel_ind, want_x_pred=False, obj_cls=None, pred_cls=None):
        """feat_fusion + gcn_backbone (reference models/AttModel.py:370-387, models/lib/gcn_backbone.py:29-53) WITHOUT the x5
        replication.  Returns x_obj [B,N,L] (and x_pred [B,K,L] when asked).  obj_cls / pred_cls (int64 [B,N] / [B,K]): class ids
        known from the loader (subgc.compact) instead of the score tensors."""
        dev = self._check_device(att_feats, obj_dist, pred_dist, rel_ind, obj_cls, pred_cls)
        L, d, w, cd = lib(), self.dims, self._weights(), self._cdims
        att_feats, rel_ind = self._f32(att_feats), self._i64(rel_ind)
        if obj_cls is None:
            obj_dist = self._f32(obj_dist)
        B = att_feats.shape[0]
        need_pred = bool(L.subgc_gcn_needs_pred(C.byref(cd), int(want_x_pred)))
        x0 = torch.empty(B, d.obj_num, d.gcn, device=dev)
        p0 = torch.empty(B, d.rel_num, d.gcn, device=dev) if need_pred else None
        x_obj = torch.empty_like(x0)
        x_pred = torch.empty(B, d.rel_num, d.gcn, device=dev) if want_x_pred else None
        wsb = L.subgc_encoder_workspace_bytes(C.byref(cd), B)
        ws = self._ws.get(wsb, dev)
        st = self._stream()
        if obj_cls is not None:
            if need_pred and pred_cls is None:
                raise _lib.SubgcError("this configuration needs the predicate classes (compact_batch(..., with_pred=True))")
            check(L.subgc_fuse_nodes_cls(C.byref(cd), C.byref(w), B, ptr(att_feats), ptr(self._i64(obj_cls)),
                                         ptr(self._i64(pred_cls)) if need_pred else None, ptr(x0), ptr(p0), ptr(ws), ws.numel(), st),
                  "subgc_fuse_nodes_cls")
        else:
            pd = self._f32(pred_dist) if need_pred else None
            check(L.subgc_fuse_nodes(C.byref(cd), C.byref(w), B, ptr(att_feats), ptr(obj_dist), ptr(pd), ptr(x0), ptr(p0), ptr(ws),
                                     ws.numel(), st), "subgc_fuse_nodes")
        check(L.subgc_gcn_forward(C.byref(cd), C.byref(w), B, ptr(x0), ptr(p0), ptr(rel_ind), ptr(x_obj), ptr(x_pred), ptr(ws),
                                  ws.numel(), st), "subgc_gcn_forward")
        self._x0, self._p0 = x0, p0
        return (x_obj, x_pred) if want_x_pred else x_obj

    def _sgpn(self, x_obj, gpn_obj_ind, att_masks, order, seq_per_img=None):
        dev = x_obj.device
        L, w, cd = lib(), self._weights(), self._cdims
        rows, _, per_half, _ = gpn_obj_ind.shape
        spi = seq_per_img or self.seq_per_img
        lay = _lib.Layout(rows, per_half, spi, order)
        n_sub = 2 * rows * per_half if order == 0 else 2 * (rows // spi) * per_half
        read_out = torch.empty(n_sub, 2 * self.dims.gcn, device=dev)
        score = torch.empty(n_sub, device=dev)
        sub_len = torch.empty(n_sub, dtype=torch.int32, device=dev)
        loss = torch.empty(1, device=dev)
        ws = self._ws.get(L.subgc_sgpn_workspace_bytes(C.byref(cd), n_sub), dev)
        check(L.subgc_sgpn_forward(C.byref(cd), C.byref(w), C.byref(lay), ptr(x_obj), ptr(gpn_obj_ind), ptr(att_masks), ptr(read_out),
                                   ptr(score), ptr(sub_len), ptr(loss), ptr(ws), ws.numel(), self._stream()), "subgc_sgpn_forward")
        return lay, n_sub, read_out, score, sub_len, loss

    def _plan(self, dev, n_rows, len_max, kind):
        """Decode plan (static buffers + graph) for a shape; a handful of shapes are kept."""
        key = (dev.index, n_rows, len_max, kind)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            plan = self._plans[key] = _DecodePlan(dev, self.dims, n_rows, len_max)
        return plan

    def _prepare(self, lay, n_rows, len_max, sel, x_obj, gpn_obj_ind, att_masks, read_out, plan=None, want_g_fc=True):
        """want_g_fc=False: the reference's intermediate `fc_feats` is not materialised (read_out_proj and fc_embed.0 as one folded
        contraction, subgc_weights.prep_fold); the returned g_fc is then None."""
        dev = x_obj.device
        L, d, w, cd = lib(), self.dims, self._weights(), self._cdims
        if plan is None:
            plan = _DecodePlan(dev, d, n_rows, len_max)
        g_fc, fc, att, p_att, masks = plan.g_fc, plan.fc, plan.att, plan.p_att, plan.masks
        if not want_g_fc and w.prep_fold.w:
            g_fc = None
        ws = self._ws.get(L.subgc_prepare_workspace_bytes(C.byref(cd), n_rows, len_max), dev)
        check(L.subgc_prepare_forward(C.byref(cd), C.byref(w), C.byref(lay), n_rows, len_max, ptr(sel), ptr(x_obj), ptr(gpn_obj_ind),
                                      ptr(att_masks), ptr(read_out), ptr(g_fc), ptr(fc), ptr(att), ptr(p_att), ptr(masks), ptr(ws),
                                      ws.numel(), self._stream()), "subgc_prepare_forward")
        return g_fc, fc, att, p_att, masks

    def _front(self, att_feats, att_masks, obj_dist, rel_ind, pred_dist, gpn_obj_ind, plan_kind=None, seq_per_img=None, obj_cls=None,
               pred_cls=None):
        """Encoder + sGPN + NMS + feature preparation for inference.  One host read-back (kept count, clip length) —
        the reference synchronises at the same two places (NMS on the host, clip_att's .max())."""
        dev = self._check_device(att_feats, att_masks, obj_dist, rel_ind, gpn_obj_ind)
        L, cd = lib(), self._cdims
        gpn_obj_ind, att_masks = self._i64(gpn_obj_ind), self._f32(att_masks)
        self._mark()
        x_obj = self.encode(att_feats, obj_dist, pred_dist, rel_ind, obj_cls=obj_cls, pred_cls=pred_cls)
        self._mark("encode")
        lay, n_sub, read_out, score, sub_len, loss = self._sgpn(x_obj, gpn_obj_ind, att_masks, order=1, seq_per_img=seq_per_img)
        n_images = lay.rows // lay.seq_per_img
        P = 2 * lay.per_half
        sel = torch.empty(n_sub, dtype=torch.int32, device=dev)
        keep = torch.empty(n_sub, dtype=torch.int64, device=dev)
        stats = torch.empty(2 + n_images, dtype=torch.int32, device=dev)
        ws = self._ws.get(L.subgc_nms_workspace_bytes(n_images, P), dev)
        g = self.gpn_layer
        check(L.subgc_subgraph_nms(C.byref(cd), C.byref(lay), ptr(score), ptr(sub_len), ptr(gpn_obj_ind), ptr(att_masks), int(bool(g.use_nms)),
                                   float(g.iou_thres), int(g.max_subgraphs), ptr(sel), ptr(keep), ptr(stats), ptr(ws), ws.numel(),
                                   self._stream()), "subgc_subgraph_nms")
        self._mark("sgpn_nms")
        stats_h = stats.cpu()
        n_rows, len_max = int(stats_h[0]), int(stats_h[1])
        sel = sel[:n_rows]
        keep = keep[:n_rows]
        self._mark()
        self._cur_plan = self._plan(dev, n_rows, len_max, plan_kind) if plan_kind is not None else None
        prep = self._prepare(lay, n_rows, len_max, sel, x_obj, gpn_obj_ind, att_masks, read_out, self._cur_plan)
        self._mark("prepare")
        sel_l = sel.long()
        self.last_gpn_loss = loss[0]
        self.last_image_of_row = torch.div(sel_l, P, rounding_mode="floor")
        self.last_x_obj = x_obj
        self.last_all_scores = score
        # the reference returns a float arange when NMS is off (models/lib/gpn.py:97) and int64 indices otherwise (:136)
        keep_ind = keep if g.use_nms else keep.to(score.dtype)
        return prep, score[sel_l], keep_ind, n_rows, len_max

    # ------------------------------------------------------------------------------------------------------------
    # reference entry points
    # ------------------------------------------------------------------------------------------------------------
    def _sample(self, fc_feats, att_feats, att_masks=None, trip_pred=None, obj_dist=None, obj_box=None, rel_ind=None, pred_fmap=None,
                pred_dist=None, gpn_obj_ind=None, gpn_pred_ind=None, gpn_nrel_ind=None, gpn_pool_mtx=None, opt={}):
        """Reference models/AttModel.py:236-326 (greedy / top-k) and :179-234 (beam).  Returns
        (seq, seqLogprobs, subgraph_score, keep_ind[, att2_weights])."""
        return self._sample_impl(dict(att_feats=att_feats, att_masks=att_masks, obj_dist=obj_dist, rel_ind=rel_ind, pred_dist=pred_dist,
                                      gpn_obj_ind=gpn_obj_ind), opt)

    def _sample_compact(self, batch, opt={}):
        """`mode='sample_compact'`: AttModel._sample (models/AttModel.py:236-326) from a subgc.compact.CompactBatch on the device --
        class ids instead of score tensors, node lists + lengths instead of masks / pooling matrices, one copy per image.  Same
        return tuple, identical results (the few index / mask tensors the kernels read are rebuilt on the device: ~100 KB)."""
        return self._sample_impl(dict(compact=batch), opt)

    @staticmethod
    def _expand_compact(batch):
        """CompactBatch -> the operands of the stages (device ops only: capturable)."""
        N = batch.sub_nodes.shape[-1]
        masks = (torch.arange(N, device=batch.sub_len.device).view(1, 1, 1, N) < batch.sub_len.unsqueeze(-1)).float()
        return dict(att_feats=batch.att_feats, att_masks=masks, obj_dist=None, rel_ind=batch.rel_ind.long(), pred_dist=None,
                    gpn_obj_ind=batch.sub_nodes.long(), seq_per_img=1, obj_cls=batch.obj_cls.long(),
                    pred_cls=None if batch.pred_cls is None else batch.pred_cls.long())

    # ------------------------------------------------------------------------------------------------------------
    # whole-step graph: encoder -> sGPN -> NMS -> prepare -> persistent decode kernel without a host round trip
    # ------------------------------------------------------------------------------------------------------------
    def _dyn_eligible(self, front, opt):
        """The reference synchronises with the host after NMS (kept rows) and in clip_att (longest sub-graph).  With the persistent
        decode kernel both numbers stay on the device (subgc_decode_sample_dyn): every stage runs at the upper bound of the row count,
        the decode kernel reads the real one.  Returns (rows_cap, n_images) or None (beam search, attention-weight output, injected
        uniforms, more than 128 rows, per-stage timing requested, graphs or the persistent kernel switched off)."""
        if (opt.get("beam_size", 1) != 1 or opt.get("return_att", 0) == 1 or opt.get("topk_uniforms", None) is not None
                or not self.use_graphs or not self.use_mega or not self.use_step_graph or self.stage_events is not None):
            return None
        w = self._weights()
        if not w.mega:
            return None
        if "compact" in front:
            rows, _, per_half, _ = front["compact"].sub_nodes.shape
            n_images = rows
        else:
            rows, _, per_half, _ = front["gpn_obj_ind"].shape
            n_images = rows // (front.get("seq_per_img") or self.seq_per_img)
        g = self.gpn_layer
        per_image = min(int(g.max_subgraphs), 2 * per_half) if g.use_nms else 2 * per_half
        rows_cap = n_images * per_image
        if rows_cap < 1 or rows_cap > 128 or rows_cap > w.mega_ctas:
            return None
        return rows_cap, n_images

    def _step_body(self, plan, front, rows_cap, n_images):
        """Launch sequence of one whole step at the row-count upper bound; every pointer it uses lives in `plan` or in `front`."""
        if "compact" in front:
            front = self._expand_compact(front["compact"])
        dev = front["att_feats"].device
        L, w, cd, T, N = lib(), self._weights(), self._cdims, self.seq_length, self.dims.obj_num
        gpn_obj_ind, att_masks = front["gpn_obj_ind"], front["att_masks"]
        x_obj = self.encode(front["att_feats"], front["obj_dist"], front["pred_dist"], front["rel_ind"], obj_cls=front.get("obj_cls"),
                            pred_cls=front.get("pred_cls"))
        lay, n_sub, read_out, score, sub_len, loss = self._sgpn(x_obj, gpn_obj_ind, att_masks, order=1, seq_per_img=front.get("seq_per_img"))
        P = 2 * lay.per_half
        sel = torch.zeros(max(n_sub, rows_cap), dtype=torch.int32, device=dev)   # rows beyond the kept count stay at sub-graph 0 (valid, unused)
        if plan.pack is None:
            # every result of the call lives in ONE buffer (typed views below): the caller's private copy is one device-to-device copy
            # instead of one per tensor (6 dependent 2 us copies at the end of a 1.8 ms call)
            spec = (("seq", torch.int64, (rows_cap, T)), ("keep", torch.int64, (n_sub,)), ("image", torch.int64, (rows_cap,)),
                    ("lps", torch.float32, (rows_cap, T)), ("sub_score", torch.float32, (rows_cap,)), ("status", torch.int32, (4,)))
            plan.pack_spec, off = [], 0
            for name, dt, shape in spec:
                nb = int(torch.tensor([], dtype=dt).element_size()) * int(torch.Size(shape).numel())
                plan.pack_spec.append((name, dt, shape, off, nb))
                off += (nb + 15) // 16 * 16
            plan.pack = torch.zeros(off, dtype=torch.uint8, device=dev)
            plan.pack_views = self._pack_views(plan.pack, plan.pack_spec)
            plan.out["seq"], plan.out["lps"] = plan.pack_views["seq"], plan.pack_views["lps"]
        pv = plan.pack_views
        keep = pv["keep"].zero_()
        stats = torch.zeros(2 + n_images, dtype=torch.int32, device=dev)
        ws = self._ws.get(L.subgc_nms_workspace_bytes(n_images, P), dev)
        g = self.gpn_layer
        check(L.subgc_subgraph_nms(C.byref(cd), C.byref(lay), ptr(score), ptr(sub_len), ptr(gpn_obj_ind), ptr(att_masks), int(bool(g.use_nms)),
                                   float(g.iou_thres), int(g.max_subgraphs), ptr(sel), ptr(keep), ptr(stats), ptr(ws), ws.numel(),
                                   self._stream()), "subgc_subgraph_nms")
        g_fc, fc, att, p_att, masks = self._prepare(lay, rows_cap, N, sel, x_obj, gpn_obj_ind, att_masks, read_out, plan, want_g_fc=False)
        o = plan.out
        check(L.subgc_decode_sample_dyn(C.byref(cd), C.byref(w), rows_cap, N, ptr(stats), 1 if self.topk_sampling else 0, float(self.topk_temp),
                                        int(self.the_k), 0, 0, ptr(o["uniforms"]), ptr(fc), ptr(att), ptr(p_att), ptr(masks), ptr(o["seq"]),
                                        ptr(o["lps"]), ptr(o["steps"]), ptr(o["ws"]), o["ws"].numel(), self._stream()), "subgc_decode_sample_dyn")
        have_flag = self._ovf_dev is not None and w.n_packs
        status = torch.cat([o["steps"], self._ovf_dev if have_flag else torch.zeros(1, dtype=torch.int32, device=dev), stats[:2]], out=pv["status"])
        sel_l = sel[:rows_cap].long()
        torch.index_select(score, 0, sel_l, out=pv["sub_score"])
        torch.div(sel_l, P, rounding_mode="floor", out=pv["image"])
        return dict(status=status, seq=o["seq"], lps=o["lps"], score=score, sub_score=pv["sub_score"], keep=keep, sel=sel_l, loss=loss, x_obj=x_obj,
                    image=pv["image"])

    @staticmethod
    def _pack_views(buf, spec):
        return {name: buf[off:off + nb].view(dt).view(shape) for name, dt, shape, off, nb in spec}

    _STEP_INPUTS = ("att_feats", "att_masks", "obj_dist", "rel_ind", "pred_dist", "gpn_obj_ind", "obj_cls", "pred_cls")

    def _sample_dyn(self, front, opt, rows_cap, n_images):
        front = dict(front)
        batch = front.get("compact")
        if batch is not None:
            from dataclasses import fields as _fields
            names = [f.name for f in _fields(batch) if getattr(batch, f.name) is not None]
            tensors = [getattr(batch, n).contiguous() for n in names]
            dev = self._check_device(*tensors)
        else:
            dev = self._check_device(front["att_feats"], front["att_masks"], front["obj_dist"], front["rel_ind"], front["gpn_obj_ind"])
            front["gpn_obj_ind"], front["att_masks"] = self._i64(front["gpn_obj_ind"]), self._f32(front["att_masks"])
            front["att_feats"], front["rel_ind"] = self._f32(front["att_feats"]), self._i64(front["rel_ind"])
            front["obj_dist"] = self._f32(front["obj_dist"])
            front["pred_dist"] = None if front.get("pred_dist") is None else self._f32(front["pred_dist"])
            names = [k for k in self._STEP_INPUTS if front.get(k) is not None]
            tensors = [front[k] for k in names]
        g = self.gpn_layer   # (_dyn_eligible validated the weight tables and dropped stale plans a moment ago)
        skey = ("step", dev.index, batch is not None, tuple(names), tuple(tuple(t.shape) for t in tensors), front.get("seq_per_img"), bool(self.topk_sampling),
                float(self.topk_temp), int(self.the_k), bool(g.use_nms), float(g.iou_thres), int(g.max_subgraphs))
        group = self._plans.get(skey)
        if group is None:
            while len(self._plans) >= 8:
                self._plans.pop(next(iter(self._plans)))
            group = self._plans[skey] = {"ptr": {}, "copy": None}
        # A captured graph holds the addresses of its inputs.  Callers that pass the same device buffers again (a loader with a fixed
        # set of staging buffers) get one zero-copy plan per buffer set; any other caller goes through one plan with its own input
        # buffers, filled by a device-to-device copy per call.
        pkey = tuple(t.data_ptr() for t in tensors)
        plan = group["ptr"].get(pkey)
        if plan is None:
            if len(group["ptr"]) < 4:
                plan = group["ptr"][pkey] = self._new_step_plan(dev, rows_cap, tensors, None)
            else:
                if group["copy"] is None:
                    group["copy"] = self._new_step_plan(dev, rows_cap, None, [torch.empty_like(t) for t in tensors])
                plan = group["copy"]
        if plan.static_in is not None:
            for dst, src in zip(plan.static_in, tensors):
                dst.copy_(src, non_blocking=True)
            tensors = plan.static_in
        if batch is not None:
            from .compact import CompactBatch
            front["compact"] = CompactBatch(**{**{f.name: None for f in _fields(batch)}, **dict(zip(names, tensors))})
        else:
            front.update(dict(zip(names, tensors)))
        if self.topk_sampling:   # the kernel maps uniforms to tokens by inverse CDF over the kept candidates
            if "seed" in opt:
                gen = torch.Generator(device=dev)
                gen.manual_seed(int(opt["seed"]))
                plan.out["uniforms"].copy_(torch.rand(self.seq_length, rows_cap, device=dev, generator=gen))
            else:
                plan.out["uniforms"].uniform_()
        plan.calls += 1
        saved_ws, self._ws = self._ws, plan.ws_obj   # the captured launches keep pointing into this plan's own workspace
        try:
            if plan.graph is not None:
                plan.graph.replay()
                _DecodePlan.replayed_launches += plan.launches
                outs = plan.outs
            elif plan.calls == 1:
                outs = self._step_body(plan, front, rows_cap, n_images)   # eager: also sets kernel attributes and sizes the workspace
            else:
                c0 = lib().subgc_launch_count()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=_capture_stream(dev)):
                    plan.outs = self._step_body(plan, front, rows_cap, n_images)
                plan.launches = int(lib().subgc_launch_count() - c0)
                plan.graph = graph
                graph.replay()
                _DecodePlan.replayed_launches += plan.launches
                outs = plan.outs
        finally:
            self._ws = saved_ws
        # private copies of the results are queued BEFORE the host waits: the launches overlap with the graph's execution
        res = self._pack_views(plan.pack.clone(), plan.pack_spec)
        # the one host round trip of the call: the whole result pack (32 KB at 128 rows) in ONE device-to-host copy -- status = (steps,
        # fp16-range flag, rows, longest sub-graph) decides the shapes returned below, and callers that want the results on the host
        # (eval loops do) find them in `last_host_results` without further copies (valid until the second next call with this shape)
        if plan.host_pack is None:
            plan.host_pack = [torch.empty(plan.pack.numel(), dtype=torch.uint8).pin_memory() for _ in range(2)]
        hp = plan.host_pack[plan.calls & 1]
        hp.copy_(plan.pack, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        host = self._pack_views(hp, plan.pack_spec)
        st_h = host["status"]
        steps, ovf, n_rows = int(st_h[0]), int(st_h[1]), int(st_h[2])
        if steps < 0:
            raise _lib.SubgcError(f"persistent decode kernel timed out (wait site {(-steps) // 1000}, CTA {(-steps) % 1000 - 1}); "
                                  "set SUBGC_MEGA=0 to decode with one launch per stage")
        if ovf != 0:
            import warnings
            self._ovf_dev.zero_()
            self.use_packed = False
            warnings.warn("subgc: an activation exceeded the fp16 range of the split-fp16 tensor-core path (|x| > 65504); "
                          "repeating the call on the fp32 (split-TF32) path, which this model keeps using from now on")
            return self._sample_impl(front, opt)
        self.last_steps = res["status"][:1]
        self.last_gpn_loss = outs["loss"][0]       # these three alias the plan's buffers: valid until the next call with this shape
        self.last_x_obj = outs["x_obj"]
        self.last_all_scores = outs["score"]
        self.last_image_of_row = res["image"][:n_rows]
        keep = res["keep"][:n_rows]
        self.last_host_results = dict(seq=host["seq"][:n_rows], seqLogprobs=host["lps"][:n_rows], subgraph_score=host["sub_score"][:n_rows],
                                      keep_ind=host["keep"][:n_rows], image=host["image"][:n_rows])
        keep_ind = keep if self.gpn_layer.use_nms else keep.to(outs["score"].dtype)
        return res["seq"][:n_rows], res["lps"][:n_rows], res["sub_score"][:n_rows], keep_ind

    def _new_step_plan(self, dev, rows_cap, keepalive, static_in):
        L, cd, T, N = lib(), self._cdims, self.seq_length, self.dims.obj_num
        plan = _DecodePlan(dev, self.dims, rows_cap, N)
        plan.ws_obj, plan.outs, plan.keepalive, plan.static_in = Workspace(), None, keepalive, static_in
        plan.pack = plan.pack_spec = plan.pack_views = plan.host_pack = None
        o = plan.out
        o["seq"] = torch.empty(rows_cap, T, dtype=torch.int64, device=dev)
        o["lps"] = torch.empty(rows_cap, T, device=dev)
        o["steps"] = torch.empty(1, dtype=torch.int32, device=dev)
        o["uniforms"] = torch.empty(T, rows_cap, device=dev) if self.topk_sampling else None
        o["ws"] = torch.empty(L.subgc_decode_workspace_bytes(C.byref(cd), rows_cap, N) + 256, dtype=torch.uint8, device=dev)
        return plan

    def _sample_impl(self, front, opt):
        if not self.test_LSTM:
            raise _lib.SubgcError("mode='sample' needs a model built with opt.test_LSTM=1 (as test.py does)")
        self.last_host_results = None
        dyn = self._dyn_eligible(front, opt)
        if dyn is not None:
            return self._sample_dyn(front, opt, *dyn)
        if "compact" in front:
            front = self._expand_compact(front["compact"])
        self.check_numerics(block=False)
        beam_size = opt.get("beam_size", 1)
        return_att = opt.get("return_att", 0) == 1
        if not opt.get("sample_max", 1) and not self.topk_sampling and beam_size == 1:
            raise NotImplementedError("sample_max=0 leaves `it` undefined in the reference (AttModel.py:304-307)")
        uniforms = opt.get("topk_uniforms", None)
        kind = ("beam", beam_size, opt.get("length_penalty", ""), int(opt.get("decoding_constraint", 0))) if beam_size > 1 else \
            ("sample", bool(self.topk_sampling), float(self.topk_temp), int(self.the_k), return_att)
        (g_fc, fc, att, p_att, masks), sub_score, keep_ind, n_rows, len_max = self._front(plan_kind=kind, **front)
        plan = self._cur_plan
        if beam_size > 1:
            res = self._beam(fc, att, p_att, masks, n_rows, len_max, opt, plan)
            if res is None:   # fp16-range overflow: repeated on the fp32 path (see _check_call)
                return self._sample_impl(front, opt)
            return res[0], res[1], sub_score, keep_ind
        dev = fc.device
        L, w, cd, T = lib(), self._weights(), self._cdims, self.seq_length
        o = plan.out
        if not o:
            o["seq"] = torch.empty(n_rows, T, dtype=torch.int64, device=dev)
            o["lps"] = torch.empty(n_rows, T, device=dev)
            o["attw"] = torch.empty(n_rows, T + 1, len_max, device=dev) if return_att else None
            o["steps"] = torch.empty(1, dtype=torch.int32, device=dev)
            o["uniforms"] = torch.empty(T, n_rows, device=dev) if self.topk_sampling else None
            o["ws"] = torch.empty(L.subgc_decode_workspace_bytes(C.byref(cd), n_rows, len_max) + 256, dtype=torch.uint8, device=dev)
        if self.topk_sampling:
            # uniforms come from torch's generator (torch.manual_seed / opt['seed'] control them); the kernel maps them to
            # tokens by inverse CDF over the k kept candidates.  (The C ABI also has its own Philox stream: uniforms = NULL.)
            if uniforms is not None:
                assert tuple(uniforms.shape) == (T, n_rows), "topk_uniforms must be [seq_length, rows]"
                o["uniforms"].copy_(uniforms.to(dev, dtype=torch.float32))
            elif "seed" in opt:
                gen = torch.Generator(device=dev)
                gen.manual_seed(int(opt["seed"]))
                o["uniforms"].copy_(torch.rand(T, n_rows, device=dev, generator=gen))
            else:
                o["uniforms"].uniform_()
        st_args = (C.byref(cd), C.byref(w), n_rows, len_max, 1 if self.topk_sampling else 0, float(self.topk_temp), int(self.the_k), 0, 0,
                   ptr(o["uniforms"]), ptr(fc), ptr(att), ptr(p_att), ptr(masks), ptr(o["seq"]), ptr(o["lps"]), ptr(o["attw"]),
                   ptr(o["steps"]), ptr(o["ws"]), o["ws"].numel())

        def launch():
            check(L.subgc_decode_sample(*st_args, self._stream()), "subgc_decode_sample")

        plan.run(launch, self.use_graphs)
        self._mark("decode")
        self.last_steps = o["steps"]
        seq, lps = o["seq"].clone(), o["lps"].clone()
        steps = self._check_call(o["steps"])
        if steps is None:   # an activation left the fp16 range: this call's results are invalid, run it again on the fp32 path
            return self._sample_impl(front, opt)
        if return_att:
            return seq, lps, sub_score, keep_ind, o["attw"][:, :steps].clone()
        return seq, lps, sub_score, keep_ind

    def _beam(self, fc, att, p_att, masks, n_sub, len_max, opt, plan):
        """Batched replacement of the per-sub-graph beam loop (reference models/AttModel.py:208-234)."""
        if opt.get("group_size", 1) != 1:
            raise NotImplementedError("diverse beam search (group_size > 1) is not part of the Sub-GC configurations")
        dev = fc.device
        L, w, cd, T = lib(), self._weights(), self._cdims, self.seq_length
        b = int(opt.get("beam_size", 10))
        pen = opt.get("length_penalty", "")
        kind, alpha = 0, 0.0
        if pen:
            name, a = pen.split("_")
            kind, alpha = {"wu": 1, "avg": 2}[name], float(a)
        o = plan.out
        if not o:
            # every result of the search in ONE device buffer (8-byte fields first), read back with one copy into pinned memory
            n_seq, n_p, n_lps = n_sub * b * T * 8, n_sub * b * 8, n_sub * b * T * 4
            o["pack"] = torch.empty(n_seq + 2 * n_p + n_lps + n_sub * 4, dtype=torch.uint8, device=dev)
            o["host"] = torch.empty(o["pack"].numel(), dtype=torch.uint8).pin_memory()

            def views(buf):
                at = 0
                out = []
                for nbytes, dt, shape in ((n_seq, torch.int64, (n_sub, b, T)), (n_p, torch.float64, (n_sub, b)), (n_p, torch.float64, (n_sub, b)),
                                          (n_lps, torch.float32, (n_sub, b, T)), (n_sub * 4, torch.int32, (n_sub,))):
                    out.append(buf[at:at + nbytes].view(dt).view(shape))
                    at += nbytes
                return out
            o["views"] = views
            o["seq"], o["p"], o["up"], o["lps"], o["cnt"] = views(o["pack"])
            o["ws"] = torch.empty(L.subgc_beam_workspace_bytes(C.byref(cd), n_sub, b, len_max) + 256, dtype=torch.uint8, device=dev)
        st_args = (C.byref(cd), C.byref(w), n_sub, len_max, b, kind, alpha, int(opt.get("decoding_constraint", 0)), ptr(fc), ptr(att),
                   ptr(p_att), ptr(masks), ptr(o["seq"]), ptr(o["lps"]), ptr(o["p"]), ptr(o["up"]), ptr(o["cnt"]), ptr(o["ws"]),
                   o["ws"].numel())

        def launch():
            check(L.subgc_decode_beam(*st_args, self._stream()), "subgc_decode_beam")

        plan.run(launch, self.use_graphs)
        self._mark("decode")
        self._arm_overflow_check()
        # the reference hands back CPU tensors and python lists here (AttModel.py:212-213,229-231): one copy, one synchronisation;
        # the private copy of the host buffer belongs to this call's results (the pinned buffer is reused by the next call)
        o["host"].copy_(o["pack"], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        seq_h, p_h, up_h, lps_h, cnt_h = o["views"](o["host"].clone())
        if self._ovf_event is not None:   # the synchronisation above covered it
            self._ovf_event.synchronize()
            self._ovf_event = None
            if int(self._ovf_host[0]) != 0:
                import warnings
                self._ovf_dev.zero_()
                self.use_packed = False
                warnings.warn("subgc: an activation exceeded the fp16 range of the split-fp16 tensor-core path (|x| > 65504); "
                              "repeating the beam search on the fp32 (split-TF32) path, which this model keeps using from now on")
                return None
        self.done_beams = _DoneBeams(seq_h, lps_h, up_h, p_h, cnt_h)
        return seq_h[:, 0].contiguous(), lps_h[:, 0].contiguous()

    def get_logprobs_state(self, it, fc_feats, att_feats, p_att_feats, att_masks, state, sg_emb=None, p_sg_emb=None, return_att=False):
        """Reference models/AttModel.py:328-341 (eval mode).  `fc_feats` may have fewer rows than `it` when consecutive
        rows share a context (beam search)."""
        dev = self._check_device(it, fc_feats, att_feats, p_att_feats, att_masks)
        L, w, cd = lib(), self._weights(), self._cdims
        S, n_ctx, len_max = it.shape[0], fc_feats.shape[0], att_feats.shape[1]
        h_in, c_in = self._f32(state[0]), self._f32(state[1])
        h_out, c_out = torch.empty_like(h_in), torch.empty_like(c_in)
        logp = torch.empty(S, self.dims.v1, device=dev)
        attw = torch.empty(S, len_max, device=dev) if return_att else None
        ws = self._ws.get(L.subgc_decode_workspace_bytes(C.byref(cd), S, len_max), dev)
        check(L.subgc_decode_step(C.byref(cd), C.byref(w), S, len_max, S // n_ctx, ptr(self._i64(it)), ptr(self._f32(fc_feats)),
                                  ptr(self._f32(att_feats)), ptr(self._f32(p_att_feats)), ptr(self._f32(att_masks)), ptr(h_in), ptr(c_in),
                                  ptr(h_out), ptr(c_out), ptr(logp), ptr(attw), ptr(ws), ws.numel(), self._stream()), "subgc_decode_step")
        if return_att:
            return logp, (h_out, c_out), attw
        return logp, (h_out, c_out)

    def _forward(self, fc_feats, att_feats, seq, att_masks=None, trip_pred=None, obj_dist=None, obj_box=None, rel_ind=None,
                 pred_fmap=None, pred_dist=None, gpn_obj_ind=None, gpn_pred_ind=None, gpn_nrel_ind=None, gpn_pool_mtx=None):
        """Reference models/AttModel.py:122-177, evaluation semantics (no dropout, no scheduled sampling): returns
        (outputs [5B, T', V+1] log-probs, gpn_loss, subgraph_score [2*5B*G, 1])."""
        dev = self._check_device(att_feats, seq, att_masks, obj_dist, rel_ind, gpn_obj_ind)
        if self.training or (torch.is_grad_enabled() and self.force_train_path):
            return self._forward_train(att_feats, seq, att_masks, obj_dist, rel_ind, gpn_obj_ind)
        L, w, cd = lib(), self._weights(), self._cdims
        gpn_obj_ind, att_masks, seq = self._i64(gpn_obj_ind), self._f32(att_masks), self._i64(seq)
        x_obj = self.encode(att_feats, obj_dist, pred_dist, rel_ind)
        lay, n_sub, read_out, score, sub_len, loss = self._sgpn(x_obj, gpn_obj_ind, att_masks, order=0)
        rows = lay.rows
        sel = torch.empty(rows, dtype=torch.int32, device=dev)
        stats = torch.empty(2, dtype=torch.int32, device=dev)
        check(L.subgc_sgpn_select_train(C.byref(lay), ptr(score), ptr(sub_len), ptr(sel), ptr(stats), self._stream()),
              "subgc_sgpn_select_train")
        len_max = int(stats.cpu()[1])
        g_fc, fc, att, p_att, masks = self._prepare(lay, rows, len_max, sel, x_obj, gpn_obj_ind, att_masks, read_out)
        n_steps = seq.shape[1] - 1
        outputs = torch.empty(rows, n_steps, self.dims.v1, device=dev)
        ws = self._ws.get(L.subgc_teacher_workspace_bytes(C.byref(cd), rows, n_steps), dev)
        check(L.subgc_decode_teacher(C.byref(cd), C.byref(w), rows, len_max, n_steps, ptr(seq), seq.shape[1], ptr(fc), ptr(att), ptr(p_att),
                                     ptr(masks), ptr(outputs), ptr(ws), ws.numel(), self._stream()), "subgc_decode_teacher")
        self.last_sel = sel
        return outputs, loss[0], score.view(-1, 1)


def _forward_train(self, att_feats, seq, att_masks, obj_dist, rel_ind, gpn_obj_ind, loss=None):
    """Training-mode AttModel._forward: CUDA forward with dropout + saved activations, gradients through _TrainStep.
    loss = (labels[:, 1:], masks[:, 1:]) (LossWrapper): returns (lang_loss, gpn_loss, score) with log-softmax + criterion fused."""
    data = dict(att_feats=self._f32(att_feats), obj_dist=self._f32(obj_dist), rel_ind=self._i64(rel_ind), labels=self._i64(seq),
                att_masks=self._f32(att_masks), gpn_obj_ind=self._i64(gpn_obj_ind))
    drop = None
    if self.training and self.dropout_enabled:
        # gpn_fc's Dropout(0.5) is active in train mode whatever drop_prob_lm is (reference models/lib/gpn.py:24-28); p = 0 only switches
        # the four language-model sites off.  The seed comes from torch's CPU generator; rank and device are mixed in so that data-parallel
        # workers seeded identically still draw different masks (as their independent CUDA generators would)
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        seed = (seed * 1000003 + rank * 8191 + (att_feats.device.index or 0) * 131) % (2 ** 31 - 1)
        drop = dict(p=float(self.drop_prob_lm), seed=seed)
    if self.training and self.ss_prob > 0:   # scheduled sampling (train.py:131: model.ss_prob is raised during training)
        data["ss"] = dict(prob=float(self.ss_prob), seed=int(torch.randint(0, 2 ** 31 - 1, (1,)).item()))
    if loss is not None and "ss" not in data:
        data["loss"] = (self._i64(loss[0]).contiguous(), self._f32(loss[1]).contiguous())
    names, params = zip(*self._named_params().items())
    outputs, gpn_loss, score = _TrainStep.apply(self, data, drop, names, *params)
    return outputs, gpn_loss, score


TopDownModel._forward_train = _forward_train


class _DoneBeams:
    """`model.done_beams` after a beam search (reference models/AttModel.py:229-231: per sub-graph the list of finished beams, best
    first, each a dict seq / logps / unaug_p / p).  Same indexing, length and iteration as the reference's list of lists; the 640 dicts
    of a 128-image batch are built when somebody looks at them (eval_utils does for `verbose_beam` only), not on every call."""

    def __init__(self, seq, logps, unaug_p, p, count):
        self._seq, self._logps = seq, logps
        self._unaug_p, self._p, self._count = unaug_p.tolist(), p.tolist(), count.tolist()
        self._cache = {}

    def __len__(self):
        return len(self._count)

    def _image(self, k):
        got = self._cache.get(k)
        if got is None:
            seqs, lps = self._seq[k].unbind(0), self._logps[k].unbind(0)
            got = [dict(seq=seqs[j], logps=lps[j], unaug_p=self._unaug_p[k][j], p=self._p[k][j]) for j in range(self._count[k])]
            self._cache[k] = got
        return got

    def __getitem__(self, k):
        if isinstance(k, slice):
            return [self._image(i) for i in range(*k.indices(len(self)))]
        n = len(self)
        if k < -n or k >= n:
            raise IndexError("done_beams index out of range")
        return self._image(k % n)

    def __iter__(self):
        return (self._image(k) for k in range(len(self)))

    def __eq__(self, other):
        return list(self) == other

    def __repr__(self):
        return repr(list(self))


class LanguageModelCriterion(nn.Module):
    """Reference misc/utils.py:111-124."""

    def forward(self, input, target, mask):
        target = target[:, :input.size(1)]
        mask = mask[:, :input.size(1)]
        output = -input.gather(2, target.unsqueeze(2)).squeeze(2) * mask
        return torch.sum(output) / torch.sum(mask)


class LossWrapper(nn.Module):
    """Reference models/loss_wrapper.py:7-27."""

    def __init__(self, model, opt):
        super().__init__()
        self.opt = opt
        self.model = model
        self.crit = LanguageModelCriterion()

    def forward(self, fc_feats, att_feats, labels, masks, att_masks, gts, gt_indices, trip_pred, obj_dist, obj_box, rel_ind, pred_fmap,
                pred_dist, gpn_obj_ind, gpn_pred_ind, gpn_nrel_ind, gpn_pool_mtx):
        m = self.model
        if (isinstance(m, TopDownModel) and m.training and m.fused_loss and att_feats.is_cuda and not (m.ss_prob > 0)
                and labels.shape[1] == masks.shape[1]):
            # same result as the two lines below, with the log-softmax and the criterion inside the decoder stage (no [rows, T, V+1] tensor)
            m._check_device(att_feats, labels, att_masks, obj_dist, rel_ind, gpn_obj_ind)
            lang_loss, gpn_loss, _ = m._forward_train(att_feats, labels, att_masks, obj_dist, rel_ind, gpn_obj_ind, loss=(labels[:, 1:], masks[:, 1:]))
            return {"gpn_loss": gpn_loss, "lang_loss": lang_loss}
        lang_output, gpn_loss, _ = self.model(fc_feats, att_feats, labels, att_masks, trip_pred, obj_dist, obj_box, rel_ind, pred_fmap,
                                              pred_dist, gpn_obj_ind, gpn_pred_ind, gpn_nrel_ind, gpn_pool_mtx)
        lang_loss = self.crit(lang_output, labels[:, 1:], masks[:, 1:]) if lang_output is not None else None
        return {"gpn_loss": gpn_loss, "lang_loss": lang_loss}


def setup(opt):
    """Reference models/__init__.py:43-59."""
    import os
    if opt.caption_model != "topdown":
        raise Exception("Caption model not supported: {}".format(opt.caption_model))
    if getattr(opt, "use_gpn", 1) == 0:   # Full-GC (train.sh:27-36): no sGPN, full scene graph, BatchNorm in the GCN units
        from .fullgc import FullGCModel
        model = FullGCModel(opt)
    else:
        model = TopDownModel(opt)
    if vars(opt).get("start_from", None) is not None:
        assert os.path.isdir(opt.start_from), " %s must be a a path" % opt.start_from
        assert os.path.isfile(os.path.join(opt.start_from, "infos_" + opt.id + ".pkl")), \
            "infos.pkl file does not exist in path %s" % opt.start_from
        model.load_state_dict(torch.load(os.path.join(opt.start_from, "model.pth")))
    return model
