"""Training step of the Sub-GC path: train-mode forward (dropout, saved activations) and the hand-written backward.

Reference: AttModel._forward (models/AttModel.py:122-177), gpn_layer.forward training branch (models/lib/gpn.py:41-81),
LossWrapper / LanguageModelCriterion (models/loss_wrapper.py:14-27, misc/utils.py:111-124).

`forward()` / `backward()` below are written against a small set of building-block operations (`CudaOps`): every
arithmetic operation is one C-ABI call into libsubgc_b200.so (contractions: subgc_gemm_nt, the same GEMM block as
inference; pointwise / gather / scatter / LSTM-cell / attention / pooling / GCN halves: train.cu).  torch is used for
storage, views, concatenation and the autograd hook-up only.  tests/emu_ops.py mirrors the same interface with torch
CPU ops so that the orchestration and the backward algebra can be checked against autograd without a GPU.

Scope: sampling_prob == 0 (no scheduled sampling), Sub-GC configuration.  Dropout uses Philox masks drawn per call
(`seed` from torch's generator); RNG streams differ from torch's, so parity tests run with dropout disabled and the
masks are tested statistically.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib, ptr

EW_MUL, EW_ADD, EW_RELU_BWD, EW_SCALE, EW_COPY = 0, 1, 2, 3, 4


DECODER_GRAD_FIELDS = {   # subgc_decoder_grads field -> state_dict name
    "logit_w": "logit.weight", "logit_b": "logit.bias", "embed": "embed.0.weight",
    "att_w_ih": "core.att_lstm.weight_ih", "att_w_hh": "core.att_lstm.weight_hh", "att_b_ih": "core.att_lstm.bias_ih", "att_b_hh": "core.att_lstm.bias_hh",
    "lang_w_ih": "core.lang_lstm.weight_ih", "lang_w_hh": "core.lang_lstm.weight_hh", "lang_b_ih": "core.lang_lstm.bias_ih",
    "lang_b_hh": "core.lang_lstm.bias_hh", "h2att_w": "core.attention.h2att.weight", "h2att_b": "core.attention.h2att.bias",
    "alpha_w": "core.attention.alpha_net.weight",
}


PREPARE_GRAD_FIELDS = {   # subgc_prepare_grads field -> state_dict name
    "fc2_w": "fc_embed.2.weight", "fc2_b": "fc_embed.2.bias", "fc0_w": "fc_embed.0.weight", "fc0_b": "fc_embed.0.bias",
    "ro1_w": "gpn_layer.read_out_proj.1.weight", "ro1_b": "gpn_layer.read_out_proj.1.bias",
    "ro0_w": "gpn_layer.read_out_proj.0.weight", "ro0_b": "gpn_layer.read_out_proj.0.bias",
    "ctx2att_w": "ctx2att.weight", "ctx2att_b": "ctx2att.bias", "att_embed_w": "att_embed.0.weight", "att_embed_b": "att_embed.0.bias",
}


class CudaOps:
    """Building blocks on CUDA tensors (fp32, contiguous).  One method == one C-ABI call."""

    def __init__(self, cdims, seq_per_img=5):
        self.cd = cdims
        self.seq_per_img = seq_per_img
        self._ws = None

    # -- plumbing ------------------------------------------------------------------------------------------------
    @staticmethod
    def _st():
        return torch.cuda.current_stream().cuda_stream

    def _wsbuf(self, nbytes, dev):
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(max(int(nbytes), 1 << 22), dtype=torch.uint8, device=dev)
        return self._ws

    def layout(self, rows, per_half, order):
        return _lib.Layout(rows, per_half, self.seq_per_img, order)

    # -- contractions ----------------------------------------------------------------------------------------------
    def linear(self, x, w, b=None, relu=False, gather=None, out=None, accumulate=False):
        """out (+)= act(x[gather] @ w.T + b).  x [*,K] (row stride = x.stride(0)), w [N,K]."""
        L = lib()
        M = x.shape[0] if gather is None else gather.shape[0]
        N, K = w.shape
        assert x.stride(1) == 1 and w.stride(1) == 1 and x.shape[1] == K
        if out is None:
            out = torch.empty(M, N, device=x.device)
        ws = self._wsbuf(L.subgc_gemm_nt_workspace_bytes(M, N, K), x.device)
        check(L.subgc_gemm_nt(M, N, K, ptr(x), x.stride(0), ptr(gather), ptr(w), w.stride(0), ptr(b), int(relu), int(accumulate), ptr(out),
                              out.stride(0), ptr(ws), ws.numel(), self._st()), "subgc_gemm_nt")
        return out

    def transpose(self, x):
        rows, cols = x.shape
        out = torch.empty(cols, rows, device=x.device)
        check(lib().subgc_transpose(rows, cols, ptr(x), x.stride(0), ptr(out), rows, self._st()), "subgc_transpose")
        return out

    def colsum(self, x, out, accumulate=True):
        check(lib().subgc_colsum(x.shape[0], x.shape[1], ptr(x), x.stride(0), ptr(out), int(accumulate), self._st()), "subgc_colsum")
        return out

    # -- pointwise -------------------------------------------------------------------------------------------------
    def _ew(self, op, a, b, scalar=0.0):
        out = torch.empty_like(a)
        check(lib().subgc_ew(op, a.numel(), ptr(a), ptr(b), ptr(out), float(scalar), self._st()), "subgc_ew")
        return out

    def mul(self, a, b):
        return self._ew(EW_MUL, a.contiguous(), b.contiguous())

    def add(self, a, b):
        return self._ew(EW_ADD, a.contiguous(), b.contiguous())

    def relu_bwd(self, y, dy):
        return self._ew(EW_RELU_BWD, y.contiguous(), dy.contiguous())

    def scale(self, a, s):
        return self._ew(EW_SCALE, a.contiguous(), None, s)

    def relu(self, a):
        out = torch.empty_like(a)
        check(lib().subgc_unary(0, a.numel(), ptr(a), ptr(out), self._st()), "subgc_unary")
        return out

    def sigmoid(self, a):
        out = torch.empty_like(a)
        check(lib().subgc_unary(1, a.numel(), ptr(a), ptr(out), self._st()), "subgc_unary")
        return out

    def dropout_mask(self, shape, p, seed, offset, device):
        m = torch.empty(shape, device=device)
        check(lib().subgc_dropout_mask(m.numel(), float(p), int(seed), int(offset), ptr(m), self._st()), "subgc_dropout_mask")
        return m

    def ss_sample(self, prev_logp, labels_col, ss_prob, seed, offset):
        """models/AttModel.py:158-167: ground-truth token or, with probability ss_prob per row, a draw from exp(prev_logp)."""
        rows, V1 = prev_logp.shape
        it = torch.empty(rows, dtype=torch.int64, device=prev_logp.device)
        check(lib().subgc_ss_sample(rows, V1, ptr(prev_logp), prev_logp.stride(0), ptr(labels_col), labels_col.stride(0), float(ss_prob), int(seed),
                                    int(offset), ptr(it), self._st()), "subgc_ss_sample")
        return it

    def gather_rows(self, src, idx, relu=False):
        out = torch.empty(idx.shape[0], src.shape[1], device=src.device)
        check(lib().subgc_gather_rows(idx.shape[0], src.shape[1], ptr(src), src.stride(0), ptr(idx), ptr(out), int(relu), self._st()),
              "subgc_gather_rows")
        return out

    def scatter_add_rows(self, src, idx, dst):
        check(lib().subgc_scatter_add_rows(src.shape[0], src.shape[1], ptr(src), src.stride(0), ptr(idx), ptr(dst), dst.stride(0), self._st()),
              "subgc_scatter_add_rows")

    # -- fused stages ----------------------------------------------------------------------------------------------
    def lstm_fwd(self, gates, c_prev):
        S, H = c_prev.shape
        h, c = torch.empty_like(c_prev), torch.empty_like(c_prev)
        check(lib().subgc_lstm_cell_train_fwd(S, H, ptr(gates), ptr(c_prev), ptr(h), ptr(c), self._st()), "subgc_lstm_cell_train_fwd")
        return h, c  # gates now holds the activations

    def lstm_bwd(self, act, c_prev, c_new, dh, dc):
        S, H = c_prev.shape
        dg, dcp = torch.empty_like(act), torch.empty_like(c_prev)
        check(lib().subgc_lstm_cell_bwd(S, H, ptr(act), ptr(c_prev), ptr(c_new), ptr(dh.contiguous()), ptr(dc), ptr(dg), ptr(dcp), self._st()),
              "subgc_lstm_cell_bwd")
        return dg, dcp

    def att_fwd(self, atth, p_att, att, masks, aw, ab):
        S, ln, H = att.shape
        AH = p_att.shape[2]
        ctx, alpha, sm = torch.empty(S, H, device=att.device), torch.empty(S, ln, device=att.device), torch.empty(S, ln, device=att.device)
        check(lib().subgc_attention_train_fwd(S, ln, H, AH, ptr(atth), ptr(p_att), ptr(att), ptr(masks), ptr(aw), ptr(ab), ptr(ctx), ptr(alpha),
                                              ptr(sm), self._st()), "subgc_attention_train_fwd")
        return ctx, alpha, sm

    def att_bwd(self, atth, p_att, att, masks, aw, alpha, sm, dctx, d_att, d_p_att):
        S, ln, H = att.shape
        AH = p_att.shape[2]
        d_atth, d_w = torch.empty(S, AH, device=att.device), torch.empty(S, AH, device=att.device)
        check(lib().subgc_attention_bwd(S, ln, H, AH, ptr(atth), ptr(p_att), ptr(att), ptr(masks), ptr(aw), ptr(alpha), ptr(sm),
                                        ptr(dctx.contiguous()), ptr(d_att), ptr(d_p_att), ptr(d_atth), ptr(d_w), self._st()),
              "subgc_attention_bwd")
        return d_atth, d_w

    # -- stage-level decoder (one C call per direction, include/subgc_b200.h: subgc_decoder_train_forward / _backward) ------------
    def decoder_forward(self, weights, tokens, fc, att3, p_att3, masks_c, m_x, m_h, outputs, loss=None):
        """tokens [T, R] int64.  Returns the saved-activation record the backward needs.  loss = (targets [T, R] int64, mask [T, R]):
        fused log-softmax + criterion, `outputs` is then None and rec["nll"] holds -logp[target] * mask per row."""
        L = lib()
        T, R = tokens.shape
        _, ln, H = att3.shape
        AH, E = p_att3.shape[2], self.cd.enc
        dev = att3.device
        f = lambda *shape: torch.empty(*shape, device=dev)
        rec = dict(tokens=tokens, fc=fc, att=att3, p_att=p_att3, masks=masks_c, m_x=m_x, m_h=m_h, xt=f(T, R, E), act1=f(T, R, 4 * H),
                   c_att=f(T + 1, R, H), h_att=f(T + 1, R, H), atth=f(T, R, AH), ctx=f(T, R, H), alpha=f(T, R, ln), sm=f(T, R, ln),
                   act2=f(T, R, 4 * H), c_lang=f(T + 1, R, H), h_lang=f(T + 1, R, H), hd=f(T, R, H) if m_h is not None else None, outputs=outputs)
        if loss is not None:
            rec.update(logits=f(T * R, self.cd.vocab1), targets=loss[0], tmask=loss[1], lse=f(T, R), nll=f(T, R))
        bufs = _lib.DecoderTrainBufs(**{k: ptr(v) for k, v in rec.items()})
        ws = self._wsbuf(L.subgc_decoder_train_workspace_bytes(C.byref(self.cd), R, ln, T), dev)
        T_total = outputs.shape[1] if outputs is not None else T
        check(L.subgc_decoder_train_forward(C.byref(self.cd), C.byref(weights), R, ln, T, T_total, C.byref(bufs), ptr(ws), ws.numel(),
                                            self._st()), "subgc_decoder_train_forward")
        rec["T_total"] = T_total
        rec.update(bufs=bufs, T=T, R=R, len=ln, weights=weights, outputs=None)   # the caller keeps a detached alias of outputs alive
        return rec

    def decoder_backward(self, rec, d_outputs, grads, coef=None):
        """grads: name -> zero-filled gradient tensor of the 14 decoder parameters.  Returns (d_fc, d_att, d_p_att).
        coef [T, R] (fused loss: mask * d(lang_loss) / sum(mask)) replaces d_outputs."""
        L = lib()
        T, R, ln = rec["T"], rec["R"], rec["len"]
        dev = rec["att"].device
        H, AH = rec["att"].shape[2], rec["p_att"].shape[2]
        if coef is not None:
            rec["coef"] = coef   # keeps the tensor alive until the launches are enqueued
            rec["bufs"].coef = ptr(coef)
        d_fc, d_att, d_patt = torch.empty(R, H, device=dev), torch.empty(R, ln, H, device=dev), torch.empty(R, ln, AH, device=dev)
        g = _lib.DecoderGrads(**{k: ptr(grads[n]) for k, n in DECODER_GRAD_FIELDS.items()})
        ws = self._wsbuf(L.subgc_decoder_train_workspace_bytes(C.byref(self.cd), R, ln, T), dev)
        check(L.subgc_decoder_train_backward(C.byref(self.cd), C.byref(rec["weights"]), R, ln, T, rec["T_total"], C.byref(rec["bufs"]),
                                             ptr(d_outputs), C.byref(g), ptr(d_fc), ptr(d_att), ptr(d_patt), ptr(ws), ws.numel(), self._st()),
              "subgc_decoder_train_backward")
        return d_fc, d_att, d_patt

    # -- stage-level front-end backward (subgc_prepare_backward / subgc_sgpn_backward / subgc_gcn_backward) -------------------
    def _front_ws(self, S):
        L = lib()
        n_sub = S["n_sub"]
        nb = L.subgc_frontend_backward_workspace_bytes(C.byref(self.cd), S["B"], S["rows"], S["len_max"], n_sub)
        return self._wsbuf(nb, S["x0"].device)

    def prepare_backward(self, S, weights, d_fc, d_att, d_patt, G):
        dev = S["x0"].device
        R, ln = S["rows"], S["len_max"]
        H = d_fc.shape[1]
        saved = _lib.PrepareSaved(m_fc=ptr(S["m_fc"]), fc_pre=ptr(S["fc_pre"]), f1=ptr(S["f1"]), g_fc=ptr(S["g_fc"]), hr=ptr(S["hr"]),
                                  read_sel=ptr(S["read_sel"]), att=ptr(S["att"]), att_pre=ptr(S["att_pre"]), m_att=ptr(S["m_att"]),
                                  x_rows=ptr(S["x_rows"]), node_row=ptr(S["node_row"]))
        g = _lib.PrepareGrads(**{k: ptr(G[n]) for k, n in PREPARE_GRAD_FIELDS.items()})
        n_nodes = S["B"] * S["N"]
        d_xobj = torch.empty(n_nodes, self.cd.gcn, device=dev)
        ws = self._front_ws(S)
        check(lib().subgc_prepare_backward(C.byref(self.cd), C.byref(weights), R, ln, n_nodes, C.byref(saved), ptr(d_fc), ptr(d_att.contiguous()),
                                           ptr(d_patt.contiguous()), C.byref(g), ptr(d_xobj), ptr(ws), ws.numel(), self._st()), "subgc_prepare_backward")
        return d_xobj

    def sgpn_backward(self, S, weights, scale, G, d_xobj):
        saved = _lib.SgpnSaved(score=ptr(S["score"]), hid=ptr(S["hid"]), hid_d=ptr(S["hid_d"]), m_gpn=ptr(S["m_gpn"]), read_out=ptr(S["read_out"]),
                               sub_len=ptr(S["sub_len"]))
        g = _lib.SgpnGrads(fc3_w=ptr(G["gpn_layer.gpn_fc.3.weight"]), fc3_b=ptr(G["gpn_layer.gpn_fc.3.bias"]),
                           fc0_w=ptr(G["gpn_layer.gpn_fc.0.weight"]), fc0_b=ptr(G["gpn_layer.gpn_fc.0.bias"]))
        ws = self._front_ws(S)
        check(lib().subgc_sgpn_backward(C.byref(self.cd), C.byref(weights), C.byref(S["lay"]), C.byref(saved), float(scale), ptr(S["x_obj"]),
                                        ptr(S["obj_ind"]), C.byref(g), ptr(d_xobj), ptr(ws), ws.numel(), self._st()), "subgc_sgpn_backward")

    def gcn_backward(self, S, weights, d_xobj, G, live):
        saved = _lib.GcnSaved()
        for l, rec in enumerate(S["layers"]):
            ls = saved.layer[l]
            ls.x_in, ls.p_in = ptr(rec["x_in"]), ptr(rec["p_in"])
            for k in ("t0", "t1", "y0", "y1", "t2", "t3", "m2", "m3"):
                setattr(ls, k, ptr(rec.get(k)))
        saved.x0, saved.att_feats, saved.cls, saved.rel_ind = ptr(S["x0"]), ptr(S["att_feats"]), ptr(S["cls"]), ptr(S["rel_ind"])
        g = _lib.GcnGrads()
        for (l, u) in live:
            pre = _unit(l, u)
            g.lft_w[l][u], g.lft_b[l][u] = ptr(G[pre + "fc_lft.weight"]), ptr(G[pre + "fc_lft.bias"])
            g.rgt_w[l][u], g.rgt_b[l][u] = ptr(G[pre + "fc_rgt.weight"]), ptr(G[pre + "fc_rgt.bias"])
        g.obj_v_w, g.obj_v_b = ptr(G["obj_v_proj.weight"]), ptr(G["obj_v_proj.bias"])
        g.obj_emb_w, g.obj_emb_b, g.sg_obj_embed = ptr(G["obj_emb_proj.weight"]), ptr(G["obj_emb_proj.bias"]), ptr(G["sg_obj_embed.weight"])
        ws = self._front_ws(S)
        check(lib().subgc_gcn_backward(C.byref(self.cd), C.byref(weights), S["B"], C.byref(saved), ptr(d_xobj), C.byref(g), ptr(ws), ws.numel(),
                                       self._st()), "subgc_gcn_backward")

    def log_softmax_fwd(self, logits, out_view):
        """out_view: [rows, V1] view with unit inner stride (e.g. outputs[:, t])."""
        check(lib().subgc_log_softmax_fwd(logits.shape[0], logits.shape[1], ptr(logits), ptr(out_view), out_view.stride(0), self._st()),
              "subgc_log_softmax_fwd")

    def log_softmax_bwd(self, logp_view, dlogp_view):
        rows, V1 = logp_view.shape
        assert logp_view.stride(0) == dlogp_view.stride(0) and dlogp_view.stride(1) == 1
        out = torch.empty(rows, V1, device=logp_view.device)
        check(lib().subgc_log_softmax_bwd(rows, V1, ptr(logp_view), ptr(dlogp_view), logp_view.stride(0), ptr(out), self._st()),
              "subgc_log_softmax_bwd")
        return out

    def class_argmax(self, dist2d, skip_first):
        cls = torch.empty(dist2d.shape[0], dtype=torch.int64, device=dist2d.device)
        check(lib().subgc_class_argmax(dist2d.shape[0], dist2d.shape[1], int(skip_first), ptr(dist2d), ptr(cls), self._st()), "subgc_class_argmax")
        return cls

    def pool(self, lay, n_sub, x_obj, obj_ind, att_masks):
        read = torch.empty(n_sub, 2 * x_obj.shape[2], device=x_obj.device)
        sub_len = torch.empty(n_sub, dtype=torch.int32, device=x_obj.device)
        check(lib().subgc_sgpn_pool(C.byref(self.cd), C.byref(lay), ptr(x_obj), ptr(obj_ind), ptr(att_masks), ptr(read), ptr(sub_len), self._st()),
              "subgc_sgpn_pool")
        return read, sub_len

    def pool_bwd(self, lay, x_obj, obj_ind, sub_len, d_read, d_x_obj):
        check(lib().subgc_sgpn_pool_bwd(C.byref(self.cd), C.byref(lay), ptr(x_obj), ptr(obj_ind), ptr(sub_len), ptr(d_read), ptr(d_x_obj),
                                        self._st()), "subgc_sgpn_pool_bwd")

    def bce(self, lay, score):
        loss = torch.empty(1, device=score.device)
        check(lib().subgc_sgpn_bce(C.byref(lay), ptr(score), ptr(loss), self._st()), "subgc_sgpn_bce")
        return loss

    def bce_bwd(self, lay, score, scale):
        dz = torch.empty_like(score)
        check(lib().subgc_bce_sigmoid_bwd(C.byref(lay), ptr(score), float(scale), ptr(dz), self._st()), "subgc_bce_sigmoid_bwd")
        return dz

    def select_train(self, lay, score, sub_len):
        sel = torch.empty(lay.rows, dtype=torch.int32, device=score.device)
        stats = torch.empty(2, dtype=torch.int32, device=score.device)
        check(lib().subgc_sgpn_select_train(C.byref(lay), ptr(score), ptr(sub_len), ptr(sel), ptr(stats), self._st()), "subgc_sgpn_select_train")
        return sel, int(stats.cpu()[1])

    def prepare_index(self, lay, sel, len_max, obj_ind, att_masks):
        n_rows = sel.shape[0]
        dev = sel.device
        node_row = torch.empty(n_rows * len_max, dtype=torch.int64, device=dev)
        masks = torch.empty(n_rows, len_max, device=dev)
        row_len = torch.empty(n_rows, dtype=torch.int32, device=dev)
        check(lib().subgc_prepare_index(C.byref(self.cd), C.byref(lay), n_rows, len_max, ptr(sel), ptr(obj_ind), ptr(att_masks), ptr(node_row),
                                        ptr(masks), ptr(row_len), self._st()), "subgc_prepare_index")
        return node_row, masks, row_len

    def fuse_nodes(self, weights, att_feats, obj_dist):
        """x0 = relu(W_v att + b_v + W_e E[cls] + b_e) through the inference kernel (subgc_fuse_nodes)."""
        L = lib()
        B, N, _ = att_feats.shape
        x0 = torch.empty(B, N, self.cd.gcn, device=att_feats.device)
        ws = self._wsbuf(L.subgc_encoder_workspace_bytes(C.byref(self.cd), B), att_feats.device)
        check(L.subgc_fuse_nodes(C.byref(self.cd), C.byref(weights), B, ptr(att_feats), ptr(obj_dist), None, ptr(x0), None, ptr(ws), ws.numel(),
                                 self._st()), "subgc_fuse_nodes")
        return x0

    def gcn_edge_fwd(self, m2, m3, rel, res):
        B, N, L_ = m2.shape
        K = rel.shape[1]
        out = torch.empty(B, K, L_, device=m2.device)
        check(lib().subgc_gcn_edge_fwd(B, N, K, L_, ptr(m2), ptr(m3), ptr(rel), ptr(res), ptr(out), self._st()), "subgc_gcn_edge_fwd")
        return out

    def gcn_node_fwd(self, m0, m1, rel, res, N):
        B, K, L_ = m0.shape
        out, y0, y1 = (torch.empty(B, N, L_, device=m0.device) for _ in range(3))
        check(lib().subgc_gcn_node_train_fwd(B, N, K, L_, ptr(m0), ptr(m1), ptr(rel), ptr(res), ptr(out), ptr(y0), ptr(y1), self._st()),
              "subgc_gcn_node_train_fwd")
        return out, y0, y1

    def gcn_node_bwd(self, dx, y0, y1, rel):
        B, N, L_ = dx.shape
        K = rel.shape[1]
        d0, d1 = torch.empty(B, K, L_, device=dx.device), torch.empty(B, K, L_, device=dx.device)
        check(lib().subgc_gcn_node_bwd(B, N, K, L_, ptr(dx.contiguous()), ptr(y0), ptr(y1), ptr(rel), ptr(d0), ptr(d1), self._st()),
              "subgc_gcn_node_bwd")
        return d0, d1

    def gcn_edge_bwd(self, dp, m2, m3, rel):
        B, N, L_ = m2.shape
        K = rel.shape[1]
        d2, d3 = torch.empty_like(m2), torch.empty_like(m3)
        check(lib().subgc_gcn_edge_bwd(B, N, K, L_, ptr(dp.contiguous()), ptr(m2), ptr(m3), ptr(rel), ptr(d2), ptr(d3), self._st()),
              "subgc_gcn_edge_bwd")
        return d2, d3


# ------------------------------------------------------------------------------------------------------------------
# liveness of GCN units (mirror of gcn_liveness in csrc/encoder.cu)
# ------------------------------------------------------------------------------------------------------------------
def gcn_liveness(layers, residual, want_x_pred=False):
    need_x, need_p = [False] * (layers + 1), [False] * (layers + 1)
    need_x[layers] = True
    need_p[layers] = bool(want_x_pred)
    for l in range(layers, 0, -1):
        if need_x[l]:
            need_p[l - 1] = True
            if l % residual == 0:
                need_x[l - residual] = True
        if need_p[l]:
            need_x[l - 1] = True
            if l % residual == 0:
                need_p[l - residual] = True
    return need_x, need_p


def _unit(l, u):
    return f"gcn_backbone.gcn.{l}.gcn_collect.collect_units.{u}."


# ------------------------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------------------------
def forward(ops, P, weights, d, data, drop=None, seq_per_img=5, ss=None, loss=None):
    """Train-mode AttModel._forward.  P: name -> parameter tensor; weights: subgc_weights struct (for subgc_fuse_nodes);
    drop: None (dropout off) or dict(p=drop_prob_lm, seed=int).  Returns (outputs, gpn_loss, score, saved)."""
    S = {"weights": weights}
    att_feats, obj_dist, rel_ind = data["att_feats"], data["obj_dist"], data["rel_ind"]
    labels, att_masks, obj_ind = data["labels"], data["att_masks"], data["gpn_obj_ind"]
    dev = att_feats.device
    B, N, A = att_feats.shape
    K = rel_ind.shape[1]
    Lg, H, AH, V1 = d.gcn, d.rnn, d.att_hid, d.v1
    rows, _, G, _ = obj_ind.shape
    mask_id = [0]

    def dmask(shape, p):
        if drop is None or p <= 0:
            return None
        mask_id[0] += 1
        return ops.dropout_mask(shape, p, drop["seed"], mask_id[0], dev)

    # ---- fusion (AttModel.py:370-387) ----
    cls = ops.class_argmax(obj_dist.reshape(B * N, -1), 1)
    x0 = ops.fuse_nodes(weights, att_feats, obj_dist)
    S.update(cls=cls, x0=x0)
    if need_pred_path(d):
        raise NotImplementedError("GCN configurations whose outputs depend on the predicate embedding are not trained by this path")

    # ---- GCN (gcn_backbone.py:29-53) ----
    need_x, need_p = gcn_liveness(d.gcn_layers, d.gcn_residual)
    x, p, x_res, p_res = x0, None, x0, None
    layers_saved = []
    for l in range(d.gcn_layers):
        boundary = (l + 1) % d.gcn_residual == 0
        rec = dict(x_in=x, p_in=p, boundary=boundary)
        x_next = p_next = None
        if need_p[l + 1]:
            x2 = x.reshape(B * N, Lg)
            t2 = ops.linear(x2, P[_unit(l, 2) + "fc_lft.weight"], P[_unit(l, 2) + "fc_lft.bias"])
            m2 = ops.linear(t2, P[_unit(l, 2) + "fc_rgt.weight"], P[_unit(l, 2) + "fc_rgt.bias"]).view(B, N, Lg)
            t3 = ops.linear(x2, P[_unit(l, 3) + "fc_lft.weight"], P[_unit(l, 3) + "fc_lft.bias"])
            m3 = ops.linear(t3, P[_unit(l, 3) + "fc_rgt.weight"], P[_unit(l, 3) + "fc_rgt.bias"]).view(B, N, Lg)
            p_next = ops.gcn_edge_fwd(m2, m3, rel_ind, p_res if boundary else None)
            rec.update(t2=t2, m2=m2, t3=t3, m3=m3)
        if need_x[l + 1]:
            p2 = p.reshape(B * K, Lg)
            t0 = ops.linear(p2, P[_unit(l, 0) + "fc_lft.weight"], P[_unit(l, 0) + "fc_lft.bias"])
            m0 = ops.linear(t0, P[_unit(l, 0) + "fc_rgt.weight"], P[_unit(l, 0) + "fc_rgt.bias"]).view(B, K, Lg)
            t1 = ops.linear(p2, P[_unit(l, 1) + "fc_lft.weight"], P[_unit(l, 1) + "fc_lft.bias"])
            m1 = ops.linear(t1, P[_unit(l, 1) + "fc_rgt.weight"], P[_unit(l, 1) + "fc_rgt.bias"]).view(B, K, Lg)
            x_next, y0, y1 = ops.gcn_node_fwd(m0, m1, rel_ind, x_res if boundary else None, N)
            rec.update(t0=t0, t1=t1, y0=y0, y1=y1)
        layers_saved.append(rec)
        x, p = x_next, p_next
        if boundary:
            x_res, p_res = x, p
    x_obj = x
    S.update(layers=layers_saved, x_obj=x_obj)

    # ---- sGPN (gpn.py:41-81) ----
    lay = ops.layout(rows, G, 0)
    n_sub = 2 * rows * G
    read_out, sub_len = ops.pool(lay, n_sub, x_obj, obj_ind, att_masks)
    hid = ops.linear(read_out, P["gpn_layer.gpn_fc.0.weight"], P["gpn_layer.gpn_fc.0.bias"], relu=True)
    m_gpn = dmask(hid.shape, 0.5)
    hid_d = hid if m_gpn is None else ops.mul(hid, m_gpn)
    z = ops.linear(hid_d, P["gpn_layer.gpn_fc.3.weight"], P["gpn_layer.gpn_fc.3.bias"])
    score = ops.sigmoid(z).view(-1)
    gpn_loss = ops.bce(lay, score)
    sel, len_max = ops.select_train(lay, score, sub_len)
    S.update(lay=lay, n_sub=n_sub, read_out=read_out, sub_len=sub_len, hid=hid, m_gpn=m_gpn, hid_d=hid_d, score=score, sel=sel)

    # ---- feature preparation (gpn.py:79, AttModel.py:348-368) ----
    node_row, masks_c, row_len = ops.prepare_index(lay, sel, len_max, obj_ind, att_masks)
    sel64 = sel.long()
    read_sel = ops.gather_rows(read_out, sel64)
    hr = ops.linear(read_sel, P["gpn_layer.read_out_proj.0.weight"], P["gpn_layer.read_out_proj.0.bias"])
    g_fc = ops.linear(hr, P["gpn_layer.read_out_proj.1.weight"], P["gpn_layer.read_out_proj.1.bias"])
    f1 = ops.linear(g_fc, P["fc_embed.0.weight"], P["fc_embed.0.bias"], relu=True)
    fc_pre = ops.linear(f1, P["fc_embed.2.weight"], P["fc_embed.2.bias"], relu=True)
    p_lm = 0.0 if drop is None else drop["p"]
    m_fc = dmask(fc_pre.shape, p_lm)
    fc = fc_pre if m_fc is None else ops.mul(fc_pre, m_fc)
    x_rows = ops.gather_rows(x_obj.view(B * N, Lg), node_row)
    att_pre = ops.linear(x_rows, P["att_embed.0.weight"], P["att_embed.0.bias"], relu=True)
    valid = (torch.arange(len_max, device=dev).view(1, -1) < row_len.view(-1, 1)).float().view(rows * len_max, 1).expand(-1, H).contiguous()
    m_att = dmask(att_pre.shape, p_lm)
    m_att = valid if m_att is None else ops.mul(m_att, valid)
    att = ops.mul(att_pre, m_att)                       # padded rows exact zeros (pack_padded_sequence)
    p_att = ops.linear(att, P["ctx2att.weight"], P["ctx2att.bias"])
    att3, p_att3 = att.view(rows, len_max, H), p_att.view(rows, len_max, AH)
    S.update(node_row=node_row, masks_c=masks_c, read_sel=read_sel, hr=hr, g_fc=g_fc, f1=f1, fc_pre=fc_pre, m_fc=m_fc, fc=fc, x_rows=x_rows,
             att_pre=att_pre, m_att=m_att, att=att3, p_att=p_att3, len_max=len_max)

    # ---- teacher-forced decoder (AttModel.py:150-177) ----
    T = labels.shape[1] - 1
    lab_h = labels.cpu()
    n_exec = T
    for i in range(1, T):
        if int(lab_h[:, i].sum()) == 0:
            n_exec = i
            break
    use_ss = ss is not None and ss["prob"] > 0.0
    if loss is not None and hasattr(ops, "decoder_forward") and not use_ss:
        # LossWrapper path (models/loss_wrapper.py:22 + misc/utils.py:115-124): log-softmax and the criterion are fused into the decoder
        # stage, the [rows, T, V1] log-probs are never written.  loss = (labels[:, 1:], masks[:, 1:]); returns the scalar lang_loss
        tgt, msk = loss
        tokens = labels[:, :n_exec].t().contiguous()
        m_x = dmask((n_exec, rows, d.enc), p_lm)
        m_h = dmask((n_exec, rows, H), p_lm)
        tmask = msk[:, :n_exec].t().contiguous().float()
        dec = ops.decoder_forward(weights, tokens, fc, att3, p_att3, masks_c, m_x, m_h, None, loss=(tgt[:, :n_exec].t().contiguous(), tmask))
        mask_total = msk[:, :T].sum()
        lang_loss = dec["nll"].sum() / mask_total
        S.update(dec=dec, outputs=None, n_exec=n_exec, B=B, N=N, K=K, rows=rows, rel_ind=rel_ind, obj_ind=obj_ind, att_feats=att_feats,
                 tmask=tmask, mask_total=mask_total)
        return lang_loss, gpn_loss, score.view(-1, 1), S
    outputs = torch.zeros(rows, T, V1, device=dev)
    if hasattr(ops, "decoder_forward") and not use_ss:
        # stage-level C entry: the whole teacher-forced loop, the batched logit contraction and the log-softmax are one call
        tokens = labels[:, :n_exec].t().contiguous()
        m_x = dmask((n_exec, rows, d.enc), p_lm)
        m_h = dmask((n_exec, rows, H), p_lm)
        dec = ops.decoder_forward(weights, tokens, fc, att3, p_att3, masks_c, m_x, m_h, outputs)
        S.update(dec=dec, outputs=outputs.detach(), n_exec=n_exec, B=B, N=N, K=K, rows=rows, rel_ind=rel_ind, obj_ind=obj_ind, att_feats=att_feats)
        return outputs, gpn_loss, score.view(-1, 1), S
    zeros = torch.zeros(rows, H, device=dev)
    h_att, c_att, h_lang, c_lang = zeros, zeros, zeros, zeros
    E = P["embed.0.weight"]
    steps = []
    for t in range(n_exec):
        if ss is not None and t >= 1 and ss["prob"] > 0.0:   # scheduled sampling (AttModel.py:158-167); the sampled token is a constant
            it = ops.ss_sample(outputs[:, t - 1], labels[:, t], ss["prob"], ss["seed"], t)
        else:
            it = labels[:, t].contiguous()
        x_relu = ops.gather_rows(E, it, relu=True)
        m_x = dmask(x_relu.shape, p_lm)
        xt = x_relu if m_x is None else ops.mul(x_relu, m_x)
        x_att = torch.cat([h_lang, fc, xt], 1)
        g1 = ops.linear(x_att, P["core.att_lstm.weight_ih"], P["core.att_lstm.bias_ih"])
        ops.linear(h_att, P["core.att_lstm.weight_hh"], P["core.att_lstm.bias_hh"], out=g1, accumulate=True)
        h_att_n, c_att_n = ops.lstm_fwd(g1, c_att)
        atth = ops.linear(h_att_n, P["core.attention.h2att.weight"], P["core.attention.h2att.bias"])
        ctx, alpha, sm = ops.att_fwd(atth, p_att3, att3, masks_c, P["core.attention.alpha_net.weight"], P["core.attention.alpha_net.bias"])
        x_lang = torch.cat([ctx, h_att_n], 1)
        g2 = ops.linear(x_lang, P["core.lang_lstm.weight_ih"], P["core.lang_lstm.bias_ih"])
        ops.linear(h_lang, P["core.lang_lstm.weight_hh"], P["core.lang_lstm.bias_hh"], out=g2, accumulate=True)
        h_lang_n, c_lang_n = ops.lstm_fwd(g2, c_lang)
        m_h = dmask(h_lang_n.shape, p_lm)
        hd = h_lang_n if m_h is None else ops.mul(h_lang_n, m_h)
        logits = ops.linear(hd, P["logit.weight"], P["logit.bias"])
        ops.log_softmax_fwd(logits, outputs[:, t])
        steps.append(dict(it=it, x_relu=x_relu, m_x=m_x, x_att=x_att, h_att_prev=h_att, act1=g1, c_att_prev=c_att, c_att=c_att_n, h_att=h_att_n,
                          atth=atth, alpha=alpha, sm=sm, x_lang=x_lang, h_lang_prev=h_lang, act2=g2, c_lang_prev=c_lang, c_lang=c_lang_n,
                          m_h=m_h, hd=hd))
        h_att, c_att, h_lang, c_lang = h_att_n, c_att_n, h_lang_n, c_lang_n
    S.update(steps=steps, outputs=outputs.detach(), n_exec=n_exec, B=B, N=N, K=K, rows=rows, rel_ind=rel_ind, obj_ind=obj_ind, att_feats=att_feats)
    return outputs, gpn_loss, score.view(-1, 1), S


def need_pred_path(d):
    _, need_p = gcn_liveness(d.gcn_layers, d.gcn_residual)
    return need_p[0]


# ------------------------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------------------------
def backward(ops, P, d, S, d_outputs, d_gpn_loss, reducer=None):
    """Gradients of every parameter given d(outputs) [rows, T, V1] and the scalar d(gpn_loss).  Returns name -> tensor
    (parameters that cannot influence the outputs are absent: the reference leaves their .grad at None).
    reducer (subgc.parallel.GradReducer, optional): gradients are allocated inside its flat buckets and each bucket's all-reduce is
    started as soon as its last gradient kernel is enqueued (decoder -> prepare -> sGPN / GCN / fusion)."""
    dev = S["x0"].device
    G = {}
    if reducer is not None:
        reducer.begin()

    def new_grad(name):
        return torch.zeros_like(P[name]) if reducer is None else reducer.alloc(name, P[name])

    def acc_w(name, dy, x):
        """grad[name] += dy^T @ x   (dy [rows, N], x [rows, K])"""
        if name not in G:
            G[name] = new_grad(name)
        ops.linear(ops.transpose(dy), ops.transpose(x), out=G[name], accumulate=True)

    def acc_b(name, dy):
        if name not in G:
            G[name] = new_grad(name)
        ops.colsum(dy, G[name], accumulate=True)

    tcache = {}

    def WT(name, lo=None, hi=None):
        """transposed (column slice of a) weight, cached for the duration of this backward pass"""
        key = (name, lo, hi)
        if key not in tcache:
            w = P[name] if lo is None else P[name][:, lo:hi]
            tcache[key] = ops.transpose(w)
        return tcache[key]

    def dx_of(dy, name, lo=None, hi=None):
        """dy @ W[:, lo:hi]"""
        return ops.linear(dy, WT(name, lo, hi))

    H, X, AH, Lg = d.rnn, d.enc, d.att_hid, d.gcn
    rows, len_max = S["rows"], S["len_max"]
    att3, p_att3, masks_c = S["att"], S["p_att"], S["masks_c"]
    if "dec" not in S:
        d_att = torch.zeros_like(att3)
        d_patt = torch.zeros_like(p_att3)
        d_fc = torch.zeros(rows, H, device=dev)
    dh_att_n = dc_att_n = dh_lang_n = dc_lang_n = None
    aw = P["core.attention.alpha_net.weight"]
    G["core.attention.alpha_net.bias"] = new_grad("core.attention.alpha_net.bias")  # softmax is shift-invariant: exactly zero
    G["embed.0.weight"] = new_grad("embed.0.weight")

    if "dec" in S:
        for n in DECODER_GRAD_FIELDS.values():
            if n not in G:
                G[n] = new_grad(n)
        if S.get("tmask") is not None:   # fused loss: d_outputs is the scalar d(lang_loss)
            coef = (S["tmask"] * (d_outputs / S["mask_total"])).contiguous()
            d_fc, d_att, d_patt = ops.decoder_backward(S["dec"], None, G, coef=coef)
        else:
            d_fc, d_att, d_patt = ops.decoder_backward(S["dec"], d_outputs, G)
    for t in (range(S["n_exec"] - 1, -1, -1) if "dec" not in S else ()):
        st = S["steps"][t]
        dlogits = ops.log_softmax_bwd(S["outputs"][:, t], d_outputs[:, t])
        acc_w("logit.weight", dlogits, st["hd"])
        acc_b("logit.bias", dlogits)
        d_h = dx_of(dlogits, "logit.weight")
        if st["m_h"] is not None:
            d_h = ops.mul(d_h, st["m_h"])
        if dh_lang_n is not None:
            d_h = ops.add(d_h, dh_lang_n)
        dg2, dc_lang_prev = ops.lstm_bwd(st["act2"], st["c_lang_prev"], st["c_lang"], d_h, dc_lang_n)
        acc_w("core.lang_lstm.weight_ih", dg2, st["x_lang"])
        acc_w("core.lang_lstm.weight_hh", dg2, st["h_lang_prev"])
        acc_b("core.lang_lstm.bias_ih", dg2)
        acc_b("core.lang_lstm.bias_hh", dg2)
        d_xl = dx_of(dg2, "core.lang_lstm.weight_ih")
        d_ctx, d_hatt = d_xl[:, :H], d_xl[:, H:]
        dh_lang_prev = dx_of(dg2, "core.lang_lstm.weight_hh")
        d_atth, d_w_rows = ops.att_bwd(st["atth"], p_att3, att3, masks_c, aw, st["alpha"], st["sm"], d_ctx, d_att, d_patt)
        acc_b("core.attention.alpha_net.weight", d_w_rows)
        acc_w("core.attention.h2att.weight", d_atth, st["h_att"])
        acc_b("core.attention.h2att.bias", d_atth)
        d_hatt = ops.add(d_hatt, dx_of(d_atth, "core.attention.h2att.weight"))
        if dh_att_n is not None:
            d_hatt = ops.add(d_hatt, dh_att_n)
        dg1, dc_att_prev = ops.lstm_bwd(st["act1"], st["c_att_prev"], st["c_att"], d_hatt, dc_att_n)
        acc_w("core.att_lstm.weight_ih", dg1, st["x_att"])
        acc_w("core.att_lstm.weight_hh", dg1, st["h_att_prev"])
        acc_b("core.att_lstm.bias_ih", dg1)
        acc_b("core.att_lstm.bias_hh", dg1)
        d_xa = dx_of(dg1, "core.att_lstm.weight_ih")
        d_fc = ops.add(d_fc, d_xa[:, H:2 * H])
        d_xt = d_xa[:, 2 * H:]
        if st["m_x"] is not None:
            d_xt = ops.mul(d_xt, st["m_x"])
        ops.scatter_add_rows(ops.relu_bwd(st["x_relu"], d_xt), st["it"], G["embed.0.weight"])
        dh_lang_n = ops.add(dh_lang_prev, d_xa[:, :H])
        dh_att_n = dx_of(dg1, "core.att_lstm.weight_hh")
        dc_att_n, dc_lang_n = dc_att_prev, dc_lang_prev

    if reducer is not None:
        reducer.bucket_done(0, G)   # logit / embed / LSTMs / attention are final: their all-reduce overlaps everything below
    if hasattr(ops, "prepare_backward") and S.get("weights") is not None:
        # stage-level C entries: one call per front-end stage (same building blocks, sequenced in C++)
        W_ = S["weights"]
        for n in PREPARE_GRAD_FIELDS.values():
            G[n] = new_grad(n)
        d_xobj = ops.prepare_backward(S, W_, d_fc, d_att, d_patt, G)
        if reducer is not None:
            reducer.bucket_done(1, G)
        for n in ("gpn_layer.gpn_fc.3.weight", "gpn_layer.gpn_fc.3.bias", "gpn_layer.gpn_fc.0.weight", "gpn_layer.gpn_fc.0.bias"):
            G[n] = new_grad(n)
        ops.sgpn_backward(S, W_, float(d_gpn_loss) / S["n_sub"], G, d_xobj)
        # which collection units receive a gradient (the reference leaves .grad of the others at None)
        Ln, R = d.gcn_layers, d.gcn_residual
        hx, hp, live = [False] * (Ln + 1), [False] * (Ln + 1), []
        hx[Ln] = True
        for l in range(Ln - 1, -1, -1):
            rec = S["layers"][l]
            if hx[l + 1] and "y0" in rec:
                if rec["boundary"]:
                    hx[l + 1 - R] = True
                hp[l] = True
                live += [(l, 0), (l, 1)]
            if hp[l + 1] and "m2" in rec:
                if rec["boundary"]:
                    hp[l + 1 - R] = True
                hx[l] = True
                live += [(l, 2), (l, 3)]
        for (l, u) in live:
            for k in ("fc_lft.weight", "fc_lft.bias", "fc_rgt.weight", "fc_rgt.bias"):
                G[_unit(l, u) + k] = new_grad(_unit(l, u) + k)
        for n in ("obj_v_proj.weight", "obj_v_proj.bias", "obj_emb_proj.weight", "obj_emb_proj.bias", "sg_obj_embed.weight"):
            G[n] = new_grad(n)
        ops.gcn_backward(S, W_, d_xobj, G, live)
        if reducer is not None:
            reducer.bucket_done(2, G)
            reducer.finish(G)
        return G
    # ---- feature preparation ----
    if S["m_fc"] is not None:
        d_fc = ops.mul(d_fc, S["m_fc"])
    d_fcp = ops.relu_bwd(S["fc_pre"], d_fc)
    acc_w("fc_embed.2.weight", d_fcp, S["f1"]); acc_b("fc_embed.2.bias", d_fcp)
    d_f1 = ops.relu_bwd(S["f1"], dx_of(d_fcp, "fc_embed.2.weight"))
    acc_w("fc_embed.0.weight", d_f1, S["g_fc"]); acc_b("fc_embed.0.bias", d_f1)
    d_gfc = dx_of(d_f1, "fc_embed.0.weight")
    acc_w("gpn_layer.read_out_proj.1.weight", d_gfc, S["hr"]); acc_b("gpn_layer.read_out_proj.1.bias", d_gfc)
    d_hr = dx_of(d_gfc, "gpn_layer.read_out_proj.1.weight")
    acc_w("gpn_layer.read_out_proj.0.weight", d_hr, S["read_sel"]); acc_b("gpn_layer.read_out_proj.0.bias", d_hr)  # read-out is detached (gpn.py:78)
    d_patt2 = d_patt.view(rows * len_max, AH)
    att2 = att3.view(rows * len_max, H)
    acc_w("ctx2att.weight", d_patt2, att2); acc_b("ctx2att.bias", d_patt2)
    d_att2 = ops.add(d_att.view(rows * len_max, H), dx_of(d_patt2, "ctx2att.weight"))
    d_attp = ops.relu_bwd(S["att_pre"], ops.mul(d_att2, S["m_att"]))
    acc_w("att_embed.0.weight", d_attp, S["x_rows"]); acc_b("att_embed.0.bias", d_attp)
    B, N, K = S["B"], S["N"], S["K"]
    d_xobj = torch.zeros(B * N, Lg, device=dev)
    ops.scatter_add_rows(dx_of(d_attp, "att_embed.0.weight"), S["node_row"], d_xobj)

    if reducer is not None:
        reducer.bucket_done(1, G)
    # ---- sGPN ----
    n_sub = S["n_sub"]
    dz = ops.bce_bwd(S["lay"], S["score"], float(d_gpn_loss) / n_sub).view(n_sub, 1)
    acc_w("gpn_layer.gpn_fc.3.weight", dz, S["hid_d"]); acc_b("gpn_layer.gpn_fc.3.bias", dz)
    d_hid = dx_of(dz, "gpn_layer.gpn_fc.3.weight")
    if S["m_gpn"] is not None:
        d_hid = ops.mul(d_hid, S["m_gpn"])
    d_hid = ops.relu_bwd(S["hid"], d_hid)
    acc_w("gpn_layer.gpn_fc.0.weight", d_hid, S["read_out"]); acc_b("gpn_layer.gpn_fc.0.bias", d_hid)
    d_read = dx_of(d_hid, "gpn_layer.gpn_fc.0.weight")
    d_xobj3 = d_xobj.view(B, N, Lg)
    ops.pool_bwd(S["lay"], S["x_obj"], S["obj_ind"], S["sub_len"], d_read, d_xobj3)

    # ---- GCN ----
    Ln, R = d.gcn_layers, d.gcn_residual
    gx, gp = [None] * (Ln + 1), [None] * (Ln + 1)
    gx[Ln] = d_xobj3

    def add_to(lst, i, g):
        lst[i] = g if lst[i] is None else ops.add(lst[i], g)

    def unit_bwd(l, u, src2, tmid, dm2):
        """backward of M = fc_rgt(fc_lft(src)); returns d(src)"""
        pre = _unit(l, u)
        acc_w(pre + "fc_rgt.weight", dm2, tmid); acc_b(pre + "fc_rgt.bias", dm2)
        dt = dx_of(dm2, pre + "fc_rgt.weight")
        acc_w(pre + "fc_lft.weight", dt, src2); acc_b(pre + "fc_lft.bias", dt)
        return dx_of(dt, pre + "fc_lft.weight")

    rel = S["rel_ind"]
    for l in range(Ln - 1, -1, -1):
        rec = S["layers"][l]
        if gx[l + 1] is not None and "y0" in rec:
            dxn = gx[l + 1]
            if rec["boundary"]:
                add_to(gx, l + 1 - R, dxn)
            dm0, dm1 = ops.gcn_node_bwd(dxn, rec["y0"], rec["y1"], rel)
            p2 = rec["p_in"].reshape(B * K, Lg)
            dsrc = ops.add(unit_bwd(l, 0, p2, rec["t0"], dm0.view(B * K, Lg)), unit_bwd(l, 1, p2, rec["t1"], dm1.view(B * K, Lg)))
            add_to(gp, l, dsrc.view(B, K, Lg))
        if gp[l + 1] is not None and "m2" in rec:
            dpn = gp[l + 1]
            if rec["boundary"]:
                add_to(gp, l + 1 - R, dpn)
            dm2, dm3 = ops.gcn_edge_bwd(dpn, rec["m2"], rec["m3"], rel)
            x2 = rec["x_in"].reshape(B * N, Lg)
            dsrc = ops.add(unit_bwd(l, 2, x2, rec["t2"], dm2.view(B * N, Lg)), unit_bwd(l, 3, x2, rec["t3"], dm3.view(B * N, Lg)))
            add_to(gx, l, dsrc.view(B, N, Lg))

    # ---- fusion ----
    d_x0 = ops.relu_bwd(S["x0"].view(B * N, Lg), gx[0].reshape(B * N, Lg))
    acc_w("obj_v_proj.weight", d_x0, S["att_feats"].reshape(B * N, -1)); acc_b("obj_v_proj.bias", d_x0)
    emb_rows = ops.gather_rows(P["sg_obj_embed.weight"], S["cls"])
    acc_w("obj_emb_proj.weight", d_x0, emb_rows); acc_b("obj_emb_proj.bias", d_x0)
    G["sg_obj_embed.weight"] = new_grad("sg_obj_embed.weight")
    ops.scatter_add_rows(dx_of(d_x0, "obj_emb_proj.weight"), S["cls"], G["sg_obj_embed.weight"])
    if reducer is not None:
        reducer.bucket_done(2, G)
        reducer.finish(G)
    return G
