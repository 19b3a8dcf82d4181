"""ctypes binding of the C ABI declared in include/subgc_b200.h (libsubgc_b200.so, built by `subgc.build`).

There is no CPU fallback: `lib()` raises if the shared library is missing or cannot be loaded, and every wrapper
raises `SubgcError` with the library's own message on a non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsubgc_b200.so")
MAX_GCN_LAYERS = 8
ABI_VERSION = 4   # include/subgc_b200.h: SUBGC_ABI_VERSION (struct layouts below mirror that header)

c_fp = C.c_void_p  # device pointers travel as plain integers


class SubgcError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("vocab1", "enc", "rnn", "att_hid", "fc_feat", "att_feat", "gcn", "low_rank", "embed",
                                         "obj_classes", "pred_classes", "gcn_layers", "gcn_residual", "pred_emb_type",
                                         "seq_length", "obj_num", "rel_num")]


class Linear(C.Structure):
    _fields_ = [("w", c_fp), ("b", c_fp)]


class Packed(C.Structure):
    _fields_ = [("w", c_fp), ("hi", c_fp), ("lo", c_fp), ("rows", C.c_int32), ("cols", C.c_int32), ("n_seg", C.c_int32),
                ("seg_col", C.c_int32 * 5)]


class Weights(C.Structure):
    _fields_ = [
        ("obj_v_proj", Linear), ("sg_obj_embed", c_fp), ("obj_emb_proj", Linear), ("sg_pred_embed", c_fp),
        ("pred_emb_prj", Linear), ("gcn_lft", (Linear * 4) * MAX_GCN_LAYERS), ("gcn_rgt", (Linear * 4) * MAX_GCN_LAYERS),
        ("gpn_fc0", Linear), ("gpn_fc3", Linear), ("read_out0", Linear), ("read_out1", Linear), ("logit", Linear),
        ("embed", c_fp), ("fc_embed0", Linear), ("fc_embed2", Linear), ("att_embed", Linear), ("ctx2att", Linear),
        ("h2att", Linear), ("alpha_net", Linear),
        ("att_w_ih", c_fp), ("att_w_hh", c_fp), ("att_b_ih", c_fp), ("att_b_hh", c_fp),
        ("lang_w_ih", c_fp), ("lang_w_hh", c_fp), ("lang_b_ih", c_fp), ("lang_b_hh", c_fp),
        ("packs", C.POINTER(Packed)), ("n_packs", C.c_int32), ("h3_overflow", c_fp), ("lang_early_w", c_fp),
        ("mega", c_fp), ("mega_bytes", C.c_uint64), ("mega_ctas", C.c_int32),
        ("gcn_fold", (Linear * 2) * MAX_GCN_LAYERS), ("gcn_fold_scale", (C.c_float * 2) * MAX_GCN_LAYERS),
        ("prep_fold", Linear),
    ]


class DecoderTrainBufs(C.Structure):   # subgc_decoder_train_bufs
    _fields_ = [(n, c_fp) for n in ("tokens", "fc", "att", "p_att", "masks", "m_x", "m_h", "xt", "act1", "c_att", "h_att", "atth", "ctx", "alpha",
                                    "sm", "act2", "c_lang", "h_lang", "hd", "outputs", "logits", "targets", "tmask", "lse", "nll", "coef")]


class DecoderGrads(C.Structure):       # subgc_decoder_grads
    _fields_ = [(n, c_fp) for n in ("logit_w", "logit_b", "embed", "att_w_ih", "att_w_hh", "att_b_ih", "att_b_hh", "lang_w_ih", "lang_w_hh",
                                    "lang_b_ih", "lang_b_hh", "h2att_w", "h2att_b", "alpha_w")]


def _ptr_struct(names):
    return [(n, c_fp) for n in names]


class PrepareSaved(C.Structure):
    _fields_ = _ptr_struct(("m_fc", "fc_pre", "f1", "g_fc", "hr", "read_sel", "att", "att_pre", "m_att", "x_rows", "node_row"))


class PrepareGrads(C.Structure):
    _fields_ = _ptr_struct(("fc2_w", "fc2_b", "fc0_w", "fc0_b", "ro1_w", "ro1_b", "ro0_w", "ro0_b", "ctx2att_w", "ctx2att_b", "att_embed_w", "att_embed_b"))


class SgpnSaved(C.Structure):
    _fields_ = _ptr_struct(("score", "hid", "hid_d", "m_gpn", "read_out", "sub_len"))


class SgpnGrads(C.Structure):
    _fields_ = _ptr_struct(("fc3_w", "fc3_b", "fc0_w", "fc0_b"))


class GcnLayerSaved(C.Structure):
    _fields_ = _ptr_struct(("x_in", "p_in", "t0", "t1", "y0", "y1", "t2", "t3", "m2", "m3"))


class GcnSaved(C.Structure):
    _fields_ = [("layer", GcnLayerSaved * MAX_GCN_LAYERS), ("x0", c_fp), ("att_feats", c_fp), ("cls", c_fp), ("rel_ind", c_fp)]


class GcnGrads(C.Structure):
    _fields_ = [("lft_w", (c_fp * 4) * MAX_GCN_LAYERS), ("lft_b", (c_fp * 4) * MAX_GCN_LAYERS), ("rgt_w", (c_fp * 4) * MAX_GCN_LAYERS),
                ("rgt_b", (c_fp * 4) * MAX_GCN_LAYERS), ("obj_v_w", c_fp), ("obj_v_b", c_fp), ("obj_emb_w", c_fp), ("obj_emb_b", c_fp),
                ("sg_obj_embed", c_fp)]


class Layout(C.Structure):
    _fields_ = [("rows", C.c_int32), ("per_half", C.c_int32), ("seq_per_img", C.c_int32), ("order", C.c_int32)]


_P = C.POINTER
_i, _sz, _f, _d, _u64 = C.c_int, C.c_size_t, C.c_float, C.c_double, C.c_uint64

# name -> (restype, argtypes); mirrors include/subgc_b200.h one to one
SIGNATURES = {
    "subgc_last_error": (C.c_char_p, []),
    "subgc_version": (_i, []),
    "subgc_launch_count": (C.c_ulonglong, []),
    "subgc_debug_att_trace": (_i, [c_fp, _i]),
    "subgc_debug_trace": (_i, [_i, c_fp, c_fp, _i]),
    "subgc_debug_mega_trace": (_i, [c_fp, _i, _i]),
    "subgc_pack_elems": (_sz, [_i, _i, C.POINTER(C.c_int32)]),
    "subgc_pack_weight": (_i, [_i, _i, c_fp, _i, _i, C.POINTER(C.c_int32), c_fp, c_fp, c_fp, c_fp]),
    "subgc_linear_packed_forward": (_i, [_i, _i, _i, c_fp, _i, c_fp, _P(Packed), c_fp, _i, c_fp, _i, c_fp, _sz, c_fp]),
    "subgc_mega_timing": (_i, [_i, _P(_f)]),
    "subgc_mega_pack_bytes": (_sz, [_P(Dims), _i]),
    "subgc_mega_pack": (_i, [_P(Dims), _P(Weights), _i, c_fp, _sz, c_fp, c_fp]),
    "subgc_linear_workspace_bytes": (_sz, [_i, _i, _i]),
    "subgc_linear_forward": (_i, [_i, _i, _i, c_fp, _i, c_fp, c_fp, _i, c_fp, _i, c_fp, _i, c_fp, _sz, c_fp]),
    "subgc_encoder_workspace_bytes": (_sz, [_P(Dims), _i]),
    "subgc_fuse_nodes": (_i, [_P(Dims), _P(Weights), _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_fuse_nodes_cls": (_i, [_P(Dims), _P(Weights), _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_gcn_forward": (_i, [_P(Dims), _P(Weights), _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_gcn_needs_pred": (_i, [_P(Dims), _i]),
    "subgc_sgpn_workspace_bytes": (_sz, [_P(Dims), _i]),
    "subgc_sgpn_forward": (_i, [_P(Dims), _P(Weights), _P(Layout), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_sgpn_select_train": (_i, [_P(Layout), c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_nms_workspace_bytes": (_sz, [_i, _i]),
    "subgc_subgraph_nms": (_i, [_P(Dims), _P(Layout), c_fp, c_fp, c_fp, c_fp, _i, _d, _i, c_fp, c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_rank_rows": (_i, [_i, c_fp, c_fp, c_fp, c_fp]),
    "subgc_prepare_workspace_bytes": (_sz, [_P(Dims), _i, _i]),
    "subgc_prepare_forward": (_i, [_P(Dims), _P(Weights), _P(Layout), _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                   c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_decode_workspace_bytes": (_sz, [_P(Dims), _i, _i]),
    "subgc_decode_step": (_i, [_P(Dims), _P(Weights), _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                               c_fp, _sz, c_fp]),
    "subgc_decode_sample": (_i, [_P(Dims), _P(Weights), _i, _i, _i, _f, _i, _u64, _u64, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                 c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_decode_sample_dyn": (_i, [_P(Dims), _P(Weights), _i, _i, c_fp, _i, _f, _i, _u64, _u64, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                     c_fp, _sz, c_fp]),
    "subgc_teacher_workspace_bytes": (_sz, [_P(Dims), _i, _i]),
    "subgc_decode_teacher": (_i, [_P(Dims), _P(Weights), _i, _i, _i, c_fp, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _sz, c_fp]),
    "subgc_beam_workspace_bytes": (_sz, [_P(Dims), _i, _i, _i]),
    "subgc_decode_beam": (_i, [_P(Dims), _P(Weights), _i, _i, _i, _i, _d, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                               c_fp, _sz, c_fp]),
    "subgc_opt_chunk_elems": (_i, []),
    "subgc_clip_adam_step": (_i, [c_fp, _i, _f, _f, _f, _f, _f, _f, _i, _i, c_fp, c_fp, c_fp]),
    # training building blocks
    "subgc_gemm_nt_workspace_bytes": (_sz, [_i, _i, _i]),
    "subgc_gemm_nt": (_i, [_i, _i, _i, c_fp, _i, c_fp, c_fp, _i, c_fp, _i, _i, c_fp, _i, c_fp, _sz, c_fp]),
    "subgc_transpose": (_i, [_i, _i, c_fp, _i, c_fp, _i, c_fp]),
    "subgc_colsum": (_i, [_i, _i, c_fp, _i, c_fp, _i, c_fp]),
    "subgc_ew": (_i, [_i, _sz, c_fp, c_fp, c_fp, _f, c_fp]),
    "subgc_dropout_mask": (_i, [_sz, _f, _u64, _u64, c_fp, c_fp]),
    "subgc_gather_rows": (_i, [_i, _i, c_fp, _i, c_fp, c_fp, _i, c_fp]),
    "subgc_ss_sample": (_i, [_i, _i, c_fp, _sz, c_fp, _i, _f, _u64, _u64, c_fp, c_fp]),
    "subgc_unary": (_i, [_i, _sz, c_fp, c_fp, c_fp]),
    "subgc_scatter_add_rows": (_i, [_i, _i, c_fp, _i, c_fp, c_fp, _i, c_fp]),
    "subgc_lstm_cell_train_fwd": (_i, [_i, _i, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_mean_nodes": (_i, [_i, _i, _i, c_fp, c_fp, c_fp]),
    "subgc_frontend_backward_workspace_bytes": (_sz, [_P(Dims), _i, _i, _i, _i]),
    "subgc_prepare_backward": (_i, [_P(Dims), _P(Weights), _i, _i, _i, _P(PrepareSaved), c_fp, c_fp, c_fp, _P(PrepareGrads), c_fp, c_fp, _sz, c_fp]),
    "subgc_sgpn_backward": (_i, [_P(Dims), _P(Weights), _P(Layout), _P(SgpnSaved), _f, c_fp, c_fp, _P(SgpnGrads), c_fp, c_fp, _sz, c_fp]),
    "subgc_gcn_backward": (_i, [_P(Dims), _P(Weights), _i, _P(GcnSaved), c_fp, _P(GcnGrads), c_fp, _sz, c_fp]),
    "subgc_decoder_train_workspace_bytes": (_sz, [_P(Dims), _i, _i, _i]),
    "subgc_decoder_train_forward": (_i, [_P(Dims), _P(Weights), _i, _i, _i, _i, _P(DecoderTrainBufs), c_fp, _sz, c_fp]),
    "subgc_decoder_train_backward": (_i, [_P(Dims), _P(Weights), _i, _i, _i, _i, _P(DecoderTrainBufs), c_fp, _P(DecoderGrads), c_fp, c_fp, c_fp,
                                          c_fp, _sz, c_fp]),
    "subgc_lstm_cell_bwd": (_i, [_i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_attention_train_fwd": (_i, [_i, _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_attention_bwd": (_i, [_i, _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_log_softmax_fwd": (_i, [_i, _i, c_fp, c_fp, _sz, c_fp]),
    "subgc_log_softmax_bwd": (_i, [_i, _i, c_fp, c_fp, _sz, c_fp, c_fp]),
    "subgc_class_argmax": (_i, [_i, _i, _i, c_fp, c_fp, c_fp]),
    "subgc_sgpn_pool": (_i, [_P(Dims), _P(Layout), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_sgpn_bce": (_i, [_P(Layout), c_fp, c_fp, c_fp]),
    "subgc_sgpn_pool_bwd": (_i, [_P(Dims), _P(Layout), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_bce_sigmoid_bwd": (_i, [_P(Layout), c_fp, _f, c_fp, c_fp]),
    "subgc_prepare_index": (_i, [_P(Dims), _P(Layout), _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_gcn_edge_fwd": (_i, [_i, _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_gcn_node_train_fwd": (_i, [_i, _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_gcn_node_bwd": (_i, [_i, _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "subgc_gcn_edge_bwd": (_i, [_i, _i, _i, _i, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
}

_lib = None


def lib():
    """Load libsubgc_b200.so (once).  Raises if it has not been built — the product path never falls back."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise SubgcError(f"{LIB_PATH} not found: build it with `python -m subgc.build` "
                             "(or __graft_entry__.build()); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.subgc_version() != ABI_VERSION:
            raise SubgcError(f"{LIB_PATH} has ABI version {handle.subgc_version()}, this package expects {ABI_VERSION}: "
                             "rebuild it with `python -m subgc.build --force`")
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().subgc_last_error()
        raise SubgcError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
