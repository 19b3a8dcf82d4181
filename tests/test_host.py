"""CPU-only checks of the host side: the C-ABI library builds, loads and exports exactly what include/subgc_b200.h
declares; the module keeps the reference's state_dict contract; and the product path refuses to run without CUDA
(no CPU fallback).  No compute entry point is called here."""
import ctypes
import os
import re

import pytest
import torch

from subgc import _lib, synth
from subgc.build import build
from subgc.config import SMALL, Dims, dims_from_opt, make_opt
from subgc.model import LossWrapper, TopDownModel, setup

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    return build()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "subgc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(subgc_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_declare_the_same_symbols():
    assert header_symbols() == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(libpath):
    handle = ctypes.CDLL(libpath)
    for name in header_symbols():
        assert hasattr(handle, name), name
    handle.subgc_version.restype = ctypes.c_int
    assert handle.subgc_version() == _lib.ABI_VERSION == 4


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.Dims) == 17 * 4
    assert ctypes.sizeof(_lib.Linear) == 16
    n_linear = 3 + 2 * 4 * _lib.MAX_GCN_LAYERS + 5 + 6
    # + packs pointer, n_packs (padded), overflow flag pointer, lang_early_w, + mega pointer, mega_bytes, mega_ctas (padded)
    # + gcn_fold (2 linears per layer) and gcn_fold_scale (2 floats per layer)
    assert ctypes.sizeof(_lib.Weights) == 16 * n_linear + 8 * (3 + 8) + 32 + 24 + _lib.MAX_GCN_LAYERS * (2 * 16 + 2 * 4) + 16
    assert ctypes.sizeof(_lib.Packed) == 3 * 8 + 8 * 4
    assert ctypes.sizeof(_lib.Layout) == 16
    assert ctypes.sizeof(_lib.DecoderTrainBufs) == 26 * 8 and ctypes.sizeof(_lib.DecoderGrads) == 14 * 8


@pytest.mark.parametrize("d", [SMALL, Dims()])
def test_state_dict_contract(d):
    m = setup(make_opt(d, test_LSTM=1))
    sd = m.state_dict()
    want = synth.param_shapes(d)
    assert list(sd.keys()) == list(want.keys())
    for k, shape in want.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    if d is not SMALL:
        assert sum(v.numel() for v in sd.values()) == 70048210  # SURVEY headline fact
    m.load_state_dict(synth.make_state_dict(d, 1), strict=True)


def test_reference_attributes_and_dispatch():
    m = setup(make_opt(SMALL, test_LSTM=1, gpn_nms_thres=0.5, gpn_max_subg=7, use_topk_sampling=1, the_k=4, topk_temp=0.7))
    assert isinstance(m, TopDownModel)
    assert m.gpn and m.seq_length == SMALL.seq_length and m.vocab_size == SMALL.vocab and m.num_layers == 2
    assert m.gpn_layer.iou_thres == 0.5 and m.gpn_layer.max_subgraphs == 7 and m.gpn_layer.use_nms
    assert m.topk_sampling and m.the_k == 4 and m.topk_temp == 0.7
    m.ss_prob = 0.25  # train.py:131 writes it
    h, c = m.init_hidden(3)
    assert h.shape == (2, 3, SMALL.rnn) and float(h.abs().sum()) == 0
    lw = LossWrapper(m, None)
    assert hasattr(lw, "crit")
    with pytest.raises(Exception):
        setup(make_opt(SMALL, caption_model="fc"))


def test_no_cpu_fallback():
    d = SMALL
    m = setup(make_opt(d, test_LSTM=1)).eval()
    data = synth.make_test_inputs(d, 1)
    with pytest.raises(_lib.SubgcError):
        m(*synth.sample_args(data), opt={"beam_size": 1}, mode="sample")
    with pytest.raises(_lib.SubgcError):
        m.get_logprobs_state(torch.zeros(2, dtype=torch.long), torch.zeros(2, d.rnn), torch.zeros(2, 3, d.rnn),
                             torch.zeros(2, 3, d.att_hid), torch.ones(2, 3), m.init_hidden(2))


def test_unsupported_variants_are_rejected():
    for over in (dict(use_gpn=0), dict(noun_fuse=0), dict(gcn_bn=1), dict(use_bn=1)):
        with pytest.raises(NotImplementedError):
            dims_from_opt(make_opt(Dims(), **over))


def test_pack_segment_descriptor_and_abi_entries():
    """Host side of the packed-weight format (include/subgc_b200.h: subgc_packed): column cuts -> seg_col array; the element count
    of a packed copy (k-block-major, every K segment padded to 64 columns) comes from the library itself (no device needed)."""
    from subgc import packing
    seg, n = packing._seg_array(3000, [1000, 2000])
    assert n == 3 and list(seg)[:4] == [0, 1000, 2000, 3000]
    seg1, n1 = packing._seg_array(300, None)
    assert n1 == 1 and list(seg1)[:2] == [0, 300]
    L = _lib.lib()
    assert L.subgc_pack_elems(4000, n, seg) == 3 * 16 * 4000 * 64      # 3 segments x ceil(1000 / 64) k-blocks x rows x 64
    assert L.subgc_pack_elems(1024, n1, seg1) == 5 * 1024 * 64
    with pytest.raises(AssertionError):
        packing._seg_array(100, [50, 40])


def test_compact_batch_host_side():
    """subgc.compact.compact_batch is loader-side host code: class ids as the reference's own arg-max (AttModel.py:374), one copy of
    the sub-graph tensors per image, lengths = ones in att_masks = trace of gpn_pool_mtx (dataloader_test.py:280-286)."""
    import torch
    from subgc import compact, synth
    from subgc.config import SMALL
    data = synth.make_test_inputs(SMALL, 4, n_images=3, per_half=2, ragged=True, ragged_edges=True)
    host = [data[k] for k in synth.SAMPLE_ARG_ORDER]
    cb = compact.compact_batch(*host, with_pred=True)
    B, N, K = 3, SMALL.obj_num, SMALL.rel_num
    assert cb.obj_cls.dtype == torch.int16 and tuple(cb.obj_cls.shape) == (B, N)
    assert torch.equal(cb.obj_cls.long(), data["obj_dist"][:, :, 1:].max(2)[1] + 1)
    assert torch.equal(cb.pred_cls.long(), data["pred_dist"][:, :, 1:].max(2)[1] + 1)
    assert cb.rel_ind.dtype == torch.uint8 and torch.equal(cb.rel_ind.long(), data["rel_ind"])
    assert tuple(cb.sub_nodes.shape) == (B, 2, 2, N) and torch.equal(cb.sub_nodes.long(), data["gpn_obj_ind"][::5])
    assert torch.equal(cb.sub_len.float(), data["att_masks"][::5].sum(-1))
    assert torch.equal(cb.sub_len.float(), torch.diagonal(data["gpn_pool_mtx"][::5], dim1=-2, dim2=-1).sum(-1))
    lean = compact.needed_only(*host)
    assert lean[0] is None and lean[8] is None and lean[10] is None and lean[11] is None and lean[12] is None and lean[1] is host[1]
    assert cb.nbytes() < sum(t.numel() * t.element_size() for t in host if t is not None) / 1.5
    # packed(): every field a view into ONE buffer (the host->device transfer of a batch is a single copy)
    pk = cb.packed()
    assert pk._buf.dtype == torch.uint8 and pk._buf.numel() >= cb.nbytes()
    for name in ("att_feats", "obj_cls", "pred_cls", "rel_ind", "sub_nodes", "sub_len"):
        a, b = getattr(cb, name), getattr(pk, name)
        assert torch.equal(a, b) and b.is_contiguous()
        assert pk._buf.data_ptr() <= b.data_ptr() < pk._buf.data_ptr() + pk._buf.numel() and b.data_ptr() % 256 == pk._buf.data_ptr() % 256
    dst = cb.packed(copy=False)
    dst.copy_(pk)
    assert torch.equal(dst._buf, pk._buf) and torch.equal(dst.sub_nodes, cb.sub_nodes)
    dst2 = cb.empty_like("cpu")
    dst2.copy_(pk)                         # field-wise path (no common buffer)
    assert torch.equal(dst2.att_feats, cb.att_feats)


def test_done_beams_is_a_lazy_list_of_lists():
    """`model.done_beams` after a beam search (reference models/AttModel.py:229-231): per sub-graph the finished beams, best first, each a
    dict seq / logps / unaug_p / p.  The product builds the dicts when they are looked at; indexing, length, iteration, slicing and
    comparison must behave like the reference's list of lists."""
    from subgc.model import _DoneBeams
    g = torch.Generator().manual_seed(3)
    n, b, T = 4, 3, 5
    seq = torch.randint(0, 50, (n, b, T), generator=g)
    lps = -torch.rand(n, b, T, generator=g)
    up = -torch.rand(n, b, generator=g, dtype=torch.float64)
    p = -torch.rand(n, b, generator=g, dtype=torch.float64)
    cnt = torch.tensor([3, 2, 0, 1], dtype=torch.int32)
    want = [[dict(seq=seq[k, j], logps=lps[k, j], unaug_p=float(up[k, j]), p=float(p[k, j])) for j in range(int(cnt[k]))] for k in range(n)]
    got = _DoneBeams(seq, lps, up, p, cnt)
    assert len(got) == n and [len(x) for x in got] == [3, 2, 0, 1]
    for k, (beams, exp) in enumerate(zip(got, want)):
        assert beams is got[k]            # built once
        for a, e in zip(beams, exp):
            assert torch.equal(a["seq"], e["seq"]) and torch.equal(a["logps"], e["logps"])
            assert a["p"] == e["p"] and a["unaug_p"] == e["unaug_p"] and isinstance(a["p"], float)
    assert got[-1] is got[n - 1] and got[1:3] == [got[1], got[2]]
    with pytest.raises(IndexError):
        got[n]
