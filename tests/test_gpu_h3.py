"""GPU tests of the split-fp16 ("h3") tensor-core path: the packed weight format, the contraction against fp64, the fp16-range
guard, and the whole path with packed weights against the fp32 (split-TF32) path and the oracle.

Bars: the packed copy reproduces the fp32 weight to 2^-21 relative (22 of 24 mantissa bits); contractions max-abs error
<= 2e-5 * max-abs(reference) like every other activation check (observed ~5e-7); token ids exact.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import subgc_oracle as O
from subgc import _lib, packing, synth
from subgc.config import SMALL, Dims, make_opt
from subgc.model import setup

pytestmark = pytest.mark.gpu
RTOL = 2e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _unpack(hi, lo, rows, cols, splits):
    """numpy restatement of the k-block-major layout of include/subgc_b200.h (subgc_packed)."""
    pts = [0] + list(splits or []) + [cols]
    h = hi.cpu().numpy().view(np.float16).astype(np.float64)
    l = lo.cpu().numpy().view(np.float16).astype(np.float64)
    out = np.zeros((rows, cols))
    kb0 = 0
    for a, b in zip(pts, pts[1:]):
        nkb = (b - a + 63) // 64
        blk_h = h[kb0 * rows * 64:(kb0 + nkb) * rows * 64].reshape(nkb, rows, 64)
        blk_l = l[kb0 * rows * 64:(kb0 + nkb) * rows * 64].reshape(nkb, rows, 64)
        full = (blk_h + blk_l / 2048.0).transpose(1, 0, 2).reshape(rows, nkb * 64)
        out[:, a:b] = full[:, :b - a]
        assert np.all(full[:, b - a:] == 0), "segment padding must be zero"
        kb0 += nkb
    return out


@pytest.mark.parametrize("rows,cols,splits", [(64, 64, None), (130, 1000, None), (400, 3000, [1000, 2000]), (96, 200, [72])])
def test_pack_round_trip(rows, cols, splits):
    g = torch.Generator().manual_seed(rows + cols)
    w = (torch.randn(rows, cols, generator=g) * torch.logspace(-6, 2, cols)[None, :]).cuda().contiguous()
    hi, lo, flag, _ = packing.pack_weight(w, splits)
    assert int(flag.item()) == 0
    got = _unpack(hi, lo, rows, cols, splits)
    ref = w.cpu().double().numpy()
    err = np.abs(got - ref)
    assert np.all(err <= np.abs(ref) * 2.0 ** -21 + 2.0 ** -36), float((err / (np.abs(ref) + 1e-30)).max())


def test_pack_flags_weights_beyond_fp16():
    w = torch.ones(64, 64, device="cuda")
    w[3, 5] = 7.0e4
    _, _, flag, _ = packing.pack_weight(w)
    assert int(flag.item()) == 1


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 512, 1000), (128, 4000, 3000), (37, 64, 72), (130, 9488, 1000), (700, 1024, 2048),
                                   (5, 4000, 3000)])
def test_packed_linear_against_fp64(M, N, K):
    L = _lib.lib()
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn(M + 2, K, generator=g)
    W = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
    b = torch.randn(N, generator=g)
    idx = torch.randint(0, M + 2, (M,), generator=g)
    ref = torch.relu(torch.nn.functional.linear(A[idx].double(), W.double(), b.double()))
    Ad, Wd, bd, idd = A.cuda(), W.cuda(), b.cuda(), idx.cuda()
    hi, lo, flag, segs = packing.pack_weight(Wd)
    pk = packing.packed_struct(Wd, hi, lo, segs)
    out = torch.full((M, N), float("nan"), device="cuda")
    ws = torch.empty(L.subgc_linear_workspace_bytes(M, N, K) + 256, dtype=torch.uint8, device="cuda")
    _lib.check(L.subgc_linear_packed_forward(M, N, K, Ad.data_ptr(), K, idd.data_ptr(), C.byref(pk), bd.data_ptr(), 1, out.data_ptr(), N,
                                             ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "linear_packed")
    o = out.cpu().double()
    assert not torch.isnan(o).any()
    assert float((o - ref).abs().max()) <= RTOL * float(ref.abs().max())


def _model(d, sd, **kw):
    m = setup(make_opt(d, test_LSTM=1, **kw))
    m.load_state_dict(sd)
    return m.cuda().eval()


@pytest.mark.parametrize("mode", ["greedy", "topk", "beam"])
def test_packed_path_matches_fp32_path_and_oracle(mode):
    """Full-size dims, several images: packed (h3) and unpacked (split-TF32) runs give the same tokens; both match the oracle."""
    d = Dims()
    sd = synth.make_state_dict(d, 11)
    data = synth.make_test_inputs(d, 11, n_images=6, per_half=2, ragged=True, ragged_edges=True)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
    args = [dev[k] for k in synth.SAMPLE_ARG_ORDER]
    kw = dict(gpn_nms_thres=0.55, gpn_max_subg=3, use_topk_sampling=1 if mode == "topk" else 0)
    T = d.seq_length
    outs = []
    for packed in (True, False):
        m = _model(d, sd, **kw)
        m.use_packed = packed
        opt = {"beam_size": 3 if mode == "beam" else 1}
        n_rows = None
        with torch.no_grad():
            if mode == "topk":
                first = m(*args, opt=dict(opt, seed=5), mode="sample")
                n_rows = first[0].shape[0]
                u = torch.rand(T, n_rows, generator=torch.Generator().manual_seed(3))
                opt["topk_uniforms"] = u
            for _ in range(3):   # eager, capture, replay
                res = m(*args, opt=opt, mode="sample")
        m.check_numerics()
        assert (m._weights().n_packs > 0) == packed
        outs.append([t.cpu() if torch.is_tensor(t) else t for t in res])
    (seq_p, lp_p, sc_p, keep_p), (seq_u, lp_u, sc_u, keep_u) = outs[0][:4], outs[1][:4]
    assert torch.equal(seq_p, seq_u) and torch.equal(keep_p, keep_u)
    assert float((lp_p - lp_u).abs().max()) <= RTOL * max(1.0, float(lp_u.abs().max()))
    okw = dict(use_nms=True, iou_thres=0.55, max_subgraphs=3)
    if mode == "topk":
        okw.update(topk=True, temp=0.6, k=3, uniforms=opt["topk_uniforms"])
    elif mode == "beam":
        okw.update(beam_size=3)
    with torch.no_grad():
        ref = O.sample(sd, d, data, **okw)
    assert torch.equal(seq_p, ref["seq"])
    assert float((lp_p - ref["seqLogprobs"]).abs().max()) <= RTOL * max(1.0, float(ref["seqLogprobs"].abs().max()))


def test_fp16_range_guard_repeats_the_call_on_the_fp32_path():
    """An embedding value beyond the fp16 range saturates in the split-fp16 copy.  The results of such a call must never reach the
    caller: the SAME call notices the device flag, warns, switches the model to the fp32 path and returns that path's results (which
    match the oracle).  Checked with the persistent decode kernel (full dims) and with the per-stage kernels (small dims)."""
    import warnings
    for d, seed in ((Dims(), 3), (SMALL, 3)):
        sd = synth.make_state_dict(d, seed)
        sd["embed.0.weight"] = sd["embed.0.weight"].clone()
        sd["embed.0.weight"][0, :] = 9.0e4          # row 0 = <bos>: fed at t = 0 of every caption
        data = synth.make_test_inputs(d, seed, n_images=2, per_half=2, ragged=True, ragged_edges=True)
        dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
        args = [dev[k] for k in synth.SAMPLE_ARG_ORDER]
        m = _model(d, sd, gpn_nms_thres=0.5, gpn_max_subg=2)
        with torch.no_grad():
            m._weights()
            had_packs = m._weights().n_packs > 0
            with warnings.catch_warnings(record=True) as caught:
                warnings.simplefilter("always")
                res = m(*args, opt={"beam_size": 1}, mode="sample")       # the FIRST call already returns valid results
            ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.5, max_subgraphs=2)
        if had_packs:
            assert any("fp16 range" in str(w.message) for w in caught), "the overflow must be reported"
            assert m.use_packed is False
        assert torch.equal(res[0].cpu(), ref["seq"])
        assert float((res[1].cpu() - ref["seqLogprobs"]).abs().max()) <= RTOL * max(1.0, float(ref["seqLogprobs"].abs().max()))
        with torch.no_grad():   # and the beam path does the same
            m2 = _model(d, sd, gpn_nms_thres=0.5, gpn_max_subg=2)
            with warnings.catch_warnings(record=True):
                warnings.simplefilter("always")
                res2 = m2(*args, opt={"beam_size": 2}, mode="sample")
            ref2 = O.sample(sd, d, data, use_nms=True, iou_thres=0.5, max_subgraphs=2, beam_size=2)
        assert torch.equal(res2[0], ref2["seq"])


@pytest.mark.parametrize("variant", ["SUBGC_FUSED_CELL", "SUBGC_MERGED", "SUBGC_FUSED_ATT", "SUBGC_NO_PDL"])
def test_opt_in_variants_match(variant):
    """Measured-but-not-default variants stay parity-green: SUBGC_FUSED_CELL=1 (gates -> cell inside the contraction, cluster split-K
    reduction over DSMEM), SUBGC_MERGED=1 (h2att + lang-early as one contraction), SUBGC_FUSED_ATT=1 (cell + h2att + attention as one
    cluster kernel), SUBGC_NO_PDL=1 (plain stream order)."""
    code = r'''
import sys, os, torch
sys.path.insert(0, os.path.join(%r, "sub-gc_b200")); sys.path.insert(0, os.path.join(%r, "oracle"))
import subgc_oracle as O
from subgc import synth
from subgc.config import Dims, make_opt
from subgc.model import setup
d = Dims(); sd = synth.make_state_dict(d, 21)
data = synth.make_test_inputs(d, 21, n_images=5, per_half=1, ragged=True, ragged_edges=True)
m = setup(make_opt(d, test_LSTM=1, gpn_nms_thres=0.75, gpn_max_subg=2)); m.load_state_dict(sd); m.cuda().eval()
args = [data[k].cuda() if data[k] is not None else None for k in synth.SAMPLE_ARG_ORDER]
with torch.no_grad():
    for _ in range(3): res = m(*args, opt={"beam_size": 1}, mode="sample")
    m.check_numerics()
    ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.75, max_subgraphs=2)
assert torch.equal(res[0].cpu(), ref["seq"]), "tokens differ"
err = float((res[1].cpu() - ref["seqLogprobs"]).abs().max())
assert err <= 2e-5 * max(1.0, float(ref["seqLogprobs"].abs().max())), err
print("OK", err)
''' % (ROOT, ROOT)
    env = dict(os.environ, **{variant: "1"})
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_benchmark_configuration_matches_oracle():
    """BASELINE config 2 at full size -- exactly what bench.py times (128 images x 36 nodes x 2048-d, one full sub-graph kept per
    image, greedy 20-token decode, the bench's seed): every token id equals the oracle's, log-probs within RTOL; three calls so
    that the compared result comes from the replayed CUDA graph."""
    d = Dims()
    sd = synth.make_state_dict(d, 2019)
    data = synth.make_test_inputs(d, 2019, n_images=128, per_half=1, ragged=False, ragged_edges=False)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
    args = [dev[k] for k in synth.SAMPLE_ARG_ORDER]
    m = _model(d, sd, gpn_nms_thres=0.75, gpn_max_subg=1)
    with torch.no_grad():
        for _ in range(3):
            seq, lps, score, keep = m(*args, opt={"beam_size": 1}, mode="sample")
        m.check_numerics()
        ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.75, max_subgraphs=1)
    assert seq.shape == (128, d.seq_length)
    assert torch.equal(keep.cpu(), ref["keep_ind"])
    assert torch.equal(seq.cpu(), ref["seq"]), int((seq.cpu() != ref["seq"]).sum())
    assert float((lps.cpu() - ref["seqLogprobs"]).abs().max()) <= RTOL * max(1.0, float(ref["seqLogprobs"].abs().max()))
    assert float((score.cpu() - ref["subgraph_score"]).abs().max()) <= RTOL


@pytest.mark.parametrize("mode,n_images", [("topk", 128), ("beam", 32)])
def test_other_baseline_configurations_match_oracle(mode, n_images):
    """BASELINE config 4 shard (top-k sampling, k=3, temp 0.6, 128 images; the same uniforms injected into oracle and kernel) and
    config 3 (beam 5; 32 images keep the oracle's CPU beam search at a few seconds): exact token ids, log-probs within RTOL."""
    d = Dims()
    sd = synth.make_state_dict(d, 2019)
    data = synth.make_test_inputs(d, 2019, n_images=n_images, per_half=1, ragged=False, ragged_edges=False)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
    args = [dev[k] for k in synth.SAMPLE_ARG_ORDER]
    m = _model(d, sd, gpn_nms_thres=0.75, gpn_max_subg=1, use_topk_sampling=1 if mode == "topk" else 0)
    opt = {"beam_size": 5 if mode == "beam" else 1}
    okw = dict(use_nms=True, iou_thres=0.75, max_subgraphs=1)
    if mode == "topk":
        u = torch.rand(d.seq_length, n_images, generator=torch.Generator().manual_seed(7))
        opt["topk_uniforms"] = u
        okw.update(topk=True, temp=0.6, k=3, uniforms=u)
    else:
        okw.update(beam_size=5)
    with torch.no_grad():
        for _ in range(3):
            res = m(*args, opt=opt, mode="sample")
        m.check_numerics()
        ref = O.sample(sd, d, data, **okw)
    seq, lps = res[0].cpu(), res[1].cpu()
    assert torch.equal(seq, ref["seq"]), int((seq != ref["seq"]).sum())
    assert float((lps - ref["seqLogprobs"]).abs().max()) <= RTOL * max(1.0, float(ref["seqLogprobs"].abs().max()))


def test_beam_attention_block_per_subgraph_equals_block_per_row(monkeypatch):
    """Beam search runs the attention of the b rows that share a context in one block per sub-graph (attention_beam_kernel: p_att / att read
    once); per row the operations are those of the block-per-row kernel, so sequences, log-probs and the finished-beam lists are identical."""
    d = Dims()
    sd = synth.make_state_dict(d, 41)
    data = synth.make_test_inputs(d, 41, n_images=6, per_half=2, ragged=True, ragged_edges=True)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}
    args = [dev[k] for k in synth.SAMPLE_ARG_ORDER]
    out = []
    for off in (False, True):
        if off:
            monkeypatch.setenv("SUBGC_NO_BEAM_ATT", "1")
        m = _model(d, sd, gpn_nms_thres=0.6, gpn_max_subg=3)
        with torch.no_grad():
            for _ in range(2):   # eager, then the captured graph
                res = m(*args, opt={"beam_size": 5}, mode="sample")
        out.append((res[0].clone(), res[1].clone(), [[(bm["seq"].clone(), bm["p"]) for bm in beams] for beams in m.done_beams]))
    (s0, l0, b0), (s1, l1, b1) = out
    assert torch.equal(s0, s1) and torch.equal(l0, l1)
    assert len(b0) == len(b1)
    for x, y in zip(b0, b1):
        assert len(x) == len(y) and all(torch.equal(a[0], c[0]) and a[1] == c[1] for a, c in zip(x, y))
