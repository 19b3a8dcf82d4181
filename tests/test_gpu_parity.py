"""GPU parity tests: the CUDA path (through the C ABI, via subgc.model) against
  (a) the committed golden fixtures produced by the REAL reference (tests/golden, oracle/make_golden.py), and
  (b) the oracle (oracle/subgc_oracle.py) on freshly seeded inputs, incl. multi-image batches and full-size shapes.

Tolerances (fp32 path, fp32 FMA accumulation in a different summation order than MKL):
  * integer outputs (token ids, beam sequences, NMS keep sets, argmax classes): exact,
  * activations / scores / log-probs: max-abs error <= RTOL * max-abs(reference), RTOL = 2e-5
    (SURVEY §8c measured the fp32-vs-fp64 noise floor of the reference itself at ~5e-7 relative).
"""
import numpy as np
import pytest
import torch

import subgc_oracle as O
from helpers import check_train_outputs, beam_sizes_in, load_golden, rebuild_test_case, rebuild_train_case, rel_err, t2n
from subgc import synth
from subgc.config import SMALL, Dims, make_opt
from subgc.model import LossWrapper, setup

pytestmark = pytest.mark.gpu
RTOL = 2e-5
TEST_CASES = ["small_test_ragged", "small_test_nms", "small_test_full", "full_test", "full_test_peaked"]


def to_dev(data):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in data.items()}


def make_model(d, sd, **opt_over):
    opt_over.setdefault("test_LSTM", 1)
    m = setup(make_opt(d, **opt_over))
    m.load_state_dict(sd)
    return m.cuda().eval()


@pytest.fixture(scope="module", params=TEST_CASES)
def case(request):
    g = load_golden(request.param)
    d, sd, data, nms = rebuild_test_case(g)
    model = make_model(d, sd, gpn_nms_thres=nms["iou_thres"], gpn_max_subg=nms["max_subgraphs"])
    return g, d, sd, data, nms, model, to_dev(data)


# ---------------------------------------------------------------------------------------------------------------
# building block
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(1, 7, 5), (37, 24, 48), (128, 4000, 1000), (4736, 1024, 300), (130, 9488, 1000), (64, 512, 2048)])
def test_linear_block(M, N, K):
    import ctypes as C
    from subgc import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(M * 131 + N)
    A = torch.randn(M + 3, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    idx = torch.randint(0, M + 3, (M,), generator=g)
    ref = torch.relu(torch.nn.functional.linear(A[idx].double(), W.double(), b.double())).float()
    Ad, Wd, bd, idd = A.cuda(), W.cuda(), b.cuda(), idx.cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    wsb = L.subgc_linear_workspace_bytes(M, N, K)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    _lib.check(L.subgc_linear_forward(M, N, K, Ad.data_ptr(), K, idd.data_ptr(), Wd.data_ptr(), K, bd.data_ptr(), 1, out.data_ptr(), N,
                                      ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "linear")
    torch.cuda.synchronize()
    assert rel_err(out.cpu(), ref) <= 5e-6


# ---------------------------------------------------------------------------------------------------------------
# golden fixtures (outputs of the real reference)
# ---------------------------------------------------------------------------------------------------------------
def test_golden_stages(case):
    g, d, sd, data, nms, model, dev = case
    with torch.no_grad():
        (g_fc, fc, att, p_att, masks), score, keep, n_rows, len_max = model._front(dev["att_feats"], dev["att_masks"], dev["obj_dist"],
                                                                                dev["rel_ind"], dev["pred_dist"], dev["gpn_obj_ind"])
        x_obj = model.last_x_obj[0]
    if "x_obj" in g.files:
        assert rel_err(t2n(model._x0[0]), g["x0"][0]) <= RTOL
        assert rel_err(t2n(x_obj), g["x_obj"]) <= RTOL
        assert rel_err(t2n(fc), g["p_fc"]) <= RTOL
        assert rel_err(t2n(att), g["p_att"]) <= RTOL
        assert rel_err(t2n(p_att), g["pp_att"]) <= RTOL
        # padded att rows are exact zeros (pack_padded_sequence semantics)
        assert np.array_equal(t2n(att) == 0, g["p_att"] == 0)
    else:
        assert rel_err(t2n(x_obj[:, :64]), g["x_obj_slice"]) <= RTOL
        assert abs(float(x_obj.double().sum()) - float(g["x_obj_sum"])) <= RTOL * float(g["x_obj_abssum"])
        assert rel_err(t2n(fc[:, :64]), g["p_fc_slice"]) <= RTOL
        assert rel_err(t2n(att[:, :, :32]), g["p_att_slice"]) <= RTOL
        assert rel_err(t2n(p_att[:, :, :32]), g["pp_att_slice"]) <= RTOL
    assert np.array_equal(t2n(keep), g["keep_ind"])
    assert rel_err(t2n(score), g["gpn_score"]) <= RTOL
    assert abs(float(model.last_gpn_loss) - float(g["gpn_loss"])) <= RTOL * 10
    assert np.array_equal(t2n(masks), g["p_mask"])
    assert rel_err(t2n(g_fc[:, :64]), g["g_fc_slice"]) <= RTOL
    # one decoder step from <bos>
    S = fc.shape[0]
    with torch.no_grad():
        lp0, st0 = model.get_logprobs_state(torch.zeros(S, dtype=torch.long, device="cuda"), fc, att, p_att, masks, model.init_hidden(S))
    if "step0_logprobs" in g.files:
        assert rel_err(t2n(lp0), g["step0_logprobs"]) <= RTOL
        assert rel_err(t2n(st0[0]), g["step0_h"]) <= RTOL and rel_err(t2n(st0[1]), g["step0_c"]) <= RTOL
    else:
        top = lp0.topk(5, 1)
        assert np.array_equal(t2n(top[1]), g["step0_top5_idx"])
        assert rel_err(t2n(top[0]), g["step0_top5_val"]) <= RTOL
        assert rel_err(t2n(st0[0][:, :, :64]), g["step0_h_slice"]) <= RTOL


def test_golden_edge_stream(case):
    """The predicate half of the encoder: p0 = W_p E_pred[argmax] + b_p (AttModel.py:382-387) and x_pred, the edge
    output of gcn_backbone (models/lib/gcn_backbone.py:29-53: layer-0 units 0/1, layer-1 units 2/3, residual on the edge stream), which
    the Sub-GC decoder never reads and encode() therefore skips unless asked."""
    g, d, sd, data, nms, model, dev = case
    with torch.no_grad():
        x_obj, x_pred = model.encode(dev["att_feats"], dev["obj_dist"], dev["pred_dist"], dev["rel_ind"], want_x_pred=True)
        p0 = model._p0
        x_obj_only = model.encode(dev["att_feats"], dev["obj_dist"], dev["pred_dist"], dev["rel_ind"])
        assert model._p0 is None, "without the edge output the predicate embedding is not computed at all (Sub-GC: dead sub-path)"
    if "x_pred" in g.files:   # fixtures of the real reference
        ref_p0, ref_xp, ref_xo = g["p0"][0], g["x_pred"], g["x_obj"]
    else:                      # full size: the oracle (pinned to the reference by the small fixtures)
        with torch.no_grad():
            rx0, rp0 = O.fuse_features(sd, d, data["att_feats"], data["obj_dist"], data["pred_dist"])
            rxo, rxp = O.gcn_encode(sd, d, rx0, rp0, data["rel_ind"])
        ref_p0, ref_xp, ref_xo = rp0[0].numpy(), rxp[0].numpy(), rxo[0].numpy()
    assert x_pred.shape == (1, d.rel_num, d.gcn)
    assert rel_err(t2n(p0[0]), ref_p0) <= RTOL
    assert rel_err(t2n(x_pred[0]), ref_xp) <= RTOL
    assert rel_err(t2n(x_obj[0]), ref_xo) <= RTOL
    assert torch.equal(x_obj, x_obj_only), "x_obj must not depend on whether the edge stream is materialised"


def test_edge_stream_multi_image_against_oracle():
    d = Dims()
    sd = synth.make_state_dict(d, 5)
    data = synth.make_test_inputs(d, 5, n_images=3, per_half=1, ragged=True, ragged_edges=True)
    model = make_model(d, sd)
    dev = to_dev(data)
    with torch.no_grad():
        x_obj, x_pred = model.encode(dev["att_feats"], dev["obj_dist"], dev["pred_dist"], dev["rel_ind"], want_x_pred=True)
        rxo, rxp = O.encode(sd, d, data["att_feats"], data["obj_dist"], data["pred_dist"], data["rel_ind"])
    assert rel_err(t2n(x_pred), rxp.numpy()) <= RTOL and rel_err(t2n(x_obj), rxo.numpy()) <= RTOL


def test_golden_greedy(case):
    g, d, sd, data, nms, model, dev = case
    with torch.no_grad():
        seq, lps, score, keep, attw = model(*synth.sample_args(dev), opt={"beam_size": 1, "return_att": 1}, mode="sample")
    assert seq.dtype == torch.int64 and seq.is_cuda
    assert np.array_equal(t2n(seq), g["greedy_seq"])
    assert rel_err(t2n(lps), g["greedy_logprobs"]) <= RTOL
    assert np.array_equal(t2n(keep), g["greedy_keep"])
    assert rel_err(t2n(score), g["greedy_score"]) <= RTOL
    assert tuple(attw.shape) == g["greedy_att_weights"].shape
    assert rel_err(t2n(attw), g["greedy_att_weights"]) <= RTOL
    # without attention weights the discarded last step is skipped (and, at full size, the loop runs as the persistent kernel, whose
    # split-K partials cover different k-ranges): same tokens, log-probs within the bar
    with torch.no_grad():
        seq2, lps2, _, _ = model(*synth.sample_args(dev), opt={"beam_size": 1}, mode="sample")
    assert torch.equal(seq, seq2)
    assert rel_err(t2n(lps2), g["greedy_logprobs"]) <= RTOL


def test_golden_beam(case):
    g, d, sd, data, nms, model, dev = case
    for b in beam_sizes_in(g):
        with torch.no_grad():
            seq, lps, score, keep = model(*synth.sample_args(dev), opt={"beam_size": b, "length_penalty": str(g["meta_length_penalty"])},
                                          mode="sample")
        assert not seq.is_cuda  # the reference returns CPU tensors on the beam path (AttModel.py:212-213)
        assert np.array_equal(t2n(seq), g[f"beam{b}_seq"])
        assert rel_err(t2n(lps), g[f"beam{b}_logprobs"]) <= RTOL
        gs, gl, gp, gu = g[f"beam{b}_beam_seq"], g[f"beam{b}_beam_logps"], g[f"beam{b}_beam_p"], g[f"beam{b}_beam_unaug_p"]
        assert len(model.done_beams) == gs.shape[0]
        for s, beams in enumerate(model.done_beams):
            n_ref = int((~np.isnan(gp[s])).sum())
            assert len(beams) == n_ref
            for j, bm in enumerate(beams):
                assert np.array_equal(t2n(bm["seq"]), gs[s, j])
                assert rel_err(t2n(bm["logps"]), gl[s, j]) <= RTOL
                assert abs(bm["p"] - gp[s, j]) <= 1e-5 * max(1.0, abs(gp[s, j]))
                assert abs(bm["unaug_p"] - gu[s, j]) <= 1e-5 * max(1.0, abs(gu[s, j]))


# ---------------------------------------------------------------------------------------------------------------
# oracle on fresh inputs
# ---------------------------------------------------------------------------------------------------------------
def test_topk_sampling_with_injected_uniforms(case):
    """RNG streams cannot be matched (SURVEY §7 hard part 4): inject the same uniforms into oracle and kernel."""
    g, d, sd, data, nms, model, dev = case
    with torch.no_grad():
        S = O.sample(sd, d, data, use_nms=True, **nms)["seq"].shape[0]
        u = torch.rand(d.seq_length, S, generator=torch.Generator().manual_seed(5))
        ref = O.sample(sd, d, data, use_nms=True, topk=True, temp=0.6, k=3, uniforms=u, **nms)
        model.topk_sampling, model.topk_temp, model.the_k = True, 0.6, 3
        try:
            seq, lps, _, _ = model(*synth.sample_args(dev), opt={"beam_size": 1, "topk_uniforms": u}, mode="sample")
            seq_p, lps_p, _, _ = model(*synth.sample_args(dev), opt={"beam_size": 1, "seed": 1234}, mode="sample")
            seq_p2, _, _, _ = model(*synth.sample_args(dev), opt={"beam_size": 1, "seed": 1234}, mode="sample")
        finally:
            model.topk_sampling = False
    assert np.array_equal(t2n(seq), t2n(ref["seq"]))
    assert rel_err(t2n(lps), t2n(ref["seqLogprobs"])) <= RTOL
    assert torch.equal(seq_p, seq_p2)                     # Philox stream is reproducible per (seed, offset)
    assert torch.isfinite(lps_p).all() and (lps_p <= 0).all()


def test_multi_image_batch_matches_oracle():
    d = SMALL
    sd = synth.make_state_dict(d, 5, logit_gain=8.0, lstm_gain=3.0, eos_bias=0.3)
    data = synth.make_test_inputs(d, 5, n_images=7, per_half=3, ragged=True, ragged_edges=True)
    model = make_model(d, sd, gpn_nms_thres=0.6, gpn_max_subg=3)
    with torch.no_grad():
        ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.6, max_subgraphs=3, return_att=True)
        seq, lps, score, keep, attw = model(*synth.sample_args(to_dev(data)), opt={"beam_size": 1, "return_att": 1}, mode="sample")
    assert np.array_equal(t2n(keep), t2n(ref["keep_ind"]))
    assert np.array_equal(t2n(model.last_image_of_row), t2n(ref["image_of_row"]))
    assert np.array_equal(t2n(seq), t2n(ref["seq"]))
    assert rel_err(t2n(lps), t2n(ref["seqLogprobs"])) <= RTOL
    assert rel_err(t2n(score), t2n(ref["subgraph_score"])) <= RTOL
    assert rel_err(t2n(model.last_x_obj), t2n(ref["x_obj"])) <= RTOL
    assert tuple(attw.shape) == tuple(ref["att_weights"].shape) and rel_err(t2n(attw), t2n(ref["att_weights"])) <= RTOL
    with torch.no_grad():
        refb = O.sample(sd, d, data, use_nms=True, iou_thres=0.6, max_subgraphs=3, beam_size=3)
        seqb, lpsb, _, _ = model(*synth.sample_args(to_dev(data)), opt={"beam_size": 3}, mode="sample")
    assert np.array_equal(t2n(seqb), t2n(refb["seq"]))
    assert rel_err(t2n(lpsb), t2n(refb["seqLogprobs"])) <= RTOL


def test_nms_off_returns_everything():
    d = SMALL
    sd = synth.make_state_dict(d, 6, logit_gain=8.0, lstm_gain=3.0)
    data = synth.make_test_inputs(d, 6, n_images=2, per_half=2, ragged=True)
    model = make_model(d, sd, sct=1)
    with torch.no_grad():
        ref = O.sample(sd, d, data, use_nms=False)
        seq, lps, score, keep = model(*synth.sample_args(to_dev(data)), opt={"beam_size": 1}, mode="sample")
    assert keep.dtype == torch.float32 and np.array_equal(t2n(keep), t2n(ref["keep_ind"]))  # gpn.py:97 quirk
    assert np.array_equal(t2n(seq), t2n(ref["seq"]))


def test_nms_golden_cases_on_device():
    """The reference's own subgraph_nms outputs (tests/golden/nms_cases.npz) through subgc_subgraph_nms."""
    import ctypes as C
    from subgc import _lib
    L = _lib.lib()
    g = load_golden("nms_cases")
    d = Dims()
    cd = _lib.Dims(d.v1, d.enc, d.rnn, d.att_hid, d.fc_feat, d.att_feat, d.gcn, d.low_rank, d.embed, d.obj_classes, d.pred_classes,
                   d.gcn_layers, d.gcn_residual, d.pred_emb_type, d.seq_length, d.obj_num, d.rel_num)
    ci = 0
    while f"c{ci}_score" in g.files:
        score, ind, mask = (torch.from_numpy(g[f"c{ci}_{k}"]) for k in ("score", "ind", "mask"))
        S, N = ind.shape
        M = S // 2
        lay = _lib.Layout(5, M, 5, 1)
        ind5 = ind.view(1, 2, M, N).expand(5, 2, M, N).contiguous().cuda()
        mask5 = mask.view(1, 2, M, N).expand(5, 2, M, N).contiguous().float().cuda()
        sub_len = mask.sum(1).int().cuda()
        sel = torch.empty(S, dtype=torch.int32, device="cuda")
        keep = torch.empty(S, dtype=torch.int64, device="cuda")
        stats = torch.empty(3, dtype=torch.int32, device="cuda")
        ws = torch.empty(L.subgc_nms_workspace_bytes(1, S) + 256, dtype=torch.uint8, device="cuda")
        sc = score.float().cuda()
        _lib.check(L.subgc_subgraph_nms(C.byref(cd), C.byref(lay), sc.data_ptr(), sub_len.data_ptr(), ind5.data_ptr(), mask5.data_ptr(), 1,
                                        float(g[f"c{ci}_thres"]), int(g[f"c{ci}_max"]), sel.data_ptr(), keep.data_ptr(), stats.data_ptr(),
                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "nms")
        n = int(stats.cpu()[0])
        assert np.array_equal(keep[:n].cpu().numpy(), g[f"c{ci}_keep"]), ci
        ci += 1
    assert ci == 6


@pytest.mark.parametrize("name", ["small_train", "small_train_refinit", "full_train"])
def test_golden_forward_mode_and_losses(name):
    """mode='forward' in eval (validation-loss branch, eval_utils.py:73-86) against the reference's outputs."""
    g = load_golden(name)
    d, sd, data = rebuild_train_case(g)
    model = make_model(d, sd, test_LSTM=0)
    dev = to_dev(data)
    with torch.no_grad():
        outputs, gpn_loss, score = model(*synth.forward_args(dev))
        lw = LossWrapper(model, None)
        res = lw(dev["fc_feats"], dev["att_feats"], dev["labels"], dev["masks"], dev["att_masks"], None, None, None, dev["obj_dist"], None,
                 dev["rel_ind"], None, dev["pred_dist"], dev["gpn_obj_ind"], dev["gpn_pred_ind"], dev["gpn_nrel_ind"], dev["gpn_pool_mtx"])
    check_train_outputs(outputs, g, RTOL)
    assert tuple(score.shape) == g["subgraph_score"].shape and rel_err(t2n(score), g["subgraph_score"]) <= RTOL
    assert abs(float(gpn_loss) - float(g["gpn_loss"])) <= RTOL * 10
    assert abs(float(res["lang_loss"]) - float(g["lang_loss"])) <= RTOL * 10 * max(1.0, abs(float(g["lang_loss"])))


# ---------------------------------------------------------------------------------------------------------------
# full-size shapes (BASELINE config 2: 128 images x 36 nodes x 2048-d): properties that need no oracle run
# ---------------------------------------------------------------------------------------------------------------
def test_full_size_batch_properties():
    d = Dims()
    sd = synth.make_state_dict(d, 41, logit_gain=8.0, lstm_gain=2.0, eos_bias=0.5)
    B = 128
    data = synth.make_test_inputs(d, 41, n_images=B, per_half=1, ragged=False)
    model = make_model(d, sd, gpn_nms_thres=0.75, gpn_max_subg=1)
    dev = to_dev(data)
    with torch.no_grad():
        seq, lps, score, keep = model(*synth.sample_args(dev), opt={"beam_size": 1}, mode="sample")
        # (1) two identical full sub-graphs per image: NMS keeps exactly one, the higher index on the exact score tie
        assert seq.shape == (B, d.seq_length) and torch.equal(keep.cpu(), torch.ones(B, dtype=torch.int64))
        # (2) batch independence: a sub-batch gives bit-identical rows
        sub = {k: (v[:16] if v is not None and v.shape[0] == B else (v[:80] if v is not None else None)) for k, v in dev.items()}
        seq_s, lps_s, score_s, _ = model(*synth.sample_args(sub), opt={"beam_size": 1}, mode="sample")
        assert torch.equal(seq_s, seq[:16]) and rel_err(t2n(lps_s), t2n(lps[:16])) <= RTOL
        # (3) finish-mask invariants: after the first 0 everything is 0; log-probs are <= 0
        s = seq.cpu()
        first0 = (s == 0).int().cumsum(1) > 0
        assert (s[first0] == 0).all() and (lps <= 0).all()
        # (4) a handful of rows against the oracle (CPU, seconds)
        sub4 = {k: (v[:2] if v is not None and v.shape[0] == B else (v[:10] if v is not None else None)) for k, v in data.items()}
        ref = O.sample(sd, d, sub4, use_nms=True, iou_thres=0.75, max_subgraphs=1)
    assert np.array_equal(t2n(seq[:2]), t2n(ref["seq"]))
    assert rel_err(t2n(lps[:2]), t2n(ref["seqLogprobs"])) <= RTOL
    assert rel_err(t2n(score[:2]), t2n(ref["subgraph_score"])) <= RTOL


# ---------------------------------------------------------------------------------------------------------------
# decoding options and edge cases (oracle on fresh inputs)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pen,constraint,beam", [("avg_0.0", 0, 2), ("wu_0.7", 1, 3), ("", 1, 4)])
def test_beam_options_match_oracle(pen, constraint, beam):
    d = SMALL
    sd = synth.make_state_dict(d, 17, logit_gain=8.0, lstm_gain=3.0, eos_bias=0.6)
    data = synth.make_test_inputs(d, 17, n_images=2, per_half=2, ragged=True, ragged_edges=True)
    model = make_model(d, sd, gpn_nms_thres=0.9, gpn_max_subg=4)
    with torch.no_grad():
        ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.9, max_subgraphs=4)
        fc, att, p_att, masks = ref["p_fc"], ref["p_att"], ref["pp_att"], ref["p_masks"]
        want = [O.beam_search_one(sd, d, fc[s:s + 1], att[s:s + 1], p_att[s:s + 1], masks[s:s + 1], beam, pen, constraint)
                for s in range(fc.shape[0])]
        seq, lps, _, _ = model(*synth.sample_args(to_dev(data)), opt={"beam_size": beam, "length_penalty": pen,
                                                                     "decoding_constraint": constraint}, mode="sample")
    assert len(model.done_beams) == len(want)
    for got, exp in zip(model.done_beams, want):
        assert len(got) == len(exp)
        for a, b in zip(got, exp):
            assert np.array_equal(t2n(a["seq"]), t2n(b["seq"]))
            assert rel_err(t2n(a["logps"]), t2n(b["logps"])) <= RTOL
            assert abs(a["p"] - b["p"]) <= 1e-5 * max(1.0, abs(b["p"]))
    assert np.array_equal(t2n(seq), np.stack([t2n(w[0]["seq"]) for w in want]))


def test_single_node_subgraphs_and_single_image():
    """Shortest possible inputs: one image, sub-graphs of 3..36 nodes clipped by NMS to one survivor, and a graph with 2 edges."""
    d = SMALL
    sd = synth.make_state_dict(d, 23, logit_gain=8.0, lstm_gain=3.0)
    data = synth.make_test_inputs(d, 23, n_images=1, per_half=1, ragged=True, ragged_edges=True)
    # shrink the first sub-graph to a single node
    data["gpn_obj_ind"][:, 0, 0, 1:] = d.obj_num - 1
    data["att_masks"][:, 0, 0, 1:] = 0
    data["gpn_pool_mtx"][:, 0, 0] = 0
    data["gpn_pool_mtx"][:, 0, 0, 0, 0] = 1
    model = make_model(d, sd, gpn_nms_thres=0.99, gpn_max_subg=2)
    with torch.no_grad():
        ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.99, max_subgraphs=2, return_att=True)
        seq, lps, score, keep, attw = model(*synth.sample_args(to_dev(data)), opt={"beam_size": 1, "return_att": 1}, mode="sample")
    assert np.array_equal(t2n(keep), t2n(ref["keep_ind"])) and np.array_equal(t2n(seq), t2n(ref["seq"]))
    assert rel_err(t2n(lps), t2n(ref["seqLogprobs"])) <= RTOL and rel_err(t2n(attw), t2n(ref["att_weights"])) <= RTOL
    assert rel_err(t2n(score), t2n(ref["subgraph_score"])) <= RTOL


def test_early_exit_when_every_caption_finishes():
    """eos_bias large: all rows emit token 0 at some step; the loop must stop exactly where the reference stops."""
    d = SMALL
    sd = synth.make_state_dict(d, 29, logit_gain=2.0, eos_bias=6.0)
    data = synth.make_test_inputs(d, 29, n_images=2, per_half=2, ragged=True)
    model = make_model(d, sd, gpn_nms_thres=0.9, gpn_max_subg=4)
    with torch.no_grad():
        ref = O.sample(sd, d, data, use_nms=True, iou_thres=0.9, max_subgraphs=4, return_att=True)
        seq, lps, _, _, attw = model(*synth.sample_args(to_dev(data)), opt={"beam_size": 1, "return_att": 1}, mode="sample")
    assert ref["att_weights"].shape[1] < d.seq_length + 1            # the oracle really stopped early
    assert tuple(attw.shape) == tuple(ref["att_weights"].shape)
    assert np.array_equal(t2n(seq), t2n(ref["seq"])) and rel_err(t2n(lps), t2n(ref["seqLogprobs"])) <= RTOL
    assert int(model.last_steps.item()) == ref["att_weights"].shape[1]


def test_topk_k1_equals_greedy_tokens():
    d = SMALL
    sd = synth.make_state_dict(d, 31, logit_gain=8.0, lstm_gain=3.0)
    data = to_dev(synth.make_test_inputs(d, 31, n_images=2, per_half=2, ragged=True))
    model = make_model(d, sd, gpn_nms_thres=0.9, gpn_max_subg=4)
    with torch.no_grad():
        seq_g, _, _, _ = model(*synth.sample_args(data), opt={"beam_size": 1}, mode="sample")
        model.topk_sampling, model.the_k, model.topk_temp = True, 1, 0.6
        seq_k, lps_k, _, _ = model(*synth.sample_args(data), opt={"beam_size": 1}, mode="sample")
        model.topk_sampling = False
    assert torch.equal(seq_g, seq_k) and (lps_k <= 0).all()
