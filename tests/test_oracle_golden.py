"""Pin the oracle (oracle/subgc_oracle.py) to outputs of the REAL reference stored in tests/golden/.

CPU-only.  The fixtures were produced by oracle/make_golden.py importing /root/reference; here the oracle is
run on the same regenerated weights/inputs and must reproduce them.  Tolerance: the oracle issues the same torch
CPU ops as the reference, so integer outputs are exact and float outputs agree to 1e-6 relative (bit-equal in
this container; the slack only allows for a different BLAS thread split on the GPU box's host).
"""
import numpy as np
import pytest
import torch

import subgc_oracle as O
from helpers import check_train_outputs, beam_sizes_in, load_golden, rebuild_test_case, rebuild_train_case, rel_err, t2n

TOL = 1e-6
TEST_CASES = ["small_test_ragged", "small_test_nms", "small_test_full", "full_test", "full_test_peaked"]


@pytest.fixture(scope="module", params=TEST_CASES)
def case(request):
    g = load_golden(request.param)
    d, sd, data, nms = rebuild_test_case(g)
    torch.manual_seed(0)
    with torch.no_grad():
        r = O.sample(sd, d, data, use_nms=True, return_att=True, **nms)
    return g, d, sd, data, nms, r


def test_encoder_and_sgpn(case):
    g, d, sd, data, nms, r = case
    if "x_obj" in g.files:
        assert rel_err(t2n(r["x_obj"][0]), g["x_obj"]) <= TOL
        x0, p0 = O.fuse_features(sd, d, data["att_feats"], data["obj_dist"], data["pred_dist"])
        assert rel_err(t2n(x0), g["x0"]) <= TOL and rel_err(t2n(p0), g["p0"]) <= TOL
        _, x_pred = O.gcn_encode(sd, d, x0, p0, data["rel_ind"])
        assert rel_err(t2n(x_pred[0]), g["x_pred"]) <= TOL
        assert rel_err(t2n(r["p_fc"]), g["p_fc"]) <= TOL
        assert rel_err(t2n(r["p_att"]), g["p_att"]) <= TOL
        assert rel_err(t2n(r["pp_att"]), g["pp_att"]) <= TOL
    else:
        assert rel_err(t2n(r["x_obj"][0][:, :64]), g["x_obj_slice"]) <= TOL
        assert abs(float(r["x_obj"][0].double().sum()) - float(g["x_obj_sum"])) <= 1e-6 * float(g["x_obj_abssum"])
        assert rel_err(t2n(r["p_att"][:, :, :32]), g["p_att_slice"]) <= TOL
        assert rel_err(t2n(r["pp_att"][:, :, :32]), g["pp_att_slice"]) <= TOL
    assert np.array_equal(t2n(r["keep_ind"]), g["keep_ind"])
    assert rel_err(t2n(r["subgraph_score"]), g["greedy_score"]) <= TOL
    assert abs(float(r["gpn_loss"]) - float(g["gpn_loss"])) <= TOL
    assert np.array_equal(t2n(r["p_masks"]), g["p_mask"])


def test_greedy(case):
    g, d, sd, data, nms, r = case
    assert np.array_equal(t2n(r["seq"]), g["greedy_seq"])
    assert rel_err(t2n(r["seqLogprobs"]), g["greedy_logprobs"]) <= TOL
    assert r["att_weights"].shape == g["greedy_att_weights"].shape
    assert rel_err(t2n(r["att_weights"]), g["greedy_att_weights"]) <= TOL


def test_topk_same_rng_stream(case):
    g, d, sd, data, nms, _ = case
    torch.manual_seed(int(g["meta_topk_seed"]))
    with torch.no_grad():
        r = O.sample(sd, d, data, use_nms=True, topk=True, temp=0.6, k=3, **nms)
    assert np.array_equal(t2n(r["seq"]), g["topk_seq"])
    assert rel_err(t2n(r["seqLogprobs"]), g["topk_logprobs"]) <= TOL


def test_beam(case):
    g, d, sd, data, nms, _ = case
    for b in beam_sizes_in(g):
        with torch.no_grad():
            r = O.sample(sd, d, data, use_nms=True, beam_size=b, length_penalty=str(g["meta_length_penalty"]), **nms)
        assert np.array_equal(t2n(r["seq"]), g[f"beam{b}_seq"])
        assert rel_err(t2n(r["seqLogprobs"]), g[f"beam{b}_logprobs"]) <= TOL
        for s, beams in enumerate(r["done_beams"]):
            for j, bm in enumerate(beams):
                assert np.array_equal(t2n(bm["seq"]), g[f"beam{b}_beam_seq"][s, j])
                assert rel_err(t2n(bm["logps"]), g[f"beam{b}_beam_logps"][s, j]) <= TOL
                assert abs(bm["p"] - g[f"beam{b}_beam_p"][s, j]) <= 1e-5 * max(1.0, abs(g[f"beam{b}_beam_p"][s, j]))


@pytest.mark.parametrize("name", ["small_train", "small_train_refinit", "full_train"])
def test_train_forward_losses_and_grads(name):
    g = load_golden(name)
    d, sd, data = rebuild_train_case(g)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    r = O.loss_wrapper(sd, d, data)
    check_train_outputs(r["outputs"], g, TOL)
    assert rel_err(t2n(r["subgraph_score"]), g["subgraph_score"]) <= TOL
    assert abs(float(r["lang_loss"]) - float(g["lang_loss"])) <= TOL * 10
    assert abs(float(r["gpn_loss"]) - float(g["gpn_loss"])) <= TOL
    (r["lang_loss"] + r["gpn_loss"]).backward()
    for k in g.files:
        if k.startswith("grad_none__"):
            n = k[len("grad_none__"):]
            assert sd[n].grad is None or float(sd[n].grad.abs().sum()) == 0.0, n
        elif k.startswith("grad__"):
            n = k[len("grad__"):]
            gr = sd[n].grad.detach().double().reshape(-1)
            ref = g[k]
            scale = max(ref[1] / max(gr.numel(), 1), 1e-12)
            assert abs(float(gr.sum()) - ref[0]) <= 2e-5 * max(ref[1], 1e-12), n
            assert abs(float(gr.abs().sum()) - ref[1]) <= 2e-5 * max(ref[1], 1e-12), n
            head = gr[:24].numpy()
            assert np.abs(head - ref[3:3 + len(head)]).max() <= 1e-4 * max(np.abs(ref[3:]).max(), scale), n


def test_nms_cases():
    g = load_golden("nms_cases")
    ci = 0
    while f"c{ci}_score" in g.files:
        keep = O.node_set_nms(g[f"c{ci}_score"], g[f"c{ci}_ind"], g[f"c{ci}_mask"], float(g[f"c{ci}_thres"]), int(g[f"c{ci}_max"]))
        assert np.array_equal(keep, g[f"c{ci}_keep"]), ci
        ci += 1
    assert ci == 6


def test_batched_images_reduce_to_single_image_calls():
    """The multi-image extension must equal per-image reference-style calls (up to the shared clip length)."""
    from subgc import synth
    from subgc.config import SMALL
    d = SMALL
    sd = synth.make_state_dict(d, 5, logit_gain=8.0, lstm_gain=3.0)
    data = synth.make_test_inputs(d, 5, n_images=3, per_half=2, ragged=True, ragged_edges=True)
    with torch.no_grad():
        rb = O.sample(sd, d, data, use_nms=True, iou_thres=0.6, max_subgraphs=3)
        row = 0
        for i in range(3):
            one = {k: (v[i:i + 1] if v is not None and v.shape[0] == 3 else (v[5 * i:5 * i + 5] if v is not None else None))
                   for k, v in data.items()}
            r1 = O.sample(sd, d, one, use_nms=True, iou_thres=0.6, max_subgraphs=3)
            n = r1["seq"].shape[0]
            assert np.array_equal(t2n(rb["keep_ind"][row:row + n]), t2n(r1["keep_ind"]))
            assert np.array_equal(t2n(rb["seq"][row:row + n]), t2n(r1["seq"]))
            assert rel_err(t2n(rb["seqLogprobs"][row:row + n]), t2n(r1["seqLogprobs"])) <= 1e-5
            row += n
        assert row == rb["seq"].shape[0]
