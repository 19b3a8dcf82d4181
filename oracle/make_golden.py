"""Generate tests/golden/*.npz by running the REAL reference (YiwuZhong/Sub-GC, /root/reference) on CPU.

TEST INFRASTRUCTURE — runs only in the build container (the reference is not available on the GPU box);
the resulting fixtures are committed.  Usage:  python oracle/make_golden.py [--only NAME]

The reference is imported unmodified with the three shims SURVEY §8c lists:
  1. misc.utils.obj_edge_vectors -> seeded N(0,1) (data/glove.6B.300d.pt is not available offline);
  2. class-name tables passed by absolute path (bundled .npy for full dims, temp files for the tiny dims);
  3. torch.Tensor.cuda -> identity (models/CaptionModel.py:129,171 call .cuda() unconditionally in beam search).
Weights and inputs come from subgc.synth (seeded), loaded with load_state_dict, so tests can regenerate the exact
same tensors from (dims, seed) without the reference; each fixture records a fingerprint of what it was fed.
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "sub-gc_b200"))
sys.path.insert(0, REF)

from subgc import synth  # noqa: E402
from subgc.config import Dims, SMALL, make_opt  # noqa: E402

import misc.utils as ref_utils  # noqa: E402  (reference)


def _fake_glove(names, wv_type="glove.6B", wv_dir="data/", wv_dim=300):
    g = torch.Generator().manual_seed(len(names))
    return torch.randn(len(names), wv_dim, generator=g)


ref_utils.obj_edge_vectors = _fake_glove
import models as ref_models  # noqa: E402  (reference models package)
import models.AttModel as ref_att  # noqa: E402
from models.loss_wrapper import LossWrapper as RefLossWrapper  # noqa: E402

ref_att.obj_edge_vectors = _fake_glove
torch.Tensor.cuda = lambda self, *a, **k: self

GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_reference(dims: Dims, seed: int, gains=None, **opt_over):
    opt = make_opt(dims, **opt_over)
    tmp = tempfile.mkdtemp()
    if dims.obj_classes == 1599 and dims.pred_classes == 21:
        opt.obj_name_path = os.path.join(REF, "data/object_names_1600-0-20.npy")
        opt.rel_name_path = os.path.join(REF, "data/predicate_names_1600-0-20.npy")
    else:
        opt.obj_name_path = os.path.join(tmp, "obj.npy")
        opt.rel_name_path = os.path.join(tmp, "rel.npy")
        np.save(opt.obj_name_path, np.array([f"o{i}" for i in range(dims.obj_classes)]))
        np.save(opt.rel_name_path, np.array([f"p{i}" for i in range(dims.pred_classes)]))
    model = ref_models.setup(opt)
    sd = synth.make_state_dict(dims, seed, **(gains or {}))
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()
    return model, opt, sd


def np_(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def beams_to_arrays(done_beams, T):
    """[S][b] dicts -> seq [S,b,T], logps [S,b,T], p [S,b], unaug_p [S,b] (padded with -1 / nan if fewer)."""
    S = len(done_beams)
    b = max(len(x) for x in done_beams)
    seq = -np.ones((S, b, T), np.int64)
    lps = np.full((S, b, T), np.nan, np.float32)
    p = np.full((S, b), np.nan, np.float64)
    up = np.full((S, b), np.nan, np.float64)
    for s, beams in enumerate(done_beams):
        for j, bm in enumerate(beams):
            seq[s, j] = np_(bm["seq"]); lps[s, j] = np_(bm["logps"]); p[s, j] = bm["p"]; up[s, j] = bm["unaug_p"]
    return dict(beam_seq=seq, beam_logps=lps, beam_p=p, beam_unaug_p=up)


GAIN_KEYS = ("gcn_std", "logit_gain", "lstm_gain", "eos_bias")


def gains_meta(gains):
    g = dict(gcn_std=-1.0, logit_gain=1.0, lstm_gain=1.0, eos_bias=0.0)
    g.update({k: v for k, v in (gains or {}).items() if v is not None})
    return np.array([g[k] for k in GAIN_KEYS], np.float64)


def run_test_case(name, dims, seed, per_half, ragged, ragged_edges, nms, beam_sizes, gains=None, topk_seed=77,
                  store_full=True, length_penalty=""):
    thres, max_subg = nms
    model, opt, sd = build_reference(dims, seed, gains, test_LSTM=1, gpn_nms_thres=thres, gpn_max_subg=max_subg)
    data = synth.make_test_inputs(dims, seed, n_images=1, per_half=per_half, ragged=ragged, ragged_edges=ragged_edges)
    args = synth.sample_args(data)
    out = dict(meta_seed=seed, meta_per_half=per_half, meta_ragged=int(ragged), meta_ragged_edges=int(ragged_edges),
               meta_nms_thres=thres, meta_nms_max=max_subg, meta_gains=gains_meta(gains),
               meta_dims=np.array(list(dims.as_dict().values()), np.int64),
               meta_fp_weights=synth.fingerprint(sd), meta_fp_inputs=synth.fingerprint([a for a in args if a is not None]),
               meta_length_penalty=length_penalty)
    with torch.no_grad():
        # --- intermediates through the reference's own sub-modules
        x0, p0 = model.feat_fusion(data["obj_dist"], data["att_feats"], data["pred_dist"])
        N, K, L = dims.obj_num, dims.rel_num, dims.gcn
        x_obj5, x_pred5 = model.gcn_backbone(1, N, K, L, x0, data["obj_dist"], p0, data["rel_ind"])
        gl, score, g_att, g_fc, g_mask, keep = model.gpn_layer(5, N, K, L, data["gpn_obj_ind"], data["gpn_pred_ind"],
                                                              data["gpn_nrel_ind"], data["gpn_pool_mtx"], x_obj5, x_pred5,
                                                              data["fc_feats"], data["att_masks"])
        p_fc, p_att, pp_att, p_mask = model._prepare_feature(g_fc, g_att, g_mask)
        S = p_fc.shape[0]
        state = model.init_hidden(S)
        lp0, st0 = model.get_logprobs_state(torch.zeros(S, dtype=torch.long), p_fc, p_att, pp_att, p_mask, state)
        if store_full:
            out.update(x0=np_(x0), p0=np_(p0), x_obj=np_(x_obj5[0]), x_pred=np_(x_pred5[0]), p_fc=np_(p_fc), p_att=np_(p_att),
                       pp_att=np_(pp_att), step0_logprobs=np_(lp0), step0_h=np_(st0[0]), step0_c=np_(st0[1]))
        else:
            out.update(x_obj_slice=np_(x_obj5[0][:, :64]), x_obj_sum=float(x_obj5[0].double().sum()),
                       x_obj_abssum=float(x_obj5[0].double().abs().sum()),
                       p_fc_slice=np_(p_fc[:, :64]), p_att_slice=np_(p_att[:, :, :32]), pp_att_slice=np_(pp_att[:, :, :32]),
                       step0_top5_val=np_(lp0.topk(5, 1)[0]), step0_top5_idx=np_(lp0.topk(5, 1)[1]),
                       step0_h_slice=np_(st0[0][:, :, :64]))
        out.update(gpn_loss=float(gl), gpn_score=np_(score), keep_ind=np_(keep), p_mask=np_(p_mask), g_fc_slice=np_(g_fc[:, :64]))
        # --- greedy with attention weights
        seq, lps, sc, kp, attw = model(*args, opt={"beam_size": 1, "return_att": 1}, mode="sample")
        out.update(greedy_seq=np_(seq), greedy_logprobs=np_(lps), greedy_score=np_(sc), greedy_keep=np_(kp), greedy_att_weights=np_(attw))
        # --- top-k sampling under a fixed torch seed (oracle uses the same torch calls => same stream)
        model.topk_sampling, model.topk_temp, model.the_k = True, 0.6, 3
        torch.manual_seed(topk_seed)
        seq, lps, sc, kp = model(*args, opt={"beam_size": 1}, mode="sample")
        out.update(topk_seq=np_(seq), topk_logprobs=np_(lps), meta_topk_seed=topk_seed)
        model.topk_sampling = False
        # --- beam search
        for b in beam_sizes:
            seq, lps, sc, kp = model(*args, opt={"beam_size": b, "length_penalty": length_penalty}, mode="sample")
            out.update({f"beam{b}_seq": np_(seq), f"beam{b}_logprobs": np_(lps)})
            out.update({f"beam{b}_{k}": v for k, v in beams_to_arrays(model.done_beams, dims.seq_length).items()})
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "S_kept", S, "greedy first tokens", out["greedy_seq"][:, :4].tolist())


def grad_summary(model):
    rows = {}
    for n, p in model.named_parameters():
        if p.grad is None:
            rows["grad_none__" + n] = np.zeros(0, np.float32)
        else:
            g = p.grad.detach().double().reshape(-1)
            head = g[:24].float().numpy()
            rows["grad__" + n] = np.concatenate([[float(g.sum()), float(g.abs().sum()), float((g * g).sum())], head]).astype(np.float64)
    return rows


def run_train_case(name, dims, seed, n_images, gpn_batch, gains=None, store_full=True):
    model, opt, sd = build_reference(dims, seed, gains)
    data = synth.make_train_inputs(dims, seed, n_images=n_images, gpn_batch=gpn_batch)
    lw = RefLossWrapper(model, opt)
    lw.eval()  # dropout off: RNG streams cannot be matched by a CUDA kernel (SURVEY §7 hard part 5)
    out = dict(meta_seed=seed, meta_n_images=n_images, meta_gpn_batch=gpn_batch, meta_gains=gains_meta(gains),
               meta_dims=np.array(list(dims.as_dict().values()), np.int64), meta_fp_weights=synth.fingerprint(sd),
               meta_fp_inputs=synth.fingerprint([v for v in synth.forward_args(data) if v is not None]))
    with torch.no_grad():
        outputs, gl, score = model(*synth.forward_args(data))
    if store_full:
        out.update(outputs=np_(outputs))
    else:   # full dimensions: [rows, 17, 9488] log-probs are too large for a fixture; a slice, checksums and the row-wise arg-max instead
        out.update(outputs_slice=np_(outputs[:, :, :48]), outputs_sum=float(outputs.double().sum()), outputs_abssum=float(outputs.double().abs().sum()),
                   outputs_argmax=np_(outputs.max(2)[1]))
    out.update(gpn_loss=float(gl), subgraph_score=np_(score))
    model.zero_grad()
    res = lw(data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None,
             data["obj_dist"], None, data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"],
             data["gpn_nrel_ind"], data["gpn_pool_mtx"])
    (res["lang_loss"] + res["gpn_loss"]).backward()
    out.update(lang_loss=float(res["lang_loss"]), gpn_loss_lw=float(res["gpn_loss"]))
    out.update(grad_summary(model))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name, "lang_loss", out["lang_loss"], "gpn_loss", out["gpn_loss"])


def run_nms_cases(name):
    """models/lib/gpn.py:108-150 called directly on random scores / node sets."""
    from models.lib.gpn import gpn_layer
    out = {}
    for ci, (S, thres, mx, seed) in enumerate([(80, 0.75, 10, 1), (80, 0.55, 1000, 2), (80, 0.3, 5, 3), (12, 0.5, 3, 4),
                                               (200, 0.6, 1000, 5), (2, 0.75, 1, 6)]):
        g = torch.Generator().manual_seed(seed)
        layer = gpn_layer(GCN_dim=8, hid_dim=4, test_LSTM=True, use_nms=True, iou_thres=thres, max_subgraphs=mx)
        N = 37
        ind = torch.full((S, N), 36, dtype=torch.int64)
        mask = torch.zeros(S, N)
        for s in range(S):
            if ci == 5:
                cnt, ids = 36, torch.arange(36)
            else:
                cnt = int(torch.randint(2, 14, (1,), generator=g))
                ids = torch.randperm(16, generator=g)[:cnt].sort().values  # small universe => many overlaps
            ind[s, :cnt] = ids
            mask[s, :cnt] = 1
        score = torch.rand(S, generator=g)
        if ci == 5:
            score[:] = 0.625  # exact tie between two identical full sub-graphs (the bench workload's situation)
        att_masks = mask.view(1, 2, S // 2, N).expand(5, 2, S // 2, N).contiguous()
        keep = layer.subgraph_nms(score, ind, att_masks)
        out.update({f"c{ci}_score": np_(score), f"c{ci}_ind": np_(ind), f"c{ci}_mask": np_(mask), f"c{ci}_thres": thres,
                    f"c{ci}_max": mx, f"c{ci}_keep": np_(keep)})
        print(name, ci, "kept", len(keep), "of", S)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)


G_DIVERSE = dict(logit_gain=8.0, lstm_gain=3.0)
G_EOS = dict(logit_gain=8.0, lstm_gain=3.0, eos_bias=0.3)

CASES = {
    "small_test_ragged": lambda: run_test_case("small_test_ragged", SMALL, 11, per_half=6, ragged=True, ragged_edges=True,
                                               nms=(0.55, 1000), beam_sizes=(2, 3), gains=G_DIVERSE),
    "small_test_nms": lambda: run_test_case("small_test_nms", SMALL, 13, per_half=8, ragged=True, ragged_edges=False,
                                            nms=(0.3, 4), beam_sizes=(3,), length_penalty="wu_0.5", gains=G_EOS),
    "small_test_full": lambda: run_test_case("small_test_full", SMALL, 12, per_half=1, ragged=False, ragged_edges=False,
                                             nms=(0.75, 1), beam_sizes=(2,), gains=G_DIVERSE),
    "small_train": lambda: run_train_case("small_train", SMALL, 21, n_images=3, gpn_batch=2, gains=G_DIVERSE),
    "small_train_refinit": lambda: run_train_case("small_train_refinit", SMALL, 22, n_images=2, gpn_batch=3,
                                                  gains=dict(gcn_std=1e-3)),
    "full_test": lambda: run_test_case("full_test", Dims(), 31, per_half=4, ragged=True, ragged_edges=True, nms=(0.75, 10),
                                       beam_sizes=(2,), store_full=False),
    "full_test_peaked": lambda: run_test_case("full_test_peaked", Dims(), 32, per_half=3, ragged=True, ragged_edges=False,
                                              nms=(0.55, 1000), beam_sizes=(5,), store_full=False,
                                              gains=dict(logit_gain=8.0, lstm_gain=2.0, eos_bias=0.5)),
    "full_train": lambda: run_train_case("full_train", Dims(), 41, n_images=2, gpn_batch=2, gains=dict(logit_gain=4.0), store_full=False),
    "nms_cases": lambda: run_nms_cases("nms_cases"),
}




# ------------------------------------------------------------------------------------------------------------------------------
# Full-GC (train.sh:27-36: use_gpn=0, noun_fuse=0, pred_emb_type=2, gcn_layers=4, gcn_residual=1, gcn_bn=1): the reference's own
# _sample on its no-sGPN branch (AttModel.py:261-271), eval mode (BatchNorm running statistics), one image
# ------------------------------------------------------------------------------------------------------------------------------
FULLGC_OPT = dict(use_gpn=0, noun_fuse=0, pred_emb_type=2, gcn_layers=4, gcn_residual=1, gcn_bn=1, test_LSTM=1)


def run_fullgc_case(name, dims, seed, topk=False):
    from subgc.fullgc import make_fullgc_state_dict
    over = dict(FULLGC_OPT)
    if topk:
        over.update(use_topk_sampling=1, topk_temp=0.6, the_k=3)
    opt = make_opt(dims, **over)
    tmp = tempfile.mkdtemp()
    opt.obj_name_path, opt.rel_name_path = os.path.join(tmp, "obj.npy"), os.path.join(tmp, "rel.npy")
    np.save(opt.obj_name_path, np.array([f"o{i}" for i in range(dims.obj_classes)]))
    np.save(opt.rel_name_path, np.array([f"p{i}" for i in range(dims.pred_classes)]))
    model = ref_models.setup(opt)
    sd = make_fullgc_state_dict(model, seed)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    model.eval()
    data = synth.make_test_inputs(dims, seed, n_images=1, per_half=1, ragged=False, ragged_edges=True)
    args = synth.sample_args(data)
    out = dict(meta_seed=seed, meta_dims=np.array(list(dims.as_dict().values()), np.int64), meta_topk=int(topk),
               meta_state_keys=np.array(list(sd.keys())), meta_fp_weights=synth.fingerprint(sd),
               meta_fp_inputs=synth.fingerprint([a for a in args if a is not None]))
    with torch.no_grad():
        x0, p0 = model.feat_fusion(data["obj_dist"], data["att_feats"], data["pred_dist"])
        N, K, L = dims.obj_num, dims.rel_num, dims.gcn
        x_obj5, _ = model.gcn_backbone(1, N, K, L, x0, data["obj_dist"], p0, data["rel_ind"])
        read_out = torch.mean(x_obj5[0:1], 1)
        g_fc = model.read_out_proj(read_out)
        out.update(x0=np_(x0), p0=np_(p0), x_obj=np_(x_obj5[0]), g_fc=np_(g_fc))
        if topk:   # the reference draws from torch's generator: record the uniforms of an equivalent inverse-CDF draw is not possible;
            torch.manual_seed(1234)   # keep the reference's own sample + its log-probs for a statistical check only
        margs = [a.clone() if torch.is_tensor(a) else a for a in args]   # the branch mutates att_masks in place (AttModel.py:269)
        seq, lps, score, keep = model(*margs, opt={"beam_size": 1}, mode="sample")
        out.update(seq=np_(seq), lps=np_(lps), score=np_(score), keep=np_(keep))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print("wrote", name, {k: np.asarray(v).shape for k, v in out.items() if not k.startswith("meta")})


FULLGC_SMALL = Dims(vocab=61, enc=24, rnn=40, att_hid=16, fc_feat=48, att_feat=48, gcn=24, low_rank=512, embed=12, obj_classes=23, pred_classes=7,
                    gcn_layers=4, gcn_residual=1, pred_emb_type=2, seq_length=8, obj_num=37, rel_num=65)
CASES["fullgc_small"] = lambda: run_fullgc_case("fullgc_small", FULLGC_SMALL, 41)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    for k, fn in CASES.items():
        if a.only is None or a.only == k:
            fn()
