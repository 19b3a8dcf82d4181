"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped path.

A CPU restatement (plain PyTorch fp32 ops, functional style over a `state_dict`) of the Sub-GC hot path of
YiwuZhong/Sub-GC: feature fusion -> GCN -> sGPN (+ node-set NMS) -> feature preparation -> top-down
attention-LSTM decoder with greedy / top-k / beam loops, and the teacher-forced forward + losses.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import it.

Parity status: PINNED against the real reference.  `oracle/make_golden.py` imports the reference from
/root/reference (build container only), runs it on `subgc.synth` weights/inputs and stores its outputs under
tests/golden/; `tests/test_oracle_golden.py` checks every function below against those files.  (The reference
ships no tests or golden vectors of its own — SURVEY §4.)

Every function cites the reference lines it restates (paths relative to the reference root).  The one
extension over the reference is multi-image inference: the reference asserts a single image per call
(models/lib/gpn.py:84); here `B` images are encoded, scored and NMS-ed per image and decoded as one batch, which
is the reference's behaviour for B = 1.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------------
# small helpers
# ----------------------------------------------------------------------------------------------------------
def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _lstm_cell(sd, name, x, h, c):
    """torch.nn.LSTMCell arithmetic (gate order i, f, g, o; two biases) — models/AttModel.py:397-398,413,423."""
    gates = F.linear(x, sd[name + ".weight_ih"], sd[name + ".bias_ih"]) + \
        F.linear(h, sd[name + ".weight_hh"], sd[name + ".bias_hh"])
    i, f, g, o = gates.chunk(4, 1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


# ----------------------------------------------------------------------------------------------------------
# encoder: fusion + GCN
# ----------------------------------------------------------------------------------------------------------
def fuse_features(sd, dims, att_feats, obj_dist, pred_dist):
    """models/AttModel.py:370-387 (noun_fuse=1; pred_emb_type 1 or 2)."""
    B, N, _ = att_feats.shape
    cls = obj_dist.reshape(-1, dims.obj_classes)[:, 1:].max(1)[1] + 1
    obj_emb = _lin(sd, "obj_emb_proj", sd["sg_obj_embed.weight"][cls]).view(B, N, dims.gcn)
    x0 = torch.relu(_lin(sd, "obj_v_proj", att_feats) + obj_emb)
    pd = pred_dist.reshape(-1, dims.pred_classes)
    pcls = (pd[:, 1:].max(1)[1] + 1) if dims.pred_emb_type == 1 else pd.max(1)[1]
    p0 = _lin(sd, "pred_emb_prj", sd["sg_pred_embed.weight"][pcls]).view(B, pred_dist.shape[1], dims.gcn)
    return x0, p0


def dense_adjacency(rel_ind, N):
    """models/lib/gcn_backbone.py:55-67: 0/1 maps [B,N,K] for the subject and the object endpoint."""
    B, K, _ = rel_ind.shape
    subj = torch.zeros(B, N, K)
    obj = torch.zeros(B, N, K)
    ones = torch.ones(B, 1, K)
    subj.scatter_(1, rel_ind[:, :, 0].reshape(B, 1, K), ones)
    obj.scatter_(1, rel_ind[:, :, 1].reshape(B, 1, K), ones)
    return subj, obj


def _collect(sd, prefix, source, adj):
    """models/lib/graph_conv_unit.py:28-36 without BN."""
    msg = _lin(sd, prefix + "fc_rgt", _lin(sd, prefix + "fc_lft", source))
    agg = torch.bmm(adj, msg)
    return torch.relu(agg / (adj.sum(2).view(agg.size(0), agg.size(1), 1) + 1e-7))


def gcn_encode(sd, dims, x0, p0, rel_ind):
    """models/lib/gcn_backbone.py:29-53 + models/lib/graph_conv.py:15-34.  Returns the un-replicated
    (x_obj [B,N,L], x_pred [B,K,L]); the reference then tiles both ×5 on the batch dim (gcn_backbone.py:50-51)."""
    x, p = x0, p0
    x_res, p_res = x0, p0
    if dims.gcn_layers:
        a_s, a_o = dense_adjacency(rel_ind, x0.shape[1])
        for l in range(dims.gcn_layers):
            pre = f"gcn_backbone.gcn.{l}.gcn_collect.collect_units."
            x_new = (_collect(sd, pre + "0.", p, a_s) + _collect(sd, pre + "1.", p, a_o)) / 2
            p_new = (_collect(sd, pre + "2.", x, a_s.transpose(1, 2)) + _collect(sd, pre + "3.", x, a_o.transpose(1, 2))) / 2
            x, p = x_new, p_new
            if (l + 1) % dims.gcn_residual == 0:
                x = x + x_res
                x_res = x
                p = p + p_res
                p_res = p
    return x, p


def encode(sd, dims, att_feats, obj_dist, pred_dist, rel_ind):
    x0, p0 = fuse_features(sd, dims, att_feats, obj_dist, pred_dist)
    return gcn_encode(sd, dims, x0, p0, rel_ind)


# ----------------------------------------------------------------------------------------------------------
# sGPN
# ----------------------------------------------------------------------------------------------------------
def _pooled_readout(x_rows, obj_ind, pool_mtx, masks):
    """models/lib/gpn.py:152-185 for a flat list of sub-graphs.
    x_rows [S,N,L] = node features of the image each sub-graph belongs to; obj_ind [S,N]; pool_mtx [S,N,N];
    masks [S,N].  Returns read_out [S,2L] = max ‖ mean (max taken over the zero-padded rows as well)."""
    S, N, L = x_rows.shape
    feats = torch.gather(x_rows, 1, obj_ind.view(S, N, 1).expand(S, N, L))
    clean = torch.bmm(pool_mtx, feats)
    return torch.cat((clean.max(1)[0], clean.sum(1) / masks.sum(-1).view(-1, 1)), -1)


def _score(sd, read_out, dropout_mask=None):
    """models/lib/gpn.py:25-31,54-55: sigmoid(W2 . drop(relu(W1 . r)))."""
    hid = torch.relu(_lin(sd, "gpn_layer.gpn_fc.0", read_out))
    if dropout_mask is not None:
        hid = hid * dropout_mask
    return torch.sigmoid(_lin(sd, "gpn_layer.gpn_fc.3", hid))


def _read_out_proj(sd, r):
    return _lin(sd, "gpn_layer.read_out_proj.1", _lin(sd, "gpn_layer.read_out_proj.0", r))


def sgpn_train(sd, dims, x_obj, gpn_obj_ind, att_masks, gpn_pool_mtx, seq_per_img=5, dropout_mask=None):
    """Training/validation branch, models/lib/gpn.py:41-81.
    x_obj [B,N,L] (un-replicated); gpn_obj_ind/att_masks [5B,2,G,N]; gpn_pool_mtx [5B,2,G,N,N].
    Returns gpn_loss, subgraph_score [2*5B*G,1] (positives first), att_feats [5B,N,L], fc_feats [5B,2L],
    att_masks [5B,N], read_out [2*5B*G,2L]."""
    b, _, G, N = gpn_obj_ind.shape
    L = x_obj.shape[-1]
    img = torch.arange(b) // seq_per_img
    # flat order = (half, sentence, g): positives then negatives (gpn.py:157-170, :174)
    ind = gpn_obj_ind.transpose(0, 1).reshape(2 * b * G, N)
    msk = att_masks.transpose(0, 1).reshape(2 * b * G, N)
    pool = gpn_pool_mtx.transpose(0, 1).reshape(2 * b * G, N, N)
    rows = x_obj[img.view(1, b, 1).expand(2, b, G).reshape(-1)]
    read_out = _pooled_readout(rows, ind, pool, msk)
    score = _score(sd, read_out, dropout_mask)
    target = torch.cat((torch.ones(b * G, 1), torch.zeros(b * G, 1)), 0)
    loss = F.binary_cross_entropy(score, target)
    pos_score = score.view(2, b, G)[0]
    pick = pos_score.argmax(-1)
    sent = torch.arange(b)
    sel_ind = gpn_obj_ind[sent, 0, pick]                       # [b,N]
    att = torch.gather(x_obj[img], 1, sel_ind.view(b, N, 1).expand(b, N, L))
    sel_mask = att_masks[sent, 0, pick]
    sel_read = read_out.view(2, b, G, -1)[0][sent, pick].detach()
    fc = _read_out_proj(sd, sel_read)
    return loss, score, att, fc, sel_mask, read_out


def node_set_nms(scores, obj_ind, masks, iou_thres, max_subgraphs):
    """models/lib/gpn.py:108-150.  scores [S] (numpy f32), obj_ind [S,N] ints, masks [S,N].
    Ties in the score sort are taken in descending index order (= reversed stable ascending argsort, which is
    what numpy's default argsort gives for the short arrays where it falls back to insertion sort)."""
    order = np.argsort(scores, kind="stable")[::-1]
    sets = [frozenset(int(v) for v, m in zip(obj_ind[i], masks[i]) if m != 0) for i in order]
    alive = np.ones(len(order), dtype=bool)
    for i in range(len(order)):
        if not alive[i]:
            continue
        for j in range(i + 1, len(order)):
            a, b = sets[i], sets[j]
            if len(a) == 0 or len(b) == 0:
                a = frozenset(range(0))  # gpn.py:145-146: "this" becomes arange(0) when either side is empty
            union = len(a | b)
            iou = (len(a & b) / float(union)) if union else 0.0
            if iou > iou_thres:
                alive[j] = False
    kept_sorted = order[alive][:max_subgraphs]
    return np.sort(kept_sorted).astype(np.int64)


def sgpn_test(sd, dims, x_obj, gpn_obj_ind, att_masks, gpn_pool_mtx, use_nms=True, iou_thres=0.75, max_subgraphs=1,
              seq_per_img=5):
    """Inference branch, models/lib/gpn.py:83-106, applied per image.
    Inputs as in sgpn_train with M sub-graphs per half; only sentence copy 0 of each image is decoded
    (gpn.py:86,89-91,94).  Returns dict with gpn_loss (BCE over all 2*5B*M scored sub-graphs, as the reference
    computes it even at test time), score [S_kept], att_feats [S_kept,N,L], fc_feats [S_kept,2L],
    att_masks [S_kept,N], keep_ind (per-image indices, concatenated; float32 arange when NMS is off as in
    gpn.py:97), image_of_row [S_kept], all_scores [B,2M]."""
    b, _, M, N = gpn_obj_ind.shape
    B = b // seq_per_img
    L = x_obj.shape[-1]
    img = torch.arange(b) // seq_per_img
    ind = gpn_obj_ind.transpose(0, 1).reshape(2 * b * M, N)
    msk = att_masks.transpose(0, 1).reshape(2 * b * M, N)
    pool = gpn_pool_mtx.transpose(0, 1).reshape(2 * b * M, N, N)
    rows = x_obj[img.view(1, b, 1).expand(2, b, M).reshape(-1)]
    read_out = _pooled_readout(rows, ind, pool, msk)
    score = _score(sd, read_out)
    target = torch.cat((torch.ones(b * M, 1), torch.zeros(b * M, 1)), 0)
    loss = F.binary_cross_entropy(score, target)
    score_c0 = score.view(2, B, seq_per_img, M)[:, :, 0].transpose(0, 1).reshape(B, 2 * M)       # (half, m) order
    read_c0 = read_out.view(2, B, seq_per_img, M, -1)[:, :, 0].transpose(0, 1).reshape(B, 2 * M, -1)
    ind_c0 = gpn_obj_ind.view(B, seq_per_img, 2, M, N)[:, 0].reshape(B, 2 * M, N)
    msk_c0 = att_masks.view(B, seq_per_img, 2, M, N)[:, 0].reshape(B, 2 * M, N)
    out = dict(gpn_loss=loss, all_scores=score_c0)
    sc, att, fc, mk, keep, owner = [], [], [], [], [], []
    for i in range(B):
        a_i = torch.gather(x_obj[i].unsqueeze(0).expand(2 * M, N, L), 1, ind_c0[i].view(2 * M, N, 1).expand(2 * M, N, L))
        f_i = _read_out_proj(sd, read_c0[i])
        if use_nms:
            k = torch.from_numpy(node_set_nms(score_c0[i].detach().numpy(), ind_c0[i].numpy(), msk_c0[i].numpy(),
                                              iou_thres, max_subgraphs))
            sc.append(score_c0[i][k]); att.append(a_i[k]); fc.append(f_i[k]); mk.append(msk_c0[i][k]); keep.append(k)
            owner.append(torch.full((len(k),), i, dtype=torch.int64))
        else:
            sc.append(score_c0[i]); att.append(a_i); fc.append(f_i); mk.append(msk_c0[i])
            keep.append(torch.arange(2 * M).float())
            owner.append(torch.full((2 * M,), i, dtype=torch.int64))
    out.update(score=torch.cat(sc), att_feats=torch.cat(att), fc_feats=torch.cat(fc), att_masks=torch.cat(mk),
               keep_ind=torch.cat(keep), image_of_row=torch.cat(owner))
    return out


# ----------------------------------------------------------------------------------------------------------
# decoder
# ----------------------------------------------------------------------------------------------------------
def prepare_features(sd, dims, fc_feats, att_feats, att_masks, drop=None):
    """models/AttModel.py:348-368 with pack_wrapper (:16-36): att_embed only on valid rows (pads exactly 0),
    ctx2att on every row up to the longest sub-graph.  `drop` optionally carries the three dropout masks
    (fc, att, —) for training-mode checks."""
    max_len = int(att_masks.long().sum(1).max())
    att_feats = att_feats[:, :max_len].contiguous()
    att_masks = att_masks[:, :max_len].contiguous()
    fc = torch.relu(_lin(sd, "fc_embed.2", torch.relu(_lin(sd, "fc_embed.0", fc_feats))))
    if drop is not None:
        fc = fc * drop["fc"]
    valid = (torch.arange(max_len).view(1, -1) < att_masks.long().sum(1).view(-1, 1)).unsqueeze(-1)
    att = torch.relu(_lin(sd, "att_embed.0", att_feats))
    if drop is not None:
        att = att * drop["att"][:, :max_len]
    att = att * valid
    p_att = _lin(sd, "ctx2att", att)
    return fc, att, p_att, att_masks


def attention(sd, h, att, p_att, masks):
    """models/AttModel.py:445-471: softmax over all rows, then mask, then renormalise."""
    att_h = _lin(sd, "core.attention.h2att", h)
    dot = torch.tanh(p_att + att_h.unsqueeze(1))
    e = F.linear(dot, sd["core.attention.alpha_net.weight"], sd["core.attention.alpha_net.bias"]).squeeze(-1)
    w = F.softmax(e, dim=1)
    if masks is not None:
        w = w * masks.float()
        w = w / w.sum(1, keepdim=True)
    return torch.bmm(w.unsqueeze(1), att).squeeze(1), w


def decoder_step(sd, it, fc, att, p_att, masks, state, drop_x=None, drop_h=None):
    """models/AttModel.py:328-341 + TopDownCore.forward :400-431.  state = (h[2,S,H], c[2,S,H]).
    Returns (logprobs [S,V+1], new state, attention weights [S,len])."""
    x = torch.relu(sd["embed.0.weight"][it])
    if drop_x is not None:
        x = x * drop_x
    h, c = state
    h_att, c_att = _lstm_cell(sd, "core.att_lstm", torch.cat([h[1], fc, x], 1), h[0], c[0])
    ctx, w = attention(sd, h_att, att, p_att, masks)
    h_lang, c_lang = _lstm_cell(sd, "core.lang_lstm", torch.cat([ctx, h_att], 1), h[1], c[1])
    out = h_lang if drop_h is None else h_lang * drop_h
    logp = F.log_softmax(_lin(sd, "logit", out), dim=1)
    return logp, (torch.stack([h_att, h_lang]), torch.stack([c_att, c_lang])), w


def init_state(dims, rows):
    return torch.zeros(2, rows, dims.rnn), torch.zeros(2, rows, dims.rnn)


def decode_greedy_or_topk(sd, dims, fc, att, p_att, masks, topk=False, temp=0.6, k=3, return_att=False,
                          uniforms=None):
    """models/AttModel.py:278-326.  `uniforms` [T,S] switches top-k sampling from torch's RNG stream
    (Categorical.sample, bit-compatible with the reference under the same manual_seed) to inverse-CDF sampling
    over the k kept tokens in descending-probability order (what the CUDA sampler does with its Philox draws)."""
    S, T = fc.shape[0], dims.seq_length
    state = init_state(dims, S)
    seq = torch.zeros(S, T, dtype=torch.int64)
    seq_lp = torch.zeros(S, T)
    weights = []
    it = torch.zeros(S, dtype=torch.int64)
    unfinished = None
    for t in range(T + 1):
        logp, state, w = decoder_step(sd, it, fc, att, p_att, masks, state)
        weights.append(w)
        if t == T:
            break
        if topk:
            q = F.log_softmax(logp / float(temp), dim=1)
            top, idx = torch.topk(q, k, dim=1)
            if uniforms is None:
                kept = torch.full_like(q, float("-inf")).scatter(1, idx, top)
                it = torch.distributions.Categorical(logits=kept).sample()
                lp = kept.gather(1, it.unsqueeze(1)).view(-1)
            else:
                pr = torch.softmax(top, dim=1)
                cdf = pr.cumsum(1)
                pos = (uniforms[t].view(-1, 1) >= cdf).sum(1).clamp(max=k - 1)
                it = idx.gather(1, pos.unsqueeze(1)).view(-1)
                lp = top.gather(1, pos.unsqueeze(1)).view(-1)
        else:
            lp, it = torch.max(logp, 1)
        unfinished = (it > 0) if t == 0 else unfinished * (it > 0)
        it = it * unfinished.type_as(it)
        seq[:, t] = it
        seq_lp[:, t] = lp
        if unfinished.sum() == 0:
            break
    if return_att:
        return seq, seq_lp, torch.stack(weights, 1)
    return seq, seq_lp


def _length_penalty(cfg):
    """misc/utils.py:242-266."""
    if cfg == "":
        return lambda length, lp: lp
    kind, alpha = cfg.split("_")
    alpha = float(alpha)
    if kind == "wu":
        return lambda length, lp: lp / (((5 + length) ** alpha) / ((5 + 1) ** alpha))
    if kind == "avg":
        return lambda length, lp: lp / length
    raise ValueError(cfg)


def beam_search_one(sd, dims, fc, att, p_att, masks, beam_size, length_penalty="", decoding_constraint=0):
    """One sub-graph: models/AttModel.py:216-231 + models/CaptionModel.py:28-176 with group_size 1.
    fc [1,H] ... are the prepared tensors of that sub-graph.  Returns the `done_beams` list (dicts with
    seq [T] int64, logps [T] f32, unaug_p, p), best first."""
    b, T = beam_size, dims.seq_length
    pen = _length_penalty(length_penalty)
    fc_b = fc.expand(b, -1)
    att_b = att.expand(b, -1, -1).contiguous()
    patt_b = p_att.expand(b, -1, -1).contiguous()
    mask_b = masks.expand(b, -1).contiguous()
    state = init_state(dims, b)
    logp, state, _ = decoder_step(sd, torch.zeros(b, dtype=torch.int64), fc_b, att_b, patt_b, mask_b, state)
    beam_seq = torch.zeros(T, b, dtype=torch.int64)
    beam_lp = torch.zeros(T, b)
    beam_sum = torch.zeros(b)
    done = []
    for t in range(T):
        lpf = logp.clone().float()
        if decoding_constraint and t > 0:
            lpf.scatter_(1, beam_seq[t - 1].unsqueeze(1), float("-inf"))
        lpf[:, -1] = lpf[:, -1] - 1000                      # UNK suppression, CaptionModel.py:131
        ys, ix = torch.sort(lpf, 1, True)
        rows = 1 if t == 0 else b
        cands = []
        for c in range(min(b, ys.size(1))):
            for q in range(rows):
                cands.append((ix[q, c], q, beam_sum[q] + ys[q, c].item(), lpf[q, ix[q, c]]))
        cands.sort(key=lambda v: -v[2])                     # stable, CaptionModel.py:69
        new_h, new_c = state[0].clone(), state[1].clone()
        prev_seq, prev_lp = beam_seq[:t].clone(), beam_lp[:t].clone()
        for v in range(b):
            tok, q, p, raw = cands[v]
            if t >= 1:
                beam_seq[:t, v] = prev_seq[:, q]
                beam_lp[:t, v] = prev_lp[:, q]
            new_h[:, v] = state[0][:, q]
            new_c[:, v] = state[1][:, q]
            beam_seq[t, v] = tok
            beam_lp[t, v] = raw
            beam_sum[v] = p
        state = (new_h, new_c)
        for v in range(b):
            if beam_seq[t, v] == 0 or t == T - 1:
                done.append(dict(seq=beam_seq[:, v].clone(), logps=beam_lp[:, v].clone(),
                                 unaug_p=beam_lp[:, v].sum().item(), p=pen(t + 1, beam_sum[v].item())))
                beam_sum[v] = -1000
        logp, state, _ = decoder_step(sd, beam_seq[t], fc_b, att_b, patt_b, mask_b, state)
    return sorted(done, key=lambda v: -v["p"])[:b]


# ----------------------------------------------------------------------------------------------------------
# whole-path entry points (mirror AttModel._sample / _forward and LossWrapper)
# ----------------------------------------------------------------------------------------------------------
def sample(sd, dims, data, beam_size=1, topk=False, temp=0.6, k=3, use_nms=True, iou_thres=0.75, max_subgraphs=1,
           return_att=False, length_penalty="", uniforms=None, seq_per_img=5):
    """AttModel._sample / _sample_sentences (models/AttModel.py:179-326) on a loader-shaped batch `data`.
    Returns dict: seq [S,T], seqLogprobs [S,T], subgraph_score [S], keep_ind [S], gpn_loss,
    (att_weights), (done_beams), image_of_row."""
    x_obj, _ = encode(sd, dims, data["att_feats"], data["obj_dist"], data["pred_dist"], data["rel_ind"])
    g = sgpn_test(sd, dims, x_obj, data["gpn_obj_ind"], data["att_masks"], data["gpn_pool_mtx"], use_nms, iou_thres,
                  max_subgraphs, seq_per_img)
    fc, att, p_att, masks = prepare_features(sd, dims, g["fc_feats"], g["att_feats"], g["att_masks"])
    out = dict(subgraph_score=g["score"], keep_ind=g["keep_ind"], gpn_loss=g["gpn_loss"], image_of_row=g["image_of_row"],
               x_obj=x_obj, all_scores=g["all_scores"], p_fc=fc, p_att=att, pp_att=p_att, p_masks=masks)
    if beam_size > 1:
        S, T = fc.shape[0], dims.seq_length
        seq = torch.zeros(S, T, dtype=torch.int64)
        lps = torch.zeros(S, T)
        beams = []
        for s in range(S):
            done = beam_search_one(sd, dims, fc[s:s + 1], att[s:s + 1], p_att[s:s + 1], masks[s:s + 1], beam_size,
                                   length_penalty)
            beams.append(done)
            seq[s] = done[0]["seq"]
            lps[s] = done[0]["logps"]
        out.update(seq=seq, seqLogprobs=lps, done_beams=beams)
        return out
    r = decode_greedy_or_topk(sd, dims, fc, att, p_att, masks, topk, temp, k, return_att, uniforms)
    out.update(seq=r[0], seqLogprobs=r[1])
    if return_att:
        out["att_weights"] = r[2]
    return out


def forward_train(sd, dims, data, seq_per_img=5, drop=None):
    """AttModel._forward (models/AttModel.py:122-177), sampling_prob = 0, dropout off unless `drop` gives
    masks.  Returns outputs [5B, T', V+1] log-probs, gpn_loss, subgraph_score, plus intermediates."""
    x_obj, _ = encode(sd, dims, data["att_feats"], data["obj_dist"], data["pred_dist"], data["rel_ind"])
    loss, score, att, fcf, masks, _ = sgpn_train(sd, dims, x_obj, data["gpn_obj_ind"], data["att_masks"],
                                                 data["gpn_pool_mtx"], seq_per_img,
                                                 None if drop is None else drop.get("gpn"))
    fc, att_e, p_att, masks_c = prepare_features(sd, dims, fcf, att, masks, drop)
    labels = data["labels"]
    rows, steps = labels.shape[0], labels.shape[1] - 1
    outputs = torch.zeros(rows, steps, dims.v1)
    state = init_state(dims, rows)
    for i in range(steps):
        if i >= 1 and labels[:, i].sum() == 0:
            break
        logp, state, _ = decoder_step(sd, labels[:, i].clone(), fc, att_e, p_att, masks_c, state,
                                      None if drop is None else drop["x"][i], None if drop is None else drop["h"][i])
        outputs[:, i] = logp
    return dict(outputs=outputs, gpn_loss=loss, subgraph_score=score, x_obj=x_obj, sel_att=att, sel_fc=fcf, sel_masks=masks)


def language_model_criterion(outputs, target, mask):
    """misc/utils.py:115-124."""
    target = target[:, :outputs.size(1)]
    mask = mask[:, :outputs.size(1)]
    picked = -outputs.gather(2, target.unsqueeze(2)).squeeze(2) * mask
    return picked.sum() / mask.sum()


def loss_wrapper(sd, dims, data, seq_per_img=5, drop=None):
    """models/loss_wrapper.py:14-27."""
    r = forward_train(sd, dims, data, seq_per_img, drop)
    lang = language_model_criterion(r["outputs"], data["labels"][:, 1:], data["masks"][:, 1:])
    return dict(gpn_loss=r["gpn_loss"], lang_loss=lang, **{k: v for k, v in r.items() if k not in ("gpn_loss",)})


# ----------------------------------------------------------------------------------------------------------
# post-decode
# ----------------------------------------------------------------------------------------------------------
BAD_ENDINGS = ['with', 'in', 'on', 'of', 'a', 'at', 'to', 'for', 'an', 'this', 'his', 'her', 'that', 'the']   # misc/utils.py:16-17


def decode_sequence(ix_to_word, seq, remove_bad_endings=False):
    """misc/utils.py:59-81, token by token as the reference does it."""
    out = []
    for i in range(seq.shape[0]):
        txt = ''
        for j in range(seq.shape[1]):
            ix = int(seq[i, j])
            if ix > 0:
                if j >= 1:
                    txt = txt + ' '
                txt = txt + ix_to_word[str(ix)]
            else:
                break
        if remove_bad_endings:
            flag = 0
            words = txt.split(' ')
            for j in range(len(words)):
                if words[-j - 1] not in BAD_ENDINGS:
                    flag = -j
                    break
            txt = ' '.join(words[0:len(words) + flag])
        out.append(txt)
    return out


def sort_by_score(seq, subgraph_score, keep_ind):
    """misc/eval_utils.py:105-110 for ONE image."""
    sorted_score, sort_ind = torch.sort(subgraph_score, descending=True, stable=True)
    return seq[sort_ind], sorted_score, keep_ind[sort_ind], sort_ind
