"""GPU tests of the training path: every CUDA building block of subgc.train.CudaOps against its torch emulation
(tests/emu_ops.py), and the end-to-end losses / parameter gradients of LossWrapper(...).backward() against the
reference's autograd results stored in tests/golden (dropout off: RNG streams cannot be matched)."""
import numpy as np
import pytest
import torch

from emu_ops import EmuOps
from helpers import load_golden, rebuild_train_case, rel_err, t2n
from subgc import _lib, synth
from subgc.config import SMALL, make_opt
from subgc.model import LossWrapper, setup
from subgc.train import CudaOps

pytestmark = pytest.mark.gpu
TOL = 2e-5


def cu(t):
    return t.cuda() if torch.is_tensor(t) else t


@pytest.fixture(scope="module")
def ops():
    d = SMALL
    cd = _lib.Dims(d.v1, d.enc, d.rnn, d.att_hid, d.fc_feat, d.att_feat, d.gcn, d.low_rank, d.embed, d.obj_classes, d.pred_classes, d.gcn_layers,
                   d.gcn_residual, d.pred_emb_type, d.seq_length, d.obj_num, d.rel_num)
    return CudaOps(cd), EmuOps(), d


def close(a, b, tol=TOL):
    assert rel_err(t2n(a), t2n(b)) <= tol


def test_gemm_transpose_colsum_pointwise(ops):
    c, e, d = ops
    g = torch.Generator().manual_seed(0)
    for (M, N, K) in [(7, 5, 3), (130, 260, 72), (64, 1000, 512), (1, 40, 300)]:
        x, w, b = torch.randn(M + 2, K, generator=g), torch.randn(N, K, generator=g), torch.randn(N, generator=g)
        idx = torch.randint(0, M + 2, (M,), generator=g)
        close(c.linear(cu(x), cu(w), cu(b), relu=True, gather=cu(idx)), e.linear(x, w, b, relu=True, gather=idx))
        out0 = torch.randn(M, N, generator=g)
        oc = cu(out0.clone())
        c.linear(cu(x[:M]), cu(w), None, out=oc, accumulate=True)
        close(oc, e.linear(x[:M], w, None, out=out0.clone(), accumulate=True))
        close(c.transpose(cu(x)), x.t())
        close(c.transpose(cu(x)[:, 1:K]), x[:, 1:K].t())
        acc = torch.randn(K, generator=g)
        close(c.colsum(cu(x), cu(acc.clone())), acc + x.sum(0))
    a, b = torch.randn(1000, generator=g), torch.randn(1000, generator=g)
    close(c.mul(cu(a), cu(b)), a * b); close(c.add(cu(a), cu(b)), a + b); close(c.relu_bwd(cu(a), cu(b)), e.relu_bwd(a, b))
    close(c.relu(cu(a)), torch.relu(a)); close(c.sigmoid(cu(a)), torch.sigmoid(a)); close(c.scale(cu(a), 0.25), a * 0.25)
    src = torch.randn(50, 33, generator=g); idx = torch.randint(0, 50, (77,), generator=g)
    close(c.gather_rows(cu(src), cu(idx), relu=True), torch.relu(src[idx]))
    dst = torch.zeros(50, 33); dc = cu(dst.clone()); upd = torch.randn(77, 33, generator=g)
    c.scatter_add_rows(cu(upd), cu(idx), dc); e.scatter_add_rows(upd, idx, dst)
    close(dc, dst, 1e-5)


def test_dropout_mask_statistics(ops):
    c, _, _ = ops
    m = c.dropout_mask((1000, 1000), 0.5, 1234, 1, torch.device("cuda"))
    vals = torch.unique(m).cpu().tolist()
    assert vals == [0.0, 2.0]
    assert abs(float(m.mean()) - 1.0) < 5e-3                      # keep prob 0.5, scale 2
    m2 = c.dropout_mask((1000, 1000), 0.5, 1234, 1, torch.device("cuda"))
    m3 = c.dropout_mask((1000, 1000), 0.5, 1234, 2, torch.device("cuda"))
    assert torch.equal(m, m2) and not torch.equal(m, m3)         # reproducible per (seed, offset)
    assert abs(float(((m > 0) & (m3 > 0)).float().mean()) - 0.25) < 5e-3   # independent streams
    m4 = c.dropout_mask((999, 7), 0.3, 5, 9, torch.device("cuda"))
    assert abs(float((m4 > 0).float().mean()) - 0.7) < 2e-2


def test_lstm_and_attention_blocks(ops):
    c, e, d = ops
    g = torch.Generator().manual_seed(1)
    S, H, AH, ln = 9, d.rnn, d.att_hid, 11
    gates, c_prev = torch.randn(S, 4 * H, generator=g), torch.randn(S, H, generator=g)
    gc = cu(gates.clone()); ge = gates.clone()
    hc, cc = c.lstm_fwd(gc, cu(c_prev)); he, ce = e.lstm_fwd(ge, c_prev)
    close(hc, he); close(cc, ce); close(gc, ge)
    dh, dc = torch.randn(S, H, generator=g), torch.randn(S, H, generator=g)
    for dcin in (None, dc):
        a1, b1 = c.lstm_bwd(gc, cu(c_prev), cc, cu(dh), cu(dcin)); a2, b2 = e.lstm_bwd(ge, c_prev, ce, dh, dcin)
        close(a1, a2); close(b1, b2)
    atth, p_att, att = torch.randn(S, AH, generator=g), torch.randn(S, ln, AH, generator=g), torch.randn(S, ln, H, generator=g)
    lens = torch.randint(1, ln + 1, (S,), generator=g); lens[0] = ln
    masks = (torch.arange(ln).view(1, -1) < lens.view(-1, 1)).float()
    aw, ab = torch.randn(1, AH, generator=g), torch.randn(1, generator=g)
    r1 = c.att_fwd(cu(atth), cu(p_att), cu(att), cu(masks), cu(aw), cu(ab)); r2 = e.att_fwd(atth, p_att, att, masks, aw, ab)
    for x, y in zip(r1, r2):
        close(x, y)
    dctx = torch.randn(S, H, generator=g)
    da0, dp0 = torch.randn(S, ln, H, generator=g), torch.randn(S, ln, AH, generator=g)
    dac, dpc = cu(da0.clone()), cu(dp0.clone()); dae, dpe = da0.clone(), dp0.clone()
    o1 = c.att_bwd(cu(atth), cu(p_att), cu(att), cu(masks), cu(aw), r1[1], r1[2], cu(dctx), dac, dpc)
    o2 = e.att_bwd(atth, p_att, att, masks, aw, r2[1], r2[2], dctx, dae, dpe)
    close(o1[0], o2[0]); close(o1[1], o2[1]); close(dac, dae); close(dpc, dpe)
    logits = torch.randn(S, d.v1, generator=g) * 3
    out = torch.zeros(S, 3, d.v1); oc = cu(out.clone())
    c.log_softmax_fwd(cu(logits), oc[:, 1]); e.log_softmax_fwd(logits, out[:, 1])
    close(oc, out)
    dl = torch.randn(S, 3, d.v1, generator=g)
    close(c.log_softmax_bwd(oc[:, 1], cu(dl)[:, 1]), e.log_softmax_bwd(out[:, 1], dl[:, 1]))


def test_graph_blocks(ops):
    c, e, d = ops
    data = synth.make_train_inputs(d, 3, n_images=3, gpn_batch=2)
    g = torch.Generator().manual_seed(2)
    B, N, K, L = 3, d.obj_num, d.rel_num, d.gcn
    rel = data["rel_ind"]
    m2, m3 = torch.randn(B, N, L, generator=g), torch.randn(B, N, L, generator=g)
    res = torch.randn(B, K, L, generator=g)
    close(c.gcn_edge_fwd(cu(m2), cu(m3), cu(rel), cu(res)), e.gcn_edge_fwd(m2, m3, rel, res))
    m0, m1 = torch.randn(B, K, L, generator=g), torch.randn(B, K, L, generator=g)
    resx = torch.randn(B, N, L, generator=g)
    r1 = c.gcn_node_fwd(cu(m0), cu(m1), cu(rel), cu(resx), N); r2 = e.gcn_node_fwd(m0, m1, rel, resx, N)
    for x, y in zip(r1, r2):
        close(x, y)
    dx, dp = torch.randn(B, N, L, generator=g), torch.randn(B, K, L, generator=g)
    for x, y in zip(c.gcn_node_bwd(cu(dx), r1[1], r1[2], cu(rel)), e.gcn_node_bwd(dx, r2[1], r2[2], rel)):
        close(x, y)
    for x, y in zip(c.gcn_edge_bwd(cu(dp), cu(m2), cu(m3), cu(rel)), e.gcn_edge_bwd(dp, m2, m3, rel)):
        close(x, y)
    rows, G = data["gpn_obj_ind"].shape[0], data["gpn_obj_ind"].shape[2]
    layc, laye = c.layout(rows, G, 0), e.layout(rows, G, 0)
    n_sub = 2 * rows * G
    x_obj = torch.randn(B, N, L, generator=g).abs()
    rc, lc = c.pool(layc, n_sub, cu(x_obj), cu(data["gpn_obj_ind"]), cu(data["att_masks"]))
    re, le = e.pool(laye, n_sub, x_obj, data["gpn_obj_ind"], data["att_masks"])
    close(rc, re); assert torch.equal(lc.cpu(), le)
    d_read = torch.randn(n_sub, 2 * L, generator=g)
    dxc, dxe = torch.zeros(B, N, L).cuda(), torch.zeros(B, N, L)
    c.pool_bwd(layc, cu(x_obj), cu(data["gpn_obj_ind"]), lc, cu(d_read), dxc); e.pool_bwd(laye, x_obj, data["gpn_obj_ind"], le, d_read, dxe)
    close(dxc, dxe, 1e-5)
    score = torch.rand(n_sub, generator=g)
    close(c.bce(layc, cu(score)), e.bce(laye, score)); close(c.bce_bwd(layc, cu(score), 0.3), e.bce_bwd(laye, score, 0.3))
    s1, l1 = c.select_train(layc, cu(score), lc); s2, l2 = e.select_train(laye, score, le)
    assert torch.equal(s1.cpu(), s2) and l1 == l2
    for x, y in zip(c.prepare_index(layc, s1, l1, cu(data["gpn_obj_ind"]), cu(data["att_masks"])),
                    e.prepare_index(laye, s2, l2, data["gpn_obj_ind"], data["att_masks"])):
        assert torch.equal(x.cpu(), y if y.dtype != torch.float32 else y) or rel_err(t2n(x), t2n(y)) == 0
    cls = c.class_argmax(cu(data["obj_dist"]).reshape(B * N, -1), 1)
    assert torch.equal(cls.cpu(), e.class_argmax(data["obj_dist"].reshape(B * N, -1), 1))


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("name", ["small_train", "small_train_refinit", "full_train"])
def test_loss_wrapper_backward_matches_reference(name, fused):
    """fused: LossWrapper's log-softmax + LanguageModelCriterion inside the decoder stage (no [rows, T, V+1] tensor) / the model's
    log-probs followed by the criterion as in models/loss_wrapper.py:22.  full_train: the reference's own autograd gradients at FULL dimensions (V = 9487, H = 1000, 2048-d features; 2 images, 10
    sentences): the training path's tensor-core contractions (split-TF32, K up to 3000, N = 9488) against something real."""
    g = load_golden(name)
    d, sd, data = rebuild_train_case(g)
    model = setup(make_opt(d))
    model.load_state_dict(sd)
    model.cuda().train()
    model.dropout_enabled = False
    model.fused_loss = fused
    lw = LossWrapper(model, None)
    dev = {k: cu(v) for k, v in data.items()}
    res = lw(dev["fc_feats"], dev["att_feats"], dev["labels"], dev["masks"], dev["att_masks"], None, None, None, dev["obj_dist"], None,
             dev["rel_ind"], None, dev["pred_dist"], dev["gpn_obj_ind"], dev["gpn_pred_ind"], dev["gpn_nrel_ind"], dev["gpn_pool_mtx"])
    assert abs(float(res["lang_loss"]) - float(g["lang_loss"])) <= 1e-5 * max(1.0, abs(float(g["lang_loss"])))
    assert abs(float(res["gpn_loss"]) - float(g["gpn_loss"])) <= 1e-5
    (res["lang_loss"] + res["gpn_loss"]).backward()
    checked = 0
    grads = dict(model.named_parameters())
    for k in g.files:
        if k.startswith("grad_none__"):
            n = k[len("grad_none__"):]
            assert grads[n].grad is None or float(grads[n].grad.abs().sum()) == 0.0, n
        elif k.startswith("grad__"):
            n = k[len("grad__"):]
            assert grads[n].grad is not None, n
            gr = grads[n].grad.double().reshape(-1).cpu()
            ref = g[k]
            if n == "core.attention.alpha_net.bias":
                assert float(gr.abs().sum()) <= 1e-6
                continue
            scale = max(ref[1] / max(gr.numel(), 1), 1e-12)
            assert abs(float(gr.sum()) - ref[0]) <= 1e-4 * max(ref[1], 1e-12), n
            assert abs(float(gr.abs().sum()) - ref[1]) <= 1e-4 * max(ref[1], 1e-12), n
            head = gr[:24].numpy()
            assert np.abs(head - ref[3:3 + len(head)]).max() <= 2e-4 * max(np.abs(ref[3:]).max(), scale), n
            checked += 1
    assert checked >= 40


def test_train_mode_with_dropout_runs_and_is_stochastic():
    d = SMALL
    sd = synth.make_state_dict(d, 9, logit_gain=4.0)
    data = {k: cu(v) for k, v in synth.make_train_inputs(d, 9, n_images=2, gpn_batch=2).items()}
    model = setup(make_opt(d))
    model.load_state_dict(sd)
    model.cuda().train()
    lw = LossWrapper(model, None)
    args = (data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
            data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])
    torch.manual_seed(0)
    a = lw(*args)
    (a["lang_loss"] + a["gpn_loss"]).backward()
    g1 = model.logit.weight.grad.clone()
    assert torch.isfinite(g1).all() and float(g1.abs().sum()) > 0
    b = lw(*args)
    assert float(a["lang_loss"]) != float(b["lang_loss"])          # different dropout masks
    torch.manual_seed(0)
    c2 = lw(*args)
    assert float(c2["lang_loss"]) == float(a["lang_loss"])         # same torch seed -> same Philox seed -> same masks
    model.eval()
    with torch.no_grad():
        e1 = lw(*args); e2 = lw(*args)
    assert float(e1["lang_loss"]) == float(e2["lang_loss"])


def test_scheduled_sampling_kernel_distribution_and_rate(ops):
    """models/AttModel.py:158-167: per row, with probability ss_prob the fed token is drawn from exp(previous log-probs).  RNG streams cannot
    match torch.multinomial's, so the kernel is checked statistically: the rate of replaced rows and a chi-square of the drawn tokens."""
    c, _, _ = ops
    V1, rows = 50, 4096
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(V1, generator=g) * 1.5
    logp = torch.log_softmax(logits, 0)
    prev = logp.unsqueeze(0).expand(rows, V1).contiguous().cuda()
    labels = torch.full((rows, 3), V1 + 7, dtype=torch.int64).cuda()          # a token id that cannot be drawn marks "ground truth kept"
    counts = torch.zeros(V1, dtype=torch.float64)
    replaced = total = 0
    for rep in range(6):
        it = c.ss_sample(prev, labels[:, 1], 0.25, 99, rep).cpu()
        drawn = it[it != V1 + 7]
        replaced += drawn.numel(); total += rows
        counts += torch.bincount(drawn, minlength=V1).double()
    rate = replaced / total
    assert abs(rate - 0.25) < 4 * (0.25 * 0.75 / total) ** 0.5, rate
    expect = torch.exp(logp).double() * replaced
    keep = expect > 5
    chi2 = float((((counts - expect) ** 2) / expect)[keep].sum())
    assert chi2 < 2.2 * int(keep.sum()), (chi2, int(keep.sum()))            # ~ p = 1e-5 for 30-50 degrees of freedom
    assert torch.equal(c.ss_sample(prev, labels[:, 1], 0.25, 99, 1), c.ss_sample(prev, labels[:, 1], 0.25, 99, 1))   # reproducible
    assert int((c.ss_sample(prev, labels[:, 1], 0.0, 99, 1) != V1 + 7).sum()) == 0
    assert int((c.ss_sample(prev, labels[:, 1], 1.0, 99, 1) == V1 + 7).sum()) == 0


def test_training_with_scheduled_sampling_runs():
    d = SMALL
    sd = synth.make_state_dict(d, 9, logit_gain=4.0)
    data = {k: cu(v) for k, v in synth.make_train_inputs(d, 9, n_images=2, gpn_batch=2).items()}
    model = setup(make_opt(d))
    model.load_state_dict(sd)
    model.cuda().train()
    model.dropout_enabled = False
    lw = LossWrapper(model, None)
    args = (data["fc_feats"], data["att_feats"], data["labels"], data["masks"], data["att_masks"], None, None, None, data["obj_dist"], None,
            data["rel_ind"], None, data["pred_dist"], data["gpn_obj_ind"], data["gpn_pred_ind"], data["gpn_nrel_ind"], data["gpn_pool_mtx"])
    base = float(lw(*args)["lang_loss"])
    model.ss_prob = 0.75                                    # train.py:131 raises it epoch by epoch
    torch.manual_seed(1)
    a = lw(*args)
    (a["lang_loss"] + a["gpn_loss"]).backward()
    assert torch.isfinite(model.logit.weight.grad).all() and float(model.logit.weight.grad.abs().sum()) > 0
    assert float(a["lang_loss"]) != base                   # sampled tokens were fed
    torch.manual_seed(1)
    assert float(lw(*args)["lang_loss"]) == float(a["lang_loss"])
    model.eval()                                            # evaluation never samples (AttModel.py:158: `self.training and ...`)
    with torch.no_grad():
        assert abs(float(lw(*args)["lang_loss"]) - base) <= 1e-5 * max(1.0, abs(base))
