"""CPU check of the training-step orchestration and backward algebra (subgc.train.forward / backward) with the torch
emulation of the CUDA building blocks, against the reference's own losses and autograd gradients (tests/golden)."""
import numpy as np
import pytest
import torch

from emu_ops import EmuOps
from helpers import check_train_outputs, load_golden, rebuild_train_case, rel_err, t2n
from subgc import train
from subgc.model import LanguageModelCriterion


@pytest.mark.parametrize("name", ["small_train", "small_train_refinit"])
def test_train_step_matches_reference_gradients(name):
    g = load_golden(name)
    d, sd, data = rebuild_train_case(g)
    ops = EmuOps()
    with torch.no_grad():
        outputs, gpn_loss, score, S = train.forward(ops, sd, sd, d, data, drop=None)
    check_train_outputs(outputs, g, 1e-5)
    assert rel_err(t2n(score), g["subgraph_score"]) <= 1e-5
    assert abs(float(gpn_loss) - float(g["gpn_loss"])) <= 1e-5
    leaf = outputs.clone().requires_grad_(True)
    lang = LanguageModelCriterion()(leaf, data["labels"][:, 1:], data["masks"][:, 1:])
    assert abs(float(lang) - float(g["lang_loss"])) <= 1e-5 * max(1.0, abs(float(g["lang_loss"])))
    lang.backward()
    with torch.no_grad():
        G = train.backward(ops, sd, d, S, leaf.grad, 1.0)
    checked = 0
    for k in g.files:
        if k.startswith("grad_none__"):
            n = k[len("grad_none__"):]
            assert n not in G or float(G[n].abs().sum()) == 0.0, n
        elif k.startswith("grad__"):
            n = k[len("grad__"):]
            assert n in G, f"missing gradient for {n}"
            gr = G[n].double().reshape(-1)
            ref = g[k]
            if n == "core.attention.alpha_net.bias":  # softmax is shift-invariant: exactly 0; the reference holds rounding noise
                assert float(gr.abs().sum()) <= 1e-6 and ref[1] <= 1e-6
                continue
            scale = max(ref[1] / max(gr.numel(), 1), 1e-12)
            assert abs(float(gr.sum()) - ref[0]) <= 1e-4 * max(ref[1], 1e-12), n
            assert abs(float(gr.abs().sum()) - ref[1]) <= 1e-4 * max(ref[1], 1e-12), n
            head = gr[:24].numpy()
            assert np.abs(head - ref[3:3 + len(head)]).max() <= 2e-4 * max(np.abs(ref[3:]).max(), scale), n
            checked += 1
    assert checked >= 40


def test_gcn_liveness_matches_the_reference_dead_units():
    need_x, need_p = train.gcn_liveness(2, 2)
    assert need_x == [True, False, True] and need_p == [False, True, False]  # L0 units 2,3 and L1 units 0,1 are live
    need_x, need_p = train.gcn_liveness(4, 1, want_x_pred=True)
    assert all(need_x) and all(need_p)
