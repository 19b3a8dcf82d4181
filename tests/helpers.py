"""Shared test plumbing: rebuild the exact tensors a golden fixture was generated from."""
import os

import numpy as np
import torch

from subgc import synth
from subgc.config import Dims

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GAIN_KEYS = ("gcn_std", "logit_gain", "lstm_gain", "eos_bias")


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


def dims_of(g):
    return Dims(**dict(zip(Dims().as_dict().keys(), (int(v) for v in g["meta_dims"]))))


def gains_of(g):
    v = dict(zip(GAIN_KEYS, (float(x) for x in g["meta_gains"])))
    if v["gcn_std"] < 0:
        v["gcn_std"] = None
    return v


def same_fingerprint(a, b):
    """Checksums are float64 sums over up to 70 M elements: the summation order depends on the host's thread count,
    so equality is up to rounding of the sum (1e-9 relative is ~1000x that noise and far below any real drift)."""
    return abs(float(a) - float(b)) <= 1e-9 * max(1.0, abs(float(b)))


def rebuild_test_case(g):
    """(dims, state_dict, data, nms kwargs) for a fixture written by make_golden.run_test_case."""
    d = dims_of(g)
    seed = int(g["meta_seed"])
    sd = synth.make_state_dict(d, seed, **gains_of(g))
    data = synth.make_test_inputs(d, seed, n_images=1, per_half=int(g["meta_per_half"]), ragged=bool(g["meta_ragged"]),
                                  ragged_edges=bool(g["meta_ragged_edges"]))
    assert same_fingerprint(synth.fingerprint(sd), g["meta_fp_weights"]), "synthetic weights drifted from the fixture"
    assert same_fingerprint(synth.fingerprint([a for a in synth.sample_args(data) if a is not None]), g["meta_fp_inputs"])
    nms = dict(iou_thres=float(g["meta_nms_thres"]), max_subgraphs=int(g["meta_nms_max"]))
    return d, sd, data, nms


def rebuild_train_case(g):
    d = dims_of(g)
    seed = int(g["meta_seed"])
    sd = synth.make_state_dict(d, seed, **gains_of(g))
    data = synth.make_train_inputs(d, seed, n_images=int(g["meta_n_images"]), gpn_batch=int(g["meta_gpn_batch"]))
    assert same_fingerprint(synth.fingerprint(sd), g["meta_fp_weights"])
    assert same_fingerprint(synth.fingerprint([v for v in synth.forward_args(data) if v is not None]), g["meta_fp_inputs"])
    return d, sd, data


def beam_sizes_in(g):
    return sorted({int(k[4:k.index("_")]) for k in g.files if k.startswith("beam") and k.endswith("_seq") and "beam_seq" not in k})


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def t2n(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def check_train_outputs(outputs, g, tol):
    """outputs [rows, T, V1] of a training-case fixture: stored in full at small dims, as a slice + checksums + row-wise arg-max at full
    dims (oracle/make_golden.py: run_train_case(store_full=False))."""
    o = t2n(outputs)
    if "outputs" in g.files:
        assert o.shape == g["outputs"].shape and rel_err(o, g["outputs"]) <= tol
        return
    assert rel_err(o[:, :, :48], g["outputs_slice"]) <= tol
    assert abs(float(np.asarray(o, np.float64).sum()) - float(g["outputs_sum"])) <= tol * float(g["outputs_abssum"])
    assert np.array_equal(o.argmax(2), g["outputs_argmax"])
