"""Post-decode step: id -> string conversion against the reference's own decode_sequence (misc/utils.py:59-81, imported
from baseline/_ref when present, else the oracle's token-by-token restatement), and the per-image score ordering (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

import subgc_oracle as O
from subgc import postdecode

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_decode():
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isfile(os.path.join(ref, "misc", "utils.py")):
        sys.path.insert(0, ref)
        try:
            import misc.utils as U
            return U.decode_sequence
        except Exception:
            return None
    return None


def _cases():
    words = ["a", "man", "with", "the", "dog", "on", "of", "in", "street", "this", "red", "his"]
    ix_to_word = {str(i + 1): w for i, w in enumerate(words)}
    g = torch.Generator().manual_seed(0)
    seq = torch.randint(0, len(words) + 1, (200, 20), generator=g)
    seq[0] = 0                       # empty sentence
    seq[1, :] = torch.tensor([1, 3, 4, 6, 7] * 4)   # only bad endings
    seq[2, 5:] = 0
    seq[3] = torch.randint(1, len(words) + 1, (20,), generator=g)   # no terminator
    seq[4, 0] = 0; seq[4, 1] = 5     # token after the terminator is ignored
    return ix_to_word, seq


@pytest.mark.parametrize("bad", [False, True])
def test_decode_sequences_matches_reference(bad, monkeypatch):
    ix_to_word, seq = _cases()
    monkeypatch.setenv("REMOVE_BAD_ENDINGS", "1" if bad else "0")
    got = postdecode.decode_sequences(ix_to_word, seq)
    assert got == O.decode_sequence(ix_to_word, seq, remove_bad_endings=bad)
    ref = _reference_decode()
    if ref is not None:
        assert got == ref(ix_to_word, seq)          # the unmodified reference function
    assert got[0] == "" and (" " not in got[4])


@pytest.mark.gpu
def test_rows_sorted_per_image_like_the_reference():
    g = torch.Generator().manual_seed(1)
    counts = [1, 4, 7, 2, 33, 1, 12]
    image = torch.cat([torch.full((c,), i) for i, c in enumerate(counts)])
    n = int(image.numel())
    score = torch.rand(n, generator=g)
    score[5] = score[6]                                  # a tie inside image 2
    seq = torch.randint(0, 50, (n, 20), generator=g)
    keep = torch.arange(n) * 3
    s_seq, s_score, s_keep, order, img = postdecode.sort_rows(seq.cuda(), score.cuda(), keep.cuda(), image.cuda())
    at = 0
    for i, c in enumerate(counts):
        sl = slice(at, at + c)
        r_seq, r_score, r_keep, r_ind = O.sort_by_score(seq[sl], score[sl], keep[sl])
        assert torch.equal(s_seq[sl].cpu(), r_seq) and torch.equal(s_score[sl].cpu(), r_score) and torch.equal(s_keep[sl].cpu(), r_keep)
        assert torch.equal(order[sl].cpu() - at, r_ind)
        at += c
    assert torch.equal(img.cpu(), image)
    ix_to_word = {str(i): f"w{i}" for i in range(1, 50)}
    entries = postdecode.collect_predictions(ix_to_word, list(range(len(counts))), seq.cuda(), score.cuda(), keep.cuda(), image.cuda())
    assert [len(e["caption"]) for e in entries] == counts
    assert entries[4]["caption"] == O.decode_sequence(ix_to_word, O.sort_by_score(seq[14:47], score[14:47], keep[14:47])[0])
