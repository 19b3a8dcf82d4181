"""Post-decode step of the caption pipeline: what the reference does between `model(..., mode='sample')` and the
`captions_*.npy` writer (misc/eval_utils.py:105-141, misc/utils.py:59-81), for a whole batch of images at once.

  * rows of an image sorted by descending sGPN score ON THE DEVICE (subgc_rank_rows), token ids / scores / kept indices gathered with it,
  * ONE device->host copy per tensor instead of an `.item()` per token (misc/utils.py:66-69 reads 20 x rows scalars),
  * id -> string through a numpy vocabulary table; a sentence ends at the first 0; REMOVE_BAD_ENDINGS as the reference applies it.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ._lib import check, lib, ptr

BAD_ENDINGS = ['with', 'in', 'on', 'of', 'a', 'at', 'to', 'for', 'an', 'this', 'his', 'her', 'that', 'the']   # misc/utils.py:16-17


def sort_rows(seq, subgraph_score, keep_ind, image_of_row, sort=True):
    """Device side of eval_utils.py:105-115.  Returns (seq, score, sorted_subgraph_ind, sort_ind, image) with the rows of every image in
    descending score order (sort=False: model.gpn is off / sct mode keeps the input order)."""
    n = subgraph_score.shape[0]
    if not sort or n == 0:
        order = torch.arange(n, device=subgraph_score.device)
    else:
        order = torch.empty(n, dtype=torch.int64, device=subgraph_score.device)
        img = image_of_row.contiguous().long()
        sc = subgraph_score.contiguous().float()
        check(lib().subgc_rank_rows(n, ptr(sc), ptr(img), ptr(order), torch.cuda.current_stream().cuda_stream), "subgc_rank_rows")
    seq_d = seq.to(order.device) if seq.device != order.device else seq      # beam search hands back CPU tensors (AttModel.py:212-213)
    return seq_d[order], subgraph_score[order], keep_ind[order], order, image_of_row[order]


class Vocab:
    """ix_to_word of the loaders (dataloader.py:69: keys are the decimal strings '1'..'V') as a numpy table indexed by token id."""

    def __init__(self, ix_to_word):
        size = max(int(k) for k in ix_to_word) + 1
        self.table = np.empty(size, dtype=object)
        self.table[:] = ""
        for k, w in ix_to_word.items():
            self.table[int(k)] = w

    def __call__(self, ids):
        return self.table[ids]


def decode_sequences(vocab, seq, remove_bad_endings=None):
    """misc/utils.py:59-81 for the whole [rows, T] tensor: one host copy, no per-token .item()."""
    if not isinstance(vocab, Vocab):
        vocab = Vocab(vocab)
    if remove_bad_endings is None:
        remove_bad_endings = bool(int(os.getenv('REMOVE_BAD_ENDINGS', '0')))
    ids = seq.detach().cpu().numpy() if torch.is_tensor(seq) else np.asarray(seq)
    ended = ids <= 0
    length = np.where(ended.any(1), ended.argmax(1), ids.shape[1])    # a sentence stops at the first token <= 0
    words = vocab(np.where(ids > 0, ids, 0))
    out = []
    for row, n in zip(words, length):
        toks = list(row[:n])
        if remove_bad_endings:
            # the reference strips trailing bad endings but keeps a sentence made only of them (flag stays 0 there)
            flag = 0
            ws = " ".join(toks).split(" ")
            for j in range(len(ws)):
                if ws[-j - 1] not in BAD_ENDINGS:
                    flag = -j
                    break
            toks = ws[0:len(ws) + flag]
        out.append(" ".join(toks))
    return out


def collect_predictions(vocab, image_ids, seq, subgraph_score, keep_ind, image_of_row, sort=True, remove_bad_endings=None):
    """eval_utils.py:105-134 for a batch: one `entry` dict per image ({'image_id', 'caption', 'subgraph_score', 'sorted_subgraph_ind'})."""
    s_seq, s_score, s_keep, order, img = sort_rows(seq, subgraph_score, keep_ind, image_of_row, sort)
    sents = decode_sequences(vocab, s_seq, remove_bad_endings)
    score_h, keep_h, img_h = s_score.cpu().numpy(), s_keep.cpu().numpy(), img.cpu().numpy()
    entries = []
    for b, image_id in enumerate(image_ids):
        rows = np.nonzero(img_h == b)[0]
        entries.append({"image_id": image_id, "caption": [sents[i] for i in rows], "subgraph_score": score_h[rows],
                        "sorted_subgraph_ind": keep_h[rows]})
    return entries
