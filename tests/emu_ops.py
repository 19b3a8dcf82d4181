"""TEST INFRASTRUCTURE: torch-CPU emulation of subgc.train.CudaOps (same method names and semantics).

It lets tests/test_train_cpu.py run subgc.train.forward / backward — the orchestration and the hand-derived backward
algebra the CUDA path executes — on CPU and compare with the reference's autograd gradients stored in tests/golden.
The product never imports this module; on the GPU box tests/test_gpu_train.py checks every CUDA building block
against these functions and the end-to-end gradients against the same golden files."""
from types import SimpleNamespace

import torch
import torch.nn.functional as F


def _slot(lay, s):
    g = s % lay.per_half
    t = s // lay.per_half
    if lay.order == 0:
        row, half = t % lay.rows, t // lay.rows
    else:
        half, row = t % 2, (t // 2) * lay.seq_per_img
    return row, half, g, row // lay.seq_per_img


class EmuOps:
    def __init__(self, seq_per_img=5):
        self.seq_per_img = seq_per_img

    def layout(self, rows, per_half, order):
        return SimpleNamespace(rows=rows, per_half=per_half, seq_per_img=self.seq_per_img, order=order)

    def linear(self, x, w, b=None, relu=False, gather=None, out=None, accumulate=False):
        xi = x if gather is None else x[gather]
        y = F.linear(xi, w, b)
        if relu:
            y = torch.relu(y)
        if out is None:
            return y
        out.copy_(out + y if accumulate else y)
        return out

    def transpose(self, x):
        return x.t().contiguous()

    def colsum(self, x, out, accumulate=True):
        s = x.sum(0).view(out.shape)
        out.copy_(out + s if accumulate else s)
        return out

    def mul(self, a, b): return a * b
    def add(self, a, b): return a + b
    def relu_bwd(self, y, dy): return torch.where(y > 0, dy, torch.zeros_like(dy))
    def scale(self, a, s): return a * s
    def relu(self, a): return torch.relu(a)
    def sigmoid(self, a): return torch.sigmoid(a)

    def dropout_mask(self, shape, p, seed, offset, device):
        g = torch.Generator().manual_seed(seed * 1000 + offset)
        return (torch.rand(shape, generator=g) >= p).float() / (1 - p)

    def ss_sample(self, prev_logp, labels_col, ss_prob, seed, offset):
        g = torch.Generator().manual_seed(seed * 1000 + offset)
        take = torch.rand(prev_logp.shape[0], generator=g) < ss_prob
        drawn = torch.multinomial(torch.exp(prev_logp), 1, generator=g).view(-1)
        return torch.where(take, drawn, labels_col)

    def gather_rows(self, src, idx, relu=False):
        y = src[idx]
        return torch.relu(y) if relu else y.clone()

    def scatter_add_rows(self, src, idx, dst):
        dst.index_add_(0, idx, src.contiguous())

    def lstm_fwd(self, gates, c_prev):
        i, f, g, o = gates.chunk(4, 1)
        i, f, g, o = torch.sigmoid(i), torch.sigmoid(f), torch.tanh(g), torch.sigmoid(o)
        c = f * c_prev + i * g
        h = o * torch.tanh(c)
        gates.copy_(torch.cat([i, f, g, o], 1))
        return h, c

    def lstm_bwd(self, act, c_prev, c_new, dh, dc):
        i, f, g, o = act.chunk(4, 1)
        tc = torch.tanh(c_new)
        dct = dh * o * (1 - tc * tc) + (0 if dc is None else dc)
        dg = torch.cat([dct * g * i * (1 - i), dct * c_prev * f * (1 - f), dct * i * (1 - g * g), dh * tc * o * (1 - o)], 1)
        return dg, dct * f

    def att_fwd(self, atth, p_att, att, masks, aw, ab):
        e = (torch.tanh(p_att + atth.unsqueeze(1)) * aw.view(1, 1, -1)).sum(-1) + ab
        sm = torch.softmax(e, 1)
        w = sm * masks
        alpha = w / w.sum(1, keepdim=True)
        return torch.bmm(alpha.unsqueeze(1), att).squeeze(1), alpha, sm

    def att_bwd(self, atth, p_att, att, masks, aw, alpha, sm, dctx, d_att, d_p_att):
        d_alpha = torch.bmm(att, dctx.unsqueeze(2)).squeeze(2)
        z = (sm * masks).sum(1, keepdim=True)
        d_w = (d_alpha - (alpha * d_alpha).sum(1, keepdim=True)) / z
        d_s = d_w * masks
        d_e = sm * (d_s - (sm * d_s).sum(1, keepdim=True))
        d_att += alpha.unsqueeze(2) * dctx.unsqueeze(1)
        u = torch.tanh(p_att + atth.unsqueeze(1))
        d_pre = d_e.unsqueeze(2) * aw.view(1, 1, -1) * (1 - u * u)
        d_p_att += d_pre
        return d_pre.sum(1), (d_e.unsqueeze(2) * u).sum(1)

    def log_softmax_fwd(self, logits, out_view):
        out_view.copy_(F.log_softmax(logits, 1))

    def log_softmax_bwd(self, logp, dlogp):
        return dlogp - torch.exp(logp) * dlogp.sum(1, keepdim=True)

    def class_argmax(self, dist2d, skip_first):
        return dist2d[:, 1:].max(1)[1] + 1 if skip_first else dist2d.max(1)[1]

    def pool(self, lay, n_sub, x_obj, obj_ind, att_masks):
        L = x_obj.shape[2]
        N = obj_ind.shape[-1]
        read = torch.zeros(n_sub, 2 * L)
        sub_len = torch.zeros(n_sub, dtype=torch.int32)
        for s in range(n_sub):
            row, half, g, img = _slot(lay, s)
            ln = int((att_masks[row, half, g] != 0).sum())
            f = x_obj[img][obj_ind[row, half, g, :ln]]
            mx = f.max(0)[0]
            if ln < N:
                mx = torch.clamp(mx, min=0)
            read[s] = torch.cat([mx, f.sum(0) / ln])
            sub_len[s] = ln
        return read, sub_len

    def pool_bwd(self, lay, x_obj, obj_ind, sub_len, d_read, d_x_obj):
        L = x_obj.shape[2]
        N = obj_ind.shape[-1]
        for s in range(d_read.shape[0]):
            row, half, g, img = _slot(lay, s)
            ln = int(sub_len[s])
            ids = obj_ind[row, half, g, :ln]
            f = x_obj[img][ids]
            best, bn = f.max(0)
            gmax = d_read[s, :L].clone()
            if ln < N:
                gmax = torch.where(best < 0, torch.zeros_like(gmax), gmax)
            grad = (d_read[s, L:] / ln).unsqueeze(0).repeat(ln, 1)
            grad[bn, torch.arange(L)] += gmax
            d_x_obj[img].index_add_(0, ids, grad)

    def _labels(self, lay, n):
        half = torch.tensor([_slot(lay, s)[1] for s in range(n)])
        return (half == 0).float()

    def bce(self, lay, score):
        return F.binary_cross_entropy(score, self._labels(lay, score.numel())).view(1)

    def bce_bwd(self, lay, score, scale):
        return (score - self._labels(lay, score.numel())) * scale

    def select_train(self, lay, score, sub_len):
        pos = score[:lay.rows * lay.per_half].view(lay.rows, lay.per_half)
        pick = pos.argmax(1)
        sel = (torch.arange(lay.rows) * lay.per_half + pick).int()
        return sel, int(sub_len[sel.long()].max())

    def prepare_index(self, lay, sel, len_max, obj_ind, att_masks):
        N = obj_ind.shape[-1]
        node_row, masks, row_len = [], [], []
        for s in sel.tolist():
            row, half, g, img = _slot(lay, s)
            node_row.append(img * N + obj_ind[row, half, g, :len_max])
            masks.append(att_masks[row, half, g, :len_max])
            row_len.append(int((att_masks[row, half, g] != 0).sum()))
        return torch.cat(node_row), torch.stack(masks), torch.tensor(row_len, dtype=torch.int32)

    def fuse_nodes(self, P, att_feats, obj_dist):
        B, N, _ = att_feats.shape
        cls = self.class_argmax(obj_dist.reshape(B * N, -1), 1)
        e = F.linear(P["sg_obj_embed.weight"][cls], P["obj_emb_proj.weight"], P["obj_emb_proj.bias"]).view(B, N, -1)
        return torch.relu(F.linear(att_feats, P["obj_v_proj.weight"], P["obj_v_proj.bias"]) + e)

    def gcn_edge_fwd(self, m2, m3, rel, res):
        B = m2.shape[0]
        bi = torch.arange(B).view(B, 1)
        d = torch.tensor(1.0) + torch.tensor(1e-7)
        out = 0.5 * (torch.relu(m2[bi, rel[:, :, 0]] / d) + torch.relu(m3[bi, rel[:, :, 1]] / d))
        return out if res is None else out + res

    def _counts(self, rel, N, col):
        B, K, _ = rel.shape
        cnt = torch.zeros(B, N)
        cnt.scatter_add_(1, rel[:, :, col], torch.ones(B, K))
        return cnt

    def gcn_node_fwd(self, m0, m1, rel, res, N):
        B, K, L = m0.shape
        s0, s1 = torch.zeros(B, N, L), torch.zeros(B, N, L)
        for b in range(B):
            s0[b].index_add_(0, rel[b, :, 0], m0[b])
            s1[b].index_add_(0, rel[b, :, 1], m1[b])
        y0 = torch.relu(s0 / (self._counts(rel, N, 0) + 1e-7).unsqueeze(2))
        y1 = torch.relu(s1 / (self._counts(rel, N, 1) + 1e-7).unsqueeze(2))
        out = (y0 + y1) * 0.5
        return (out if res is None else out + res), y0, y1

    def gcn_node_bwd(self, dx, y0, y1, rel):
        B, N, L = dx.shape
        bi = torch.arange(B).view(B, 1)
        c0 = (self._counts(rel, N, 0) + 1e-7).unsqueeze(2)
        c1 = (self._counts(rel, N, 1) + 1e-7).unsqueeze(2)
        g0 = torch.where(y0 > 0, 0.5 * dx / c0, torch.zeros_like(dx))
        g1 = torch.where(y1 > 0, 0.5 * dx / c1, torch.zeros_like(dx))
        return g0[bi, rel[:, :, 0]], g1[bi, rel[:, :, 1]]

    def gcn_edge_bwd(self, dp, m2, m3, rel):
        B, N, L = m2.shape
        a0, a1 = torch.zeros(B, N, L), torch.zeros(B, N, L)
        for b in range(B):
            a0[b].index_add_(0, rel[b, :, 0], dp[b])
            a1[b].index_add_(0, rel[b, :, 1], dp[b])
        d = torch.tensor(1.0) + torch.tensor(1e-7)
        return torch.where(m2 > 0, 0.5 * a0 / d, torch.zeros_like(a0)), torch.where(m3 > 0, 0.5 * a1 / d, torch.zeros_like(a1))
