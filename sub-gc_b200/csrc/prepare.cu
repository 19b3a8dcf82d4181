// Decoder feature preparation for the selected sub-graphs.
//
//   subgc_prepare_forward <- gpn_layer.read_out_proj (reference models/lib/gpn.py:35-36,79,95),
//                            AttModel.clip_att / _prepare_feature (reference models/AttModel.py:348-368) and
//                            pack_wrapper (reference models/AttModel.py:16-36).
//
// The reference gathers [S,37,1024] node rows, clips to the longest sub-graph, sorts / packs the rows to run
// att_embed on valid rows only and pads back with zeros.  Here the node gather is folded into the GEMM's A-row
// index (no gathered copy), and the epilogue writes exact zeros for rows beyond a sub-graph's length.
#include "common.cuh"

namespace subgc {

// row index (into x_obj viewed as [B*N, L]) of node n of selected sub-graph r, and the clipped mask copy
__global__ void __launch_bounds__(256) prepare_index_kernel(const subgc_subgraph_layout lay, const int* __restrict__ sel, int n_rows, int len_max,
                                                            int N, const long long* __restrict__ obj_ind, const float* __restrict__ att_masks,
                                                            long long* __restrict__ node_row, float* __restrict__ masks,
                                                            int* __restrict__ row_len) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * len_max) return;
    int r = idx / len_max, n = idx - r * len_max;
    int image;
    int slot = subgraph_slot(lay, sel[r], &image);
    node_row[idx] = (long long)image * N + obj_ind[(size_t)slot * N + n];
    masks[idx] = att_masks[(size_t)slot * N + n];
    if (n == 0) {  // length = number of ones in the full mask row (AttModel.py:351, pack_wrapper :33)
        int len = 0;
        for (int j = 0; j < N; ++j) len += (att_masks[(size_t)slot * N + j] != 0.f) ? 1 : 0;
        row_len[r] = len;
    }
}

}  // namespace subgc

using namespace subgc;

static size_t max2(size_t a, size_t b) { return a > b ? a : b; }

extern "C" size_t subgc_prepare_workspace_bytes(const subgc_dims* d, int n_rows, int len_max) {
    if (!d || n_rows <= 0 || len_max <= 0) return 0;
    size_t rows = (size_t)n_rows * len_max;
    size_t b = align_up(rows * 8, 256) + align_up((size_t)n_rows * 4, 256);  // node_row, row_len
    b += align_up((size_t)n_rows * d->att_hid * 4, 256);                   // read_out_proj hidden
    b += align_up((size_t)n_rows * d->fc_feat * 4, 256);                   // fc_embed hidden
    b += 2 * align_up((size_t)n_rows * d->fc_feat * 2, 256) + 2 * align_up(rows * d->rnn * 2, 256) + 1024;   // split-fp16 copies of fc hidden / att
    size_t g = gemm_workspace_bytes(n_rows, d->att_hid, 2 * d->gcn);
    g = max2(g, gemm_workspace_bytes(n_rows, d->fc_feat, 2 * d->gcn));
    g = max2(g, gemm_workspace_bytes(n_rows, 2 * d->gcn, d->att_hid));
    g = max2(g, gemm_workspace_bytes(n_rows, d->fc_feat, d->att_feat));
    g = max2(g, gemm_workspace_bytes(n_rows, d->rnn, d->fc_feat));
    g = max2(g, gemm_workspace_bytes((int)rows, d->rnn, d->gcn));
    g = max2(g, gemm_workspace_bytes((int)rows, d->att_hid, d->rnn));
    return b + align_up(g, 256) + 1024;
}

extern "C" int subgc_prepare_forward(const subgc_dims* d, const subgc_weights* w, const subgc_subgraph_layout* lay, int n_rows, int len_max,
                                     const int32_t* sel, const float* x_obj, const int64_t* gpn_obj_ind, const float* att_masks,
                                     const float* read_out, float* g_fc, float* fc, float* att, float* p_att, float* masks, void* ws_,
                                     size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && lay && sel && x_obj && gpn_obj_ind && att_masks && read_out && fc && att && p_att && masks,
                    "subgc_prepare_forward: null argument");
    SUBGC_CHECK_ARG(g_fc || w->prep_fold.w, "subgc_prepare_forward: g_fc may only be NULL when w->prep_fold is set");
    SUBGC_CHECK_ARG(n_rows > 0 && len_max > 0 && len_max <= d->obj_num, "subgc_prepare_forward: bad n_rows/len_max (%d, %d)", n_rows, len_max);
    SUBGC_CHECK_ARG(d->att_feat == 2 * d->gcn, "subgc_prepare_forward: fc_embed input (att_feat_size) must equal 2*gcn_dim");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int L = d->gcn, AH = d->att_hid, H = d->rnn, FC = d->fc_feat, N = d->obj_num;
    const int rows = n_rows * len_max;
    Workspace ws(ws_, ws_bytes);
    long long* node_row = ws.take<long long>(rows);
    int* row_len = ws.take<int>(n_rows);
    float* hid = ws.take<float>((size_t)n_rows * AH);
    float* fch = ws.take<float>((size_t)n_rows * FC);
    if (!ws.ok()) { set_error("subgc_prepare_forward: workspace too small"); return SUBGC_E_WORKSPACE; }
    prepare_index_kernel<<<(rows + 255) / 256, 256, 0, st>>>(*lay, sel, n_rows, len_max, N, reinterpret_cast<const long long*>(gpn_obj_ind),
                                                             att_masks, node_row, masks, row_len);
    SUBGC_LAUNCH_CHECK();
    GemmProblem p;
    p.wts = w;
    // split-fp16 copies handed from a contraction to the next one (written by the producer's epilogue / split-K reduction)
    const bool c16 = w->packs != nullptr && w->n_packs > 0 && (FC & 7) == 0 && (H & 7) == 0;
    unsigned short* f16_hi = c16 ? ws.take<unsigned short>((size_t)n_rows * FC) : nullptr;
    unsigned short* f16_lo = c16 ? ws.take<unsigned short>((size_t)n_rows * FC) : nullptr;
    unsigned short* a16_hi = c16 ? ws.take<unsigned short>((size_t)rows * H) : nullptr;
    unsigned short* a16_lo = c16 ? ws.take<unsigned short>((size_t)rows * H) : nullptr;
    if (!ws.ok()) { set_error("subgc_prepare_forward: workspace too small"); return SUBGC_E_WORKSPACE; }
    if (!g_fc) {
        // read_out_proj.0 / .1 and fc_embed.0 have nothing between them: one folded contraction, then ReLU (w->prep_fold)
        p = GemmProblem(); p.wts = w;
        p.M = n_rows; p.N = FC; p.nseg = 1;
        p.seg[0] = make_seg(read_out, 2 * L, w->prep_fold.w, 2 * L, 2 * L);
        p.seg[0].gather32 = sel;
        p.epi.bias = w->prep_fold.b; p.epi.relu = 1;
        p.epi.c16_hi = f16_hi; p.epi.c16_lo = f16_lo; p.epi.ld16 = FC;
        p.C = fch; p.ldc = FC;
        SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    } else {
    // read_out_proj: 2L -> AH -> 2L (no activation)
    p = GemmProblem(); p.wts = w;
    p.M = n_rows; p.N = AH; p.nseg = 1;
    p.seg[0] = make_seg(read_out, 2 * L, w->read_out0.w, 2 * L, 2 * L);
    p.seg[0].gather32 = sel;
    p.epi.bias = w->read_out0.b;
    p.C = hid; p.ldc = AH;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    p = GemmProblem(); p.wts = w;
    p.M = n_rows; p.N = 2 * L; p.nseg = 1;
    p.seg[0] = make_seg(hid, AH, w->read_out1.w, AH, AH);
    p.epi.bias = w->read_out1.b;
    p.C = g_fc; p.ldc = 2 * L;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    // fc_embed: relu(W2 relu(W1 g_fc + b1) + b2)
    p = GemmProblem(); p.wts = w;
    p.M = n_rows; p.N = FC; p.nseg = 1;
    p.seg[0] = make_seg(g_fc, 2 * L, w->fc_embed0.w, d->att_feat, d->att_feat);
    p.epi.bias = w->fc_embed0.b; p.epi.relu = 1;
    p.epi.c16_hi = f16_hi; p.epi.c16_lo = f16_lo; p.epi.ld16 = FC;
    p.C = fch; p.ldc = FC;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    }
    p = GemmProblem(); p.wts = w;
    p.M = n_rows; p.N = H; p.nseg = 1;
    p.seg[0] = make_seg(fch, FC, w->fc_embed2.w, FC, FC);
    if (f16_hi) { p.seg[0].A16_hi = f16_hi; p.seg[0].A16_lo = f16_lo; p.seg[0].lda16 = FC; }
    p.epi.bias = w->fc_embed2.b; p.epi.relu = 1;
    p.C = fc; p.ldc = H;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    // att_embed on valid rows (pads exactly zero), straight from x_obj through the node index
    p = GemmProblem(); p.wts = w;
    p.M = rows; p.N = H; p.nseg = 1;
    p.seg[0] = make_seg(x_obj, L, w->att_embed.w, L, L);
    p.seg[0].gather = node_row;
    p.epi.bias = w->att_embed.b; p.epi.relu = 1;
    p.epi.group = len_max; p.epi.group_len = row_len;
    p.epi.c16_hi = a16_hi; p.epi.c16_lo = a16_lo; p.epi.ld16 = H;
    p.C = att; p.ldc = H;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    // ctx2att on every row up to len_max (padded rows give the bias)
    p = GemmProblem(); p.wts = w;
    p.M = rows; p.N = AH; p.nseg = 1;
    p.seg[0] = make_seg(att, H, w->ctx2att.w, H, H);
    if (a16_hi) { p.seg[0].A16_hi = a16_hi; p.seg[0].A16_lo = a16_lo; p.seg[0].lda16 = H; }
    p.epi.bias = w->ctx2att.b;
    p.C = p_att; p.ldc = AH;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    return SUBGC_OK;
}

extern "C" int subgc_prepare_index(const subgc_dims* d, const subgc_subgraph_layout* lay, int n_rows, int len_max, const int32_t* sel,
                                   const int64_t* gpn_obj_ind, const float* att_masks, int64_t* node_row, float* masks, int32_t* row_len,
                                   subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && lay && sel && gpn_obj_ind && att_masks && node_row && masks && row_len && n_rows > 0 && len_max > 0,
                    "subgc_prepare_index: bad arguments");
    const int rows = n_rows * len_max;
    prepare_index_kernel<<<(rows + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        *lay, sel, n_rows, len_max, d->obj_num, reinterpret_cast<const long long*>(gpn_obj_ind), att_masks, reinterpret_cast<long long*>(node_row),
        masks, row_len);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}
