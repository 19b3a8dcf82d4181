/*
 * subgc_b200 — C ABI of the B200-native Sub-GC hot path (sm_100a).
 *
 * The reference (YiwuZhong/Sub-GC) is pure Python/PyTorch and has no FFI; the boundary it exposes is the
 * nn.Module surface `models.setup(opt)` -> `AttModel.forward(mode=...)` (reference models/__init__.py:43-59,
 * models/CaptionModel.py:21-26) and `LossWrapper` (models/loss_wrapper.py:7-27).  This header is the native layer a
 * replacement of that path binds to (ctypes in `sub-gc_b200/subgc/_lib.py`; SURVEY.md §8b lists the contract):
 *
 *   - every entry point is `extern "C"`, returns int (0 = ok, non-zero = SUBGC_E_*; text via subgc_last_error()),
 *   - takes raw DEVICE pointers + explicit sizes + a cudaStream_t, never allocates, frees or synchronises,
 *   - keeps no global mutable state (thread-local error string only), so DataParallel-style threads and
 *     one-process-per-GPU both work, and every call can be captured into a CUDA graph by the caller,
 *   - scratch memory is caller-provided; size it with the matching *_workspace_bytes() query.
 *
 * All floating-point tensors are fp32 row-major; index tensors are int64 exactly as the reference loaders emit
 * them (dataloaders/dataloader.py:194-206, dataloaders/dataloader_test.py:191-203).
 * Each function names the reference code it replaces (file:line relative to the reference root).
 */
#ifndef SUBGC_B200_H
#define SUBGC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUBGC_ABI_VERSION 4
#define SUBGC_MAX_GCN_LAYERS 8

typedef void* subgc_stream_t; /* cudaStream_t */

enum {
    SUBGC_OK = 0,
    SUBGC_E_INVALID = 1,   /* bad argument (null pointer, size, unsupported option) */
    SUBGC_E_WORKSPACE = 2, /* workspace too small */
    SUBGC_E_CUDA = 3,      /* a CUDA runtime call / launch failed */
    SUBGC_E_UNSUPPORTED = 4
};

/* Dimensions: the `opt` attributes AttModel.__init__ reads (models/AttModel.py:44-69). */
typedef struct subgc_dims {
    int32_t vocab1;       /* vocab_size + 1 (logit rows)            */
    int32_t enc;          /* input_encoding_size X                  */
    int32_t rnn;          /* rnn_size H                             */
    int32_t att_hid;      /* att_hid_size AH (also sGPN hidden)     */
    int32_t fc_feat;      /* fc_feat_size                           */
    int32_t att_feat;     /* att_feat_size A (== 2*gcn)             */
    int32_t gcn;          /* gcn_dim L                              */
    int32_t low_rank;     /* GCN unit low-rank width R (512)        */
    int32_t embed;        /* embed_dim E                            */
    int32_t obj_classes;  /* C (1599)                               */
    int32_t pred_classes; /* P (21)                                 */
    int32_t gcn_layers;
    int32_t gcn_residual;
    int32_t pred_emb_type;
    int32_t seq_length;   /* T                                      */
    int32_t obj_num;      /* N (37)                                 */
    int32_t rel_num;      /* K (65)                                 */
} subgc_dims;

typedef struct subgc_linear {
    const float* w; /* [out, in] as nn.Linear stores it */
    const float* b; /* [out] */
} subgc_linear;

/* Split-fp16 copy of one weight matrix for the tensor cores (subgc_pack_weight): w[r,c] = hi + lo * 2^-11 with hi, lo IEEE
 * fp16, i.e. the same 4 bytes per weight as the fp32 tensor it mirrors and 22 of its 24 mantissa bits.  Layout: k-block-major
 * [kb][rows][32] -- for a block of 32 consecutive columns all rows are contiguous (64 bytes each), so the [tile rows x 32] tile a
 * CTA streams per step is ONE contiguous run in HBM (DRAM page locality; a row-major layout gives 64-byte pieces 2-12 KB apart,
 * measured at < 50 % of the HBM rate).  The columns may be cut into K segments (the un-concatenated LSTM inputs
 * [h_lang | fc | x_t]): every segment starts on a k-block boundary and is zero-padded to a multiple of 32 columns.
 * The fp32 tensor stays the source of truth (and the lookup key): contractions whose weight operand is exactly one segment of
 * `w` read the packed copy instead; anything else takes the fp32 (split-TF32) path. */
#define SUBGC_PACK_MAX_SEG 4
typedef struct subgc_packed {
    const float* w;      /* the fp32 [rows, cols] contiguous tensor this was packed from */
    const uint16_t* hi;  /* [kb_total][rows][32] fp16 */
    const uint16_t* lo;  /* same shape, scaled by 2^11 */
    int32_t rows, cols;
    int32_t n_seg;                           /* 1..SUBGC_PACK_MAX_SEG */
    int32_t seg_col[SUBGC_PACK_MAX_SEG + 1]; /* 0 = seg_col[0] < ... < seg_col[n_seg] = cols */
} subgc_packed;

/* Device pointers to the reference state_dict tensors, un-repacked (key names in comments). */
typedef struct subgc_weights {
    subgc_linear obj_v_proj;                           /* obj_v_proj.{weight[L,A],bias}                         */
    const float* sg_obj_embed;                         /* sg_obj_embed.weight [C,E]                             */
    subgc_linear obj_emb_proj;                         /* obj_emb_proj [L,E]                                    */
    const float* sg_pred_embed;                        /* sg_pred_embed.weight [P,E]                            */
    subgc_linear pred_emb_prj;                         /* pred_emb_prj [L,E]                                    */
    subgc_linear gcn_lft[SUBGC_MAX_GCN_LAYERS][4];     /* gcn_backbone.gcn.l.gcn_collect.collect_units.u.fc_lft [R,L] */
    subgc_linear gcn_rgt[SUBGC_MAX_GCN_LAYERS][4];     /* ...fc_rgt [L,R]                                       */
    subgc_linear gpn_fc0;                              /* gpn_layer.gpn_fc.0 [AH,2L]                            */
    subgc_linear gpn_fc3;                              /* gpn_layer.gpn_fc.3 [1,AH]                             */
    subgc_linear read_out0;                            /* gpn_layer.read_out_proj.0 [AH,2L]                     */
    subgc_linear read_out1;                            /* gpn_layer.read_out_proj.1 [2L,AH]                     */
    subgc_linear logit;                                /* logit [V1,H]                                          */
    const float* embed;                                /* embed.0.weight [V1,X]                                 */
    subgc_linear fc_embed0;                            /* fc_embed.0 [FC,A]                                     */
    subgc_linear fc_embed2;                            /* fc_embed.2 [H,FC]                                     */
    subgc_linear att_embed;                            /* att_embed.0 [H,L]                                     */
    subgc_linear ctx2att;                              /* ctx2att [AH,H]                                        */
    subgc_linear h2att;                                /* core.attention.h2att [AH,H]                           */
    subgc_linear alpha_net;                            /* core.attention.alpha_net [1,AH]                       */
    const float* att_w_ih; const float* att_w_hh;      /* core.att_lstm.weight_ih [4H,X+2H], weight_hh [4H,H]   */
    const float* att_b_ih; const float* att_b_hh;
    const float* lang_w_ih; const float* lang_w_hh;    /* core.lang_lstm.weight_ih [4H,2H], weight_hh [4H,H]    */
    const float* lang_b_ih; const float* lang_b_hh;
    const subgc_packed* packs;                         /* HOST array of packed copies (nullable)                 */
    int32_t n_packs;
    int32_t* h3_overflow;                              /* DEVICE flag (nullable): OR-ed with 1 when an activation fed to the
                                                          split-fp16 path did not fit fp16 (|x| > 65504, saturated): the
                                                          results of that call are invalid, re-run without packs          */
    const float* lang_early_w;                         /* optional derived tensor [AH+4H, 2H] (fp32, contiguous):
                                                          rows 0..AH   = [ core.attention.h2att.weight | 0 ],
                                                          rows AH..    = [ core.lang_lstm.weight_ih[:, H:2H] | core.lang_lstm.weight_hh ].
                                                          When set, a decode step contracts everything that depends only on
                                                          (h_att(t), h_lang(t-1)) in ONE launch before the attention (h2att + two
                                                          thirds of the language-LSTM gates) and only the ctx segment after it. */
    const void* mega;                                  /* DEVICE (nullable): schedule tables + "stream pack" of the decoder weights for the
                                                          persistent decode kernel, built by subgc_mega_pack for `mega_ctas` CTAs.  When
                                                          set, subgc_decode_sample runs its whole loop as ONE cooperative launch
                                                          (<= 128 rows, no attention-weight output); otherwise one launch per stage.   */
    uint64_t mega_bytes;
    int32_t mega_ctas;                                 /* CTAs the schedule was built for (= SMs of the device, one CTA each)           */
    /* Optional derived tensors (inference): a _Collection_Unit applies fc_rgt(fc_lft(.)) with nothing in between
     * (models/lib/graph_conv_unit.py:29-31), so the two units of a direction fold into ONE weight
     *   gcn_fold[l][dir].w = 2^s [W_rgt(u) W_lft(u) ; W_rgt(u') W_lft(u')]  [2L, L],  .b = 2^s [W_rgt b_lft + b_rgt ; ...]  [2L]
     * with (u, u') = (0, 1) for dir 0 (node <- edges) and (2, 3) for dir 1 (edge <- nodes); gcn_fold_scale = 2^s (> 0 when set; a
     * power of two that keeps the folded products of small weights inside the normal fp16 range of the split-fp16 copy, undone
     * exactly when the messages are consumed).  The same FLOPs as the low-rank pair, one contraction instead of four per direction. */
    subgc_linear gcn_fold[SUBGC_MAX_GCN_LAYERS][2];
    float gcn_fold_scale[SUBGC_MAX_GCN_LAYERS][2];
    /* read_out_proj.0 -> read_out_proj.1 -> fc_embed.0 are three Linears in a row (models/lib/gpn.py:35-36, models/AttModel.py:109):
     *   prep_fold.w = W_fc0 W_ro1 W_ro0  [FC, 2L],  prep_fold.b = W_fc0 (W_ro1 b_ro0 + b_ro1) + b_fc0  [FC]
     * used by subgc_prepare_forward when the caller does not ask for g_fc (the reference's intermediate `fc_feats`). */
    subgc_linear prep_fold;
} subgc_weights;

/* How sub-graph s of a flat list maps onto the loader tensors gpn_obj_ind / att_masks [rows,2,per_half,N].
 *   order 0 (training, models/lib/gpn.py:157-170): s = (half*rows + row)*per_half + g      (positives first)
 *   order 1 (inference, models/lib/gpn.py:86-94):  s = (image*2 + half)*per_half + g, row = image*seq_per_img
 * image(s) = row / seq_per_img.  n_sub = 2*rows*per_half (order 0) or 2*(rows/seq_per_img)*per_half (order 1). */
typedef struct subgc_subgraph_layout {
    int32_t rows;        /* leading dim of the loader tensors (5B)  */
    int32_t per_half;    /* G (training) or M (inference)           */
    int32_t seq_per_img; /* 5                                       */
    int32_t order;       /* 0 training, 1 inference                 */
} subgc_subgraph_layout;

const char* subgc_last_error(void);
int subgc_version(void);
/* kernels launched so far by this process through this library, all threads (instrumentation for bench.py's gpu_launches) */
unsigned long long subgc_launch_count(void);
/* debugging aid (env SUBGC_ATT_TRACE=1, synchronises): per-block stage time stamps [n_blocks][8] of the last fused att-phase launch */
int subgc_debug_att_trace(unsigned long long* host_out, int n_blocks);
/* debugging aid (env SUBGC_TRACE=1, synchronises): globaltimer stamps of the decode-loop launches, per launch
 * [first block start, last block end, first block past its dependency wait, last block past it, 4 kernel-specific marks].
 * op 0 restart slot numbering, op 1 reset stamps, op 2 copy out up to n slots (stamps [n][8], kernel ids [n]); returns slots used */
int subgc_debug_trace(int op, unsigned long long* stamps, int* ids, int n);
/* debugging aid (env SUBGC_MEGA_TRACE=1, synchronises): [n_cta][n_steps][48] globaltimer stamps of the last persistent decode launch */
int subgc_debug_mega_trace(unsigned long long* host_out, int n_cta, int n_steps);

/* Packs an fp32 weight matrix [rows, cols] (leading dim ldw) into the split-fp16 form of subgc_packed, columns cut at
 * seg_col[0..n_seg].  Each output array holds subgc_pack_elems(rows, n_seg, seg_col) fp16 values.  `overflow` (device,
 * nullable) is OR-ed with 1 when a weight does not fit fp16 (|w| > 65504): such a tensor must not be registered.
 * Replaces nothing in the reference: nn.Linear / nn.LSTMCell keep fp32 weights (models/AttModel.py:393-398). */
size_t subgc_pack_elems(int rows, int n_seg, const int32_t* seg_col);
int subgc_pack_weight(int rows, int cols, const float* w, int ldw, int n_seg, const int32_t* seg_col, uint16_t* hi,
                      uint16_t* lo, int32_t* overflow, subgc_stream_t stream);

/* Persistent decode kernel (csrc/mega_decode.cu): the greedy / top-k loop of AttModel._sample (models/AttModel.py:278-326) as one
 * cooperative launch with one CTA per SM.  Every CTA streams a fixed slice of the four per-step weight matrices
 * (core.att_lstm, core.attention.h2att, core.lang_lstm, logit: models/AttModel.py:393-398,438-443,87) from its own contiguous
 * region of the "stream pack": split-fp16 tiles (as subgc_packed: v = hi + lo * 2^-11) already in the shared-memory layout of the
 * tensor core, in the order the CTA contracts them.  subgc_mega_pack_bytes: size of tables + pack for these dims and `n_cta` CTAs
 * (0: dims not supported by the kernel).  subgc_mega_pack: builds both into `buf` (device, 1024-byte aligned) from the fp32
 * parameters in `w`; `overflow` (device, nullable) is OR-ed with 1 when a weight does not fit fp16.  Re-pack after the parameters
 * change.  Not capturable (copies the tables from host memory). */
size_t subgc_mega_pack_bytes(const subgc_dims* d, int n_cta);
/* Bench / profiling: with `on`, every eager (not graph-captured) launch of the persistent kernel by this thread is bracketed by CUDA events
 * on its stream; last_ms (nullable) receives the duration of the most recent one (blocks until it finished; -1 if none). */
int subgc_mega_timing(int on, float* last_ms);
int subgc_mega_pack(const subgc_dims* d, const subgc_weights* w, int n_cta, void* buf, size_t bytes, int32_t* overflow,
                    subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Building block: C[M,N] = act((A[gather] . W^T + bias + addend) / div), the nn.Linear contraction every stage
 * of the path is made of (exported for unit tests; fp32 FMA accumulation).
 * ------------------------------------------------------------------------------------------------------- */
size_t subgc_linear_workspace_bytes(int M, int N, int K);
int subgc_linear_forward(int M, int N, int K, const float* A, int lda, const int64_t* a_gather /*nullable*/,
                         const float* W, int ldw, const float* bias /*nullable*/, int relu, float* C, int ldc,
                         void* ws, size_t ws_bytes, subgc_stream_t stream);
/* Same contraction with the weight given as a packed copy (pk->w must be the [N,K] tensor itself): split-fp16 tensor-core path. */
int subgc_linear_packed_forward(int M, int N, int K, const float* A, int lda, const int64_t* a_gather /*nullable*/,
                                const subgc_packed* pk, const float* bias /*nullable*/, int relu, float* C, int ldc,
                                void* ws, size_t ws_bytes, subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Encoder.
 * subgc_fuse_nodes  replaces AttModel.feat_fusion (models/AttModel.py:370-387):
 *    x0[b,n] = relu(W_v att[b,n] + b_v + W_e E_obj[1+argmax(obj_dist[b,n,1:])] + b_e)
 *    p0[b,k] = W_p E_pred[cls(pred_dist[b,k])] + b_p          (skipped when p0 == NULL)
 * subgc_gcn_forward replaces gcn_backbone.forward/make_map (models/lib/gcn_backbone.py:29-67),
 *    _GraphConvolutionLayer.forward (models/lib/graph_conv.py:15-34) and _Collection_Unit.forward
 *    (models/lib/graph_conv_unit.py:28-36) with an edge-list gather / segment-mean instead of the dense
 *    adjacency bmm.  Outputs are NOT tiled x5 (gcn_backbone.py:50-51): consumers index image = row/5.
 *    x_pred == NULL skips every unit whose result cannot reach x_obj (for Sub-GC: half of them).
 * ------------------------------------------------------------------------------------------------------- */
size_t subgc_encoder_workspace_bytes(const subgc_dims* d, int n_images);
int subgc_fuse_nodes(const subgc_dims* d, const subgc_weights* w, int n_images, const float* att_feats,
                     const float* obj_dist, const float* pred_dist, float* x0, float* p0 /*nullable*/,
                     void* ws, size_t ws_bytes, subgc_stream_t stream);
/* Same fusion with the class ids already known (loader-side input compaction): obj_cls [n_images, N] = 1 + argmax_first(
 * obj_dist[:, :, 1:]) and pred_cls [n_images, K] (nullable with p0) as int64, i.e. what AttModel.py:374,382-385 derives from the score
 * tensors the loaders ship (dataloaders/dataloader_test.py:234-273) -- 30 MB of fp32 scores per 128 images become 9.5 KB of ids. */
int subgc_fuse_nodes_cls(const subgc_dims* d, const subgc_weights* w, int n_images, const float* att_feats,
                         const int64_t* obj_cls, const int64_t* pred_cls /*nullable*/, float* x0, float* p0 /*nullable*/,
                         void* ws, size_t ws_bytes, subgc_stream_t stream);
int subgc_gcn_forward(const subgc_dims* d, const subgc_weights* w, int n_images, const float* x0,
                      const float* p0 /*nullable iff not needed*/, const int64_t* rel_ind, float* x_obj,
                      float* x_pred /*nullable*/, void* ws, size_t ws_bytes, subgc_stream_t stream);
/* 1 if p0 (the predicate embedding) influences x_obj / x_pred for this configuration, else 0. */
int subgc_gcn_needs_pred(const subgc_dims* d, int want_x_pred);

/* ---------------------------------------------------------------------------------------------------------
 * sGPN.  subgc_sgpn_forward replaces gpn_layer.extract_subgraph_feats + graph_pooling + gpn_fc + sigmoid
 * (models/lib/gpn.py:50-57,152-185) for the n_sub sub-graphs described by `lay`:
 *    read_out[s] = [max_n F ‖ sum_n F / len],  F = rows of x_obj[image(s)] listed in gpn_obj_ind, zero beyond len
 *    score[s]    = sigmoid(w2 . relu(W1 read_out[s] + b1) + b2)          (eval-mode: no dropout)
 *    sub_len[s]  = number of ones in att_masks (the loader's pooling matrix is diag(first len ones))
 *    bce_loss    = mean BCE(score, 1 for the first half of the list, 0 for the second)   (nullable)
 * ------------------------------------------------------------------------------------------------------- */
size_t subgc_sgpn_workspace_bytes(const subgc_dims* d, int n_sub);
int subgc_sgpn_forward(const subgc_dims* d, const subgc_weights* w, const subgc_subgraph_layout* lay,
                       const float* x_obj, const int64_t* gpn_obj_ind, const float* att_masks, float* read_out,
                       float* score, int32_t* sub_len, float* bce_loss /*nullable*/, void* ws, size_t ws_bytes,
                       subgc_stream_t stream);

/* Training-mode selection (models/lib/gpn.py:64-81): for each sentence row pick argmax_first over its per_half
 * positive scores.  sel[row] = flat sub-graph id; stats[0] = n rows, stats[1] = max length of the picks. */
int subgc_sgpn_select_train(const subgc_subgraph_layout* lay, const float* score, const int32_t* sub_len,
                            int32_t* sel, int32_t* stats, subgc_stream_t stream);

/* Inference-mode NMS, per image (models/lib/gpn.py:108-150): greedy suppression in descending score order
 * (ties: higher index first) of sub-graphs whose node-set IoU with a kept one exceeds iou_thres (compared in
 * double like the reference), at most max_subgraphs survivors per image, emitted in ascending original index.
 * use_nms == 0 keeps everything (gpn.py:97).  Outputs: sel [n_sub] flat ids of kept sub-graphs, images in order,
 * compacted; keep_ind [n_sub] the per-image index of each kept one; stats[0] = total kept, stats[1] = max length
 * among kept, stats[2 + i] = kept count of image i. */
size_t subgc_nms_workspace_bytes(int n_images, int per_image);
int subgc_subgraph_nms(const subgc_dims* d, const subgc_subgraph_layout* lay, const float* score,
                       const int32_t* sub_len, const int64_t* gpn_obj_ind, const float* att_masks, int use_nms,
                       double iou_thres, int max_subgraphs, int32_t* sel, int64_t* keep_ind, int32_t* stats,
                       void* ws, size_t ws_bytes, subgc_stream_t stream);

/* Post-decode ordering (misc/eval_utils.py:105-110: `torch.sort(subgraph_score, descending=True)` per image, then seq / keep_ind are
 * indexed with it).  Rows of an image are contiguous; order [n_rows] lists, image by image, the rows by descending score (ties: lower
 * row first).  score [n_rows], image_of_row [n_rows] int64 (ascending). */
int subgc_rank_rows(int n_rows, const float* score, const int64_t* image_of_row, int64_t* order, subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Decoder feature preparation: replaces gpn read_out_proj (models/lib/gpn.py:79,95), AttModel.clip_att /
 * _prepare_feature / pack_wrapper (models/AttModel.py:16-36,348-368) for the n_rows selected sub-graphs `sel`:
 *    g_fc  = W_b(W_a read_out[sel] + b_a) + b_b                      [n_rows, 2L]   (reference `fc_feats`; NULL: not wanted --
 *            with w->prep_fold the three Linears up to fc_embed.0 are then ONE contraction)
 *    fc    = relu(W2 relu(W1 g_fc + b1) + b2)                        [n_rows, H]
 *    att   = relu(W_a x_obj[image, ids[n]] + b_a) for n < len else 0 [n_rows, len_max, H]
 *    p_att = W_c att + b_c                                           [n_rows, len_max, AH]
 *    masks = att_masks[sel, :len_max]                                [n_rows, len_max]
 * ------------------------------------------------------------------------------------------------------- */
size_t subgc_prepare_workspace_bytes(const subgc_dims* d, int n_rows, int len_max);
int subgc_prepare_forward(const subgc_dims* d, const subgc_weights* w, const subgc_subgraph_layout* lay,
                          int n_rows, int len_max, const int32_t* sel, const float* x_obj,
                          const int64_t* gpn_obj_ind, const float* att_masks, const float* read_out,
                          float* g_fc, float* fc, float* att, float* p_att, float* masks, void* ws,
                          size_t ws_bytes, subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Decoder.  One step = AttModel.get_logprobs_state (models/AttModel.py:328-341) = embed+ReLU, TopDownCore.forward
 * (:400-431: att-LSTM, Attention.forward :445-471, lang-LSTM) and logit + log_softmax.
 * rows_per_ctx > 1 lets consecutive rows share one (att, p_att, masks, fc) entry (beam search: the beams of a
 * sub-graph).  State tensors are [2, n_rows, H] (index 0 attention LSTM, 1 language LSTM), in/out distinct.
 * ------------------------------------------------------------------------------------------------------- */
size_t subgc_decode_workspace_bytes(const subgc_dims* d, int n_rows, int len_max);
int subgc_decode_step(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int rows_per_ctx,
                      const int64_t* it, const float* fc, const float* att, const float* p_att,
                      const float* masks, const float* h_in, const float* c_in, float* h_out, float* c_out,
                      float* logprobs /*[n_rows,V1]*/, float* att_weights /*nullable [n_rows,len_max]*/,
                      void* ws, size_t ws_bytes, subgc_stream_t stream);

/* Whole greedy / top-k loop of AttModel._sample (models/AttModel.py:278-326): T+1 decoder steps, finish masks and
 * the all-finished early exit evaluated on the device (no host sync).
 *   mode 0: greedy argmax (first index on ties).
 *   mode 1: top-k sampling: q = log_softmax(logp / temp); keep the k largest; draw from softmax(kept) by
 *           inverse CDF over the kept tokens in descending order using uniforms[t*n_rows + r] if given, else a
 *           Philox4x32-10 stream keyed by (seed, offset, t, r); seqLogprobs holds q[token] (un-renormalised).
 * Outputs: seq [n_rows,T] int64, seq_logprobs [n_rows,T], att_weights (nullable) [n_rows,T+1,len_max],
 * steps_done[0] = decoder steps executed (<= T+1).  Columns after an early exit stay zero. */
int subgc_decode_sample(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int mode,
                        float temp, int top_k, uint64_t seed, uint64_t offset, const float* uniforms /*nullable*/,
                        const float* fc, const float* att, const float* p_att, const float* masks,
                        int64_t* seq, float* seq_logprobs, float* att_weights /*nullable*/,
                        int32_t* steps_done, void* ws, size_t ws_bytes, subgc_stream_t stream);

/* Same loop with the row count and the attention length decided ON THE DEVICE: counts (device int32[2]) = (rows kept by
 * subgc_subgraph_nms = its stats[0], longest kept sub-graph = stats[1]) is read by the kernel; rows_cap / len_cap are upper bounds and
 * the strides of fc / att / p_att / masks / seq / seq_logprobs.  The reference synchronises with the host at exactly these two points
 * (NMS on the host, models/lib/gpn.py:114; clip_att's .max(), models/AttModel.py:351); here encoder -> sGPN -> NMS -> prepare -> decode
 * is one uninterrupted stream of launches (capturable as one CUDA graph).  Rows >= counts[0] of the outputs stay zero.  uniforms
 * (nullable) is [T, rows_cap].  Needs the persistent decode kernel (w->mega): SUBGC_E_UNSUPPORTED otherwise. */
int subgc_decode_sample_dyn(const subgc_dims* d, const subgc_weights* w, int rows_cap, int len_cap, const int32_t* counts,
                            int mode, float temp, int top_k, uint64_t seed, uint64_t offset, const float* uniforms /*nullable*/,
                            const float* fc, const float* att, const float* p_att, const float* masks, int64_t* seq,
                            float* seq_logprobs, int32_t* steps_done, void* ws, size_t ws_bytes, subgc_stream_t stream);

/* Teacher-forced decoding of AttModel._forward (models/AttModel.py:150-177, sampling_prob == 0, eval mode): step i
 * is fed tokens[:, i]; outputs[:, i] = log-probs; from the first column i >= 1 that is entirely zero on, the loop
 * stops and the remaining outputs stay zero (AttModel.py:170-171), evaluated on the device.
 * tokens [n_rows, ld_tok] int64 (labels), outputs [n_rows, n_steps, V1]. */
size_t subgc_teacher_workspace_bytes(const subgc_dims* d, int n_rows, int n_steps);
int subgc_decode_teacher(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int n_steps,
                         const int64_t* tokens, int ld_tok, const float* fc, const float* att, const float* p_att,
                         const float* masks, float* outputs, void* ws, size_t ws_bytes, subgc_stream_t stream);

/* Batched beam search: AttModel._sample_sentences + CaptionModel.beam_search with group_size 1
 * (models/AttModel.py:208-234, models/CaptionModel.py:43-94,97-176) for n_sub sub-graphs at once, beam_size rows
 * each.  length_penalty: 0 none, 1 wu_alpha, 2 avg (misc/utils.py:242-266).
 * Outputs (per sub-graph, the reference's `done_beams` list in its final order, best first):
 *   done_seq [n_sub, beam, T] int64, done_logps [n_sub, beam, T], done_p [n_sub, beam] (double),
 *   done_unaug_p [n_sub, beam] (double), done_count [n_sub]. */
size_t subgc_beam_workspace_bytes(const subgc_dims* d, int n_sub, int beam_size, int len_max);
int subgc_decode_beam(const subgc_dims* d, const subgc_weights* w, int n_sub, int len_max, int beam_size,
                      int length_penalty, double lp_alpha, int decoding_constraint, const float* fc,
                      const float* att, const float* p_att, const float* masks, int64_t* done_seq,
                      float* done_logps, double* done_p, double* done_unaug_p, int32_t* done_count, void* ws,
                      size_t ws_bytes, subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Training building blocks (train-mode forward pieces and the backward halves).  The host side composes them
 * into AttModel._forward + LossWrapper + backward (models/AttModel.py:122-177, models/loss_wrapper.py:14-27,
 * misc/utils.py:111-124); see sub-gc_b200/subgc/train.py for the exact sequence.
 * ------------------------------------------------------------------------------------------------------- */
/* C[M,N] (+)= act(A[gather] . W^T + bias): nn.Linear forward; with swapped / transposed operands the two backward
 * products dX = dY . W and dW += dY^T . X (operands are K-major: materialise transposes with subgc_transpose). */
size_t subgc_gemm_nt_workspace_bytes(int M, int N, int K);
int subgc_gemm_nt(int M, int N, int K, const float* A, int lda, const int64_t* a_gather /*nullable*/, const float* W,
                  int ldw, const float* bias /*nullable*/, int relu, int accumulate, float* C, int ldc, void* ws,
                  size_t ws_bytes, subgc_stream_t stream);
int subgc_transpose(int rows, int cols, const float* in, int ld_in, float* out, int ld_out, subgc_stream_t stream);
int subgc_colsum(int rows, int cols, const float* in, int ld, float* out, int accumulate, subgc_stream_t stream);
/* op: 0 out=a*b, 1 out=a+b, 2 out=(a>0?b:0) (ReLU backward, a = forward output), 3 out=a*scalar, 4 out=a */
int subgc_ew(int op, size_t n, const float* a, const float* b /*nullable*/, float* out, float scalar, subgc_stream_t stream);
/* mask[i] = (u_i >= p) ? 1/(1-p) : 0 with Philox4x32-10 uniforms keyed by (seed, offset, i): nn.Dropout(p) */
int subgc_dropout_mask(size_t n, float p, uint64_t seed, uint64_t offset, float* mask, subgc_stream_t stream);
int subgc_gather_rows(int n_rows, int cols, const float* src, int ld_src, const int64_t* idx, float* out, int relu,
                      subgc_stream_t stream);
/* Scheduled sampling (models/AttModel.py:158-167): it[r] = labels[r * ld_lab] or, with probability ss_prob per row, a token drawn from
 * exp(prev_logp[r, :]) (the previous step's log-probs, leading dim ld) by inverse CDF; Philox4x32-10 keyed by (seed; row, offset). */
int subgc_ss_sample(int rows, int V1, const float* prev_logp, size_t ld, const int64_t* labels, int ld_lab, float ss_prob,
                    uint64_t seed, uint64_t offset, int64_t* it, subgc_stream_t stream);
/* op 0: relu, 1: sigmoid */
int subgc_unary(int op, size_t n, const float* a, float* out, subgc_stream_t stream);
int subgc_scatter_add_rows(int n_rows, int cols, const float* src, int ld_src, const int64_t* idx, float* dst, int ld_dst,
                           subgc_stream_t stream);
/* nn.LSTMCell pointwise part; `gates` (pre-activation incl. biases) is overwritten by the activations (i,f,g,o) */
int subgc_lstm_cell_train_fwd(int S, int H, float* gates, const float* c_prev, float* h_out, float* c_out,
                              subgc_stream_t stream);
int subgc_lstm_cell_bwd(int S, int H, const float* act, const float* c_prev, const float* c_new, const float* dh,
                        const float* dc_in /*nullable*/, float* dgates, float* dc_prev, subgc_stream_t stream);
/* Attention.forward (models/AttModel.py:445-471) keeping the pre-mask softmax `sm` and the final weights `alpha` */
int subgc_attention_train_fwd(int S, int len, int H, int AH, const float* atth, const float* p_att, const float* att,
                              const float* masks, const float* alpha_w, const float* alpha_b, float* ctx, float* alpha,
                              float* sm, subgc_stream_t stream);
/* d_att and d_p_att are accumulated (+=); d_atth [S,AH] and d_w_rows [S,AH] (per-row partials of alpha_net.weight's
 * gradient) are overwritten */
int subgc_attention_bwd(int S, int len, int H, int AH, const float* atth, const float* p_att, const float* att,
                        const float* masks, const float* alpha_w, const float* alpha, const float* sm, const float* dctx,
                        float* d_att, float* d_p_att, float* d_atth, float* d_w_rows, subgc_stream_t stream);
int subgc_log_softmax_fwd(int rows, int V1, const float* logits, float* logp, size_t ld_out, subgc_stream_t stream);
int subgc_log_softmax_bwd(int rows, int V1, const float* logp, const float* dlogp, size_t ld, float* dlogits,
                          subgc_stream_t stream);
int subgc_class_argmax(int rows, int n_classes, int skip_first, const float* dist, int64_t* cls, subgc_stream_t stream);
/* Full-GC read-out (models/AttModel.py:146,200,265): out [B, L] = mean over the N nodes of x [B, N, L]. */
int subgc_mean_nodes(int B, int N, int L, const float* x, float* out, subgc_stream_t stream);
int subgc_sgpn_pool(const subgc_dims* d, const subgc_subgraph_layout* lay, const float* x_obj, const int64_t* gpn_obj_ind,
                    const float* att_masks, float* read_out, int32_t* sub_len, subgc_stream_t stream);
int subgc_sgpn_bce(const subgc_subgraph_layout* lay, const float* score, float* loss, subgc_stream_t stream);
int subgc_sgpn_pool_bwd(const subgc_dims* d, const subgc_subgraph_layout* lay, const float* x_obj,
                        const int64_t* gpn_obj_ind, const int32_t* sub_len, const float* d_read_out, float* d_x_obj,
                        subgc_stream_t stream);
int subgc_bce_sigmoid_bwd(const subgc_subgraph_layout* lay, const float* score, float scale, float* dz,
                          subgc_stream_t stream);
int subgc_prepare_index(const subgc_dims* d, const subgc_subgraph_layout* lay, int n_rows, int len_max, const int32_t* sel,
                        const int64_t* gpn_obj_ind, const float* att_masks, int64_t* node_row, float* masks,
                        int32_t* row_len, subgc_stream_t stream);
int subgc_gcn_edge_fwd(int B, int N, int K, int L, const float* m_subj, const float* m_obj, const int64_t* rel_ind,
                       const float* res /*nullable*/, float* out, subgc_stream_t stream);
int subgc_gcn_node_train_fwd(int B, int N, int K, int L, const float* m_subj, const float* m_obj, const int64_t* rel_ind,
                             const float* res /*nullable*/, float* out, float* y0, float* y1, subgc_stream_t stream);
int subgc_gcn_node_bwd(int B, int N, int K, int L, const float* dx, const float* y0, const float* y1,
                       const int64_t* rel_ind, float* dm_subj, float* dm_obj, subgc_stream_t stream);
int subgc_gcn_edge_bwd(int B, int N, int K, int L, const float* dp, const float* m_subj, const float* m_obj,
                       const int64_t* rel_ind, float* dm_subj, float* dm_obj, subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Stage-level training entries: the teacher-forced decoder of AttModel._forward (models/AttModel.py:150-177:
 * embed -> TopDownCore (:400-431) -> Attention (:445-471) -> logit + log_softmax (:339-340)) and its backward, ONE call each.
 * The recurrence is sequenced in C++; the logit contraction, log-softmax, d(logits) and every weight / bias gradient are
 * batched over the T executed steps (dW = dY_all^T . X_all is one contraction over K = T * R per weight block).
 * All pointers DEVICE, fp32 unless noted, contiguous.  R rows (sentences), T executed steps, T_total = labels.size(1) - 1.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct subgc_decoder_train_bufs {
    /* inputs */
    const int64_t* tokens; /* [T, R] token fed at step t (labels[:, t])                                         */
    const float* fc;       /* [R, H] fc_embed output after dropout                                              */
    const float* att;      /* [R, len, H] att_embed output (padded rows exact zeros)                            */
    const float* p_att;    /* [R, len, AH] ctx2att(att)                                                         */
    const float* masks;    /* [R, len]                                                                          */
    const float* m_x;      /* nullable [T, R, E] dropout mask of the word embedding (values 0 or 1/(1-p))       */
    const float* m_h;      /* nullable [T, R, H] dropout mask of h_lang before the logit                        */
    /* saved by the forward, read by the backward */
    float* xt;             /* [T, R, E] relu(E[token]) * m_x                                                    */
    float* act1;           /* [T, R, 4H] attention-LSTM gate activations (i, f, g, o after their non-linearities) */
    float* c_att;          /* [T+1, R, H] cell states, slot 0 = init_hidden zeros                               */
    float* h_att;          /* [T+1, R, H]                                                                       */
    float* atth;           /* [T, R, AH] h2att(h_att)                                                           */
    float* ctx;            /* [T, R, H] attention context                                                       */
    float* alpha;          /* [T, R, len] attention weights                                                     */
    float* sm;             /* [T, R, len] softmax before masking                                                */
    float* act2;           /* [T, R, 4H] language-LSTM gate activations                                         */
    float* c_lang;         /* [T+1, R, H]                                                                       */
    float* h_lang;         /* [T+1, R, H]                                                                       */
    float* hd;             /* [T, R, H] h_lang * m_h (unused when m_h is NULL)                                  */
    float* outputs;        /* nullable [R, T_total, V1] log-probabilities; steps >= T are left untouched        */
    /* fused log-softmax + LanguageModelCriterion (misc/utils.py:115-124), used when the caller is LossWrapper
     * (models/loss_wrapper.py:22): with `nll` set the log-probs need not exist (outputs may be NULL)            */
    float* logits;         /* nullable [T*R, V1]: where the logits of all steps are kept (required for the fused loss) */
    const int64_t* targets;/* [T, R] labels[:, t+1]                                                             */
    const float* tmask;    /* [T, R] masks[:, t+1]                                                              */
    float* lse;            /* [T, R] log-sum-exp of every row (saved for the backward)                          */
    float* nll;            /* [T, R] -logp[target] * mask                                                       */
    const float* coef;     /* backward with d_outputs == NULL: [T, R] mask * d(lang_loss) / sum(mask)           */
} subgc_decoder_train_bufs;

typedef struct subgc_decoder_grads { /* gradients, same shapes as the parameters; ACCUMULATED into (caller zero-fills) */
    float* logit_w; float* logit_b; float* embed;
    float* att_w_ih; float* att_w_hh; float* att_b_ih; float* att_b_hh;
    float* lang_w_ih; float* lang_w_hh; float* lang_b_ih; float* lang_b_hh;
    float* h2att_w; float* h2att_b; float* alpha_w;
} subgc_decoder_grads;

size_t subgc_decoder_train_workspace_bytes(const subgc_dims* d, int R, int len, int T);
int subgc_decoder_train_forward(const subgc_dims* d, const subgc_weights* w, int R, int len, int T, int T_total,
                                const subgc_decoder_train_bufs* b, void* ws, size_t ws_bytes, subgc_stream_t stream);
/* d_fc [R, H], d_att [R, len, H], d_p_att [R, len, AH]: gradients of the decoder's inputs (overwritten).
 * d_outputs [R, T_total, V1], or NULL when the forward ran the fused loss (b->coef then carries d(lang_loss)). */
int subgc_decoder_train_backward(const subgc_dims* d, const subgc_weights* w, int R, int len, int T, int T_total,
                                 const subgc_decoder_train_bufs* b, const float* d_outputs, const subgc_decoder_grads* g,
                                 float* d_fc, float* d_att, float* d_p_att, void* ws, size_t ws_bytes, subgc_stream_t stream);

/* Front-end backward stages.  `*_saved`: train-mode forward activations; `*_grads`: parameter gradients, ACCUMULATED
 * into (caller zero-fills).  One workspace query covers the three calls. */
typedef struct subgc_prepare_train_saved {
    const float* m_fc;       /* nullable [R, H] dropout mask of fc_embed's output                                 */
    const float* fc_pre;     /* [R, H] fc_embed output before dropout                                             */
    const float* f1;         /* [R, FC] relu(fc_embed.0)                                                          */
    const float* g_fc;       /* [R, 2L] read_out_proj output                                                      */
    const float* hr;         /* [R, AH] read_out_proj.0 output                                                    */
    const float* read_sel;   /* [R, 2L] pooled read-out of the selected sub-graphs (detached, gpn.py:78)          */
    const float* att;        /* [R*len, H] att_embed output after mask                                            */
    const float* att_pre;    /* [R*len, H] relu(att_embed.0)                                                      */
    const float* m_att;      /* [R*len, H] dropout mask x valid-row mask                                          */
    const float* x_rows;     /* [R*len, L] gathered node features                                                 */
    const int64_t* node_row; /* [R*len] row of x_obj [B*N, L] each attention row was gathered from                */
} subgc_prepare_train_saved;
typedef struct subgc_prepare_grads {
    float* fc2_w; float* fc2_b; float* fc0_w; float* fc0_b; float* ro1_w; float* ro1_b; float* ro0_w; float* ro0_b;
    float* ctx2att_w; float* ctx2att_b; float* att_embed_w; float* att_embed_b;
} subgc_prepare_grads;
typedef struct subgc_sgpn_train_saved {
    const float* score;      /* [n_sub] sigmoid output                                                            */
    const float* hid;        /* [n_sub, AH] relu(gpn_fc.0)                                                        */
    const float* hid_d;      /* [n_sub, AH] after Dropout(0.5) (= hid when m_gpn is NULL)                         */
    const float* m_gpn;      /* nullable [n_sub, AH]                                                              */
    const float* read_out;   /* [n_sub, 2L]                                                                       */
    const int32_t* sub_len;  /* [n_sub]                                                                           */
} subgc_sgpn_train_saved;
typedef struct subgc_sgpn_grads { float* fc3_w; float* fc3_b; float* fc0_w; float* fc0_b; } subgc_sgpn_grads;
typedef struct subgc_gcn_layer_saved { /* NULL where the unit pair is dead (gcn liveness) */
    const float* x_in; const float* p_in;   /* [B, N, L] / [B, K, L] streams entering the layer                   */
    const float* t0; const float* t1;       /* [B*K, R] fc_lft outputs of units 0, 1                              */
    const float* y0; const float* y1;       /* [B, N, L] pre-ReLU segment means of units 0, 1                     */
    const float* t2; const float* t3;       /* [B*N, R] fc_lft outputs of units 2, 3                              */
    const float* m2; const float* m3;       /* [B, N, L] messages of units 2, 3                                   */
} subgc_gcn_layer_saved;
typedef struct subgc_gcn_train_saved {
    subgc_gcn_layer_saved layer[SUBGC_MAX_GCN_LAYERS];
    const float* x0;         /* [B, N, L] fused node features                                                     */
    const float* att_feats;  /* [B, N, A]                                                                         */
    const int64_t* cls;      /* [B*N] object class ids                                                            */
    const int64_t* rel_ind;  /* [B, K, 2]                                                                         */
} subgc_gcn_train_saved;
typedef struct subgc_gcn_grads {
    float* lft_w[SUBGC_MAX_GCN_LAYERS][4]; float* lft_b[SUBGC_MAX_GCN_LAYERS][4];
    float* rgt_w[SUBGC_MAX_GCN_LAYERS][4]; float* rgt_b[SUBGC_MAX_GCN_LAYERS][4];
    float* obj_v_w; float* obj_v_b; float* obj_emb_w; float* obj_emb_b; float* sg_obj_embed;
} subgc_gcn_grads;
size_t subgc_frontend_backward_workspace_bytes(const subgc_dims* d, int n_images, int R, int len, int n_sub);
int subgc_prepare_backward(const subgc_dims* d, const subgc_weights* w, int R, int len, int n_nodes,
                           const subgc_prepare_train_saved* s, float* d_fc, const float* d_att, const float* d_p_att,
                           const subgc_prepare_grads* g, float* d_x_obj, void* ws, size_t ws_bytes, subgc_stream_t stream);
int subgc_sgpn_backward(const subgc_dims* d, const subgc_weights* w, const subgc_subgraph_layout* lay,
                        const subgc_sgpn_train_saved* s, float scale, const float* x_obj, const int64_t* gpn_obj_ind,
                        const subgc_sgpn_grads* g, float* d_x_obj, void* ws, size_t ws_bytes, subgc_stream_t stream);
int subgc_gcn_backward(const subgc_dims* d, const subgc_weights* w, int n_images, const subgc_gcn_train_saved* s,
                       const float* d_x_obj, const subgc_gcn_grads* g, void* ws, size_t ws_bytes, subgc_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimiser step: utils.clip_gradient_norm(optimizer, clip) (misc/utils.py:174-200) + torch.optim.Adam.step()
 * (misc/utils.py:236, train.py:107,163-164) over every parameter in two multi-tensor passes.  `chunks` is a DEVICE table of
 * n_chunks entries {float* p; const float* g; float* m; float* v; int32 n; int32 pad} (24 + 8 bytes), each covering at most
 * subgc_opt_chunk_elems() consecutive elements of one parameter.  partial [n_chunks] is scratch; norm_out[0] = total gradient
 * norm, norm_out[1] = clip coefficient (both stay on the device: no host sync).  step = 1, 2, ... (bias correction).
 * write_grad != 0 also stores the scaled gradients back, as the reference's in-place p.grad.mul_(norm) does. */
int subgc_opt_chunk_elems(void);
int subgc_clip_adam_step(const void* chunks, int n_chunks, float clip_norm, float lr, float beta1, float beta2, float eps,
                         float weight_decay, int step, int write_grad, float* partial, float* norm_out, subgc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SUBGC_B200_H */
