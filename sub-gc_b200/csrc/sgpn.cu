// sGPN: sub-graph read-out, scorer MLP, BCE loss, training-time selection and inference-time node-set NMS.
//
//   subgc_sgpn_forward      <- gpn_layer.extract_subgraph_feats + graph_pooling + gpn_fc + sigmoid + BCELoss
//                              (reference models/lib/gpn.py:50-57,152-185)
//   subgc_sgpn_select_train <- gpn_layer.forward training branch (reference models/lib/gpn.py:64-81)
//   subgc_subgraph_nms      <- gpn_layer.subgraph_nms + cal_node_iou (reference models/lib/gpn.py:108-150)
//
// The reference gathers [2*5B*G, 37, 1024] node rows, multiplies them with a 37x37 diagonal 0/1 matrix to zero the
// padded rows and then pools; here a sub-graph is (image, node ids, length) and the pooling kernel reads the node
// rows of x_obj in place (no x5 replication, no gather buffer, no pooling matrix).  The NMS runs on 64-bit node
// masks with popcounts — exact for node-set IoU — instead of the O(S^2) Python/numpy loop.
#include "common.cuh"

namespace subgc {

// one block per sub-graph: read_out[s] = [ max over the (zero-padded) rows | sum over rows / len ]
__global__ void __launch_bounds__(256) sgpn_pool_kernel(const subgc_subgraph_layout lay, const float* __restrict__ x_obj,
                                                        const long long* __restrict__ obj_ind, const float* __restrict__ att_masks,
                                                        float* __restrict__ read_out, int* __restrict__ sub_len, int N, int L) {
    extern __shared__ int s_ids[];  // [N]
    __shared__ int s_len;
    const int s = blockIdx.x;
    int image;
    const int slot = subgraph_slot(lay, s, &image);
    if (threadIdx.x == 0) {
        int len = 0;
        for (int n = 0; n < N; ++n) len += (__ldg(att_masks + (size_t)slot * N + n) != 0.f) ? 1 : 0;
        s_len = len;
        sub_len[s] = len;
    }
    for (int n = threadIdx.x; n < N; n += blockDim.x) s_ids[n] = (int)obj_ind[(size_t)slot * N + n];
    __syncthreads();
    const int len = s_len;
    const float* xi = x_obj + (size_t)image * N * L;
    const float flen = (float)len;
    if ((L & 3) == 0) {   // 16-byte columns, six independent row loads in flight, sum kept in ascending row order
        const int L4 = L >> 2;
        for (int c4 = threadIdx.x; c4 < L4; c4 += blockDim.x) {
            const float m0 = (len < N) ? 0.f : -INFINITY;
            float4 mx = make_float4(m0, m0, m0, m0), sm = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int n = 0; n < len; n += 6) {
                float4 v[6];
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    v[u] = (n + u < len) ? __ldg(reinterpret_cast<const float4*>(xi + (size_t)s_ids[n + u] * L) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 6; ++u) {
                    if (n + u < len) {
                        mx.x = fmaxf(mx.x, v[u].x); mx.y = fmaxf(mx.y, v[u].y); mx.z = fmaxf(mx.z, v[u].z); mx.w = fmaxf(mx.w, v[u].w);
                        sm.x += v[u].x; sm.y += v[u].y; sm.z += v[u].z; sm.w += v[u].w;
                    }
                }
            }
            reinterpret_cast<float4*>(read_out + (size_t)s * 2 * L)[c4] = mx;
            reinterpret_cast<float4*>(read_out + (size_t)s * 2 * L + L)[c4] = make_float4(sm.x / flen, sm.y / flen, sm.z / flen, sm.w / flen);
        }
        return;
    }
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float mx = (len < N) ? 0.f : -INFINITY;  // rows beyond len are zeros in the reference and take part in the max
        float sm = 0.f;
        for (int n = 0; n < len; ++n) {
            float v = __ldg(xi + (size_t)s_ids[n] * L + c);
            mx = fmaxf(mx, v);
            sm += v;
        }
        read_out[(size_t)s * 2 * L + c] = mx;
        read_out[(size_t)s * 2 * L + L + c] = sm / flen;
    }
}

// one warp per sub-graph: score = sigmoid(w2 . hid + b2)
__global__ void __launch_bounds__(256) sgpn_score_kernel(const float* __restrict__ hid, const float* __restrict__ w2, const float* __restrict__ b2,
                                                         float* __restrict__ score, int n_sub, int AH) {
    int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_sub) return;
    const int lane = threadIdx.x & 31;
    float a = 0.f;
    for (int j = lane; j < AH; j += 32) a = fmaf(__ldg(hid + (size_t)s * AH + j), __ldg(w2 + j), a);
    a = warp_sum(a);
    if (lane == 0) score[s] = sigmoidf_(a + __ldg(b2));
}

// single block: mean BCE with torch's log clamp at -100; label 1 for sub-graphs of half 0, 0 for half 1
__global__ void __launch_bounds__(1024) sgpn_bce_kernel(const subgc_subgraph_layout lay, const float* __restrict__ score, int n_sub,
                                                        float* __restrict__ loss) {
    __shared__ float red[32];
    float a = 0.f;
    for (int s = threadIdx.x; s < n_sub; s += blockDim.x) {
        int half = (lay.order == 0) ? (s / lay.per_half) / lay.rows : (s / lay.per_half) % 2;
        float p = score[s];
        float t = half == 0 ? fmaxf(logf(p), -100.f) : fmaxf(logf(1.f - p), -100.f);
        a -= t;
    }
    float tot = block_sum(a, red);
    if (threadIdx.x == 0) loss[0] = tot / (float)n_sub;
}

// training-time pick: per sentence row argmax_first over its positive sub-graphs
__global__ void __launch_bounds__(256) sgpn_select_train_kernel(const subgc_subgraph_layout lay, const float* __restrict__ score,
                                                                const int* __restrict__ sub_len, int* __restrict__ sel, int* __restrict__ stats) {
    __shared__ int s_max;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    int local = 0;
    for (int row = threadIdx.x; row < lay.rows; row += blockDim.x) {
        int base = row * lay.per_half;  // order 0, half 0
        float bv = score[base];
        int bi = 0;
        for (int g = 1; g < lay.per_half; ++g) {
            float v = score[base + g];
            if (v > bv) { bv = v; bi = g; }
        }
        sel[row] = base + bi;
        local = max(local, sub_len[base + bi]);
    }
    atomicMax(&s_max, local);
    __syncthreads();
    if (threadIdx.x == 0) { stats[0] = lay.rows; stats[1] = s_max; }
}

// ---- NMS ---------------------------------------------------------------------------------------------------------
// one block per image.  P = 2*per_half candidates; dynamic smem: order[P] (int), mask[P] (u64), alive[P] (u8)
__global__ void __launch_bounds__(256) nms_image_kernel(const subgc_subgraph_layout lay, const float* __restrict__ score,
                                                        const int* __restrict__ sub_len, const long long* __restrict__ obj_ind,
                                                        const float* __restrict__ att_masks, int N, int use_nms, double thres, int max_keep,
                                                        unsigned char* __restrict__ kept /*[n_images*P]*/, int* __restrict__ counts /*[n_images]*/) {
    extern __shared__ unsigned long long s_mem[];
    const int P = 2 * lay.per_half;
    unsigned long long* s_mask = s_mem;                       // [P]
    int* s_order = reinterpret_cast<int*>(s_mask + P);        // [P]
    unsigned char* s_alive = reinterpret_cast<unsigned char*>(s_order + P);  // [P], indexed by sorted position
    __shared__ int s_nkept, s_cur;
    const int img = blockIdx.x;
    const int base = img * P;
    if (!use_nms) {
        for (int i = threadIdx.x; i < P; i += blockDim.x) kept[base + i] = 1;
        if (threadIdx.x == 0) counts[img] = P;
        return;
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int slot = subgraph_slot(lay, base + i, nullptr);
        unsigned long long m = 0;
        for (int n = 0; n < N; ++n)
            if (__ldg(att_masks + (size_t)slot * N + n) != 0.f) m |= 1ull << (unsigned)obj_ind[(size_t)slot * N + n];
        s_mask[i] = m;
        // rank in descending score order, ties: higher original index first (reversed stable ascending argsort)
        const float si = score[base + i];
        int rank = 0;
        for (int j = 0; j < P; ++j) {
            float sj = score[base + j];
            rank += (sj > si || (sj == si && j > i)) ? 1 : 0;
        }
        s_order[rank] = i;
        kept[base + i] = 0;
    }
    for (int i = threadIdx.x; i < P; i += blockDim.x) s_alive[i] = 1;
    if (threadIdx.x == 0) { s_nkept = 0; s_cur = 0; }
    __syncthreads();
    // greedy walk: position `cur` is kept iff still alive; it then suppresses every later position with IoU > thres
    while (true) {
        __syncthreads();
        int cur = s_cur;
        if (cur >= P || s_nkept >= max_keep) break;
        if (!s_alive[cur]) {
            __syncthreads();
            if (threadIdx.x == 0) s_cur = cur + 1;
            continue;
        }
        const unsigned long long a = s_mask[s_order[cur]];
        const int na = __popcll(a);
        for (int j = cur + 1 + threadIdx.x; j < P; j += blockDim.x) {
            if (!s_alive[j]) continue;
            const unsigned long long b = s_mask[s_order[j]];
            const int nb = __popcll(b);
            double iou = 0.0;
            if (na > 0 && nb > 0) iou = (double)__popcll(a & b) / (double)__popcll(a | b);
            if (iou > thres) s_alive[j] = 0;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            kept[base + s_order[cur]] = 1;
            s_nkept = s_nkept + 1;
            s_cur = cur + 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) counts[img] = s_nkept;
}

// single block: exclusive scan of per-image counts, then compaction in (image, ascending original index) order
__global__ void __launch_bounds__(256) nms_compact_kernel(const subgc_subgraph_layout lay, const unsigned char* __restrict__ kept,
                                                          const int* __restrict__ counts, const int* __restrict__ sub_len, int n_images,
                                                          int* __restrict__ sel, long long* __restrict__ keep_ind, int* __restrict__ stats,
                                                          int* __restrict__ offsets /*[n_images]*/) {
    __shared__ int s_maxlen;
    const int P = 2 * lay.per_half;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 0; i < n_images; ++i) { offsets[i] = run; run += counts[i]; stats[2 + i] = counts[i]; }
        stats[0] = run;
        s_maxlen = 0;
    }
    __syncthreads();
    int local = 0;
    for (int img = threadIdx.x; img < n_images; img += blockDim.x) {
        int o = offsets[img];
        for (int i = 0; i < P; ++i) {
            if (kept[img * P + i]) {
                sel[o] = img * P + i;
                keep_ind[o] = i;
                local = max(local, sub_len[img * P + i]);
                ++o;
            }
        }
    }
    atomicMax(&s_maxlen, local);
    __syncthreads();
    if (threadIdx.x == 0) stats[1] = s_maxlen;
}

// Post-decode ordering (reference misc/eval_utils.py:105-110): the captions of an image are listed by descending sGPN score.  Rows of
// one image are contiguous (subgc_subgraph_nms emits images in order); order[first + rank] = row with rank = number of rows of the
// same image that come before it (higher score, or equal score and lower row index: a stable descending sort).
__global__ void __launch_bounds__(256) rank_rows_kernel(const float* __restrict__ score, const long long* __restrict__ image_of_row, int n_rows,
                                                        long long* __restrict__ order) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const long long img = image_of_row[i];
    const float si = score[i];
    int first = i, rank = 0;
    while (first > 0 && image_of_row[first - 1] == img) --first;
    for (int j = first; j < n_rows && image_of_row[j] == img; ++j) {
        const float sj = score[j];
        rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
    }
    order[first + rank] = i;
}

static int check_layout(const subgc_subgraph_layout* l) {
    SUBGC_CHECK_ARG(l != nullptr, "layout is null");
    SUBGC_CHECK_ARG(l->rows > 0 && l->per_half > 0 && l->seq_per_img > 0 && (l->order == 0 || l->order == 1), "bad sub-graph layout");
    SUBGC_CHECK_ARG(l->rows % l->seq_per_img == 0, "layout rows must be a multiple of seq_per_img");
    return SUBGC_OK;
}

}  // namespace subgc

using namespace subgc;

extern "C" size_t subgc_sgpn_workspace_bytes(const subgc_dims* d, int n_sub) {
    if (!d || n_sub <= 0) return 0;
    return align_up((size_t)n_sub * d->att_hid * 4, 256) + align_up(gemm_workspace_bytes(n_sub, d->att_hid, 2 * d->gcn), 256) + 1024;
}

extern "C" int subgc_sgpn_forward(const subgc_dims* d, const subgc_weights* w, const subgc_subgraph_layout* lay, const float* x_obj,
                                  const int64_t* gpn_obj_ind, const float* att_masks, float* read_out, float* score, int32_t* sub_len,
                                  float* bce_loss, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_layout(lay));
    SUBGC_CHECK_ARG(d && w && x_obj && gpn_obj_ind && att_masks && read_out && score && sub_len, "subgc_sgpn_forward: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_sub = subgraph_count(*lay), N = d->obj_num, L = d->gcn, AH = d->att_hid;
    Workspace ws(ws_, ws_bytes);
    float* hid = ws.take<float>((size_t)n_sub * AH);
    if (!ws.ok()) { set_error("subgc_sgpn_forward: workspace too small"); return SUBGC_E_WORKSPACE; }
    sgpn_pool_kernel<<<n_sub, 256, N * sizeof(int), st>>>(*lay, x_obj, reinterpret_cast<const long long*>(gpn_obj_ind), att_masks, read_out,
                                                          sub_len, N, L);
    SUBGC_LAUNCH_CHECK();
    GemmProblem p;
    p.wts = w;
    p.M = n_sub; p.N = AH; p.nseg = 1;
    p.seg[0] = make_seg(read_out, 2 * L, w->gpn_fc0.w, 2 * L, 2 * L);
    p.epi.bias = w->gpn_fc0.b;
    p.epi.relu = 1;
    p.C = hid; p.ldc = AH;
    SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    sgpn_score_kernel<<<(n_sub + 7) / 8, 256, 0, st>>>(hid, w->gpn_fc3.w, w->gpn_fc3.b, score, n_sub, AH);
    SUBGC_LAUNCH_CHECK();
    if (bce_loss) {
        sgpn_bce_kernel<<<1, 1024, 0, st>>>(*lay, score, n_sub, bce_loss);
        SUBGC_LAUNCH_CHECK();
    }
    return SUBGC_OK;
}

extern "C" int subgc_sgpn_select_train(const subgc_subgraph_layout* lay, const float* score, const int32_t* sub_len, int32_t* sel,
                                       int32_t* stats, subgc_stream_t stream) {
    SUBGC_TRY(check_layout(lay));
    SUBGC_CHECK_ARG(lay->order == 0, "subgc_sgpn_select_train: needs the training layout (order 0)");
    SUBGC_CHECK_ARG(score && sub_len && sel && stats, "subgc_sgpn_select_train: null argument");
    sgpn_select_train_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(*lay, score, sub_len, sel, stats);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" size_t subgc_nms_workspace_bytes(int n_images, int per_image) {
    if (n_images <= 0 || per_image <= 0) return 0;
    return align_up((size_t)n_images * per_image, 256) + 2 * align_up((size_t)n_images * 4, 256) + 1024;
}

extern "C" int subgc_subgraph_nms(const subgc_dims* d, const subgc_subgraph_layout* lay, const float* score, const int32_t* sub_len,
                                  const int64_t* gpn_obj_ind, const float* att_masks, int use_nms, double iou_thres, int max_subgraphs,
                                  int32_t* sel, int64_t* keep_ind, int32_t* stats, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_layout(lay));
    SUBGC_CHECK_ARG(lay->order == 1, "subgc_subgraph_nms: needs the inference layout (order 1)");
    SUBGC_CHECK_ARG(d && score && sub_len && gpn_obj_ind && att_masks && sel && keep_ind && stats, "subgc_subgraph_nms: null argument");
    SUBGC_CHECK_ARG(d->obj_num <= 64, "subgc_subgraph_nms: node masks hold at most 64 nodes (obj_num = %d)", d->obj_num);
    SUBGC_CHECK_ARG(!use_nms || max_subgraphs >= 1, "subgc_subgraph_nms: max_subgraphs must be >= 1");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_images = lay->rows / lay->seq_per_img, P = 2 * lay->per_half;
    Workspace ws(ws_, ws_bytes);
    unsigned char* kept = ws.take<unsigned char>((size_t)n_images * P);
    int* counts = ws.take<int>(n_images);
    int* offsets = ws.take<int>(n_images);
    if (!ws.ok()) { set_error("subgc_subgraph_nms: workspace too small"); return SUBGC_E_WORKSPACE; }
    size_t smem = (size_t)P * (8 + 4 + 1) + 16;
    SUBGC_CHECK_ARG(smem <= 200 * 1024, "subgc_subgraph_nms: too many sub-graphs per image (%d)", P);
    if (smem > 48 * 1024) SUBGC_CUDA(cudaFuncSetAttribute(nms_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_image_kernel<<<n_images, 256, smem, st>>>(*lay, score, sub_len, reinterpret_cast<const long long*>(gpn_obj_ind), att_masks, d->obj_num,
                                                  use_nms, iou_thres, max_subgraphs, kept, counts);
    SUBGC_LAUNCH_CHECK();
    nms_compact_kernel<<<1, 256, 0, st>>>(*lay, kept, counts, sub_len, n_images, sel, reinterpret_cast<long long*>(keep_ind), stats, offsets);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_rank_rows(int n_rows, const float* score, const int64_t* image_of_row, int64_t* order, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(n_rows > 0 && score && image_of_row && order, "subgc_rank_rows: bad arguments");
    rank_rows_kernel<<<(n_rows + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(score, reinterpret_cast<const long long*>(image_of_row), n_rows,
                                                                                       reinterpret_cast<long long*>(order));
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_sgpn_pool(const subgc_dims* d, const subgc_subgraph_layout* lay, const float* x_obj, const int64_t* gpn_obj_ind,
                               const float* att_masks, float* read_out, int32_t* sub_len, subgc_stream_t stream) {
    SUBGC_TRY(check_layout(lay));
    SUBGC_CHECK_ARG(d && x_obj && gpn_obj_ind && att_masks && read_out && sub_len, "subgc_sgpn_pool: null argument");
    const int n_sub = subgraph_count(*lay);
    sgpn_pool_kernel<<<n_sub, 256, d->obj_num * sizeof(int), static_cast<cudaStream_t>(stream)>>>(
        *lay, x_obj, reinterpret_cast<const long long*>(gpn_obj_ind), att_masks, read_out, sub_len, d->obj_num, d->gcn);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_sgpn_bce(const subgc_subgraph_layout* lay, const float* score, float* loss, subgc_stream_t stream) {
    SUBGC_TRY(check_layout(lay));
    SUBGC_CHECK_ARG(score && loss, "subgc_sgpn_bce: null argument");
    sgpn_bce_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(*lay, score, subgraph_count(*lay), loss);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}
