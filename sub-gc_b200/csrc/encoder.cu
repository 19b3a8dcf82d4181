// Encoder of the Sub-GC path: feature fusion and GCN message passing over ragged per-image scene graphs.
//
//   subgc_fuse_nodes   <- AttModel.feat_fusion                     (reference models/AttModel.py:370-387)
//   subgc_gcn_forward  <- gcn_backbone.forward / make_map          (reference models/lib/gcn_backbone.py:29-67)
//                         _GraphConvolutionLayer.forward           (reference models/lib/graph_conv.py:15-34)
//                         _Collection_Unit.forward                 (reference models/lib/graph_conv_unit.py:28-36)
//
// The reference materialises a dense 0/1 adjacency [B,N,K,2] and multiplies with it (bmm).  Each edge has exactly one
// subject and one object, so  adj^T . msg  is a row gather (edge <- node) and  adj . msg  is a segment sum over the
// edges incident to a node, taken here in ascending edge order (the order of the dense product).  The mean divisor
// is the fp32 value (count + 1e-7) the reference divides by.  Units whose result cannot reach a requested output are
// skipped (for Sub-GC's 2 layers / residual 2 that is half of them, SURVEY headline fact 2).
#include "common.cuh"

namespace subgc {

// cls[row] = off + argmax_first(dist[row, off:])   (torch.max first-index tie-break), one warp per row
__global__ void __launch_bounds__(256) class_argmax_kernel(const float* __restrict__ dist, int rows, int C, int off,
                                                           long long* __restrict__ cls) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* p = dist + (size_t)row * C;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll 8
    for (int j = off + lane; j < C; j += 32) {   // eight independent loads in flight per lane
        float v = __ldg(p + j);
        if (v > bv || bi == 0x7fffffff) { bv = v; bi = j; }  // strictly greater keeps the first index within a lane
    }
    warp_argmax(bv, bi);
    if (lane == 0) cls[row] = (bi == 0x7fffffff) ? off : bi;
}

// p_new[b,k,:] = 0.5 * ( relu(M2[b, s_k, :] / (1+1e-7)) + relu(M3[b, o_k, :] / (1+1e-7)) ) (+ residual)
__global__ void __launch_bounds__(256) gcn_edge_update_kernel(const float* __restrict__ m_subj, const float* __restrict__ m_obj,
                                                              const long long* __restrict__ rel_ind, const float* __restrict__ res,
                                                              float* __restrict__ out, int B, int N, int K, int L,
                                                              unsigned short* __restrict__ o16_hi = nullptr, unsigned short* __restrict__ o16_lo = nullptr,
                                                              int ld16 = 0, int* __restrict__ overflow = nullptr, int ldm = 0, float mscale = 1.f) {
    // o16_hi / o16_lo (nullable, [B*K, ld16]): split-fp16 copy of the new edge features for the contractions of the next layer
    // ldm: row stride of m_subj / m_obj (0 = L; 2L when both units of the direction came out of one folded contraction)
    const int bk = blockIdx.x;  // b*K + k
    const int b = bk / K;
    const long long s = rel_ind[(size_t)bk * 2], o = rel_ind[(size_t)bk * 2 + 1];
    const float d = 1.f + 1e-7f;
    if (ldm == 0) ldm = L;
    // mscale: messages of a folded contraction arrive multiplied by a power of two (exact), undone here
    const float* ms = m_subj + ((size_t)b * N + s) * ldm;
    const float* mo = m_obj + ((size_t)b * N + o) * ldm;
    if ((L & 3) == 0) {
        const int L4 = L >> 2;
        for (int c4 = threadIdx.x; c4 < L4; c4 += blockDim.x) {
            float4 a = __ldg(reinterpret_cast<const float4*>(ms) + c4), bb = __ldg(reinterpret_cast<const float4*>(mo) + c4);
            a.x *= mscale; a.y *= mscale; a.z *= mscale; a.w *= mscale; bb.x *= mscale; bb.y *= mscale; bb.z *= mscale; bb.w *= mscale;
            float4 v;
            v.x = 0.5f * (fmaxf(a.x / d, 0.f) + fmaxf(bb.x / d, 0.f));
            v.y = 0.5f * (fmaxf(a.y / d, 0.f) + fmaxf(bb.y / d, 0.f));
            v.z = 0.5f * (fmaxf(a.z / d, 0.f) + fmaxf(bb.z / d, 0.f));
            v.w = 0.5f * (fmaxf(a.w / d, 0.f) + fmaxf(bb.w / d, 0.f));
            if (res) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(res + (size_t)bk * L) + c4);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            reinterpret_cast<float4*>(out + (size_t)bk * L)[c4] = v;
            if (o16_hi) {
                unsigned short hh[4], hl[4];
                int ovf = 0;
                split_f16(v.x, hh[0], hl[0], ovf); split_f16(v.y, hh[1], hl[1], ovf); split_f16(v.z, hh[2], hl[2], ovf); split_f16(v.w, hh[3], hl[3], ovf);
                const size_t o = (size_t)bk * ld16 + 4 * c4;
                *reinterpret_cast<uint2*>(o16_hi + o) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
                *reinterpret_cast<uint2*>(o16_lo + o) = make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16));
                if (ovf && overflow) atomicOr(overflow, 1);
            }
        }
        return;
    }
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float v = 0.5f * (fmaxf(__ldg(ms + c) * mscale / d, 0.f) + fmaxf(__ldg(mo + c) * mscale / d, 0.f));
        if (res) v += __ldg(res + (size_t)bk * L + c);
        out[(size_t)bk * L + c] = v;
        if (o16_hi) split_f16_store(v, o16_hi, o16_lo, (size_t)bk * ld16 + c, overflow);
    }
}

// x_new[b,n,:] = 0.5 * ( relu(sum_{k: s_k = n} M0[b,k,:] / (cnt_s + 1e-7)) + relu(sum_{k: o_k = n} M1[b,k,:] / (cnt_o + 1e-7)) )
__global__ void __launch_bounds__(256) gcn_node_update_kernel(const float* __restrict__ m_subj, const float* __restrict__ m_obj,
                                                              const long long* __restrict__ rel_ind, const float* __restrict__ res,
                                                              float* __restrict__ out, int B, int N, int K, int L, int ldm = 0, float mscale = 1.f) {
    extern __shared__ int s_list[];  // [2][K] edge lists of this node | [2][K] (subject, object) of the image's edges
    if (ldm == 0) ldm = L;           // row stride of m_subj / m_obj
    __shared__ int s_cnt[2];
    int* s_raw = s_list + 2 * K;
    const int bn = blockIdx.x;
    const int b = bn / N, n = bn - b * N;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {   // one coalesced read of the edge list instead of a serial scan of global memory
        s_raw[k] = (int)rel_ind[((size_t)b * K + k) * 2];
        s_raw[K + k] = (int)rel_ind[((size_t)b * K + k) * 2 + 1];
    }
    __syncthreads();
    if (threadIdx.x == 0 || threadIdx.x == 32) {  // K is small (65): a serial scan (of shared memory) keeps the ascending edge order
        const int which = threadIdx.x >> 5;
        int c = 0;
        for (int k = 0; k < K; ++k)
            if (s_raw[which * K + k] == n) s_list[which * K + c++] = k;
        s_cnt[which] = c;
    }
    __syncthreads();
    const int cs = s_cnt[0], co = s_cnt[1];
    const float ds = (float)cs + 1e-7f, dob = (float)co + 1e-7f;
    const float* ms = m_subj + (size_t)b * K * ldm;
    const float* mo = m_obj + (size_t)b * K * ldm;
    if ((L & 3) == 0) {
        const int L4 = L >> 2;
        for (int c4 = threadIdx.x; c4 < L4; c4 += blockDim.x) {
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            for (int i = 0; i < cs; i += 4) {   // four independent 16-byte loads in flight, added in ascending edge order
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = (i + u < cs) ? __ldg(reinterpret_cast<const float4*>(ms + (size_t)s_list[i + u] * ldm) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u < cs) { a0.x += v[u].x; a0.y += v[u].y; a0.z += v[u].z; a0.w += v[u].w; }
            }
            for (int i = 0; i < co; i += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    v[u] = (i + u < co) ? __ldg(reinterpret_cast<const float4*>(mo + (size_t)s_list[K + i + u] * ldm) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u < co) { a1.x += v[u].x; a1.y += v[u].y; a1.z += v[u].z; a1.w += v[u].w; }
            }
            float4 v;
            a0.x *= mscale; a0.y *= mscale; a0.z *= mscale; a0.w *= mscale; a1.x *= mscale; a1.y *= mscale; a1.z *= mscale; a1.w *= mscale;
            v.x = (fmaxf(a0.x / ds, 0.f) + fmaxf(a1.x / dob, 0.f)) * 0.5f;
            v.y = (fmaxf(a0.y / ds, 0.f) + fmaxf(a1.y / dob, 0.f)) * 0.5f;
            v.z = (fmaxf(a0.z / ds, 0.f) + fmaxf(a1.z / dob, 0.f)) * 0.5f;
            v.w = (fmaxf(a0.w / ds, 0.f) + fmaxf(a1.w / dob, 0.f)) * 0.5f;
            if (res) {
                const float4 rr = __ldg(reinterpret_cast<const float4*>(res + (size_t)bn * L) + c4);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            reinterpret_cast<float4*>(out + (size_t)bn * L)[c4] = v;
        }
        return;
    }
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        for (int i = 0; i < cs; ++i) a0 += __ldg(ms + (size_t)s_list[i] * ldm + c);
        for (int i = 0; i < co; ++i) a1 += __ldg(mo + (size_t)s_list[K + i] * ldm + c);
        float v = (fmaxf(a0 * mscale / ds, 0.f) + fmaxf(a1 * mscale / dob, 0.f)) * 0.5f;
        if (res) v += __ldg(res + (size_t)bn * L + c);
        out[(size_t)bn * L + c] = v;
    }
}

// Aggregate-then-transform for the node <- edges direction with folded unit weights: a unit's message is LINEAR in its source row
// (graph_conv_unit.py:29-35: collect = A (W src + b), collect / (rowsum(A) + 1e-7)), so the segment mean over the edges of a node is taken
// BEFORE the contraction: agg_s[b,n] = sum_{k: s_k = n} p[b,k] / (c_s + 1e-7) (ascending edge order), likewise agg_o; the contraction then
// runs over 37 node rows per image instead of 65 edge rows (-43 % FLOPs for the layer), and the bias enters as b * c / (c + 1e-7).
// m2 != nullptr (p == nullptr): the edge stream is not materialised at all -- p[b,k] = 0.5 * (relu(Ma[b,s_k] * msc / (1+1e-7)) +
// relu(Mb[b,o_k] * msc / (1+1e-7))) is evaluated on the fly from the node messages m2 [B*N, 2L] = [Ma | Mb] of the previous layer's
// edge <- node direction (same summation order as gcn_edge_update_kernel followed by this kernel; the division by 1 + 1e-7 is a multiplication).
__global__ void __launch_bounds__(256) gcn_edge_aggregate_kernel(const float* __restrict__ p, const float* __restrict__ m2, float msc,
                                                                 const long long* __restrict__ rel_ind, float* __restrict__ agg_s,
                                                                 float* __restrict__ agg_o, float* __restrict__ ratio, unsigned short* __restrict__ s16_hi,
                                                                 unsigned short* __restrict__ s16_lo, unsigned short* __restrict__ o16_hi,
                                                                 unsigned short* __restrict__ o16_lo, int* __restrict__ overflow, int N, int K, int L) {
    extern __shared__ int s_list[];  // [2][K] edge lists of this node (ascending) | [2][K] the OTHER end node of those edges
    __shared__ int s_cnt[2];
    __shared__ int s_wcnt[2][8];
    int* s_other = s_list + 2 * K;
    const int bn = blockIdx.x;
    const int b = bn / N, n = bn - b * N;
    // order-preserving compaction of the edges whose subject / object is this node: ballot + prefix counts (K <= 256)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int k = threadIdx.x;
    const int ek_s = k < K ? (int)rel_ind[((size_t)b * K + k) * 2] : -1, ek_o = k < K ? (int)rel_ind[((size_t)b * K + k) * 2 + 1] : -1;
    const bool fs = k < K && ek_s == n;
    const bool fo = k < K && ek_o == n;
    const unsigned ms_ = __ballot_sync(0xffffffffu, fs), mo_ = __ballot_sync(0xffffffffu, fo);
    if (lane == 0) { s_wcnt[0][wid] = __popc(ms_); s_wcnt[1][wid] = __popc(mo_); }
    __syncthreads();
    {
        int off_s = 0, off_o = 0;
        for (int w2 = 0; w2 < wid; ++w2) { off_s += s_wcnt[0][w2]; off_o += s_wcnt[1][w2]; }
        const unsigned lt = (1u << lane) - 1u;
        if (fs) { s_list[off_s + __popc(ms_ & lt)] = k; s_other[off_s + __popc(ms_ & lt)] = ek_o; }
        if (fo) { s_list[K + off_o + __popc(mo_ & lt)] = k; s_other[K + off_o + __popc(mo_ & lt)] = ek_s; }
        if (threadIdx.x == 0) {
            int cs_ = 0, co_ = 0;
            for (int w2 = 0; w2 < 8; ++w2) { cs_ += s_wcnt[0][w2]; co_ += s_wcnt[1][w2]; }
            s_cnt[0] = cs_; s_cnt[1] = co_;
        }
    }
    __syncthreads();
    const int cs = s_cnt[0], co = s_cnt[1];
    const float ds = (float)cs + 1e-7f, dob = (float)co + 1e-7f;
    if (threadIdx.x == 0) { ratio[2 * (size_t)bn] = (float)cs / ds; ratio[2 * (size_t)bn + 1] = (float)co / dob; }
    const float* pb = p ? p + (size_t)b * K * L : nullptr;
    const float* mb = m2 ? m2 + (size_t)b * N * 2 * L : nullptr;
    // the unit's divisor for one source row is 1 + 1e-7 (= 1 + 2^-23 in fp32): folded with the power-of-two message scale into one factor
    // (x * mf and (x * msc) / (1 + 2^-23) are both correctly rounded values of reals 2^-46 x apart)
    const float mf = (float)((double)msc / (double)(1.f + 1e-7f));
    const int L4 = L >> 2;   // L % 4 == 0 (checked by the caller)
    for (int c4 = threadIdx.x; c4 < L4; c4 += blockDim.x) {
        float4 own_a = make_float4(0.f, 0.f, 0.f, 0.f), own_b = own_a;   // relu(Ma[n] * msc / d1), relu(Mb[n] * msc / d1)
        if (mb) {
            const float4 ta = __ldg(reinterpret_cast<const float4*>(mb + (size_t)n * 2 * L) + c4), tb = __ldg(reinterpret_cast<const float4*>(mb + (size_t)n * 2 * L + L) + c4);
            own_a = make_float4(fmaxf(ta.x * mf, 0.f), fmaxf(ta.y * mf, 0.f), fmaxf(ta.z * mf, 0.f), fmaxf(ta.w * mf, 0.f));
            own_b = make_float4(fmaxf(tb.x * mf, 0.f), fmaxf(tb.y * mf, 0.f), fmaxf(tb.z * mf, 0.f), fmaxf(tb.w * mf, 0.f));
        }
#pragma unroll
        for (int which = 0; which < 2; ++which) {
            const int cnt = which ? co : cs;
            const int* lst = s_list + which * K;
            const int* oth = s_other + which * K;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = 0; i < cnt; i += 8) {   // eight independent 16-byte loads in flight, added in ascending edge order
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (i + u >= cnt) { v[u] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
                    if (pb) v[u] = __ldg(reinterpret_cast<const float4*>(pb + (size_t)lst[i + u] * L) + c4);
                    else    v[u] = __ldg(reinterpret_cast<const float4*>(mb + (size_t)oth[i + u] * 2 * L + (which ? 0 : L)) + c4);   // the other end's message
                }
                if (!pb) {   // edge value = 0.5 * (relu(Ma[s_k]) + relu(Mb[o_k])): this node is s_k (which = 0) or o_k (which = 1)
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        if (i + u >= cnt) continue;
                        const float4 t = make_float4(fmaxf(v[u].x * mf, 0.f), fmaxf(v[u].y * mf, 0.f), fmaxf(v[u].z * mf, 0.f), fmaxf(v[u].w * mf, 0.f));
                        const float4 sa = which ? t : own_a, sb = which ? own_b : t;
                        v[u] = make_float4(0.5f * (sa.x + sb.x), 0.5f * (sa.y + sb.y), 0.5f * (sa.z + sb.z), 0.5f * (sa.w + sb.w));
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (i + u < cnt) { a.x += v[u].x; a.y += v[u].y; a.z += v[u].z; a.w += v[u].w; }
            }
            const float dv = which ? dob : ds;
            a.x /= dv; a.y /= dv; a.z /= dv; a.w /= dv;
            if (agg_s) reinterpret_cast<float4*>((which ? agg_o : agg_s) + (size_t)bn * L)[c4] = a;   // fp32 copy only when a contraction reads it
            unsigned short* h16 = which ? o16_hi : s16_hi;
            if (h16) {
                unsigned short* l16 = which ? o16_lo : s16_lo;
                unsigned short hh[4], hl[4];
                int ovf = 0;
                split_f16(a.x, hh[0], hl[0], ovf); split_f16(a.y, hh[1], hl[1], ovf); split_f16(a.z, hh[2], hl[2], ovf); split_f16(a.w, hh[3], hl[3], ovf);
                const size_t o = (size_t)bn * L + 4 * c4;
                *reinterpret_cast<uint2*>(h16 + o) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
                *reinterpret_cast<uint2*>(l16 + o) = make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16));
                if (ovf && overflow) atomicOr(overflow, 1);
            }
        }
    }
}
// x_new[bn, c] = 0.5 * ( relu((Y_s + b0 * r_s) * inv) + relu((Y_o + b1 * r_o) * inv) ) (+ residual);  y [B*N, 2L] = [Y_s | Y_o], bias [2L]
__global__ void __launch_bounds__(256) gcn_node_finish_kernel(const float* __restrict__ y, const float* __restrict__ bias, const float* __restrict__ ratio,
                                                              float inv, const float* __restrict__ res, float* __restrict__ out, int L) {
    const size_t bn = blockIdx.x;
    const float rs = ratio[2 * bn], ro = ratio[2 * bn + 1];
    const float* yr = y + bn * 2 * L;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        const float vs = fmaxf((yr[c] + __ldg(bias + c) * rs) * inv, 0.f);
        const float vo = fmaxf((yr[L + c] + __ldg(bias + L + c) * ro) * inv, 0.f);
        float v = (vs + vo) * 0.5f;
        if (res) v += res[bn * L + c];
        out[bn * L + c] = v;
    }
}

static int check_dims(const subgc_dims* d) {
    SUBGC_CHECK_ARG(d != nullptr, "dims is null");
    SUBGC_CHECK_ARG(d->gcn > 0 && d->low_rank > 0 && d->att_feat > 0 && d->embed > 0 && d->obj_num > 1 && d->rel_num > 0 &&
                        d->obj_classes > 1 && d->pred_classes > 1,
                    "bad encoder dims");
    SUBGC_CHECK_ARG(d->gcn_layers >= 0 && d->gcn_layers <= SUBGC_MAX_GCN_LAYERS && d->gcn_residual >= 1, "bad gcn_layers/gcn_residual");
    return SUBGC_OK;
}

// which layer inputs are live: need_x[l] / need_p[l] = the node / edge stream entering layer l (l == layers: outputs)
static void gcn_liveness(const subgc_dims* d, int want_x_pred, bool* need_x, bool* need_p) {
    const int Ln = d->gcn_layers, R = d->gcn_residual;
    for (int l = 0; l <= Ln; ++l) need_x[l] = need_p[l] = false;
    need_x[Ln] = true;
    need_p[Ln] = want_x_pred != 0;
    for (int l = Ln; l >= 1; --l) {
        if (need_x[l]) {
            need_p[l - 1] = true;                  // units 0,1 read the edge stream
            if (l % R == 0) need_x[l - R] = true;  // residual anchor
        }
        if (need_p[l]) {
            need_x[l - 1] = true;                  // units 2,3 read the node stream
            if (l % R == 0) need_p[l - R] = true;
        }
    }
}

static size_t gcn_ws_bytes(const subgc_dims* d, int B) {
    const size_t rows = (size_t)B * (d->obj_num > d->rel_num ? d->obj_num : d->rel_num);
    size_t b = 0;
    b += align_up(rows * d->low_rank * 4, 256);          // T
    b += 2 * align_up(rows * d->gcn * 4, 256);           // two message buffers
    b += (size_t)d->gcn_layers * (align_up((size_t)B * d->obj_num * d->gcn * 4, 256) + align_up((size_t)B * d->rel_num * d->gcn * 4, 256));
    size_t g1 = gemm_workspace_bytes((int)rows, d->low_rank, d->gcn), g2 = gemm_workspace_bytes((int)rows, d->gcn, d->low_rank);
    const size_t g3 = gemm_workspace_bytes((int)rows, 2 * d->gcn, d->gcn);   // folded direction: both units in one contraction
    if (g3 > g1) g1 = g3;
    b += align_up(g1 > g2 ? g1 : g2, 256) + 1024;
    b += 2 * align_up(rows * (size_t)((d->gcn + 7) & ~7) * 2, 256) + 1024;   // split-fp16 copy of a layer input shared by two units
    b += 2 * align_up(rows * (size_t)((d->low_rank + 7) & ~7) * 2, 256) + 1024;   // split-fp16 copy of T written by the fc_lft contraction
    b += 4 * align_up((size_t)B * d->rel_num * (size_t)((d->gcn + 7) & ~7) * 2, 256) + 1024;   // split-fp16 copies (two sets) of the edge stream written by the edge update kernel
    // aggregate-then-transform (folded node <- edges direction): two aggregated inputs, their split copies, count ratios
    b += 2 * align_up((size_t)B * d->obj_num * d->gcn * 4, 256) + 4 * align_up((size_t)B * d->obj_num * d->gcn * 2, 256) + align_up((size_t)B * d->obj_num * 8, 256) + 1024;
    return b;
}

static size_t fuse_ws_bytes(const subgc_dims* d, int B) {
    size_t b = 0;
    b += align_up((size_t)B * d->obj_num * 8, 256) + align_up((size_t)B * d->rel_num * 8, 256);
    size_t g1 = gemm_workspace_bytes(B * d->obj_num, d->gcn, d->att_feat + d->embed), g2 = gemm_workspace_bytes(B * d->rel_num, d->gcn, d->embed);
    b += align_up(g1 > g2 ? g1 : g2, 256) + 1024;
    return b;
}

static int linear(const subgc_weights* w, const float* A, int M, int K, const subgc_linear& lin, int N, float* C, Workspace& ws, cudaStream_t st,
                  const unsigned short* a16_hi = nullptr, const unsigned short* a16_lo = nullptr, int lda16 = 0, unsigned short* c16_hi = nullptr,
                  unsigned short* c16_lo = nullptr, int ld16 = 0) {
    GemmProblem p;
    p.wts = w;
    p.M = M; p.N = N; p.nseg = 1;
    p.seg[0] = make_seg(A, K, lin.w, K, K);
    if (a16_hi) { p.seg[0].A16_hi = a16_hi; p.seg[0].A16_lo = a16_lo; p.seg[0].lda16 = lda16; }   // shared split copy of A
    p.epi.bias = lin.b;
    p.epi.c16_hi = c16_hi; p.epi.c16_lo = c16_lo; p.epi.ld16 = ld16;                               // split copy of the result for the next contraction
    p.C = C; p.ldc = N;
    return launch_gemm(p, ws.cursor(), ws.remaining(), st);
}

}  // namespace subgc

using namespace subgc;

extern "C" size_t subgc_encoder_workspace_bytes(const subgc_dims* d, int n_images) {
    if (!d || n_images <= 0) return 0;
    size_t a = fuse_ws_bytes(d, n_images), b = gcn_ws_bytes(d, n_images);
    return a > b ? a : b;
}

extern "C" int subgc_gcn_needs_pred(const subgc_dims* d, int want_x_pred) {
    if (!d || d->gcn_layers < 0 || d->gcn_layers > SUBGC_MAX_GCN_LAYERS || d->gcn_residual < 1) return 1;
    bool nx[SUBGC_MAX_GCN_LAYERS + 1], np[SUBGC_MAX_GCN_LAYERS + 1];
    gcn_liveness(d, want_x_pred, nx, np);
    return np[0] ? 1 : 0;
}

// obj_cls / pred_cls (nullable, int64 device): class ids computed by the caller (loader-side compaction); when null they are taken
// from the score tensors here
static int fuse_nodes_impl(const subgc_dims* d, const subgc_weights* w, int n_images, const float* att_feats, const float* obj_dist,
                           const float* pred_dist, const long long* obj_cls, const long long* pred_cls, float* x0, float* p0, void* ws_,
                           size_t ws_bytes, cudaStream_t st) {
    Workspace ws(ws_, ws_bytes);
    const int rows_n = n_images * d->obj_num, rows_k = n_images * d->rel_num;
    long long* cls = ws.take<long long>(rows_n);
    long long* pcls = ws.take<long long>(rows_k);
    if (!ws.ok()) { set_error("subgc_fuse_nodes: workspace too small"); return SUBGC_E_WORKSPACE; }
    if (!obj_cls) {
        class_argmax_kernel<<<(rows_n + 7) / 8, 256, 0, st>>>(obj_dist, rows_n, d->obj_classes, 1, cls);
        SUBGC_LAUNCH_CHECK();
        obj_cls = cls;
    }
    {
        GemmProblem p;
        p.wts = w;
        p.M = rows_n; p.N = d->gcn; p.nseg = 2;
        p.seg[0] = make_seg(att_feats, d->att_feat, w->obj_v_proj.w, d->att_feat, d->att_feat);
        p.seg[1] = make_seg(w->sg_obj_embed, d->embed, w->obj_emb_proj.w, d->embed, d->embed);
        p.seg[1].gather = obj_cls;
        p.epi.bias = w->obj_v_proj.b;
        p.epi.bias2 = w->obj_emb_proj.b;
        p.epi.relu = 1;
        p.C = x0; p.ldc = d->gcn;
        SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    }
    if (p0) {
        if (!pred_cls) {
            class_argmax_kernel<<<(rows_k + 7) / 8, 256, 0, st>>>(pred_dist, rows_k, d->pred_classes, d->pred_emb_type == 1 ? 1 : 0, pcls);
            SUBGC_LAUNCH_CHECK();
            pred_cls = pcls;
        }
        GemmProblem p;
        p.wts = w;
        p.M = rows_k; p.N = d->gcn; p.nseg = 1;
        p.seg[0] = make_seg(w->sg_pred_embed, d->embed, w->pred_emb_prj.w, d->embed, d->embed);
        p.seg[0].gather = pred_cls;
        p.epi.bias = w->pred_emb_prj.b;
        p.C = p0; p.ldc = d->gcn;
        SUBGC_TRY(launch_gemm(p, ws.cursor(), ws.remaining(), st));
    }
    return SUBGC_OK;
}

extern "C" int subgc_fuse_nodes(const subgc_dims* d, const subgc_weights* w, int n_images, const float* att_feats, const float* obj_dist,
                                const float* pred_dist, float* x0, float* p0, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_dims(d));
    SUBGC_CHECK_ARG(w && att_feats && obj_dist && x0 && n_images > 0, "subgc_fuse_nodes: null argument");
    SUBGC_CHECK_ARG(p0 == nullptr || pred_dist != nullptr, "subgc_fuse_nodes: p0 requested without pred_dist");
    SUBGC_CHECK_ARG(d->pred_emb_type == 1 || d->pred_emb_type == 2, "subgc_fuse_nodes: pred_emb_type must be 1 or 2");
    return fuse_nodes_impl(d, w, n_images, att_feats, obj_dist, pred_dist, nullptr, nullptr, x0, p0, ws_, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int subgc_fuse_nodes_cls(const subgc_dims* d, const subgc_weights* w, int n_images, const float* att_feats, const int64_t* obj_cls,
                                    const int64_t* pred_cls, float* x0, float* p0, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_dims(d));
    SUBGC_CHECK_ARG(w && att_feats && obj_cls && x0 && n_images > 0, "subgc_fuse_nodes_cls: null argument");
    SUBGC_CHECK_ARG(p0 == nullptr || pred_cls != nullptr, "subgc_fuse_nodes_cls: p0 requested without pred_cls");
    return fuse_nodes_impl(d, w, n_images, att_feats, nullptr, nullptr, reinterpret_cast<const long long*>(obj_cls),
                           reinterpret_cast<const long long*>(pred_cls), x0, p0, ws_, ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int subgc_gcn_forward(const subgc_dims* d, const subgc_weights* w, int n_images, const float* x0, const float* p0,
                                 const int64_t* rel_ind, float* x_obj, float* x_pred, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_dims(d));
    SUBGC_CHECK_ARG(w && x0 && rel_ind && x_obj && n_images > 0, "subgc_gcn_forward: null argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int B = n_images, N = d->obj_num, K = d->rel_num, L = d->gcn, R = d->low_rank, Ln = d->gcn_layers;
    bool need_x[SUBGC_MAX_GCN_LAYERS + 1], need_p[SUBGC_MAX_GCN_LAYERS + 1];
    gcn_liveness(d, x_pred != nullptr, need_x, need_p);
    SUBGC_CHECK_ARG(!need_p[0] || p0 != nullptr, "subgc_gcn_forward: this configuration needs the predicate embedding p0");
    const size_t xn = (size_t)B * N * L, pn = (size_t)B * K * L;
    if (Ln == 0) {
        SUBGC_CUDA(cudaMemcpyAsync(x_obj, x0, xn * 4, cudaMemcpyDeviceToDevice, st));
        if (x_pred) SUBGC_CUDA(cudaMemcpyAsync(x_pred, p0, pn * 4, cudaMemcpyDeviceToDevice, st));
        return SUBGC_OK;
    }
    Workspace ws(ws_, ws_bytes);
    const size_t rows_max = (size_t)B * (N > K ? N : K);
    float* T = ws.take<float>(rows_max * R);
    float* M2 = ws.take<float>(2 * rows_max * L);   // [rows, L] x 2 (one per unit), or [rows, 2L] when a direction is one folded contraction
    float* Ma = M2;
    float* Mb = M2 ? M2 + rows_max * L : nullptr;
    // both units of a direction contract the same layer input: split it once (split-fp16 path), in a region reserved up front
    const size_t split_bytes = 2 * align_up(rows_max * (size_t)((L + 7) & ~7) * 2, 256) + 1024;
    char* split_region = ws.take<char>(split_bytes);
    // split-fp16 copies that their producers write themselves (only when the packed path is in use): T by the fc_lft contraction's
    // epilogue, the edge stream of the next layer by the edge update kernel
    const bool fuse16 = w->packs != nullptr && w->n_packs > 0 && (R & 7) == 0 && (L & 7) == 0;
    unsigned short* t16_hi = fuse16 ? ws.take<unsigned short>(rows_max * R) : nullptr;
    unsigned short* t16_lo = fuse16 ? ws.take<unsigned short>(rows_max * R) : nullptr;
    // two sets: a layer's edge update writes the NEXT layer's copy while its node update still reads the current one
    unsigned short* n16_hi[2] = {fuse16 ? ws.take<unsigned short>((size_t)B * K * L) : nullptr, fuse16 ? ws.take<unsigned short>((size_t)B * K * L) : nullptr};
    unsigned short* n16_lo[2] = {fuse16 ? ws.take<unsigned short>((size_t)B * K * L) : nullptr, fuse16 ? ws.take<unsigned short>((size_t)B * K * L) : nullptr};
    const unsigned short *cur16_hi = nullptr, *cur16_lo = nullptr;   // split copy of the current edge stream p, if its producer wrote one
    bool skip_edges = false;   // the previous layer left its edge stream un-materialised (node messages still in M2)
    float skip_scale = 1.f;
    // aggregate-then-transform buffers (folded node <- edges direction)
    static const bool agg_off = getenv("SUBGC_GCN_AGG") != nullptr && getenv("SUBGC_GCN_AGG")[0] == '0';
    bool any_fold0 = false;
    for (int l = 0; l < Ln; ++l) any_fold0 = any_fold0 || (w->gcn_fold[l][0].w != nullptr && w->gcn_fold_scale[l][0] > 0.f);
    const bool agg = any_fold0 && !agg_off && (L & 3) == 0;
    float* agg_s = agg ? ws.take<float>(xn) : nullptr;
    float* agg_o = agg ? ws.take<float>(xn) : nullptr;
    float* agg_ratio = agg ? ws.take<float>((size_t)2 * B * N) : nullptr;
    unsigned short* ag16[4] = {nullptr, nullptr, nullptr, nullptr};   // s_hi, s_lo, o_hi, o_lo
    if (agg && fuse16)
        for (int i = 0; i < 4; ++i) ag16[i] = ws.take<unsigned short>(xn);
    const float* x = x0;
    const float* p = p0;
    const float* x_res = x0;
    const float* p_res = p0;
    const long long* rel = reinterpret_cast<const long long*>(rel_ind);
    for (int l = 0; l < Ln; ++l) {
        const bool last = (l == Ln - 1), boundary = ((l + 1) % d->gcn_residual == 0);
        bool next16 = false, skip_next = false;
        float* x_next = nullptr;
        float* p_next = nullptr;
        if (need_x[l + 1]) x_next = last ? x_obj : ws.take<float>(xn);
        if (need_p[l + 1]) p_next = last ? x_pred : ws.take<float>(pn);
        if (!ws.ok() || !T || !Ma || !Mb) { set_error("subgc_gcn_forward: workspace too small"); return SUBGC_E_WORKSPACE; }
        if (p_next) {  // units 2,3: edge <- node (graph_conv.py:29-33)
            const unsigned short *xh = nullptr, *xl = nullptr;
            int xld = 0;
            if (split_region) { Workspace sw(split_region, split_bytes); if (!h3_presplit(x, B * N, L, L, w, sw, st, &xh, &xl, &xld)) xh = xl = nullptr; }
            // fc_rgt(fc_lft(.)) has no non-linearity in between (graph_conv_unit.py:29-31): when the host supplies the folded weights
            // [W_rgt W_lft of unit 2 ; of unit 3] (x 2^s, undone exactly by the update kernel), the direction is ONE contraction
            const subgc_linear& fold = w->gcn_fold[l][1];
            const bool folded = fold.w != nullptr && w->gcn_fold_scale[l][1] > 0.f;
            if (folded) {
                SUBGC_TRY(linear(w, x, B * N, L, fold, 2 * L, M2, ws, st, xh, xl, xld));
            } else {
                SUBGC_TRY(linear(w, x, B * N, L, w->gcn_lft[l][2], R, T, ws, st, xh, xl, xld, t16_hi, t16_lo, R));
                SUBGC_TRY(linear(w, T, B * N, R, w->gcn_rgt[l][2], L, Ma, ws, st, t16_hi, t16_lo, R));
                SUBGC_TRY(linear(w, x, B * N, L, w->gcn_lft[l][3], R, T, ws, st, xh, xl, xld, t16_hi, t16_lo, R));
                SUBGC_TRY(linear(w, T, B * N, R, w->gcn_rgt[l][3], L, Mb, ws, st, t16_hi, t16_lo, R));
            }
            next16 = fuse16 && !last && need_x[l + 2];   // layer l+1 contracts this edge stream (its units 0, 1 produce x of layer l+2)
            if (agg && !last && w->gcn_fold[l + 1][0].w != nullptr && w->gcn_fold_scale[l + 1][0] > 0.f) next16 = false;   // ... unless it aggregates first
            // The edge stream p(l+1) need not exist when its only reader is the next (= last) layer's aggregation, which can evaluate an
            // edge's value from the node messages in M2 itself: no residual at this layer, no x_pred output, no later residual anchor
            static const bool noskip = getenv("SUBGC_GCN_KEEP_EDGES") != nullptr;
            skip_next = folded && agg && !noskip && !boundary && x_pred == nullptr && l + 1 == Ln - 1 && need_x[l + 2] && !need_p[l + 2] &&
                         w->gcn_fold[l + 1][0].w != nullptr && w->gcn_fold_scale[l + 1][0] > 0.f;
            skip_scale = folded ? 1.f / w->gcn_fold_scale[l][1] : 1.f;
            if (!skip_next) {
                gcn_edge_update_kernel<<<B * K, 256, 0, st>>>(folded ? M2 : Ma, folded ? M2 + L : Mb, rel, boundary ? p_res : nullptr, p_next, B, N, K, L,
                                                              next16 ? n16_hi[l & 1] : nullptr, next16 ? n16_lo[l & 1] : nullptr, L, w->h3_overflow,
                                                              folded ? 2 * L : L, folded ? 1.f / w->gcn_fold_scale[l][1] : 1.f);
                SUBGC_LAUNCH_CHECK();
            }
        }
        if (x_next) {  // units 0,1: node <- edges (graph_conv.py:22-26)
            const subgc_linear& fold = w->gcn_fold[l][0];
            const bool folded = fold.w != nullptr && w->gcn_fold_scale[l][0] > 0.f;
            const unsigned short *ph = nullptr, *pl = nullptr;
            int pld = 0;
            if (folded && agg) {
                // the aggregation kernel reads the fp32 edge stream: no split copy of p is needed
            } else if (cur16_hi) {   // the edge update of the previous layer already wrote the split copy of this input
                ph = cur16_hi; pl = cur16_lo; pld = L;
            } else if (split_region) {
                Workspace sw(split_region, split_bytes);
                if (!h3_presplit(p, B * K, L, L, w, sw, st, &ph, &pl, &pld)) ph = pl = nullptr;
            }
            if (folded && agg) {
                // aggregate first (37 node rows per image instead of 65 edge rows), one contraction per unit, bias / ReLU / mean of both after
                GemmProblem gp[2];
                bool all16 = ag16[0] != nullptr;
                for (int u = 0; u < 2; ++u) {
                    gp[u].M = B * N; gp[u].N = L; gp[u].nseg = 1;
                    gp[u].seg[0] = make_seg(u ? agg_o : agg_s, L, fold.w + (size_t)u * L * L, L, L);
                    if (ag16[0]) { gp[u].seg[0].A16_hi = ag16[2 * u]; gp[u].seg[0].A16_lo = ag16[2 * u + 1]; gp[u].seg[0].lda16 = L; }
                    gp[u].C = M2 + (size_t)u * L; gp[u].ldc = 2 * L;
                    resolve_packs(gp[u], w);
                    all16 = all16 && h3_eligible(gp[u]);   // the contraction will read the split copy only
                }
                SUBGC_CHECK_ARG(K <= 256, "subgc_gcn_forward: at most 256 edges per image");
                gcn_edge_aggregate_kernel<<<B * N, 256, 4 * K * sizeof(int), st>>>(skip_edges ? nullptr : p, skip_edges ? M2 : nullptr, skip_scale, rel,
                                                                                   all16 ? nullptr : agg_s, all16 ? nullptr : agg_o, agg_ratio, ag16[0],
                                                                                   ag16[1], ag16[2], ag16[3], w->h3_overflow, N, K, L);
                SUBGC_LAUNCH_CHECK();
                for (int u = 0; u < 2; ++u) SUBGC_TRY(launch_gemm(gp[u], ws.cursor(), ws.remaining(), st));
                gcn_node_finish_kernel<<<B * N, 256, 0, st>>>(M2, fold.b, agg_ratio, 1.f / w->gcn_fold_scale[l][0], boundary ? x_res : nullptr, x_next, L);
                SUBGC_LAUNCH_CHECK();
            } else if (folded) {
                SUBGC_TRY(linear(w, p, B * K, L, fold, 2 * L, M2, ws, st, ph, pl, pld));
            } else {
                SUBGC_TRY(linear(w, p, B * K, L, w->gcn_lft[l][0], R, T, ws, st, ph, pl, pld, t16_hi, t16_lo, R));
                SUBGC_TRY(linear(w, T, B * K, R, w->gcn_rgt[l][0], L, Ma, ws, st, t16_hi, t16_lo, R));
                SUBGC_TRY(linear(w, p, B * K, L, w->gcn_lft[l][1], R, T, ws, st, ph, pl, pld, t16_hi, t16_lo, R));
                SUBGC_TRY(linear(w, T, B * K, R, w->gcn_rgt[l][1], L, Mb, ws, st, t16_hi, t16_lo, R));
            }
            if (!(folded && agg)) {
                gcn_node_update_kernel<<<B * N, 256, 4 * K * sizeof(int), st>>>(folded ? M2 : Ma, folded ? M2 + L : Mb, rel, boundary ? x_res : nullptr, x_next,
                                                                                B, N, K, L, folded ? 2 * L : L, folded ? 1.f / w->gcn_fold_scale[l][0] : 1.f);
                SUBGC_LAUNCH_CHECK();
            }
        }
        cur16_hi = next16 ? n16_hi[l & 1] : nullptr; cur16_lo = next16 ? n16_lo[l & 1] : nullptr;
        skip_edges = skip_next;
        x = x_next; p = p_next;
        if (boundary) { x_res = x; p_res = p; }
    }
    return SUBGC_OK;
}

// Full-GC read-out (models/AttModel.py:146,200,265: torch.mean(att_feats, 1)): out[b, c] = mean over the N nodes of x[b, n, c]
__global__ void __launch_bounds__(256) mean_nodes_kernel(const float* __restrict__ x, float* __restrict__ out, int N, int L) {
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float a = 0.f;
        for (int n = 0; n < N; ++n) a += x[((size_t)b * N + n) * L + c];   // ascending node order, as torch's reduction over a short dim
        out[(size_t)b * L + c] = a / (float)N;
    }
}
extern "C" int subgc_mean_nodes(int B, int N, int L, const float* x, float* out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(x && out && B > 0 && N > 0 && L > 0, "subgc_mean_nodes: bad arguments");
    mean_nodes_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, N, L);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_class_argmax(int rows, int n_classes, int skip_first, const float* dist, int64_t* cls, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(dist && cls && rows > 0 && n_classes > skip_first, "subgc_class_argmax: bad arguments");
    class_argmax_kernel<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(dist, rows, n_classes, skip_first ? 1 : 0,
                                                                                      reinterpret_cast<long long*>(cls));
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_gcn_edge_fwd(int B, int N, int K, int L, const float* m_subj, const float* m_obj, const int64_t* rel_ind, const float* res,
                                  float* out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(m_subj && m_obj && rel_ind && out && B > 0, "subgc_gcn_edge_fwd: bad arguments");
    gcn_edge_update_kernel<<<B * K, 256, 0, static_cast<cudaStream_t>(stream)>>>(m_subj, m_obj, reinterpret_cast<const long long*>(rel_ind), res, out,
                                                                               B, N, K, L);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}
