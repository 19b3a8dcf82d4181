// Top-down attention-LSTM decoder: one step, the greedy / top-k sampling loop and batched beam search.
//
//   subgc_decode_step   <- AttModel.get_logprobs_state            (reference models/AttModel.py:328-341)
//                          TopDownCore.forward                    (reference models/AttModel.py:400-431)
//                          Attention.forward                      (reference models/AttModel.py:445-471)
//   subgc_decode_sample <- AttModel._sample, beam_size == 1       (reference models/AttModel.py:278-326)
//   subgc_decode_beam   <- AttModel._sample_sentences             (reference models/AttModel.py:208-234)
//                          CaptionModel.beam_search / beam_step   (reference models/CaptionModel.py:43-94,97-176)
//
// One decoder step is: att-LSTM gates (one contraction over the un-concatenated [h_lang | fc | relu(E[it])] and
// h_att segments) -> LSTM cell -> h2att -> fused attention (tanh / alpha / softmax / mask / renormalise / context,
// warp-shuffle reductions) -> lang-LSTM gates -> LSTM cell -> logit.  The loops keep token feedback, finish masks,
// the all-finished early exit and (for beam search) candidate ranking, history / state re-ordering and the
// done-beam lists on the device: no host synchronisation anywhere, the whole loop is graph-capturable.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace subgc {

// ---- split-K reduction + biases + LSTM cell (torch.nn.LSTMCell gate order i, f, g, o) ---------------------------
// gates[r, g*H + j] = sum_z part[z][r][g*H + j] + b_ih + b_hh (z ascending: deterministic); c' = s(f) c + s(i) tanh(g); h' = s(o) tanh(c')
// (<= 42 registers: six blocks per SM, so the whole grid is resident next to the contraction's CTAs and is released at once)
__global__ void __launch_bounds__(256, 6) lstm_reduce_cell_kernel(const float* __restrict__ part, int splits, const float* __restrict__ b_ih,
                                                               const float* __restrict__ b_hh, const float* __restrict__ c_prev,
                                                               const long long* __restrict__ parent, float* __restrict__ h_out,
                                                               float* __restrict__ c_out, int S, int H, const int* __restrict__ active,
                                                               const float* __restrict__ addend, int add_div,
                                                               unsigned short* __restrict__ h16_hi, unsigned short* __restrict__ h16_lo, int Hp,
                                                               TraceSlot trace, const float* __restrict__ part_b, int splits_b, int ld_b) {
    // part_b (nullable): a second set of split-K partials [splits_b][S][ld_b] whose first 4H columns (from the given pointer) belong
    // to the same gates (the merged h2att | lang-early contraction); added after `part`, in split order
    // h16_hi / h16_lo (nullable, [S, Hp]): the split-fp16 copy of h' the next contractions read as their activation operand
    // addend != nullptr: a pre-computed [S / add_div, 4H] term (the step-invariant fc segment with both biases folded in)
    // replaces b_ih + b_hh
    trace_begin(trace);
    pdl_trigger();
    pdl_wait();
    trace_released(trace);
    if (active != nullptr && *active == 0) return;
    const size_t zs = (size_t)S * 4 * H;
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * H) { trace_end(trace); return; }
    int r = idx / H, j = idx - r * H;
    const float* g = part + (size_t)r * 4 * H + j;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int z = 0; z < splits; ++z) {  // one hidden unit per thread keeps ~128 K threads in flight; 8 independent loads each
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += g[(size_t)z * zs + (size_t)q * H];
    }
    if (part_b) {
        const float* gb = part_b + (size_t)r * ld_b + j;
        const size_t zb = (size_t)S * ld_b;
#pragma unroll 2
        for (int z = 0; z < splits_b; ++z) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] += gb[(size_t)z * zb + (size_t)q * H];
        }
    }
    if (addend) {
        const float* a = addend + (size_t)(r / add_div) * 4 * H + j;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += a[(size_t)q * H];
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = (acc[q] + __ldg(b_ih + q * H + j)) + __ldg(b_hh + q * H + j);
    }
    long long pr = parent ? parent[r] : r;
    float c = sigmoidf_(acc[1]) * c_prev[(size_t)pr * H + j] + sigmoidf_(acc[0]) * tanhf(acc[2]);
    c_out[idx] = c;
    const float hv = sigmoidf_(acc[3]) * tanhf(c);
    h_out[idx] = hv;
    if (h16_hi) split_f16_store(hv, h16_hi, h16_lo, (size_t)r * Hp + j);
    trace_end(trace);
}

// ---- fused attention: one block per decode row -------------------------------------------------------------------
// atth = sum_z part[z][r] + b_h (h2att split-K partials reduced here); e_n = w . tanh(p_att[n] + atth) + b ;
// alpha = softmax_n(e) ; alpha *= mask ; alpha /= sum(alpha) ; ctx = sum_n alpha_n att[n]
// tanh for the attention scores: 18 K evaluations per row and step make libdevice's tanhf (~30 instructions) a third of the
// attention kernel.  (1 - e) / (1 + e) with e = exp(-2|x|) from the ex2 unit: absolute error <= ~2e-7 (tanhf: ~6e-8), far below the
// 2e-5 bar once weighted by alpha_net (|w| ~ 0.04) and summed.
__device__ __forceinline__ float tanh_score(float x) {
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

constexpr int kMaxBeamAtt = 8;    // beams per block of attention_beam_kernel (= kMaxBeam)
constexpr int kAttThreads = 1024;  // one block per row; the row is latency-bound (28 MB of att / p_att per step over 128 rows), so every
                                   // warp slot of the SM is used to keep loads in flight

__device__ __forceinline__ void attention_tail(float* s_e, float* s_c, int r, int cr, const float* __restrict__ att, const float* __restrict__ masks,
                                               float* __restrict__ ctx, float* __restrict__ att_w, int att_w_stride, int len_max, int H,
                                               unsigned short* __restrict__ c16_hi, unsigned short* __restrict__ c16_lo, int Hp,
                                               TraceSlot trace = TraceSlot{nullptr, 0}, int* overflow = nullptr, const float* s_mask = nullptr);

// body shared by attention_kernel and the fused att-phase kernel: expects s_h (atth of this row) and s_w (alpha_net weight)
// filled and synchronised
__device__ __forceinline__ void attention_row_body(float* s_h, float* s_w, float* s_e, float* s_c, int r, int cr, const float* __restrict__ p_att,
                                                   const float* __restrict__ att, const float* __restrict__ masks,
                                                   const float* __restrict__ alpha_b, float* __restrict__ ctx, float* __restrict__ att_w,
                                                   int att_w_stride, int len_max, int H, int AH, unsigned short* __restrict__ c16_hi,
                                                   unsigned short* __restrict__ c16_lo, int Hp) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* pa = p_att + (size_t)cr * len_max * AH;
    if ((AH & 3) == 0) {
        const int AH4 = AH >> 2;
        const float4* h4 = reinterpret_cast<const float4*>(s_h);
        const float4* w4 = reinterpret_cast<const float4*>(s_w);
        for (int n = wid; n < len_max; n += nw) {
            const float4* pr = reinterpret_cast<const float4*>(pa + (size_t)n * AH);
            float a = 0.f;
#pragma unroll 4
            for (int j4 = lane; j4 < AH4; j4 += 32) {
                const float4 v = __ldg(pr + j4), hh = h4[j4], ww = w4[j4];
                a = fmaf(ww.x, tanhf(v.x + hh.x), a);
                a = fmaf(ww.y, tanhf(v.y + hh.y), a);
                a = fmaf(ww.z, tanhf(v.z + hh.z), a);
                a = fmaf(ww.w, tanhf(v.w + hh.w), a);
            }
            a = warp_sum(a);
            if (lane == 0) s_e[n] = a + __ldg(alpha_b);
        }
    } else {
        for (int n = wid; n < len_max; n += nw) {
            float a = 0.f;
            for (int j = lane; j < AH; j += 32) a = fmaf(s_w[j], tanhf(__ldg(pa + (size_t)n * AH + j) + s_h[j]), a);
            a = warp_sum(a);
            if (lane == 0) s_e[n] = a + __ldg(alpha_b);
        }
    }
    __syncthreads();
    attention_tail(s_e, s_c, r, cr, att, masks, ctx, att_w, att_w_stride, len_max, H, c16_hi, c16_lo, Hp);
}

// softmax / mask / renormalise over the scores s_e[0..len_max) (filled and synchronised) and the context vector
__device__ __forceinline__ void attention_tail(float* s_e, float* s_c, int r, int cr, const float* __restrict__ att, const float* __restrict__ masks,
                                               float* __restrict__ ctx, float* __restrict__ att_w, int att_w_stride, int len_max, int H,
                                               unsigned short* __restrict__ c16_hi, unsigned short* __restrict__ c16_lo, int Hp, TraceSlot trace,
                                               int* overflow, const float* s_mask) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (wid == 0) {  // len_max <= 64: one warp finishes the softmax / mask / renormalise (two-stage, as the reference)
        float m = -INFINITY;
        for (int n = lane; n < len_max; n += 32) m = fmaxf(m, s_e[n]);
        m = warp_max(m);
        float sum = 0.f;
        for (int n = lane; n < len_max; n += 32) sum += expf(s_e[n] - m);
        sum = warp_sum(sum);
        float msum = 0.f;
        for (int n = lane; n < len_max; n += 32) {
            float wv = expf(s_e[n] - m) / sum;
            wv = wv * (s_mask ? s_mask[n] : __ldg(masks + (size_t)cr * len_max + n));
            s_e[n] = wv;
            msum += wv;
        }
        msum = warp_sum(msum);
        for (int n = lane; n < len_max; n += 32) {
            float wv = s_e[n] / msum;
            s_e[n] = wv;
            if (att_w) att_w[(size_t)r * att_w_stride + n] = wv;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_mark(trace, 2);   // softmax done
    // context: four thread groups take interleaved node subsets (n = g, g+4, ...), partials combined in fixed order
    const float* af = att + (size_t)cr * len_max * H;
    const int grp = threadIdx.x >> 8, tg = threadIdx.x & 255;
    if ((H & 3) == 0) {
        const int H4 = H >> 2;
        for (int j4 = tg; j4 < H4; j4 += 256) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 9
            for (int n = grp; n < len_max; n += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(af + (size_t)n * H) + j4);
                const float wv = s_e[n];
                a.x = fmaf(wv, v.x, a.x); a.y = fmaf(wv, v.y, a.y); a.z = fmaf(wv, v.z, a.z); a.w = fmaf(wv, v.w, a.w);
            }
            reinterpret_cast<float4*>(s_c + (size_t)grp * H)[j4] = a;
        }
    } else {
        for (int j = tg; j < H; j += 256) {
            float a = 0.f;
            for (int n = grp; n < len_max; n += 4) a = fmaf(s_e[n], __ldg(af + (size_t)n * H + j), a);
            s_c[(size_t)grp * H + j] = a;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        const float cv = ((s_c[j] + s_c[H + j]) + s_c[2 * H + j]) + s_c[3 * H + j];
        ctx[(size_t)r * H + j] = cv;
        if (c16_hi) split_f16_store(cv, c16_hi, c16_lo, (size_t)r * Hp + j, overflow);
    }
}

// (<= 40 registers: 1024 threads fit next to a contraction CTA, which then prefetches its weight ring while this kernel runs)
__global__ void __maxnreg__(48) attention_kernel(const float* __restrict__ atth_part, int splits, const float* __restrict__ h2att_b,
                                                                const float* __restrict__ p_att, const float* __restrict__ att,
                                                                const float* __restrict__ masks, const float* __restrict__ alpha_w,
                                                                const float* __restrict__ alpha_b, float* __restrict__ ctx, float* __restrict__ att_w,
                                                                int att_w_stride, int S, int len_max, int H, int AH, int rows_per_ctx,
                                                                const int* __restrict__ active, unsigned short* __restrict__ c16_hi,
                                                                unsigned short* __restrict__ c16_lo, int Hp, TraceSlot trace, int* overflow,
                                                                int ld_part, int early_ok) {
    // early_ok: p_att / masks were written before this launch sequence began (the decode loops: their buffers come from the prepare
    // stage and memset nodes separate it from the loop), so they may be read before the dependency wait.  A single step through the
    // C ABI (subgc_decode_step) may directly follow the kernel that produced them: there they are only read after the wait.
    trace_begin(trace);
    pdl_trigger();
    extern __shared__ float s_att[];  // [AH] atth | [AH] alpha_w | [64] e | [4][H] context partials (first 256 floats: score partials)
    float* s_h = s_att;
    float* s_w = s_att + AH;
    float* s_e = s_att + 2 * AH;
    float* s_c = s_att + 2 * AH + 64;
    const int r = blockIdx.x, cr = r / rows_per_ctx;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // p_att / att / alpha_net do not depend on this step.  Before the dependency wait: the p_att quads this thread scores go into
    // registers (work item = (node, 32-quad slice); a warp owns items wid, wid + nw, ...), att is requested into L2.
    constexpr int kMaxItems = 5;
    const int AH4 = AH >> 2, Q = AH4 >> 5;
    const int items = len_max * Q;
    const bool fast = (AH & 127) == 0 && items <= kMaxItems * nw && items <= 256;
    float4 pv[kMaxItems];
    if (fast && early_ok) {
        const float4* pa4 = reinterpret_cast<const float4*>(p_att + (size_t)cr * len_max * AH);
#pragma unroll
        for (int k = 0; k < kMaxItems; ++k) {
            const int item = wid + k * nw;
            pv[k] = item < items ? __ldg(pa4 + (size_t)(item / Q) * AH4 + (item % Q) * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    {
        const char* af = reinterpret_cast<const char*>(att + (size_t)cr * len_max * H);
        const size_t nb_a = (size_t)len_max * H * 4;
        for (size_t o = (size_t)threadIdx.x * 128; o < nb_a; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(af + o));
        if (!fast) {
            const char* pa = reinterpret_cast<const char*>(p_att + (size_t)cr * len_max * AH);
            const size_t nb_p = (size_t)len_max * AH * 4;
            for (size_t o = (size_t)threadIdx.x * 128; o < nb_p; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + o));
        }
    }
    for (int j = threadIdx.x; j < AH; j += blockDim.x) s_w[j] = __ldg(alpha_w + j);
    __shared__ float s_mask[64];   // the row's mask does not depend on the step either
    if (early_ok && threadIdx.x < len_max) s_mask[threadIdx.x] = __ldg(masks + (size_t)cr * len_max + threadIdx.x);
    pdl_wait();
    trace_released(trace);
    if (active != nullptr && *active == 0) return;
    if (!early_ok) {
        if (threadIdx.x < len_max) s_mask[threadIdx.x] = masks[(size_t)cr * len_max + threadIdx.x];
        if (fast) {
            const float4* pa4 = reinterpret_cast<const float4*>(p_att + (size_t)cr * len_max * AH);
#pragma unroll
            for (int k = 0; k < kMaxItems; ++k) {
                const int item = wid + k * nw;
                pv[k] = item < items ? pa4[(size_t)(item / Q) * AH4 + (item % Q) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    for (int j = threadIdx.x; j < AH; j += blockDim.x) {
        float a = 0.f;
        for (int z = 0; z < splits; ++z) a += atth_part[((size_t)z * S + r) * ld_part + j];
        s_h[j] = a + __ldg(h2att_b + j);
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_mark(trace, 0);   // atth reduced
    if (!fast) {
        attention_row_body(s_h, s_w, s_e, s_c, r, cr, p_att, att, masks, alpha_b, ctx, att_w, att_w_stride, len_max, H, AH, c16_hi, c16_lo, Hp);
        trace_end(trace);
        return;
    }
    {
        const float4* h4 = reinterpret_cast<const float4*>(s_h);
        const float4* w4 = reinterpret_cast<const float4*>(s_w);
#pragma unroll
        for (int k = 0; k < kMaxItems; ++k) {
            const int item = wid + k * nw;
            if (item >= items) continue;
            const int j4 = (item % Q) * 32 + lane;
            const float4 v = pv[k], hh = h4[j4], ww = w4[j4];
            float a = ww.x * tanh_score(v.x + hh.x);
            a = fmaf(ww.y, tanh_score(v.y + hh.y), a);
            a = fmaf(ww.z, tanh_score(v.z + hh.z), a);
            a = fmaf(ww.w, tanh_score(v.w + hh.w), a);
            a = warp_sum(a);
            if (lane == 0) s_c[item] = a;   // score partial of (node, slice); combined below in slice order
        }
    }
    __syncthreads();
    if (threadIdx.x < len_max) {
        float e = 0.f;
        for (int q = 0; q < Q; ++q) e += s_c[threadIdx.x * Q + q];
        s_e[threadIdx.x] = e + __ldg(alpha_b);
    }
    __syncthreads();
    if (threadIdx.x == 0) trace_mark(trace, 1);   // scores done
    attention_tail(s_e, s_c, r, cr, att, masks, ctx, att_w, att_w_stride, len_max, H, c16_hi, c16_lo, Hp, trace, overflow, s_mask);
    trace_end(trace);
}

// ---- attention of a beam-search step: one block per SUB-GRAPH ------------------------------------------------------------------
// The b beams of a sub-graph attend over the same p_att / att rows (rows_per_ctx = b).  attention_kernel runs one block per decode row
// and reads them b times (640 blocks at beam 5: 68 us per step, L2- and latency-bound); here the block of a sub-graph reads every p_att
// quad and every att row once and applies it to its b rows.  Per row the operations and their order are those of attention_kernel's fast
// path (same split-K reduce, score partials per (node, slice) combined in slice order, the same one-warp softmax, four node groups
// combined in the same fixed order), so the results are identical.
constexpr int kBeamAttItems = 5;   // (node, 32-quad slice) items per warp: len_max * AH / 128 <= 160
__global__ void __launch_bounds__(kAttThreads, 1) attention_beam_kernel(const float* __restrict__ atth_part, int splits, const float* __restrict__ h2att_b,
                                                                       const float* __restrict__ p_att, const float* __restrict__ att,
                                                                       const float* __restrict__ masks, const float* __restrict__ alpha_w,
                                                                       const float* __restrict__ alpha_b, float* __restrict__ ctx, float* __restrict__ att_w,
                                                                       int att_w_stride, int S, int len_max, int H, int AH, int b,
                                                                       const int* __restrict__ active, unsigned short* __restrict__ c16_hi,
                                                                       unsigned short* __restrict__ c16_lo, int Hp, int* overflow, int ld_part, int early_ok) {
    pdl_trigger();
    extern __shared__ float s_att[];   // [b][AH] atth | [AH] alpha_w | [b][64] e | [64] mask | [b][256] score partials | [4][b][H] context partials
    float* s_h = s_att;
    float* s_w = s_h + (size_t)b * AH;
    float* s_e = s_w + AH;
    float* s_mask = s_e + (size_t)b * 64;
    float* s_sp = s_mask + 64;
    float* s_c = s_sp + (size_t)b * 256;
    const int cr = blockIdx.x, r0 = cr * b;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int AH4 = AH >> 2, Q = AH4 >> 5;
    const int items = len_max * Q;
    const float4* pa4 = reinterpret_cast<const float4*>(p_att + (size_t)cr * len_max * AH);
    float4 pv[kBeamAttItems];
    if (early_ok) {   // p_att / masks do not depend on this step (see attention_kernel)
#pragma unroll
        for (int k = 0; k < kBeamAttItems; ++k) {
            const int item = wid + k * nw;
            pv[k] = item < items ? __ldg(pa4 + (size_t)(item / Q) * AH4 + (item % Q) * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (threadIdx.x < len_max) s_mask[threadIdx.x] = __ldg(masks + (size_t)cr * len_max + threadIdx.x);
    }
    {
        const char* af = reinterpret_cast<const char*>(att + (size_t)cr * len_max * H);
        const size_t nb_a = (size_t)len_max * H * 4;
        for (size_t o = (size_t)threadIdx.x * 128; o < nb_a; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(af + o));
    }
    for (int j = threadIdx.x; j < AH; j += blockDim.x) s_w[j] = __ldg(alpha_w + j);
    pdl_wait();
    if (active != nullptr && *active == 0) return;
    if (!early_ok) {
#pragma unroll
        for (int k = 0; k < kBeamAttItems; ++k) {
            const int item = wid + k * nw;
            pv[k] = item < items ? pa4[(size_t)(item / Q) * AH4 + (item % Q) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (threadIdx.x < len_max) s_mask[threadIdx.x] = masks[(size_t)cr * len_max + threadIdx.x];
    }
    for (int i = threadIdx.x; i < b * AH; i += blockDim.x) {
        const int q = i / AH, j = i - q * AH;
        float a = 0.f;
        for (int z = 0; z < splits; ++z) a += atth_part[((size_t)z * S + r0 + q) * ld_part + j];
        s_h[i] = a + __ldg(h2att_b + j);
    }
    __syncthreads();
    {
        const float4* w4 = reinterpret_cast<const float4*>(s_w);
#pragma unroll
        for (int k = 0; k < kBeamAttItems; ++k) {
            const int item = wid + k * nw;
            if (item >= items) continue;
            const int j4 = (item % Q) * 32 + lane;
            const float4 v = pv[k], ww = w4[j4];
            for (int q = 0; q < b; ++q) {
                const float4 hh = reinterpret_cast<const float4*>(s_h + (size_t)q * AH)[j4];
                float a = ww.x * tanh_score(v.x + hh.x);
                a = fmaf(ww.y, tanh_score(v.y + hh.y), a);
                a = fmaf(ww.z, tanh_score(v.z + hh.z), a);
                a = fmaf(ww.w, tanh_score(v.w + hh.w), a);
                a = warp_sum(a);
                if (lane == 0) s_sp[q * 256 + item] = a;
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < b * len_max; i += blockDim.x) {
        const int q = i / len_max, n = i - q * len_max;
        float e = 0.f;
        for (int z = 0; z < Q; ++z) e += s_sp[q * 256 + n * Q + z];
        s_e[q * 64 + n] = e + __ldg(alpha_b);
    }
    __syncthreads();
    if (wid < b) {   // warp q: softmax / mask / renormalise of beam q (two-stage, as the reference; the sequence of attention_tail)
        float* e = s_e + wid * 64;
        float m = -INFINITY;
        for (int n = lane; n < len_max; n += 32) m = fmaxf(m, e[n]);
        m = warp_max(m);
        float sum = 0.f;
        for (int n = lane; n < len_max; n += 32) sum += expf(e[n] - m);
        sum = warp_sum(sum);
        float msum = 0.f;
        for (int n = lane; n < len_max; n += 32) {
            float wv = expf(e[n] - m) / sum;
            wv = wv * s_mask[n];
            e[n] = wv;
            msum += wv;
        }
        msum = warp_sum(msum);
        for (int n = lane; n < len_max; n += 32) {
            const float wv = e[n] / msum;
            e[n] = wv;
            if (att_w) att_w[(size_t)(r0 + wid) * att_w_stride + n] = wv;
        }
    }
    __syncthreads();
    // context: four thread groups take interleaved node subsets (n = g, g + 4, ...); every att quad is loaded once for all beams
    const float* af = att + (size_t)cr * len_max * H;
    const int grp = threadIdx.x >> 8, tg = threadIdx.x & 255, H4 = H >> 2;
    for (int j4 = tg; j4 < H4; j4 += 256) {
        float4 acc[kMaxBeamAtt];
#pragma unroll
        for (int q = 0; q < kMaxBeamAtt; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 3
        for (int n = grp; n < len_max; n += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(af + (size_t)n * H) + j4);
#pragma unroll
            for (int q = 0; q < kMaxBeamAtt; ++q) {
                if (q < b) {
                    const float wv = s_e[q * 64 + n];
                    acc[q].x = fmaf(wv, v.x, acc[q].x); acc[q].y = fmaf(wv, v.y, acc[q].y); acc[q].z = fmaf(wv, v.z, acc[q].z); acc[q].w = fmaf(wv, v.w, acc[q].w);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kMaxBeamAtt; ++q)
            if (q < b) reinterpret_cast<float4*>(s_c + ((size_t)grp * b + q) * H)[j4] = acc[q];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < b * H; i += blockDim.x) {
        const int q = i / H, j = i - q * H;
        const size_t gs = (size_t)b * H;
        const float cv = ((s_c[(size_t)q * H + j] + s_c[gs + (size_t)q * H + j]) + s_c[2 * gs + (size_t)q * H + j]) + s_c[3 * gs + (size_t)q * H + j];
        ctx[(size_t)(r0 + q) * H + j] = cv;
        if (c16_hi) split_f16_store(cv, c16_hi, c16_lo, (size_t)(r0 + q) * Hp + j, overflow);
    }
}
static size_t attention_beam_smem(int b, int AH, int H) { return (size_t)(b * AH + AH + b * 64 + 64 + b * 256 + 4 * b * H) * sizeof(float); }
static bool attention_beam_ok(int S, int b, int len_max, int H, int AH) {
    return b > 1 && b <= kMaxBeamAtt && S % b == 0 && (AH & 127) == 0 && (H & 3) == 0 && len_max <= 64 && len_max * (AH >> 7) <= kBeamAttItems * (kAttThreads / 32) &&
           attention_beam_smem(b, AH, H) <= 200 * 1024;
}

// ---- fused attention phase of a decode step: one cluster kernel instead of cell + h2att GEMM + attention --------------------
// The three stages are tiny but strictly dependent (att-LSTM cell -> W_h h_att -> attention); as separate kernels they cost
// three launch / fill / tail latencies.  Here block r (one per decode row) belongs to a cluster of 8 consecutive rows and
//   1. reduces the att-LSTM split-K partials of row r and applies the LSTM cell                  -> h_att[r], c_att[r]
//   -- cluster barrier --
//   2. computes atth[rows of the cluster, its 1/8 slice of the AH columns] (W_h slice streamed once from L2, the 8 h rows in
//      shared memory) and writes each value into the shared memory of the block that owns the row (DSMEM)
//   -- cluster barrier --
//   3. runs the attention of row r (tanh / alpha / softmax / mask / renormalise / context)         -> ctx[r]
constexpr int kAttCluster = 8;
struct AttPhaseArgs {
    const float* part; int splits;               // att-LSTM gate partials [splits][S][4H]
    const float* b_ih; const float* b_hh;        // used when fc_pre == nullptr
    const float* fc_pre;                         // [S / rows_per_ctx, 4H] hoisted fc segment + biases (nullable)
    const float* c_prev; float* h_out; float* c_out;   // layer-0 state rows [S, H]
    const long long* parent;                     // nullable: previous-state row of each row (beam re-ordering)
    const float* w_h; const float* b_h;          // h2att [AH, H], [AH]
    const float* p_att; const float* att; const float* masks; const float* alpha_w; const float* alpha_b;
    float* ctx; float* att_w; int att_w_stride;
    int S, len_max, H, AH, cols_per_block, rows_per_ctx;
    const int* active;
    unsigned short *h16_hi, *h16_lo, *c16_hi, *c16_lo;   // nullable split-fp16 copies of h_att / ctx, [S, Hp]
    int Hp;
    unsigned long long* trace;   // debug (SUBGC_ATT_TRACE=1): [grid][8] globaltimer stamps per block; nullptr in normal operation
};
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define ATT_STAMP(i) do { if (a.trace != nullptr && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 8 + (i)] = gtime(); } while (0)

__global__ void __launch_bounds__(kAttThreads, 1) att_phase_kernel(const AttPhaseArgs a) {
    pdl_trigger();
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ float s_att[];  // [AH] atth | [AH] alpha_w | [64] e | [8][H] h rows of the cluster (stage 2) / [4][H] context partials
    float* s_h = s_att;
    float* s_w = s_att + a.AH;
    float* s_e = s_att + 2 * a.AH;
    float* s_c = s_att + 2 * a.AH + 64;
    const int r = blockIdx.x, H = a.H, AH = a.AH, S = a.S;
    const bool valid = r < S;
    const int crank = (int)cluster.block_rank();
    const int row0 = r - crank;
    const int cr = r / a.rows_per_ctx;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    ATT_STAMP(0);
    // The row's attention operands (p_att 74 KB, att 144 KB) do not depend on this step: ask for them in L2 now, they are
    // evicted by the weight streams between steps and stage 3 is otherwise a chain of DRAM-latency rounds.
    if (valid) {
        const char* pa = reinterpret_cast<const char*>(a.p_att + (size_t)cr * a.len_max * AH);
        const char* af = reinterpret_cast<const char*>(a.att + (size_t)cr * a.len_max * H);
        const size_t nb_p = (size_t)a.len_max * AH * 4, nb_a = (size_t)a.len_max * H * 4;
        for (size_t o = (size_t)threadIdx.x * 128; o < nb_p; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + o));
        for (size_t o = (size_t)threadIdx.x * 128; o < nb_a; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(af + o));
    }
    pdl_wait();
    if (a.active != nullptr && *a.active == 0) return;   // uniform across the grid: nobody reaches the barriers
    // ---- 1. split-K reduce + biases + LSTM cell of row r
    if (valid) {
        const size_t zs = (size_t)S * 4 * H;
        const long long pr = a.parent ? a.parent[r] : r;
        for (int j = threadIdx.x; j < H; j += blockDim.x) {
            const float* g = a.part + (size_t)r * 4 * H + j;
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 9
            for (int z = 0; z < a.splits; ++z) {   // all splits' loads in flight at once
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += g[(size_t)z * zs + (size_t)q * H];
            }
            if (a.fc_pre) {
                const float* ad = a.fc_pre + (size_t)cr * 4 * H + j;
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] += ad[(size_t)q * H];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = (acc[q] + __ldg(a.b_ih + q * H + j)) + __ldg(a.b_hh + q * H + j);
            }
            const float c = sigmoidf_(acc[1]) * a.c_prev[(size_t)pr * H + j] + sigmoidf_(acc[0]) * tanhf(acc[2]);
            a.c_out[(size_t)r * H + j] = c;
            const float hv = sigmoidf_(acc[3]) * tanhf(c);
            a.h_out[(size_t)r * H + j] = hv;
            if (a.h16_hi) split_f16_store(hv, a.h16_hi, a.h16_lo, (size_t)r * a.Hp + j);
        }
    }
    ATT_STAMP(1);
    cluster.sync();
    ATT_STAMP(2);
    // ---- 2. h2att: this block's column slice for the rows of the cluster
    {
        const int nrows = max(0, min(kAttCluster, S - row0));
        const float* hsrc = a.h_out + (size_t)row0 * H;   // the cluster's rows are consecutive: one contiguous run, written in stage 1
        for (int idx = threadIdx.x; idx < nrows * H; idx += blockDim.x) s_c[idx] = __ldcg(hsrc + idx);
        for (int idx = nrows * H + threadIdx.x; idx < kAttCluster * H; idx += blockDim.x) s_c[idx] = 0.f;
        __syncthreads();
        const int c0 = crank * a.cols_per_block;
        const int nc = max(0, min(a.cols_per_block, AH - c0));
        for (int cp = wid * 2; cp < nc; cp += nw * 2) {
            const int colA = c0 + cp, colB = min(c0 + cp + 1, AH - 1);
            const bool hasB = cp + 1 < nc;
            const float* wa = a.w_h + (size_t)colA * H;
            const float* wb = a.w_h + (size_t)colB * H;
            float accA[kAttCluster], accB[kAttCluster];
#pragma unroll
            for (int i = 0; i < kAttCluster; ++i) { accA[i] = 0.f; accB[i] = 0.f; }
            if ((H & 3) == 0) {
                // The stage is instruction-bound (65 MFMA over the grid), so every shared-memory read feeds 8 FMAs: a lane owns 4
                // consecutive k (one 16-byte weight load per column, one 16-byte read per h row)
                const int H4 = H >> 2;
                const float4* wa4 = reinterpret_cast<const float4*>(wa);
                const float4* wb4 = reinterpret_cast<const float4*>(wb);
                float4 a0 = lane < H4 ? __ldg(wa4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 b0 = lane < H4 ? __ldg(wb4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                for (int q0 = lane; q0 < H4; q0 += 32) {   // the next quad of both columns is requested before this one is consumed
                    const int q1 = q0 + 32;
                    const float4 an = q1 < H4 ? __ldg(wa4 + q1) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bn = q1 < H4 ? __ldg(wb4 + q1) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < kAttCluster; ++i) {
                        const float4 h0 = reinterpret_cast<const float4*>(s_c + i * H)[q0];
                        accA[i] = fmaf(a0.x, h0.x, accA[i]); accA[i] = fmaf(a0.y, h0.y, accA[i]);
                        accA[i] = fmaf(a0.z, h0.z, accA[i]); accA[i] = fmaf(a0.w, h0.w, accA[i]);
                        accB[i] = fmaf(b0.x, h0.x, accB[i]); accB[i] = fmaf(b0.y, h0.y, accB[i]);
                        accB[i] = fmaf(b0.z, h0.z, accB[i]); accB[i] = fmaf(b0.w, h0.w, accB[i]);
                    }
                    a0 = an; b0 = bn;
                }
            } else {
                for (int k = lane; k < H; k += 32) {
                    const float va = __ldg(wa + k), vb = __ldg(wb + k);
#pragma unroll
                    for (int i = 0; i < kAttCluster; ++i) {
                        const float hv = s_c[i * H + k];
                        accA[i] = fmaf(va, hv, accA[i]);
                        accB[i] = fmaf(vb, hv, accB[i]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kAttCluster; ++i) {
                const float sa = warp_sum(accA[i]), sb = warp_sum(accB[i]);
                if (lane == 0 && i < nrows) {
                    float* peer = cluster.map_shared_rank(s_h, i);   // atth row of the block that owns row (row0 + i)
                    peer[colA] = sa + __ldg(a.b_h + colA);
                    if (hasB) peer[colB] = sb + __ldg(a.b_h + colB);
                }
            }
        }
    }
    ATT_STAMP(3);
    cluster.sync();
    ATT_STAMP(4);
    if (!valid) return;
    // ---- 3. attention of row r
    for (int j = threadIdx.x; j < AH; j += blockDim.x) s_w[j] = __ldg(a.alpha_w + j);
    __syncthreads();
    attention_row_body(s_h, s_w, s_e, s_c, r, cr, a.p_att, a.att, a.masks, a.alpha_b, a.ctx, a.att_w, a.att_w_stride, a.len_max, H, AH, a.c16_hi,
                       a.c16_lo, a.Hp);
    __syncthreads();
    ATT_STAMP(5);
}

// ---- row-wise log_softmax (materialised log-probs: get_logprobs_state API and beam search) ----------------------
__global__ void __launch_bounds__(256) log_softmax_kernel(const float* __restrict__ logits, float* __restrict__ logp, int V1,
                                                          size_t out_stride, const int* __restrict__ active) {
    pdl_trigger();
    pdl_wait();
    if (active != nullptr && *active == 0) return;
    __shared__ float red[32];
    const float* x = logits + (size_t)blockIdx.x * V1;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < V1; j += blockDim.x) m = fmaxf(m, x[j]);
    m = block_max(m, red);
    float s = 0.f;
    for (int j = threadIdx.x; j < V1; j += blockDim.x) s += expf(x[j] - m);
    s = block_sum(s, red);
    const float lz = logf(s);
    for (int j = threadIdx.x; j < V1; j += blockDim.x) logp[(size_t)blockIdx.x * out_stride + j] = (x[j] - m) - lz;
}

// ---- greedy / top-k selection + finish bookkeeping: one block per row ------------------------------------------
struct Philox {
    static __device__ __forceinline__ void round(unsigned (&c)[4], unsigned k0, unsigned k1) {
        unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    // Philox4x32-10 keyed by `seed`, counter (offset, t, r); returns a uniform in [0, 1)
    static __device__ float uniform(unsigned long long seed, unsigned long long offset, unsigned t, unsigned r) {
        unsigned c[4] = {(unsigned)offset, (unsigned)(offset >> 32), t, r};
        unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
        for (int i = 0; i < 10; ++i) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        return (float)(c[0] >> 8) * (1.0f / 16777216.0f);
    }
};

constexpr int kMaxTopK = 16;

struct SelectArgs {
    const float* logits;   // [S, V1] materialised logits, or (splits > 0) split-K partials [splits][S][V1] without bias
    int splits;
    const float* bias;     // logit bias, added here when reading partials
    int V1, T, t, S;
    int mode;              // 0 greedy, 1 top-k sampling
    float temp;
    int top_k;
    unsigned long long seed, offset;
    const float* uniforms; // nullable [T, S]
    long long* it;         // [S] token fed to the next step
    int* unfinished;       // [S]
    long long* seq;        // [S, T]
    float* seq_lp;         // [S, T]
    int* count;            // [T + 1] unfinished rows after step t (count[t] doubles as the `active` flag of step t+1)
    const int* active;
    const float* embed;    // [V1, X] word embedding; the next step's input row relu(E[it]) is written to xt [S, X] here
    float* xt;
    int X;
    unsigned short *xt16_hi, *xt16_lo;   // nullable split-fp16 copy of xt, [S, Xp]
    int Xp;
    TraceSlot trace;
    int* overflow;                       // nullable device flag: an embedding value did not fit the fp16 split
};

constexpr int kSelectThreads = 1024;  // one block per row: the row is latency-bound, so use every warp slot of the SM

__global__ void __launch_bounds__(kSelectThreads) select_kernel(const SelectArgs a) {
    trace_begin(a.trace);
    pdl_trigger();
    pdl_wait();
    trace_released(a.trace);
    if (a.active != nullptr && *a.active == 0) return;
    __shared__ float redv[32];
    __shared__ int redi[32];
    __shared__ float s_topv[kMaxTopK];
    __shared__ int s_topi[kMaxTopK];
    extern __shared__ float s_row[];  // [V1] the logit row (reduced from the partials)
    const int r = blockIdx.x;
    if (a.splits > 0) {
        const size_t zs = (size_t)a.S * a.V1;
        const float* p0 = a.logits + (size_t)r * a.V1;
        for (int j0 = threadIdx.x; j0 < a.V1; j0 += 4 * blockDim.x) {  // 4 columns x splits independent loads in flight
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            for (int z = 0; z < a.splits; ++z) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * blockDim.x;
                    if (j < a.V1) v[u] += p0[(size_t)z * zs + j];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * blockDim.x;
                if (j < a.V1) s_row[j] = v[u] + __ldg(a.bias + j);
            }
        }
    } else {
        for (int j = threadIdx.x; j < a.V1; j += blockDim.x) s_row[j] = a.logits[(size_t)r * a.V1 + j];
    }
    __syncthreads();
    const float* x = s_row;
    // max (first index) and log-sum-exp of the logits
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < a.V1; j += blockDim.x) {
        float v = x[j];
        if (v > bv || bi == 0x7fffffff) { bv = v; bi = j; }
    }
    block_argmax(bv, bi, redv, redi);
    const float m = bv;
    float s = 0.f;
    for (int j = threadIdx.x; j < a.V1; j += blockDim.x) s += expf(x[j] - m);
    s = block_sum(s, redv);
    const float lz = logf(s);
    int tok;
    float lp;
    if (a.mode == 0) {
        tok = bi;
        lp = (m - m) - lz;  // log_softmax value at the arg-max
    } else {
        // q = log_softmax(logp / temp) (AttModel.py:296), logp = (x - m) - lz
        const float ym = ((m - m) - lz) / a.temp;
        float s2 = 0.f;
        for (int j = threadIdx.x; j < a.V1; j += blockDim.x) s2 += expf(((x[j] - m) - lz) / a.temp - ym);
        s2 = block_sum(s2, redv);
        const float lz2 = logf(s2);
        const int k = a.top_k;
        for (int c = 0; c < k; ++c) {  // k rounds of arg-max with exclusion (descending q, lower index first on ties)
            float cv = -INFINITY;
            int ci = 0x7fffffff;
            for (int j = threadIdx.x; j < a.V1; j += blockDim.x) {
                bool taken = false;
                for (int e = 0; e < c; ++e) taken |= (s_topi[e] == j);
                if (taken) continue;
                float q = (((x[j] - m) - lz) / a.temp - ym) - lz2;
                if (q > cv || ci == 0x7fffffff) { cv = q; ci = j; }
            }
            block_argmax(cv, ci, redv, redi);
            if (threadIdx.x == 0) { s_topv[c] = cv; s_topi[c] = ci; }
            __syncthreads();
        }
        // Categorical over the kept tokens: inverse CDF in descending-probability order
        float u = a.uniforms ? a.uniforms[(size_t)a.t * a.S + r] : Philox::uniform(a.seed, a.offset, (unsigned)a.t, (unsigned)r);
        float den = 0.f;
        for (int c = 0; c < k; ++c) den += expf(s_topv[c] - s_topv[0]);
        float cdf = 0.f;
        int pos = 0;
        for (int c = 0; c < k; ++c) {
            cdf += expf(s_topv[c] - s_topv[0]) / den;
            if (u >= cdf) pos = c + 1;
        }
        if (pos > k - 1) pos = k - 1;
        tok = s_topi[pos];
        lp = s_topv[pos];
    }
    __shared__ int s_it;
    if (threadIdx.x == 0) {
        int unf = (a.t == 0 ? 1 : a.unfinished[r]) && (tok > 0);
        long long it = unf ? tok : 0;
        a.it[r] = it;
        a.unfinished[r] = unf;
        a.seq[(size_t)r * a.T + a.t] = it;
        a.seq_lp[(size_t)r * a.T + a.t] = lp;
        if (unf) atomicAdd(a.count + a.t, 1);
        s_it = (int)it;
    }
    if (a.xt != nullptr) {  // embed + ReLU of the token fed to the next step (AttModel.py:332), fused here
        __syncthreads();
        const float* e = a.embed + (size_t)s_it * a.X;
        for (int j = threadIdx.x; j < a.X; j += blockDim.x) {
            const float xv = fmaxf(__ldg(e + j), 0.f);
            a.xt[(size_t)r * a.X + j] = xv;
            if (a.xt16_hi) split_f16_store(xv, a.xt16_hi, a.xt16_lo, (size_t)r * a.Xp + j, a.overflow);
        }
    }
    trace_end(a.trace);
}

// Same selection with the logit row held in registers (V1 <= 10 x 1024): no dynamic shared memory and <= 40 registers, so the
// kernel fits next to the CTAs of the following contraction (they prefetch their weight ring meanwhile).
constexpr int kSelVals = 10;
__global__ void __maxnreg__(40) select_reg_kernel(const SelectArgs a) {
    trace_begin(a.trace);
    pdl_trigger();
    pdl_wait();
    trace_released(a.trace);
    if (a.active != nullptr && *a.active == 0) return;
    __shared__ float redv[32];
    __shared__ int redi[32];
    __shared__ float s_topv[kMaxTopK];
    __shared__ int s_topi[kMaxTopK];
    const int r = blockIdx.x, tid = threadIdx.x;
    float v[kSelVals];
#pragma unroll
    for (int i = 0; i < kSelVals; ++i) v[i] = 0.f;
    if (a.splits > 0) {
        const size_t zs = (size_t)a.S * a.V1;
        const float* p0 = a.logits + (size_t)r * a.V1;
        for (int z = 0; z < a.splits; ++z) {
#pragma unroll
            for (int i = 0; i < kSelVals; ++i) {
                const int j = tid + i * kSelectThreads;
                if (j < a.V1) v[i] += p0[(size_t)z * zs + j];
            }
        }
#pragma unroll
        for (int i = 0; i < kSelVals; ++i) {
            const int j = tid + i * kSelectThreads;
            if (j < a.V1) v[i] += __ldg(a.bias + j);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kSelVals; ++i) {
            const int j = tid + i * kSelectThreads;
            if (j < a.V1) v[i] = a.logits[(size_t)r * a.V1 + j];
        }
    }
    // max (first index) and log-sum-exp of the logits
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kSelVals; ++i) {
        const int j = tid + i * kSelectThreads;
        if (j < a.V1 && (v[i] > bv || bi == 0x7fffffff)) { bv = v[i]; bi = j; }
    }
    block_argmax(bv, bi, redv, redi);
    const float m = bv;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kSelVals; ++i) {
        const int j = tid + i * kSelectThreads;
        if (j < a.V1) s += expf(v[i] - m);
    }
    s = block_sum(s, redv);
    const float lz = logf(s);
    int tok;
    float lp;
    if (a.mode == 0) {
        tok = bi;
        lp = (m - m) - lz;  // log_softmax value at the arg-max
    } else {
        const float ym = ((m - m) - lz) / a.temp;
        float s2 = 0.f;
#pragma unroll
        for (int i = 0; i < kSelVals; ++i) {
            const int j = tid + i * kSelectThreads;
            if (j < a.V1) s2 += expf(((v[i] - m) - lz) / a.temp - ym);
        }
        s2 = block_sum(s2, redv);
        const float lz2 = logf(s2);
        const int k = a.top_k;
        for (int c = 0; c < k; ++c) {  // k rounds of arg-max with exclusion (descending q, lower index first on ties)
            float cv = -INFINITY;
            int ci = 0x7fffffff;
#pragma unroll
            for (int i = 0; i < kSelVals; ++i) {
                const int j = tid + i * kSelectThreads;
                if (j >= a.V1) continue;
                bool taken = false;
                for (int e = 0; e < c; ++e) taken |= (s_topi[e] == j);
                if (taken) continue;
                const float q = (((v[i] - m) - lz) / a.temp - ym) - lz2;
                if (q > cv || ci == 0x7fffffff) { cv = q; ci = j; }
            }
            block_argmax(cv, ci, redv, redi);
            if (tid == 0) { s_topv[c] = cv; s_topi[c] = ci; }
            __syncthreads();
        }
        float u = a.uniforms ? a.uniforms[(size_t)a.t * a.S + r] : Philox::uniform(a.seed, a.offset, (unsigned)a.t, (unsigned)r);
        float den = 0.f;
        for (int c = 0; c < k; ++c) den += expf(s_topv[c] - s_topv[0]);
        float cdf = 0.f;
        int pos = 0;
        for (int c = 0; c < k; ++c) {
            cdf += expf(s_topv[c] - s_topv[0]) / den;
            if (u >= cdf) pos = c + 1;
        }
        if (pos > k - 1) pos = k - 1;
        tok = s_topi[pos];
        lp = s_topv[pos];
    }
    __shared__ int s_it;
    if (tid == 0) {
        int unf = (a.t == 0 ? 1 : a.unfinished[r]) && (tok > 0);
        long long it = unf ? tok : 0;
        a.it[r] = it;
        a.unfinished[r] = unf;
        a.seq[(size_t)r * a.T + a.t] = it;
        a.seq_lp[(size_t)r * a.T + a.t] = lp;
        if (unf) atomicAdd(a.count + a.t, 1);
        s_it = (int)it;
    }
    if (a.xt != nullptr) {  // embed + ReLU of the token fed to the next step (AttModel.py:332), fused here
        __syncthreads();
        const float* e = a.embed + (size_t)s_it * a.X;
        for (int j = tid; j < a.X; j += blockDim.x) {
            const float xv = fmaxf(__ldg(e + j), 0.f);
            a.xt[(size_t)r * a.X + j] = xv;
            if (a.xt16_hi) split_f16_store(xv, a.xt16_hi, a.xt16_lo, (size_t)r * a.Xp + j, a.overflow);
        }
    }
    trace_end(a.trace);
}

__global__ void steps_done_kernel(const int* __restrict__ count, int T, int* __restrict__ steps_done) {
    int steps = T + 1;
    for (int t = 0; t < T; ++t)
        if (count[t] == 0) { steps = t + 1; break; }
    steps_done[0] = steps;
}

// teacher forcing: tok[i][r] = tokens[r, i]; flag[i] = 1 while no earlier column i >= 1 was all-zero (AttModel.py:170-171)
__global__ void __launch_bounds__(256) teacher_columns_kernel(const long long* __restrict__ tokens, int ld_tok, int S, int n_steps,
                                                              long long* __restrict__ tok_cols, int* __restrict__ flags) {
    __shared__ int s_any;
    int alive = 1;
    for (int i = 0; i < n_steps; ++i) {
        if (threadIdx.x == 0) s_any = 0;
        __syncthreads();
        int any = 0;
        for (int r = threadIdx.x; r < S; r += blockDim.x) {
            long long v = tokens[(size_t)r * ld_tok + i];
            tok_cols[(size_t)i * S + r] = v;
            any |= (v != 0);
        }
        if (any) atomicOr(&s_any, 1);
        __syncthreads();
        if (i >= 1 && s_any == 0) alive = 0;
        if (threadIdx.x == 0) flags[i] = alive;
        __syncthreads();
    }
}

// ---- one decoder step ---------------------------------------------------------------------------------------------
struct StepScratch {
    float* gates;   // [S, 4H]
    float* atth;    // [S, AH]
    float* ctx;     // [S, H]
    void* gemm_ws;
    size_t gemm_ws_bytes;
    void* gemm_ws2;        // second contraction workspace: the merged [h2att | lang-early] partials live across the attention
    size_t gemm_ws2_bytes;
};

static size_t step_gemm_ws_bytes(const subgc_dims* d, int S) {
    size_t g = gemm_workspace_bytes(S, 4 * d->rnn, d->enc + 3 * d->rnn);
    size_t t = gemm_workspace_bytes(S, 4 * d->rnn, 3 * d->rnn);
    if (t > g) g = t;
    t = gemm_workspace_bytes(S, d->att_hid, d->rnn);
    if (t > g) g = t;
    t = gemm_workspace_bytes(S, d->vocab1, d->rnn);
    if (t > g) g = t;
    return align_up(g, 256);
}

static size_t step_gemm_ws2_bytes(const subgc_dims* d, int S) { return align_up(gemm_workspace_bytes(S, d->att_hid + 4 * d->rnn, 2 * d->rnn), 256); }

static size_t step_scratch_bytes(const subgc_dims* d, int S) {
    return align_up((size_t)S * 4 * d->rnn * 4, 256) + align_up((size_t)S * d->att_hid * 4, 256) + align_up((size_t)S * d->rnn * 4, 256) +
           step_gemm_ws_bytes(d, S) + step_gemm_ws2_bytes(d, S) + 512;
}

static bool take_step_scratch(const subgc_dims* d, int S, Workspace& ws, StepScratch& sc) {
    sc.gates = ws.take<float>((size_t)S * 4 * d->rnn);
    sc.atth = ws.take<float>((size_t)S * d->att_hid);
    sc.ctx = ws.take<float>((size_t)S * d->rnn);
    sc.gemm_ws_bytes = step_gemm_ws_bytes(d, S);
    sc.gemm_ws = ws.take<char>(sc.gemm_ws_bytes);
    sc.gemm_ws2_bytes = step_gemm_ws2_bytes(d, S);
    sc.gemm_ws2 = ws.take<char>(sc.gemm_ws2_bytes);
    return ws.ok();
}

// Ablation aid (timing experiments only, results are wrong when set): SUBGC_SKIP = bitmask of decode-step kernels NOT to launch:
// 1 att-LSTM GEMM, 2 att cell / fused att phase, 4 h2att GEMM, 8 attention, 16 lang GEMM, 32 lang cell, 64 logit GEMM, 128 select
static int skip_mask() {
    static int m = -1;
    if (m < 0) { const char* e = getenv("SUBGC_SKIP"); m = e ? atoi(e) : 0; }
    return m;
}

// debugging aid (SUBGC_ATT_TRACE=1): per-block stage time stamps of the most recent att-phase launch, read back by subgc_debug_att_trace
static unsigned long long* att_trace_buffer() {
    static unsigned long long* buf = nullptr;
    static int on = -1;
    if (on < 0) {
        on = getenv("SUBGC_ATT_TRACE") != nullptr ? 1 : 0;
        if (on) { cudaMalloc(&buf, 1024 * 8 * sizeof(unsigned long long)); cudaMemset(buf, 0, 1024 * 8 * sizeof(unsigned long long)); }
    }
    return buf;
}

// The fused att-phase kernel runs as clusters of 8 row-blocks (distributed shared memory carries the h2att results).
static bool att_phase_fusable(size_t smem) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("SUBGC_FUSED_ATT");   // opt-in: measured slower than the three separate kernels under PDL
        mode = (e && e[0] == '1') ? 1 : 0;
    }
    if (mode != 1 || smem > 48 * 1024) return false;
    int dev = 0, clus = 0;   // cluster launch support is a property of the current device, not of the process
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&clus, cudaDevAttrClusterLaunch, dev);
    return clus != 0;
}

// The loop's small kernels run next to the contraction's CTAs (which hold ~193 KB of shared memory): an SM only hosts kernels with
// the same shared-memory / L1 split, so they ask for the maximum shared-memory carve-out as well.
static void prefer_smem_carveout() {
    static DeviceOnce once;
    once.run([]() -> cudaError_t {   // a preference only: a failure costs overlap, not correctness, and is retried at the next call
        cudaError_t e = cudaFuncSetAttribute(lstm_reduce_cell_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(select_reg_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(log_softmax_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        return e;
    });
}

// upto: 0 = whole step, 1 = stop after the attention (the reference's discarded last step, only its attention
// weights are observable).  parent (nullable) re-maps the previous-state rows (beam re-ordering).
// raw_logits != nullptr: the logit contraction leaves its split-K partials (no bias) for a fused consumer
// (select_kernel); otherwise `logits` [S, V1] is materialised with the bias applied.
// Split-fp16 copies of the step's activations, written by the kernels that produce them (cells, attention, selection) so that the
// h3 contractions read them through TMA without a conversion pass.  Row-identity only: unused when rows are re-mapped (beam search).
struct Step16 {
    unsigned short *hin_hi = nullptr, *hin_lo = nullptr;     // [2, S, Hp] previous state
    unsigned short *hout_hi = nullptr, *hout_lo = nullptr;   // [2, S, Hp] new state
    unsigned short *ctx_hi = nullptr, *ctx_lo = nullptr;     // [S, Hp]
    unsigned short *xt_hi = nullptr, *xt_lo = nullptr;       // [S, Xp] relu(E[it]) of the token fed to this step (valid when xt != nullptr)
    int Hp = 0, Xp = 0;
};
static void set_a16(GemmSeg& g, const unsigned short* hi, const unsigned short* lo, int ld) { g.A16_hi = hi; g.A16_lo = lo; g.lda16 = ld; }

static int launch_step(const subgc_dims* d, const subgc_weights* w, int S, int len_max, int rows_per_ctx, const long long* it, const float* xt,
                       const long long* parent, const float* fc, const float* att, const float* p_att, const float* masks,
                       const float* h_in, const float* c_in, float* h_out, float* c_out, float* logits, RawPartials* raw_logits, float* att_w,
                       int att_w_stride, const StepScratch& sc, const int* active, int upto, cudaStream_t st, const float* fc_pre = nullptr,
                       const Step16* h16 = nullptr) {
    // fc_pre != nullptr: W_ih[:, H:2H] fc + b_ih + b_hh was computed once for the whole loop (launch_fc_pre): the fc segment
    // (a quarter of the att-LSTM weights) is not streamed again at every step
    const int H = d->rnn, X = d->enc, AH = d->att_hid, V1 = d->vocab1;
    prefer_smem_carveout();
    const size_t SH = (size_t)S * H;
    const int pw_blocks = (int)((SH + 255) / 256);
    GemmProblem p;
    p.wts = w;
    RawPartials rp;
    // attention LSTM: gates = W_ih [h_lang | fc | relu(E[it])] + b_ih + W_hh h_att + b_hh   (AttModel.py:410-413)
    p.M = S; p.N = 4 * H;
    int ns = 0;
    // split-fp16 copies: what this step produces (h_att, ctx, h_lang) is always written / read as such; the previous state only when
    // rows are not re-mapped (beam search gathers them by `parent`, which goes through split_rows_kernel)
    const bool out16 = h16 != nullptr && h16->hout_hi != nullptr;
    const bool in16 = out16 && parent == nullptr && h16->hin_hi != nullptr;
    const size_t SHp = out16 ? (size_t)S * h16->Hp : 0;
    p.seg[ns] = make_seg(h_in + SH, H, w->att_w_ih, X + 2 * H, H);
    if (in16) set_a16(p.seg[ns], h16->hin_hi + SHp, h16->hin_lo + SHp, h16->Hp);
    p.seg[ns++].gather = parent;
    if (!fc_pre) {
        p.seg[ns] = make_seg(fc, H, w->att_w_ih + H, X + 2 * H, H);
        p.seg[ns++].a_row_div = rows_per_ctx;
    }
    if (xt) {  // relu(E[it]) already materialised by the previous step's selection kernel
        p.seg[ns] = make_seg(xt, X, w->att_w_ih + 2 * H, X + 2 * H, X);
        if (in16 && h16->xt_hi) set_a16(p.seg[ns], h16->xt_hi, h16->xt_lo, h16->Xp);
        ++ns;
    } else {
        p.seg[ns] = make_seg(w->embed, X, w->att_w_ih + 2 * H, X + 2 * H, X);
        p.seg[ns].gather = it;
        p.seg[ns++].relu_a = 1;
    }
    p.seg[ns] = make_seg(h_in, H, w->att_w_hh, H, H);
    if (in16) set_a16(p.seg[ns], h16->hin_hi, h16->hin_lo, h16->Hp);
    p.seg[ns++].gather = parent;
    p.nseg = ns;
    p.active = active;
    const int skip = skip_mask();
    rp.part = static_cast<const float*>(sc.gemm_ws); rp.splits = 1;
    bool cell_fused = false;   // gates -> cell inside the contraction (h3 path with packed weights), else partials + cell kernel
    if (!(skip & 1)) {
        CellEpilogue ce;
        ce.H = H; ce.c_prev = c_in; ce.parent = parent; ce.addend = fc_pre; ce.add_div = rows_per_ctx; ce.b_ih = w->att_b_ih; ce.b_hh = w->att_b_hh;
        ce.h_out = h_out; ce.c_out = c_out;
        if (out16) { ce.h16_hi = h16->hout_hi; ce.h16_lo = h16->hout_lo; ce.Hp = h16->Hp; }
        SUBGC_TRY(launch_gemm_cell(p, ce, sc.gemm_ws, sc.gemm_ws_bytes, st, &cell_fused));
        if (!cell_fused) SUBGC_TRY(launch_gemm_raw(p, sc.gemm_ws, sc.gemm_ws_bytes, st, &rp));
    }
    // opt-in (SUBGC_MERGED=1): parity-green, but measured 0.04 ms slower per loop -- half of the merged contraction's activation tiles
    // (h_att) were written by the kernel right before it, and reads of just-written lines take ~2.5 us instead of ~1.2 us to land
    static const bool merge_off = !(getenv("SUBGC_MERGED") != nullptr && getenv("SUBGC_MERGED")[0] == '1');
    bool merged = w->lang_early_w != nullptr && !merge_off && upto == 0;   // [h2att | lang-early] as one contraction before the attention
    RawPartials rp_early{nullptr, 0};
    size_t smem = (size_t)(2 * AH + 64 + 4 * H) * sizeof(float);
    const size_t smem_fused = (size_t)(2 * AH + 64 + kAttCluster * H) * sizeof(float);
    if (!cell_fused && !merged && att_phase_fusable(smem_fused)) {
        // cell + h2att + attention as one kernel (one block per row, clusters of 8 rows, two cluster barriers)
        AttPhaseArgs fa;
        fa.part = rp.part; fa.splits = rp.splits; fa.b_ih = w->att_b_ih; fa.b_hh = w->att_b_hh; fa.fc_pre = fc_pre; fa.c_prev = c_in;
        fa.h_out = h_out; fa.c_out = c_out; fa.parent = parent; fa.w_h = w->h2att.w; fa.b_h = w->h2att.b; fa.p_att = p_att; fa.att = att;
        fa.masks = masks; fa.alpha_w = w->alpha_net.w; fa.alpha_b = w->alpha_net.b; fa.ctx = sc.ctx; fa.att_w = att_w;
        fa.att_w_stride = att_w_stride; fa.S = S; fa.len_max = len_max; fa.H = H; fa.AH = AH;
        fa.cols_per_block = (AH + kAttCluster - 1) / kAttCluster; fa.rows_per_ctx = rows_per_ctx; fa.active = active;
        fa.h16_hi = out16 ? h16->hout_hi : nullptr; fa.h16_lo = out16 ? h16->hout_lo : nullptr;
        fa.c16_hi = out16 ? h16->ctx_hi : nullptr; fa.c16_lo = out16 ? h16->ctx_lo : nullptr; fa.Hp = out16 ? h16->Hp : 0;
        fa.trace = att_trace_buffer();
        if (!(skip & 2)) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((S + kAttCluster - 1) / kAttCluster * kAttCluster);
            cfg.blockDim = dim3(kAttThreads);
            cfg.dynamicSmemBytes = smem_fused;
            cfg.stream = st;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = kAttCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
            cfg.attrs = at; cfg.numAttrs = 2;
            SUBGC_CUDA(cudaLaunchKernelEx(&cfg, att_phase_kernel, fa));
            count_launch();
        }
    } else {
        if (!(skip & 2) && !cell_fused) launch_pdl(lstm_reduce_cell_kernel, dim3(pw_blocks), dim3(256), (size_t)0, st, rp.part, rp.splits, w->att_b_ih, w->att_b_hh, c_in, parent, h_out, c_out, S,
                                                                            H, active, fc_pre, rows_per_ctx, out16 ? h16->hout_hi : nullptr,
                                                                            out16 ? h16->hout_lo : nullptr, out16 ? h16->Hp : 0, next_trace_slot(2), (const float*)nullptr, 0, 0);
        SUBGC_LAUNCH_CHECK();
        // attention (AttModel.py:445-471); the h2att partials are reduced inside the attention kernel
        p = GemmProblem(); p.wts = w;
        if (merged) {
            // everything that depends only on (h_att(t), h_lang(t-1)): h2att and two thirds of the language-LSTM gates, one launch
            p.M = S; p.N = AH + 4 * H; p.nseg = 2;
            p.seg[0] = make_seg(h_out, H, w->lang_early_w, 2 * H, H);
            p.seg[1] = make_seg(h_in + SH, H, w->lang_early_w + H, 2 * H, H);
            p.seg[1].gather = parent;
            if (out16) set_a16(p.seg[0], h16->hout_hi, h16->hout_lo, h16->Hp);
            if (in16) set_a16(p.seg[1], h16->hin_hi + SHp, h16->hin_lo + SHp, h16->Hp);
            p.active = active;
            if (!(skip & 4)) SUBGC_TRY(launch_gemm_raw(p, sc.gemm_ws2, sc.gemm_ws2_bytes, st, &rp_early));
            rp = rp_early;
        } else {
        p.M = S; p.N = AH; p.nseg = 1;
        p.seg[0] = make_seg(h_out, H, w->h2att.w, H, H);
        if (out16) set_a16(p.seg[0], h16->hout_hi, h16->hout_lo, h16->Hp);
        p.active = active;
        if (!(skip & 4)) SUBGC_TRY(launch_gemm_raw(p, sc.gemm_ws, sc.gemm_ws_bytes, st, &rp));
        }
        if (!(skip & 8) && attention_beam_ok(S, rows_per_ctx, len_max, H, AH) && !getenv("SUBGC_NO_BEAM_ATT")) {
            static DeviceOnce once_ba;
            SUBGC_CUDA(once_ba.run([]() { return cudaFuncSetAttribute(attention_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); }));
            launch_pdl(attention_beam_kernel, dim3(S / rows_per_ctx), dim3(kAttThreads), attention_beam_smem(rows_per_ctx, AH, H), st, rp.part, rp.splits, w->h2att.b,
                       p_att, att, masks, w->alpha_net.w, w->alpha_net.b, sc.ctx, att_w, att_w_stride, S, len_max, H, AH, rows_per_ctx, active,
                       out16 ? h16->ctx_hi : nullptr, out16 ? h16->ctx_lo : nullptr, out16 ? h16->Hp : 0, (int*)(out16 ? w->h3_overflow : nullptr),
                       merged ? AH + 4 * H : AH, fc_pre != nullptr ? 1 : 0);
        } else
        if (!(skip & 8)) launch_pdl(attention_kernel, dim3(S), dim3(kAttThreads), smem, st, rp.part, rp.splits, w->h2att.b, p_att, att, masks, w->alpha_net.w, w->alpha_net.b,
                                                                        sc.ctx, att_w, att_w_stride, S, len_max, H, AH, rows_per_ctx, active,
                                                                        out16 ? h16->ctx_hi : nullptr, out16 ? h16->ctx_lo : nullptr, out16 ? h16->Hp : 0, next_trace_slot(3),
                                                                        (int*)(out16 ? w->h3_overflow : nullptr), merged ? AH + 4 * H : AH,
                                                                        fc_pre != nullptr ? 1 : 0);
        SUBGC_LAUNCH_CHECK();
    }
    if (upto == 1) return SUBGC_OK;
    // language LSTM on [ctx | h_att] (AttModel.py:421-423)
    p = GemmProblem(); p.wts = w;
    p.M = S; p.N = 4 * H; p.nseg = merged ? 1 : 3;
    p.seg[0] = make_seg(sc.ctx, H, w->lang_w_ih, 2 * H, H);
    p.seg[1] = make_seg(h_out, H, w->lang_w_ih + H, 2 * H, H);
    p.seg[2] = make_seg(h_in + SH, H, w->lang_w_hh, H, H);
    p.seg[2].gather = parent;
    if (out16) {
        set_a16(p.seg[0], h16->ctx_hi, h16->ctx_lo, h16->Hp);
        set_a16(p.seg[1], h16->hout_hi, h16->hout_lo, h16->Hp);
    }
    if (in16) set_a16(p.seg[2], h16->hin_hi + SHp, h16->hin_lo + SHp, h16->Hp);
    p.active = active;
    cell_fused = false;
    if (merged) {   // only the ctx segment is left; the cell adds the lang-early partials of the merged contraction
        if (!(skip & 16)) SUBGC_TRY(launch_gemm_raw(p, sc.gemm_ws, sc.gemm_ws_bytes, st, &rp));
    } else if (!(skip & 16)) {
        CellEpilogue ce;
        ce.H = H; ce.c_prev = c_in + SH; ce.parent = parent; ce.b_ih = w->lang_b_ih; ce.b_hh = w->lang_b_hh;
        ce.h_out = h_out + SH; ce.c_out = c_out + SH;
        if (out16) { ce.h16_hi = h16->hout_hi + SHp; ce.h16_lo = h16->hout_lo + SHp; ce.Hp = h16->Hp; }
        SUBGC_TRY(launch_gemm_cell(p, ce, sc.gemm_ws, sc.gemm_ws_bytes, st, &cell_fused));
        if (!cell_fused) SUBGC_TRY(launch_gemm_raw(p, sc.gemm_ws, sc.gemm_ws_bytes, st, &rp));
    }
    if (!(skip & 32) && !cell_fused) launch_pdl(lstm_reduce_cell_kernel, dim3(pw_blocks), dim3(256), (size_t)0, st, rp.part, rp.splits, w->lang_b_ih, w->lang_b_hh, c_in + SH, parent, h_out + SH,
                                                                         c_out + SH, S, H, active, nullptr, 1, out16 ? h16->hout_hi + SHp : nullptr,
                                                                         out16 ? h16->hout_lo + SHp : nullptr, out16 ? h16->Hp : 0, next_trace_slot(2),
                                                                         merged ? rp_early.part + AH : (const float*)nullptr, merged ? rp_early.splits : 0,
                                                                         merged ? AH + 4 * H : 0);
    SUBGC_LAUNCH_CHECK();
    // logit (AttModel.py:336,340); eval mode: dropout is the identity
    p = GemmProblem(); p.wts = w;
    p.M = S; p.N = V1; p.nseg = 1;
    p.seg[0] = make_seg(h_out + SH, H, w->logit.w, H, H);
    if (out16) set_a16(p.seg[0], h16->hout_hi + SHp, h16->hout_lo + SHp, h16->Hp);
    p.active = active;
    if (raw_logits) {
        raw_logits->part = static_cast<const float*>(sc.gemm_ws); raw_logits->splits = 1;
        return (skip & 64) ? SUBGC_OK : launch_gemm_raw(p, sc.gemm_ws, sc.gemm_ws_bytes, st, raw_logits);
    }
    p.epi.bias = w->logit.b;
    p.C = logits; p.ldc = V1;
    return launch_gemm(p, sc.gemm_ws, sc.gemm_ws_bytes, st);
}

// ping-pong storage of the split-fp16 activation copies of a decode loop
struct Step16Bufs {
    unsigned short* h[2][2];   // [ping-pong][hi | lo] -> [2, S, Hp]
    unsigned short* ctx[2];    // [hi | lo] -> [S, Hp]
    unsigned short* xt[2];     // [hi | lo] -> [S, Xp]
    int Hp, Xp;
};
static size_t step16_bytes(const subgc_dims* d, int S) {
    const size_t Hp = (d->rnn + 7) & ~7, Xp = (d->enc + 7) & ~7;
    return 4 * align_up(2 * (size_t)S * Hp * 2, 256) + 2 * align_up((size_t)S * Hp * 2, 256) + 2 * align_up((size_t)S * Xp * 2, 256) + 256;
}
static bool take_step16(const subgc_dims* d, int S, Workspace& ws, Step16Bufs& b) {
    b.Hp = (d->rnn + 7) & ~7; b.Xp = (d->enc + 7) & ~7;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) b.h[i][j] = ws.take<unsigned short>(2 * (size_t)S * b.Hp);
    for (int j = 0; j < 2; ++j) b.ctx[j] = ws.take<unsigned short>((size_t)S * b.Hp);
    for (int j = 0; j < 2; ++j) b.xt[j] = ws.take<unsigned short>((size_t)S * b.Xp);
    return ws.ok();
}
static Step16 step16_of(const Step16Bufs& b, int in, int out, bool has_xt) {
    Step16 s;
    s.hin_hi = b.h[in][0]; s.hin_lo = b.h[in][1]; s.hout_hi = b.h[out][0]; s.hout_lo = b.h[out][1];
    s.ctx_hi = b.ctx[0]; s.ctx_lo = b.ctx[1];
    s.xt_hi = has_xt ? b.xt[0] : nullptr; s.xt_lo = has_xt ? b.xt[1] : nullptr;
    s.Hp = b.Hp; s.Xp = b.Xp;
    return s;
}
// the split copies only pay off when the contractions take the h3 path (packed LSTM / logit weights present)
static bool use_step16(const subgc_weights* w) { return w->packs != nullptr && w->n_packs > 0 && getenv("SUBGC_NO_STEP16") == nullptr; }

// fc_pre[c, :] = W_ih[:, H:2H] fc[c] + b_ih + b_hh for every context row c (step-invariant part of the att-LSTM gates)
static int launch_fc_pre(const subgc_dims* d, const subgc_weights* w, int n_ctx, const float* fc, const StepScratch& sc, cudaStream_t st) {
    const int H = d->rnn, X = d->enc;
    GemmProblem p;
    p.wts = w;
    p.M = n_ctx; p.N = 4 * H; p.nseg = 1;
    p.seg[0] = make_seg(fc, H, w->att_w_ih + H, X + 2 * H, H);
    p.epi.bias = w->att_b_ih; p.epi.bias2 = w->att_b_hh;
    p.C = sc.gates; p.ldc = 4 * H;
    return launch_gemm(p, sc.gemm_ws, sc.gemm_ws_bytes, st);
}

static int check_decode_args(const subgc_dims* d, const subgc_weights* w, int S, int len_max, const char* who) {
    SUBGC_CHECK_ARG(d && w, "%s: null dims/weights", who);
    SUBGC_CHECK_ARG(d->rnn > 0 && d->enc > 0 && d->att_hid > 0 && d->vocab1 > 1 && d->seq_length > 0, "%s: bad decoder dims", who);
    SUBGC_CHECK_ARG(S > 0 && len_max > 0 && len_max <= 64, "%s: bad n_rows/len_max (%d, %d)", who, S, len_max);
    SUBGC_CHECK_ARG((size_t)(2 * d->att_hid + 64 + 4 * d->rnn) * 4 <= 48 * 1024, "%s: att_hid / rnn_size too large for the attention kernel", who);
    return SUBGC_OK;
}

// ---- beam search step: one block per sub-graph -------------------------------------------------------------------
constexpr int kMaxBeam = 8;

struct BeamArgs {
    const float* logits;     // [n_sub*b, V1]
    int V1, T, t, b;
    int length_penalty;      // 0 none, 1 wu, 2 avg
    double lp_alpha;
    int decoding_constraint;
    int* seq_prev; int* seq_next;       // [n_sub, b, T] histories (ping-pong)
    float* lp_prev; float* lp_next;     // [n_sub, b, T]
    float* sum;                         // [n_sub, b] cumulative scores
    long long* it;                      // [n_sub*b] next input tokens
    long long* parent;                  // [n_sub*b] state row each new beam continues
    long long* done_seq; float* done_logps; double* done_p; double* done_unaug_p; int* done_count;  // outputs
    int* done_total;                    // [n_sub] appended so far
    const float* ys_in; const int* ix_in;  // nullable [n_sub*b, kMaxBeam]: per-row top-b (value, token) from beam_topb_kernel
};

// Per-row top-b of the UNK-suppressed log-probs, one 1024-thread block per decode row with the logit row in registers (V1 <= 10 x 1024):
// split-K reduce of the logit partials + bias, log-softmax statistics, b rounds of arg-max with exclusion (descending value, lower index
// first on ties: the order torch.sort gives the reference, CaptionModel.py:131-135).  beam_step_kernel then only merges b x b candidates.
struct BeamTopArgs {
    const float* logits; int splits; const float* bias;   // partials [splits][S][V1] (+ bias) or materialised logits (splits == 0)
    int V1, S, T, t, b, decoding_constraint;
    const int* seq_prev;     // [n_sub, b, T]
    float* ys; int* ix;      // [S, kMaxBeam]
};
__global__ void __maxnreg__(48) beam_topb_kernel(const BeamTopArgs a) {
    pdl_trigger();
    pdl_wait();
    __shared__ float redv[32];
    const int r = blockIdx.x, tid = threadIdx.x, q = r % a.b;
    if (a.t == 0 && q != 0) return;   // the <bos> step expands one row per sub-graph
    float v[kSelVals];
#pragma unroll
    for (int i = 0; i < kSelVals; ++i) v[i] = 0.f;
    if (a.splits > 0) {
        const size_t zs = (size_t)a.S * a.V1;
        const float* p0 = a.logits + (size_t)r * a.V1;
        for (int z = 0; z < a.splits; ++z) {
#pragma unroll
            for (int i = 0; i < kSelVals; ++i) {
                const int j = tid + i * kSelectThreads;
                if (j < a.V1) v[i] += p0[(size_t)z * zs + j];
            }
        }
#pragma unroll
        for (int i = 0; i < kSelVals; ++i) {
            const int j = tid + i * kSelectThreads;
            if (j < a.V1) v[i] += __ldg(a.bias + j);
        }
    } else {
#pragma unroll
        for (int i = 0; i < kSelVals; ++i) {
            const int j = tid + i * kSelectThreads;
            if (j < a.V1) v[i] = a.logits[(size_t)r * a.V1 + j];
        }
    }
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < kSelVals; ++i)
        if (tid + i * kSelectThreads < a.V1) m = fmaxf(m, v[i]);
    m = block_max(m, redv);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kSelVals; ++i)
        if (tid + i * kSelectThreads < a.V1) s += expf(v[i] - m);
    s = block_sum(s, redv);
    const float lz = logf(s);
    const int banned = (a.decoding_constraint && a.t > 0) ? a.seq_prev[(size_t)r * a.T + a.t - 1] : -1;
#pragma unroll
    for (int i = 0; i < kSelVals; ++i) {
        const int j = tid + i * kSelectThreads;
        float lp = (v[i] - m) - lz;
        if (j == banned) lp = -INFINITY;
        if (j == a.V1 - 1) lp = lp - 1000.f;  // UNK suppression (CaptionModel.py:131)
        v[i] = lp;
    }
    // top-b of the row, best first, first index on ties: every warp takes the b best of its own elements (b rounds of a shuffle
    // arg-max, no block barrier), then warp 0 merges the sorted per-warp lists (each round the list heads compete, the winner's list
    // moves up).  One barrier instead of three per round: the kernel is one block of 32 warps per row and was barrier-bound (78 us per
    // step at 640 rows); a vocabulary index belongs to exactly one warp, so the result is the same list as b block-wide rounds.
    __shared__ float cand_v[32][kMaxBeam];
    __shared__ int cand_i[32][kMaxBeam];
    const int lane = tid & 31, wid = tid >> 5;
    unsigned taken = 0;
    for (int c = 0; c < a.b; ++c) {
        float cv = -INFINITY;
        int ci = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < kSelVals; ++i) {
            const int j = tid + i * kSelectThreads;
            if (j >= a.V1 || ((taken >> i) & 1u)) continue;
            if (v[i] > cv || ci == 0x7fffffff) { cv = v[i]; ci = j; }
        }
        warp_argmax(cv, ci);
        if ((ci % kSelectThreads) == tid && ci != 0x7fffffff) taken |= 1u << (ci / kSelectThreads);
        if (lane == 0) { cand_v[wid][c] = cv; cand_i[wid][c] = ci; }
    }
    __syncthreads();
    if (wid == 0) {
        float lv[kMaxBeam];
        int li[kMaxBeam];
#pragma unroll
        for (int c = 0; c < kMaxBeam; ++c) {
            const bool have = c < a.b && lane < kSelectThreads / 32;
            lv[c] = have ? cand_v[lane][c] : -INFINITY;
            li[c] = have ? cand_i[lane][c] : 0x7fffffff;
        }
        for (int c = 0; c < a.b; ++c) {
            float wv = lv[0];
            int wi = li[0];
            warp_argmax(wv, wi);
            if (wi != 0x7fffffff && li[0] == wi) {   // this lane's head won: its list moves up
#pragma unroll
                for (int k = 0; k + 1 < kMaxBeam; ++k) { lv[k] = lv[k + 1]; li[k] = li[k + 1]; }
                lv[kMaxBeam - 1] = -INFINITY; li[kMaxBeam - 1] = 0x7fffffff;
            }
            if (lane == 0) { a.ys[(size_t)r * kMaxBeam + c] = wv; a.ix[(size_t)r * kMaxBeam + c] = wi; }
        }
    }
}

__global__ void __launch_bounds__(256) beam_step_kernel(const BeamArgs a) {
    __shared__ float redv[32];
    __shared__ int redi[32];
    __shared__ float s_ys[kMaxBeam][kMaxBeam];
    __shared__ int s_ix[kMaxBeam][kMaxBeam];
    __shared__ int s_q[kMaxBeam], s_tok[kMaxBeam];
    __shared__ float s_p[kMaxBeam], s_r[kMaxBeam];
    pdl_trigger();
    pdl_wait();
    const int sg = blockIdx.x, b = a.b, t = a.t, T = a.T, V1 = a.V1;
    const int rows = (t == 0) ? 1 : b;
    const int* seq_prev = a.seq_prev + (size_t)sg * b * T;
    int* seq_next = a.seq_next + (size_t)sg * b * T;
    const float* lp_prev = a.lp_prev + (size_t)sg * b * T;
    float* lp_next = a.lp_next + (size_t)sg * b * T;
    if (a.ix_in != nullptr) {   // per-row top-b already computed (beam_topb_kernel)
        if (threadIdx.x < rows * b) {
            const int q = threadIdx.x / b, c = threadIdx.x % b;
            s_ys[q][c] = a.ys_in[((size_t)sg * b + q) * kMaxBeam + c];
            s_ix[q][c] = a.ix_in[((size_t)sg * b + q) * kMaxBeam + c];
        }
        __syncthreads();
    }
    for (int q = 0; q < (a.ix_in != nullptr ? 0 : rows); ++q) {
        const float* x = a.logits + ((size_t)sg * b + q) * V1;
        float m = -INFINITY;
        for (int j = threadIdx.x; j < V1; j += blockDim.x) m = fmaxf(m, x[j]);
        m = block_max(m, redv);
        float s = 0.f;
        for (int j = threadIdx.x; j < V1; j += blockDim.x) s += expf(x[j] - m);
        s = block_sum(s, redv);
        const float lz = logf(s);
        const int banned = (a.decoding_constraint && t > 0) ? seq_prev[q * T + t - 1] : -1;
        for (int c = 0; c < b; ++c) {
            float cv = -INFINITY;
            int ci = 0x7fffffff;
            for (int j = threadIdx.x; j < V1; j += blockDim.x) {
                bool taken = false;
                for (int e = 0; e < c; ++e) taken |= (s_ix[q][e] == j);
                if (taken) continue;
                float v = (x[j] - m) - lz;
                if (j == banned) v = -INFINITY;
                if (j == V1 - 1) v = v - 1000.f;  // UNK suppression (CaptionModel.py:131)
                if (v > cv || ci == 0x7fffffff) { cv = v; ci = j; }
            }
            block_argmax(cv, ci, redv, redi);
            if (threadIdx.x == 0) { s_ys[q][c] = cv; s_ix[q][c] = ci; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        // candidates c-major / q-minor, stable descending sort by p = fl32(sum[q] + ys[q][c]) (CaptionModel.py:61-69)
        float cp[kMaxBeam * kMaxBeam];
        unsigned char cq[kMaxBeam * kMaxBeam], cc[kMaxBeam * kMaxBeam];
        int n = 0;
        for (int c = 0; c < b; ++c)
            for (int q = 0; q < rows; ++q) {
                float p = a.sum[sg * b + q] + s_ys[q][c];
                int pos = n;
                while (pos > 0 && cp[pos - 1] < p) { cp[pos] = cp[pos - 1]; cq[pos] = cq[pos - 1]; cc[pos] = cc[pos - 1]; --pos; }
                cp[pos] = p; cq[pos] = (unsigned char)q; cc[pos] = (unsigned char)c;
                ++n;
            }
        for (int v = 0; v < b; ++v) {
            s_q[v] = cq[v]; s_tok[v] = s_ix[cq[v]][cc[v]]; s_p[v] = cp[v]; s_r[v] = s_ys[cq[v]][cc[v]];
        }
    }
    __syncthreads();
    // fork histories (beam v continues beam s_q[v]) and append the new token
    for (int idx = threadIdx.x; idx < b * T; idx += blockDim.x) {
        int v = idx / T, i = idx - v * T;
        int sv = 0;
        float lv = 0.f;
        if (i < t) { sv = seq_prev[s_q[v] * T + i]; lv = lp_prev[s_q[v] * T + i]; }
        else if (i == t) { sv = s_tok[v]; lv = s_r[v]; }
        seq_next[idx] = sv;
        lp_next[idx] = lv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int v = 0; v < b; ++v) {
            a.sum[sg * b + v] = s_p[v];
            a.it[sg * b + v] = s_tok[v];
            a.parent[sg * b + v] = (long long)sg * b + s_q[v];
        }
        for (int v = 0; v < b; ++v) {
            if (s_tok[v] == 0 || t == T - 1) {  // finished beam -> done list (CaptionModel.py:149-162)
                double p = (double)a.sum[sg * b + v];
                const int len = t + 1;
                if (a.length_penalty == 1) p = p / (pow(5.0 + (double)len, a.lp_alpha) / pow(6.0, a.lp_alpha));
                else if (a.length_penalty == 2) p = p / (double)len;
                float un = 0.f;
                for (int i = 0; i < T; ++i) un += lp_next[v * T + i];
                // stable insertion into the descending-p list, truncated to b entries
                int cnt = a.done_total[sg] < b ? a.done_total[sg] : b;
                int pos = cnt;
                while (pos > 0 && a.done_p[sg * b + pos - 1] < p) --pos;
                if (pos < b) {
                    int last = cnt < b ? cnt : b - 1;
                    for (int e = last; e > pos; --e) {
                        for (int i = 0; i < T; ++i) {
                            a.done_seq[((size_t)sg * b + e) * T + i] = a.done_seq[((size_t)sg * b + e - 1) * T + i];
                            a.done_logps[((size_t)sg * b + e) * T + i] = a.done_logps[((size_t)sg * b + e - 1) * T + i];
                        }
                        a.done_p[sg * b + e] = a.done_p[sg * b + e - 1];
                        a.done_unaug_p[sg * b + e] = a.done_unaug_p[sg * b + e - 1];
                    }
                    for (int i = 0; i < T; ++i) {
                        a.done_seq[((size_t)sg * b + pos) * T + i] = seq_next[v * T + i];
                        a.done_logps[((size_t)sg * b + pos) * T + i] = lp_next[v * T + i];
                    }
                    a.done_p[sg * b + pos] = p;
                    a.done_unaug_p[sg * b + pos] = (double)un;
                }
                a.done_total[sg] += 1;
                a.done_count[sg] = a.done_total[sg] < b ? a.done_total[sg] : b;
                a.sum[sg * b + v] = -1000.f;  // don't continue beams from finished sequences
            }
        }
    }
}

}  // namespace subgc

using namespace subgc;

extern "C" int subgc_debug_att_trace(unsigned long long* host_out, int n_blocks) {
    unsigned long long* buf = att_trace_buffer();
    if (!buf || !host_out || n_blocks > 1024) return SUBGC_E_INVALID;
    cudaDeviceSynchronize();
    return cudaMemcpy(host_out, buf, (size_t)n_blocks * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? SUBGC_OK : SUBGC_E_CUDA;
}

extern "C" int subgc_log_softmax_fwd(int rows, int V1, const float* logits, float* logp, size_t ld_out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(logits && logp && rows > 0 && V1 > 0 && ld_out >= (size_t)V1, "subgc_log_softmax_fwd: bad arguments");
    launch_pdl(log_softmax_kernel, dim3(rows), dim3(256), (size_t)0, static_cast<cudaStream_t>(stream), logits, logp, V1, ld_out, (const int*)nullptr);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" size_t subgc_decode_workspace_bytes(const subgc_dims* d, int n_rows, int len_max) {
    if (!d || n_rows <= 0) return 0;
    (void)len_max;
    const size_t S = n_rows, H = d->rnn;
    size_t b = step_scratch_bytes(d, n_rows);
    b += 4 * align_up(2 * S * H * 4, 256);                 // h / c ping-pong
    b += align_up(S * d->vocab1 * 4, 256);                 // logits
    b += align_up(S * 8, 256) + align_up(S * 4, 256) + align_up(S * d->enc * 4, 256);  // it, unfinished, xt
    b += align_up((size_t)(d->seq_length + 2) * 4, 256);   // count
    b += step16_bytes(d, n_rows);                          // split-fp16 activation copies
    if (n_rows <= 128) b += mega_decode_scratch_bytes_max(d);   // activation tiles / partials / counters of the persistent kernel
    return b + 1024;
}

extern "C" int subgc_decode_step(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int rows_per_ctx, const int64_t* it,
                                 const float* fc, const float* att, const float* p_att, const float* masks, const float* h_in,
                                 const float* c_in, float* h_out, float* c_out, float* logprobs, float* att_weights, void* ws_,
                                 size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_decode_args(d, w, n_rows, len_max, "subgc_decode_step"));
    SUBGC_CHECK_ARG(it && fc && att && p_att && masks && h_in && c_in && h_out && c_out && logprobs, "subgc_decode_step: null argument");
    SUBGC_CHECK_ARG(rows_per_ctx >= 1 && n_rows % rows_per_ctx == 0, "subgc_decode_step: n_rows must be a multiple of rows_per_ctx");
    SUBGC_CHECK_ARG(h_in != h_out && c_in != c_out, "subgc_decode_step: state in/out must be distinct buffers");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace ws(ws_, ws_bytes);
    StepScratch sc;
    bool ok = take_step_scratch(d, n_rows, ws, sc);
    float* logits = ws.take<float>((size_t)n_rows * d->vocab1);
    if (!ok || !ws.ok()) { set_error("subgc_decode_step: workspace too small"); return SUBGC_E_WORKSPACE; }
    SUBGC_TRY(launch_step(d, w, n_rows, len_max, rows_per_ctx, reinterpret_cast<const long long*>(it), nullptr, nullptr, fc, att, p_att, masks, h_in,
                          c_in, h_out, c_out, logits, nullptr, att_weights, len_max, sc, nullptr, 0, st));
    launch_pdl(log_softmax_kernel, dim3(n_rows), dim3(256), (size_t)0, st, (const float*)logits, logprobs, (int)d->vocab1, (size_t)d->vocab1, (const int*)nullptr);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

static int decode_sample_impl(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int mode, float temp, int top_k,
                              uint64_t seed, uint64_t offset, const float* uniforms, const float* fc, const float* att,
                              const float* p_att, const float* masks, int64_t* seq, float* seq_logprobs, float* att_weights,
                              int32_t* steps_done, void* ws_, size_t ws_bytes, subgc_stream_t stream, const int32_t* counts) {
    SUBGC_TRY(check_decode_args(d, w, n_rows, len_max, "subgc_decode_sample"));
    SUBGC_CHECK_ARG(fc && att && p_att && masks && seq && seq_logprobs && steps_done, "subgc_decode_sample: null argument");
    SUBGC_CHECK_ARG(mode == 0 || mode == 1, "subgc_decode_sample: mode must be 0 (greedy) or 1 (top-k)");
    SUBGC_CHECK_ARG(mode == 0 || (top_k >= 1 && top_k <= kMaxTopK && top_k <= d->vocab1 && temp > 0.f),
                    "subgc_decode_sample: top_k must be in [1, %d] and temp > 0", kMaxTopK);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = n_rows, H = d->rnn, T = d->seq_length, V1 = d->vocab1;
    Workspace ws(ws_, ws_bytes);
    StepScratch sc;
    bool ok = take_step_scratch(d, S, ws, sc);
    float* hbuf[2] = {ws.take<float>(2 * (size_t)S * H), ws.take<float>(2 * (size_t)S * H)};
    float* cbuf[2] = {ws.take<float>(2 * (size_t)S * H), ws.take<float>(2 * (size_t)S * H)};
    float* logits = ws.take<float>((size_t)S * V1);
    long long* it = ws.take<long long>(S);
    int* unfinished = ws.take<int>(S);
    int* count = ws.take<int>(T + 2);
    float* xt = ws.take<float>((size_t)S * d->enc);
    Step16Bufs b16;
    const bool s16 = use_step16(w);
    if (s16) ok = take_step16(d, S, ws, b16) && ok;
    if (!ok || !ws.ok()) { set_error("subgc_decode_sample: workspace too small"); return SUBGC_E_WORKSPACE; }
    SUBGC_CUDA(cudaMemsetAsync(seq, 0, (size_t)S * T * 8, st));
    SUBGC_CUDA(cudaMemsetAsync(seq_logprobs, 0, (size_t)S * T * 4, st));
    if (mega_decode_eligible(d, w, S, len_max, att_weights)) {   // the whole loop as one persistent kernel (mega_decode.cu)
        SUBGC_TRY(launch_fc_pre(d, w, S, fc, sc, st));
        return launch_mega_decode(d, w, S, len_max, mode, temp, top_k, seed, offset, uniforms, sc.gates, att, p_att, masks, seq, seq_logprobs,
                                  steps_done, ws, st, counts);
    }
    if (s16) {
        SUBGC_CUDA(cudaMemsetAsync(b16.h[0][0], 0, 2 * (size_t)S * b16.Hp * 2, st));   // fp16 zeros: split copy of the zero state
        SUBGC_CUDA(cudaMemsetAsync(b16.h[0][1], 0, 2 * (size_t)S * b16.Hp * 2, st));
    }
    SUBGC_CUDA(cudaMemsetAsync(hbuf[0], 0, 2 * (size_t)S * H * 4, st));   // init_hidden (AttModel.py:343-346)
    SUBGC_CUDA(cudaMemsetAsync(cbuf[0], 0, 2 * (size_t)S * H * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(it, 0, (size_t)S * 8, st));                // <bos>
    SUBGC_CUDA(cudaMemsetAsync(count, 0, (size_t)(T + 2) * 4, st));
    if (att_weights) SUBGC_CUDA(cudaMemsetAsync(att_weights, 0, (size_t)S * (T + 1) * len_max * 4, st));
    if ((size_t)V1 * sizeof(float) > 48 * 1024) {
        SUBGC_CHECK_ARG((size_t)V1 * sizeof(float) <= 200 * 1024, "subgc_decode_sample: vocabulary too large for the selection kernel");
        SUBGC_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(V1 * sizeof(float))));
    }
    SUBGC_TRY(launch_fc_pre(d, w, S, fc, sc, st));
    if (counts) {
        set_error("subgc_decode_sample_dyn: device-side row counts need the persistent decode kernel (w->mega, <= 128 rows, no attention weights)");
        return SUBGC_E_UNSUPPORTED;
    }
    for (int t = 0; t <= T; ++t) {
        const int* active = (t == 0) ? nullptr : count + (t - 1);
        const int in = t & 1, out = in ^ 1;
        float* aw = att_weights ? att_weights + (size_t)t * len_max : nullptr;
        const Step16 h16v = s16 ? step16_of(b16, in, out, t > 0) : Step16();
        const Step16* h16 = s16 ? &h16v : nullptr;
        if (t == T) {
            // the reference runs this step and discards its log-probs (AttModel.py:292-293); only the attention
            // weights are observable, so the step stops there (and is skipped entirely when they are not requested)
            if (att_weights)
                SUBGC_TRY(launch_step(d, w, S, len_max, 1, it, t > 0 ? xt : nullptr, nullptr, fc, att, p_att, masks, hbuf[in], cbuf[in], hbuf[out], cbuf[out], logits,
                                      nullptr, aw, (T + 1) * len_max, sc, active, 1, st, sc.gates, h16));
            break;
        }
        RawPartials rl;
        SUBGC_TRY(launch_step(d, w, S, len_max, 1, it, t > 0 ? xt : nullptr, nullptr, fc, att, p_att, masks, hbuf[in], cbuf[in], hbuf[out], cbuf[out], logits, &rl, aw,
                              (T + 1) * len_max, sc, active, 0, st, sc.gates, h16));
        SelectArgs a;
        a.logits = rl.part; a.splits = rl.splits; a.bias = w->logit.b; a.V1 = V1; a.T = T; a.t = t; a.S = S; a.mode = mode; a.temp = temp; a.top_k = top_k; a.seed = seed;
        a.offset = offset; a.uniforms = uniforms; a.it = it; a.unfinished = unfinished; a.seq = reinterpret_cast<long long*>(seq);
        a.seq_lp = seq_logprobs; a.count = count; a.active = active;
        a.embed = w->embed; a.xt = xt; a.X = d->enc;
        a.xt16_hi = s16 ? b16.xt[0] : nullptr; a.xt16_lo = s16 ? b16.xt[1] : nullptr; a.Xp = s16 ? b16.Xp : 0;
        a.trace = next_trace_slot(4);
        a.overflow = s16 ? w->h3_overflow : nullptr;
        if (!(skip_mask() & 128)) {
            if (V1 <= kSelVals * kSelectThreads) launch_pdl(select_reg_kernel, dim3(S), dim3(kSelectThreads), (size_t)0, st, a);
            else launch_pdl(select_kernel, dim3(S), dim3(kSelectThreads), (size_t)V1 * sizeof(float), st, a);
        }
        SUBGC_LAUNCH_CHECK();
    }
    steps_done_kernel<<<1, 1, 0, st>>>(count, T, steps_done);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_decode_sample(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int mode, float temp, int top_k,
                                   uint64_t seed, uint64_t offset, const float* uniforms, const float* fc, const float* att,
                                   const float* p_att, const float* masks, int64_t* seq, float* seq_logprobs, float* att_weights,
                                   int32_t* steps_done, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    return decode_sample_impl(d, w, n_rows, len_max, mode, temp, top_k, seed, offset, uniforms, fc, att, p_att, masks, seq, seq_logprobs, att_weights,
                              steps_done, ws_, ws_bytes, stream, nullptr);
}

extern "C" int subgc_decode_sample_dyn(const subgc_dims* d, const subgc_weights* w, int rows_cap, int len_cap, const int32_t* counts, int mode,
                                       float temp, int top_k, uint64_t seed, uint64_t offset, const float* uniforms, const float* fc, const float* att,
                                       const float* p_att, const float* masks, int64_t* seq, float* seq_logprobs, int32_t* steps_done, void* ws_,
                                       size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(counts != nullptr, "subgc_decode_sample_dyn: null counts");
    return decode_sample_impl(d, w, rows_cap, len_cap, mode, temp, top_k, seed, offset, uniforms, fc, att, p_att, masks, seq, seq_logprobs, nullptr,
                              steps_done, ws_, ws_bytes, stream, counts);
}

extern "C" size_t subgc_teacher_workspace_bytes(const subgc_dims* d, int n_rows, int n_steps) {
    if (!d || n_rows <= 0 || n_steps <= 0) return 0;
    return subgc_decode_workspace_bytes(d, n_rows, 0) + align_up((size_t)n_rows * n_steps * 8, 256) + align_up((size_t)n_steps * 4, 256) + 512;
}

extern "C" int subgc_decode_teacher(const subgc_dims* d, const subgc_weights* w, int n_rows, int len_max, int n_steps, const int64_t* tokens,
                                    int ld_tok, const float* fc, const float* att, const float* p_att, const float* masks, float* outputs,
                                    void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_TRY(check_decode_args(d, w, n_rows, len_max, "subgc_decode_teacher"));
    SUBGC_CHECK_ARG(tokens && fc && att && p_att && masks && outputs && n_steps > 0 && ld_tok >= n_steps, "subgc_decode_teacher: bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int S = n_rows, H = d->rnn, V1 = d->vocab1;
    Workspace ws(ws_, ws_bytes);
    StepScratch sc;
    bool ok = take_step_scratch(d, S, ws, sc);
    float* hbuf[2] = {ws.take<float>(2 * (size_t)S * H), ws.take<float>(2 * (size_t)S * H)};
    float* cbuf[2] = {ws.take<float>(2 * (size_t)S * H), ws.take<float>(2 * (size_t)S * H)};
    float* logits = ws.take<float>((size_t)S * V1);
    long long* tok_cols = ws.take<long long>((size_t)S * n_steps);
    int* flags = ws.take<int>(n_steps);
    if (!ok || !ws.ok()) { set_error("subgc_decode_teacher: workspace too small"); return SUBGC_E_WORKSPACE; }
    SUBGC_CUDA(cudaMemsetAsync(hbuf[0], 0, 2 * (size_t)S * H * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(cbuf[0], 0, 2 * (size_t)S * H * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(outputs, 0, (size_t)S * n_steps * V1 * 4, st));  // steps after the early break stay zero
    teacher_columns_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const long long*>(tokens), ld_tok, S, n_steps, tok_cols, flags);
    SUBGC_LAUNCH_CHECK();
    SUBGC_TRY(launch_fc_pre(d, w, S, fc, sc, st));
    for (int i = 0; i < n_steps; ++i) {
        const int in = i & 1, out = in ^ 1;
        SUBGC_TRY(launch_step(d, w, S, len_max, 1, tok_cols + (size_t)i * S, nullptr, nullptr, fc, att, p_att, masks, hbuf[in], cbuf[in], hbuf[out],
                              cbuf[out], logits, nullptr, nullptr, 0, sc, flags + i, 0, st, sc.gates));
        launch_pdl(log_softmax_kernel, dim3(S), dim3(256), (size_t)0, st, (const float*)logits, outputs + (size_t)i * V1, V1, (size_t)n_steps * V1, (const int*)(flags + i));
        SUBGC_LAUNCH_CHECK();
    }
    return SUBGC_OK;
}

extern "C" size_t subgc_beam_workspace_bytes(const subgc_dims* d, int n_sub, int beam_size, int len_max) {
    if (!d || n_sub <= 0 || beam_size <= 0) return 0;
    (void)len_max;
    const size_t S = (size_t)n_sub * beam_size, H = d->rnn, T = d->seq_length;
    size_t b = step_scratch_bytes(d, (int)S);
    b += 4 * align_up(2 * S * H * 4, 256);
    b += align_up(S * d->vocab1 * 4, 256);
    b += 2 * align_up(S * 8, 256);                          // it, parent
    b += 4 * align_up(S * T * 4, 256);                      // histories
    b += align_up(S * 4, 256) + align_up((size_t)n_sub * 4, 256);
    b += 2 * align_up(S * kMaxBeam * 4, 256);                // per-row top-b candidates
    b += step16_bytes(d, (int)S);                            // split-fp16 copies of this step's h_att / ctx / h_lang
    return b + 1024;
}

extern "C" int subgc_decode_beam(const subgc_dims* d, const subgc_weights* w, int n_sub, int len_max, int beam_size, int length_penalty,
                                 double lp_alpha, int decoding_constraint, const float* fc, const float* att, const float* p_att,
                                 const float* masks, int64_t* done_seq, float* done_logps, double* done_p, double* done_unaug_p,
                                 int32_t* done_count, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(beam_size >= 1 && beam_size <= kMaxBeam, "subgc_decode_beam: beam_size must be in [1, %d]", kMaxBeam);
    SUBGC_CHECK_ARG(n_sub > 0, "subgc_decode_beam: n_sub must be positive");
    SUBGC_TRY(check_decode_args(d, w, n_sub * beam_size, len_max, "subgc_decode_beam"));
    SUBGC_CHECK_ARG(fc && att && p_att && masks && done_seq && done_logps && done_p && done_unaug_p && done_count,
                    "subgc_decode_beam: null argument");
    SUBGC_CHECK_ARG(length_penalty >= 0 && length_penalty <= 2, "subgc_decode_beam: unknown length penalty %d", length_penalty);
    SUBGC_CHECK_ARG(beam_size <= d->vocab1, "subgc_decode_beam: beam_size exceeds the vocabulary");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int b = beam_size, S = n_sub * b, H = d->rnn, T = d->seq_length, V1 = d->vocab1;
    Workspace ws(ws_, ws_bytes);
    StepScratch sc;
    bool ok = take_step_scratch(d, S, ws, sc);
    float* hbuf[2] = {ws.take<float>(2 * (size_t)S * H), ws.take<float>(2 * (size_t)S * H)};
    float* cbuf[2] = {ws.take<float>(2 * (size_t)S * H), ws.take<float>(2 * (size_t)S * H)};
    float* logits = ws.take<float>((size_t)S * V1);
    long long* it = ws.take<long long>(S);
    long long* parent = ws.take<long long>(S);
    int* seqb[2] = {ws.take<int>((size_t)S * T), ws.take<int>((size_t)S * T)};
    float* lpb[2] = {ws.take<float>((size_t)S * T), ws.take<float>((size_t)S * T)};
    float* sum = ws.take<float>(S);
    int* done_total = ws.take<int>(n_sub);
    float* top_ys = ws.take<float>((size_t)S * kMaxBeam);
    int* top_ix = ws.take<int>((size_t)S * kMaxBeam);
    Step16Bufs b16;
    const bool s16 = use_step16(w);
    if (s16) ok = take_step16(d, S, ws, b16) && ok;
    if (!ok || !ws.ok()) { set_error("subgc_decode_beam: workspace too small"); return SUBGC_E_WORKSPACE; }
    Step16 h16v = s16 ? step16_of(b16, 0, 1, false) : Step16();   // only this step's outputs are read as split copies:
    h16v.hin_hi = nullptr; h16v.hin_lo = nullptr;                  // the previous state is re-mapped by `parent` (and never written in split form)
    const Step16* h16 = s16 ? &h16v : nullptr;
    const bool fast_top = V1 <= kSelVals * kSelectThreads;   // logit row fits the register-resident top-b kernel
    RawPartials rl{nullptr, 0};
    SUBGC_CUDA(cudaMemsetAsync(hbuf[0], 0, 2 * (size_t)S * H * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(cbuf[0], 0, 2 * (size_t)S * H * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(it, 0, (size_t)S * 8, st));
    SUBGC_CUDA(cudaMemsetAsync(seqb[0], 0, (size_t)S * T * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(lpb[0], 0, (size_t)S * T * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(sum, 0, (size_t)S * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(done_total, 0, (size_t)n_sub * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(done_count, 0, (size_t)n_sub * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(done_seq, 0, (size_t)S * T * 8, st));
    SUBGC_CUDA(cudaMemsetAsync(done_logps, 0, (size_t)S * T * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(done_p, 0, (size_t)S * 8, st));
    SUBGC_CUDA(cudaMemsetAsync(done_unaug_p, 0, (size_t)S * 8, st));
    // <bos> step on b identical rows per sub-graph (AttModel.py:216-227)
    SUBGC_TRY(launch_fc_pre(d, w, n_sub, fc, sc, st));
    SUBGC_TRY(launch_step(d, w, S, len_max, b, it, nullptr, nullptr, fc, att, p_att, masks, hbuf[0], cbuf[0], hbuf[1], cbuf[1], logits,
                          fast_top ? &rl : nullptr, nullptr, 0, sc, nullptr, 0, st, sc.gates, h16));
    for (int t = 0; t < T; ++t) {
        BeamArgs a;
        a.ys_in = nullptr; a.ix_in = nullptr;
        if (fast_top) {
            BeamTopArgs ta;
            ta.logits = rl.part; ta.splits = rl.splits; ta.bias = w->logit.b; ta.V1 = V1; ta.S = S; ta.T = T; ta.t = t; ta.b = b;
            ta.decoding_constraint = decoding_constraint; ta.seq_prev = seqb[t & 1]; ta.ys = top_ys; ta.ix = top_ix;
            launch_pdl(beam_topb_kernel, dim3(S), dim3(kSelectThreads), (size_t)0, st, ta);
            SUBGC_LAUNCH_CHECK();
            a.ys_in = top_ys; a.ix_in = top_ix;
        }
        a.logits = logits; a.V1 = V1; a.T = T; a.t = t; a.b = b; a.length_penalty = length_penalty; a.lp_alpha = lp_alpha;
        a.decoding_constraint = decoding_constraint;
        a.seq_prev = seqb[t & 1]; a.seq_next = seqb[(t & 1) ^ 1]; a.lp_prev = lpb[t & 1]; a.lp_next = lpb[(t & 1) ^ 1];
        a.sum = sum; a.it = it; a.parent = parent;
        a.done_seq = reinterpret_cast<long long*>(done_seq); a.done_logps = done_logps; a.done_p = done_p; a.done_unaug_p = done_unaug_p;
        a.done_count = done_count; a.done_total = done_total;
        launch_pdl(beam_step_kernel, dim3(n_sub), dim3(256), (size_t)0, st, a);
        SUBGC_LAUNCH_CHECK();
        if (t == T - 1) break;  // the reference's final get_logprobs_state result is never read (CaptionModel.py:170-171)
        const int in = (t + 1) & 1, out = in ^ 1;
        SUBGC_TRY(launch_step(d, w, S, len_max, b, it, nullptr, parent, fc, att, p_att, masks, hbuf[in], cbuf[in], hbuf[out], cbuf[out], logits,
                              fast_top ? &rl : nullptr, nullptr, 0, sc, nullptr, 0, st, sc.gates, h16));
    }
    return SUBGC_OK;
}
