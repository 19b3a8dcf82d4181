// Training-side building blocks of the Sub-GC path: the backward halves of the decoder / sGPN / GCN stages and the
// train-mode (dropout, saved activations) variants of the forward kernels.  The host side (subgc/train.py) composes
// them into AttModel._forward + LossWrapper + backward (reference models/AttModel.py:122-177, models/loss_wrapper.py:14-27,
// misc/utils.py:111-124); every contraction goes through the same GEMM block as inference (C = A . W^T, operands
// K-major — transposed operands are materialised with subgc_transpose so that the tensor-core path applies).
#include "common.cuh"

namespace subgc {

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows, int cols, int ld_in, float* __restrict__ out,
                                                        int ld_out) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        int r = r0 + i, c = c0 + tx;
        if (r < rows && c < cols) tile[i][tx] = in[(size_t)r * ld_in + c];
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, r = r0 + tx;
        if (r < rows && c < cols) out[(size_t)c * ld_out + r] = tile[tx][i];
    }
}

// out[c] (+)= sum_r in[r, c]; one block per 32 columns, fixed summation order
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ in, int rows, int cols, int ld, float* __restrict__ out,
                                                     int accumulate) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (c < cols)
        for (int r = ty; r < rows; r += 8) a += in[(size_t)r * ld + c];
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        out[c] = accumulate ? out[c] + t : t;
    }
}

// generic element-wise ops (grid-stride)
enum { EW_MUL = 0, EW_ADD = 1, EW_RELU_BWD = 2, EW_SCALE = 3, EW_COPY = 4 };
__global__ void __launch_bounds__(256) ew_kernel(int op, size_t n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                 float scalar) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v;
        switch (op) {
            case EW_MUL: v = a[i] * b[i]; break;
            case EW_ADD: v = a[i] + b[i]; break;
            case EW_RELU_BWD: v = a[i] > 0.f ? b[i] : 0.f; break;  // a = forward output, b = upstream gradient
            case EW_SCALE: v = a[i] * scalar; break;
            default: v = a[i]; break;
        }
        out[i] = v;
    }
}

// Philox4x32-10 dropout mask: mask[i] = (u >= p) ? 1/(1-p) : 0, keyed by (seed, offset, i)
__device__ __forceinline__ void philox_round(unsigned (&c)[4], unsigned k0, unsigned k1) {
    unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
__global__ void __launch_bounds__(256) dropout_mask_kernel(size_t n, float p, unsigned long long seed, unsigned long long offset,
                                                           float* __restrict__ mask) {
    const float keep = 1.f / (1.f - p);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < n; q += (size_t)gridDim.x * blockDim.x) {
        unsigned c[4] = {(unsigned)q, (unsigned)(q >> 32), (unsigned)offset, (unsigned)(offset >> 32)};
        unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
        for (int i = 0; i < 10; ++i) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        for (int j = 0; j < 4; ++j) {
            size_t i = q * 4 + j;
            if (i < n) mask[i] = ((float)(c[j] >> 8) * (1.0f / 16777216.0f) >= p) ? keep : 0.f;
        }
    }
}

// Scheduled sampling (reference models/AttModel.py:158-167): with probability ss_prob the token fed to step i is drawn from
// exp(outputs[:, i-1]) instead of the ground truth.  One block per row: u1 decides, u2 is mapped through the inverse CDF of the row's
// probabilities (torch.multinomial draws from the same distribution; the RNG streams differ).  Philox4x32-10 keyed by (seed; row, offset).
__global__ void __launch_bounds__(256) ss_sample_kernel(const float* __restrict__ prev_logp, size_t ld, int V1, const long long* __restrict__ labels,
                                                        int ld_lab, float ss_prob, unsigned long long seed, unsigned long long offset,
                                                        long long* __restrict__ it) {
    __shared__ float s_part[256];
    __shared__ float s_u[2];
    const int r = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        unsigned c[4] = {(unsigned)r, 0x5353u, (unsigned)offset, (unsigned)(offset >> 32)};
        unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
        for (int i = 0; i < 10; ++i) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        s_u[0] = (float)(c[0] >> 8) * (1.0f / 16777216.0f);
        s_u[1] = (float)(c[1] >> 8) * (1.0f / 16777216.0f);
    }
    __syncthreads();
    const long long truth = labels[(size_t)r * ld_lab];
    if (!(s_u[0] < ss_prob)) {   // uniform decision per row, so the whole block takes the same branch
        if (tid == 0) it[r] = truth;
        return;
    }
    const float* lp = prev_logp + (size_t)r * ld;
    const int per = (V1 + 255) / 256, lo = tid * per, hi = min(V1, lo + per);   // contiguous slice per thread: the CDF runs in index order
    float s = 0.f;
    for (int j = lo; j < hi; ++j) s += expf(lp[j]);
    s_part[tid] = s;
    __syncthreads();
    if (tid == 0) {
        float tot = 0.f;
        for (int i = 0; i < 256; ++i) tot += s_part[i];
        const float target = s_u[1] * tot;
        float run = 0.f;
        int blk = 255;
        for (int i = 0; i < 256; ++i) {
            if (run + s_part[i] > target) { blk = i; break; }
            run += s_part[i];
        }
        int tok = min(V1, (blk + 1) * per) - 1;
        for (int j = blk * per; j < min(V1, (blk + 1) * per); ++j) {
            run += expf(lp[j]);
            if (run > target) { tok = j; break; }
        }
        it[r] = tok;
    }
}

// dst[idx[r], :] += src[r, :]   (embedding / node-feature gradient; atomics: order-independent up to fp32 rounding)
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, int n_rows,
                                                               int cols, int ld_src, float* __restrict__ dst, int ld_dst) {
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    float* d = dst + (size_t)idx[r] * ld_dst;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) atomicAdd(d + c, src[(size_t)r * ld_src + c]);
}

// LSTM cell, training flavour: gates (pre-activation, biases included) are overwritten with their activations
// (i, f, o: sigmoid; g: tanh), which is what the backward pass needs.
__global__ void __launch_bounds__(256) lstm_cell_train_fwd_kernel(float* __restrict__ gates, const float* __restrict__ c_prev,
                                                                  float* __restrict__ h_out, float* __restrict__ c_out, int S, int H) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * H) return;
    int r = idx / H, j = idx - r * H;
    float* g = gates + (size_t)r * 4 * H;
    float i = sigmoidf_(g[j]), f = sigmoidf_(g[H + j]), gg = tanhf(g[2 * H + j]), o = sigmoidf_(g[3 * H + j]);
    float c = f * c_prev[idx] + i * gg;
    g[j] = i; g[H + j] = f; g[2 * H + j] = gg; g[3 * H + j] = o;
    c_out[idx] = c;
    h_out[idx] = o * tanhf(c);
}

// dgates (pre-activation) and dc_prev from dh, dc (nullable), the saved activations and cell states
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const float* __restrict__ act, const float* __restrict__ c_prev,
                                                            const float* __restrict__ c_new, const float* __restrict__ dh,
                                                            const float* __restrict__ dc_in, float* __restrict__ dgates,
                                                            float* __restrict__ dc_prev, int S, int H) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * H) return;
    int r = idx / H, j = idx - r * H;
    const float* a = act + (size_t)r * 4 * H;
    float i = a[j], f = a[H + j], g = a[2 * H + j], o = a[3 * H + j];
    float tc = tanhf(c_new[idx]);
    float dhv = dh[idx];
    float dc = dhv * o * (1.f - tc * tc) + (dc_in ? dc_in[idx] : 0.f);
    float* dg = dgates + (size_t)r * 4 * H;
    dg[j] = dc * g * i * (1.f - i);
    dg[H + j] = dc * c_prev[idx] * f * (1.f - f);
    dg[2 * H + j] = dc * i * (1.f - g * g);
    dg[3 * H + j] = dhv * tc * o * (1.f - o);
    dc_prev[idx] = dc * f;
}

// Attention forward, training flavour: also stores the pre-mask softmax `sm` [S, len] next to alpha.
__global__ void __launch_bounds__(256) attention_train_fwd_kernel(const float* __restrict__ atth, const float* __restrict__ p_att,
                                                                  const float* __restrict__ att, const float* __restrict__ masks,
                                                                  const float* __restrict__ alpha_w, const float* __restrict__ alpha_b,
                                                                  float* __restrict__ ctx, float* __restrict__ alpha, float* __restrict__ sm,
                                                                  int len, int H, int AH) {
    extern __shared__ float s_a[];  // [AH] atth | [AH] w | [len] e
    float* s_h = s_a; float* s_w = s_a + AH; float* s_e = s_a + 2 * AH;
    const int r = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = threadIdx.x; j < AH; j += blockDim.x) { s_h[j] = atth[(size_t)r * AH + j]; s_w[j] = alpha_w[j]; }
    __syncthreads();
    const float* pa = p_att + (size_t)r * len * AH;
    for (int n = wid; n < len; n += nw) {
        float a = 0.f;
        for (int j = lane; j < AH; j += 32) a = fmaf(s_w[j], tanhf(pa[(size_t)n * AH + j] + s_h[j]), a);
        a = warp_sum(a);
        if (lane == 0) s_e[n] = a + alpha_b[0];
    }
    __syncthreads();
    if (wid == 0) {
        float m = -INFINITY;
        for (int n = lane; n < len; n += 32) m = fmaxf(m, s_e[n]);
        m = warp_max(m);
        float sum = 0.f;
        for (int n = lane; n < len; n += 32) sum += expf(s_e[n] - m);
        sum = warp_sum(sum);
        float msum = 0.f;
        for (int n = lane; n < len; n += 32) {
            float sv = expf(s_e[n] - m) / sum;
            sm[(size_t)r * len + n] = sv;
            float wv = sv * masks[(size_t)r * len + n];
            s_e[n] = wv;
            msum += wv;
        }
        msum = warp_sum(msum);
        for (int n = lane; n < len; n += 32) {
            float wv = s_e[n] / msum;
            s_e[n] = wv;
            alpha[(size_t)r * len + n] = wv;
        }
    }
    __syncthreads();
    const float* af = att + (size_t)r * len * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float a = 0.f;
        for (int n = 0; n < len; ++n) a = fmaf(s_e[n], af[(size_t)n * H + j], a);
        ctx[(size_t)r * H + j] = a;
    }
}

// Attention backward for one row per block.
//   d_alpha_n = dctx . att_n ; d_w_n = (d_alpha_n - sum_k alpha_k d_alpha_k) / Z ; d_s_n = d_w_n m_n ; d_e_n = s_n (d_s_n - sum_k s_k d_s_k)
//   d_att_n += alpha_n dctx ; u_nj = tanh(p_att_nj + atth_j) ; d_pre_nj = d_e_n w_j (1 - u^2) ; d_p_att_nj += d_pre_nj ;
//   d_atth_j = sum_n d_pre_nj ; d_w_row_j = sum_n d_e_n u_nj  (per-row partial of alpha_net.weight's gradient)
constexpr int kAttBwdParts = 4;   // blocks per row: 160 rows alone leave most of the 148 SMs with one 8-warp block (measured 146 us per launch)
__device__ __forceinline__ float fast_tanh_(float x) {   // ex2-based, |err| <= 2e-7 (the forward's scores use the same form)
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
__global__ void __launch_bounds__(256) attention_bwd_kernel(const float* __restrict__ atth, const float* __restrict__ p_att,
                                                            const float* __restrict__ att, const float* __restrict__ masks,
                                                            const float* __restrict__ alpha_w, const float* __restrict__ alpha,
                                                            const float* __restrict__ sm, const float* __restrict__ dctx,
                                                            float* __restrict__ d_att, float* __restrict__ d_p_att, float* __restrict__ d_atth,
                                                            float* __restrict__ d_w_rows, int len, int H, int AH, int ld_dctx) {
    extern __shared__ float s_b[];  // [len] d_e | [len] alpha | [2 * 256] partials of the two node halves
    float* s_de = s_b; float* s_al = s_b + len; float* s_part = s_b + 2 * len;
    const int r = blockIdx.x, part = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const float* af = att + (size_t)r * len * H;
    const float* dc = dctx + (size_t)r * ld_dctx;   // dctx may be a column block of a wider gradient
    // every block of the row needs d(e): the len dot products d_alpha_n = <dctx, att_n> and the softmax / mask / renormalise backward
    for (int n = wid; n < len; n += nw) {
        float a = 0.f;
        for (int j = lane; j < H; j += 32) a = fmaf(dc[j], af[(size_t)n * H + j], a);
        a = warp_sum(a);
        if (lane == 0) { s_de[n] = a; s_al[n] = alpha[(size_t)r * len + n]; }
    }
    __syncthreads();
    if (wid == 0) {
        float z = 0.f, dot = 0.f;
        for (int n = lane; n < len; n += 32) {
            z += sm[(size_t)r * len + n] * masks[(size_t)r * len + n];
            dot += s_al[n] * s_de[n];
        }
        z = warp_sum(z); dot = warp_sum(dot);
        float dot2 = 0.f;
        for (int n = lane; n < len; n += 32) {
            float ds = (s_de[n] - dot) / z * masks[(size_t)r * len + n];
            s_de[n] = ds;
            dot2 += sm[(size_t)r * len + n] * ds;
        }
        dot2 = warp_sum(dot2);
        for (int n = lane; n < len; n += 32) s_de[n] = sm[(size_t)r * len + n] * (s_de[n] - dot2);
    }
    __syncthreads();
    // d_att += alpha_n dctx: this block's quarter of the row's len x H elements
    float* da = d_att + (size_t)r * len * H;
    const int total = len * H, chunk = (total + kAttBwdParts - 1) / kAttBwdParts;
    const int hi = min(total, (part + 1) * chunk);
    for (int idx = part * chunk + threadIdx.x; idx < hi; idx += blockDim.x) {
        const int n = idx / H, j = idx - n * H;
        da[idx] += s_al[n] * dc[j];
    }
    // attention-hidden columns: this block's quarter of AH, the two thread halves take the two halves of the nodes
    const float* pa = p_att + (size_t)r * len * AH;
    float* dpa = d_p_att + (size_t)r * len * AH;
    const int jw = (AH + kAttBwdParts - 1) / kAttBwdParts;            // columns per block
    const int half = threadIdx.x >> 7, tj = threadIdx.x & 127;
    const int n0 = half ? (len + 1) / 2 : 0, n1 = half ? len : (len + 1) / 2;
    for (int j0 = part * jw; j0 < min(AH, (part + 1) * jw); j0 += 128) {
        const int j = j0 + tj;
        const bool live = j < min(AH, (part + 1) * jw);
        float dh = 0.f, dw = 0.f;
        if (live) {
            const float hj = atth[(size_t)r * AH + j], wj = alpha_w[j];
#pragma unroll 6
            for (int n = n0; n < n1; ++n) {
                const float u = fast_tanh_(pa[(size_t)n * AH + j] + hj);
                const float dpre = s_de[n] * wj * (1.f - u * u);
                dpa[(size_t)n * AH + j] += dpre;
                dh += dpre;
                dw = fmaf(s_de[n], u, dw);
            }
        }
        s_part[half * 256 + tj] = dh; s_part[half * 256 + 128 + tj] = dw;
        __syncthreads();
        if (half == 0 && live) {
            d_atth[(size_t)r * AH + j] = s_part[tj] + s_part[256 + tj];
            d_w_rows[(size_t)r * AH + j] = s_part[128 + tj] + s_part[256 + 128 + tj];
        }
        __syncthreads();
    }
}

// dlogits = dlogp - exp(logp) * rowsum(dlogp)
__global__ void __launch_bounds__(256) log_softmax_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ dlogp, size_t ld,
                                                              float* __restrict__ dlogits, int V1) {
    __shared__ float red[32];
    const float* lp = logp + (size_t)blockIdx.x * ld;
    const float* dl = dlogp + (size_t)blockIdx.x * ld;
    float s = 0.f;
    for (int j = threadIdx.x; j < V1; j += blockDim.x) s += dl[j];
    s = block_sum(s, red);
    for (int j = threadIdx.x; j < V1; j += blockDim.x) dlogits[(size_t)blockIdx.x * V1 + j] = dl[j] - expf(lp[j]) * s;
}

// sGPN pooling backward: read_out = [max over rows | mean over rows]; d_x_obj[image, ids[n], c] += ...
__global__ void __launch_bounds__(256) sgpn_pool_bwd_kernel(const subgc_subgraph_layout lay, const float* __restrict__ x_obj,
                                                            const long long* __restrict__ obj_ind, const int* __restrict__ sub_len,
                                                            const float* __restrict__ d_read, float* __restrict__ d_x_obj, int N, int L) {
    extern __shared__ int s_ids[];
    const int s = blockIdx.x;
    int image;
    const int slot = subgraph_slot(lay, s, &image);
    for (int n = threadIdx.x; n < N; n += blockDim.x) s_ids[n] = (int)obj_ind[(size_t)slot * N + n];
    __syncthreads();
    const int len = sub_len[s];
    const float* xi = x_obj + (size_t)image * N * L;
    float* dxi = d_x_obj + (size_t)image * N * L;
    const float inv = 1.f / (float)len;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float best = -INFINITY;
        int bn = -1;
        for (int n = 0; n < len; ++n) {
            float v = xi[(size_t)s_ids[n] * L + c];
            if (v > best) { best = v; bn = n; }
        }
        const float dmax = d_read[(size_t)s * 2 * L + c], dmean = d_read[(size_t)s * 2 * L + L + c] * inv;
        const bool pad_wins = (len < N) && (0.f > best);  // a zero padding row holds the max: no gradient reaches x_obj
        for (int n = 0; n < len; ++n) {
            float g = dmean + ((n == bn && !pad_wins) ? dmax : 0.f);
            atomicAdd(dxi + (size_t)s_ids[n] * L + c, g);
        }
    }
}

// GCN, training flavour of the node update: also stores the two branch outputs y0 = relu(S0/d0), y1 = relu(S1/d1)
__global__ void __launch_bounds__(256) gcn_node_train_fwd_kernel(const float* __restrict__ m_subj, const float* __restrict__ m_obj,
                                                                 const long long* __restrict__ rel_ind, const float* __restrict__ res,
                                                                 float* __restrict__ out, float* __restrict__ y0, float* __restrict__ y1, int N,
                                                                 int K, int L) {
    extern __shared__ int s_list[];
    __shared__ int s_cnt[2];
    const int bn = blockIdx.x, b = bn / N, n = bn - b * N;
    if (threadIdx.x == 0) {
        int cs = 0, co = 0;
        for (int k = 0; k < K; ++k) {
            long long s = rel_ind[((size_t)b * K + k) * 2], o = rel_ind[((size_t)b * K + k) * 2 + 1];
            if (s == n) s_list[cs++] = k;
            if (o == n) s_list[K + co++] = k;
        }
        s_cnt[0] = cs; s_cnt[1] = co;
    }
    __syncthreads();
    const int cs = s_cnt[0], co = s_cnt[1];
    const float ds = (float)cs + 1e-7f, dob = (float)co + 1e-7f;
    const float* ms = m_subj + (size_t)b * K * L;
    const float* mo = m_obj + (size_t)b * K * L;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        for (int i = 0; i < cs; ++i) a0 += ms[(size_t)s_list[i] * L + c];
        for (int i = 0; i < co; ++i) a1 += mo[(size_t)s_list[K + i] * L + c];
        float v0 = fmaxf(a0 / ds, 0.f), v1 = fmaxf(a1 / dob, 0.f);
        y0[(size_t)bn * L + c] = v0;
        y1[(size_t)bn * L + c] = v1;
        float v = (v0 + v1) * 0.5f;
        if (res) v += res[(size_t)bn * L + c];
        out[(size_t)bn * L + c] = v;
    }
}

// backward of the node update w.r.t. the per-edge messages: dM0[b,k] = 0.5 dx[b,s_k] 1[y0[b,s_k] > 0] / (cnt_s(s_k) + 1e-7), same for M1 / o_k
__global__ void __launch_bounds__(256) gcn_node_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ y0, const float* __restrict__ y1,
                                                           const long long* __restrict__ rel_ind, float* __restrict__ dm_subj,
                                                           float* __restrict__ dm_obj, int N, int K, int L) {
    __shared__ float s_d[2];
    const int bk = blockIdx.x, b = bk / K;
    const long long s = rel_ind[(size_t)bk * 2], o = rel_ind[(size_t)bk * 2 + 1];
    if (threadIdx.x == 0) {
        int cs = 0, co = 0;
        for (int k = 0; k < K; ++k) {
            cs += rel_ind[((size_t)b * K + k) * 2] == s;
            co += rel_ind[((size_t)b * K + k) * 2 + 1] == o;
        }
        s_d[0] = (float)cs + 1e-7f; s_d[1] = (float)co + 1e-7f;
    }
    __syncthreads();
    const float ds = s_d[0], dob = s_d[1];
    const size_t rs = ((size_t)b * N + s) * L, ro = ((size_t)b * N + o) * L;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        dm_subj[(size_t)bk * L + c] = (y0[rs + c] > 0.f) ? 0.5f * dx[rs + c] / ds : 0.f;
        dm_obj[(size_t)bk * L + c] = (y1[ro + c] > 0.f) ? 0.5f * dx[ro + c] / dob : 0.f;
    }
}

// backward of the edge update w.r.t. the per-node messages: dM2[b,n] = (0.5/(1+1e-7)) 1[M2[b,n] > 0] sum_{k: s_k = n} dp[b,k], same for M3 / o_k
__global__ void __launch_bounds__(256) gcn_edge_bwd_kernel(const float* __restrict__ dp, const float* __restrict__ m_subj, const float* __restrict__ m_obj,
                                                           const long long* __restrict__ rel_ind, float* __restrict__ dm_subj,
                                                           float* __restrict__ dm_obj, int N, int K, int L) {
    extern __shared__ int s_list[];
    __shared__ int s_cnt[2];
    const int bn = blockIdx.x, b = bn / N, n = bn - b * N;
    if (threadIdx.x == 0) {
        int cs = 0, co = 0;
        for (int k = 0; k < K; ++k) {
            long long s = rel_ind[((size_t)b * K + k) * 2], o = rel_ind[((size_t)b * K + k) * 2 + 1];
            if (s == n) s_list[cs++] = k;
            if (o == n) s_list[K + co++] = k;
        }
        s_cnt[0] = cs; s_cnt[1] = co;
    }
    __syncthreads();
    const int cs = s_cnt[0], co = s_cnt[1];
    const float d = 1.f + 1e-7f;
    const float* dpb = dp + (size_t)b * K * L;
    for (int c = threadIdx.x; c < L; c += blockDim.x) {
        float a0 = 0.f, a1 = 0.f;
        for (int i = 0; i < cs; ++i) a0 += dpb[(size_t)s_list[i] * L + c];
        for (int i = 0; i < co; ++i) a1 += dpb[(size_t)s_list[K + i] * L + c];
        dm_subj[(size_t)bn * L + c] = (m_subj[(size_t)bn * L + c] > 0.f) ? 0.5f * a0 / d : 0.f;
        dm_obj[(size_t)bn * L + c] = (m_obj[(size_t)bn * L + c] > 0.f) ? 0.5f * a1 / d : 0.f;
    }
}

// d_z = (p - t) * scale for BCE(sigmoid(z)); t = 1 for half 0 sub-graphs, 0 for half 1
__global__ void __launch_bounds__(256) bce_sigmoid_bwd_kernel(const subgc_subgraph_layout lay, const float* __restrict__ score, int n_sub,
                                                              float scale, float* __restrict__ dz) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sub) return;
    int half = (lay.order == 0) ? (s / lay.per_half) / lay.rows : (s / lay.per_half) % 2;
    dz[s] = (score[s] - (half == 0 ? 1.f : 0.f)) * scale;
}

static int ew_blocks(size_t n) {
    size_t b = (n + 255) / 256;
    return (int)(b > (size_t)kNumSMs * 16 ? (size_t)kNumSMs * 16 : (b ? b : 1));
}

}  // namespace subgc

using namespace subgc;
#define ST static_cast<cudaStream_t>(stream)

extern "C" size_t subgc_gemm_nt_workspace_bytes(int M, int N, int K) { return gemm_workspace_bytes(M, N, K) + 256; }

extern "C" int subgc_gemm_nt(int M, int N, int K, const float* A, int lda, const int64_t* a_gather, const float* W, int ldw, const float* bias,
                             int relu, int accumulate, float* C, int ldc, void* ws, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0, "subgc_gemm_nt: bad arguments");
    GemmProblem p;
    p.M = M; p.N = N; p.nseg = 1;
    p.seg[0] = make_seg(A, lda, W, ldw, K);
    p.seg[0].gather = reinterpret_cast<const long long*>(a_gather);
    p.epi.bias = bias; p.epi.relu = relu; p.epi.accumulate = accumulate;
    p.C = C; p.ldc = ldc;
    return launch_gemm(p, ws, ws_bytes, ST);
}

extern "C" int subgc_transpose(int rows, int cols, const float* in, int ld_in, float* out, int ld_out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(in && out && rows > 0 && cols > 0, "subgc_transpose: bad arguments");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, 256, 0, ST>>>(in, rows, cols, ld_in, out, ld_out);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_colsum(int rows, int cols, const float* in, int ld, float* out, int accumulate, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(in && out && rows >= 0 && cols > 0, "subgc_colsum: bad arguments");
    colsum_kernel<<<(cols + 31) / 32, 256, 0, ST>>>(in, rows, cols, ld, out, accumulate);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_ew(int op, size_t n, const float* a, const float* b, float* out, float scalar, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(a && out && op >= 0 && op <= 4, "subgc_ew: bad arguments");
    if (n == 0) return SUBGC_OK;
    ew_kernel<<<ew_blocks(n), 256, 0, ST>>>(op, n, a, b, out, scalar);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_dropout_mask(size_t n, float p, uint64_t seed, uint64_t offset, float* mask, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(mask && p >= 0.f && p < 1.f, "subgc_dropout_mask: bad arguments");
    if (n == 0) return SUBGC_OK;
    dropout_mask_kernel<<<ew_blocks((n + 3) / 4), 256, 0, ST>>>(n, p, seed, offset, mask);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_ss_sample(int rows, int V1, const float* prev_logp, size_t ld, const int64_t* labels, int ld_lab, float ss_prob,
                               uint64_t seed, uint64_t offset, int64_t* it, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(rows > 0 && V1 > 0 && prev_logp && labels && it && ld >= (size_t)V1 && ss_prob >= 0.f && ss_prob <= 1.f, "subgc_ss_sample: bad arguments");
    ss_sample_kernel<<<rows, 256, 0, ST>>>(prev_logp, ld, V1, reinterpret_cast<const long long*>(labels), ld_lab, ss_prob, seed, offset,
                                           reinterpret_cast<long long*>(it));
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_scatter_add_rows(int n_rows, int cols, const float* src, int ld_src, const int64_t* idx, float* dst, int ld_dst,
                                      subgc_stream_t stream) {
    SUBGC_CHECK_ARG(src && idx && dst && cols > 0, "subgc_scatter_add_rows: bad arguments");
    if (n_rows <= 0) return SUBGC_OK;
    scatter_add_rows_kernel<<<n_rows, 256, 0, ST>>>(src, reinterpret_cast<const long long*>(idx), n_rows, cols, ld_src, dst, ld_dst);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_lstm_cell_train_fwd(int S, int H, float* gates, const float* c_prev, float* h_out, float* c_out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(gates && c_prev && h_out && c_out && S > 0 && H > 0, "subgc_lstm_cell_train_fwd: bad arguments");
    lstm_cell_train_fwd_kernel<<<(S * H + 255) / 256, 256, 0, ST>>>(gates, c_prev, h_out, c_out, S, H);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_lstm_cell_bwd(int S, int H, const float* act, const float* c_prev, const float* c_new, const float* dh, const float* dc_in,
                                   float* dgates, float* dc_prev, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(act && c_prev && c_new && dh && dgates && dc_prev && S > 0 && H > 0, "subgc_lstm_cell_bwd: bad arguments");
    lstm_cell_bwd_kernel<<<(S * H + 255) / 256, 256, 0, ST>>>(act, c_prev, c_new, dh, dc_in, dgates, dc_prev, S, H);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_attention_train_fwd(int S, int len, int H, int AH, const float* atth, const float* p_att, const float* att,
                                         const float* masks, const float* alpha_w, const float* alpha_b, float* ctx, float* alpha, float* sm,
                                         subgc_stream_t stream) {
    SUBGC_CHECK_ARG(atth && p_att && att && masks && alpha_w && alpha_b && ctx && alpha && sm && S > 0 && len > 0 && len <= 64,
                    "subgc_attention_train_fwd: bad arguments");
    attention_train_fwd_kernel<<<S, 256, (size_t)(2 * AH + len) * 4, ST>>>(atth, p_att, att, masks, alpha_w, alpha_b, ctx, alpha, sm, len, H, AH);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_attention_bwd(int S, int len, int H, int AH, const float* atth, const float* p_att, const float* att, const float* masks,
                                   const float* alpha_w, const float* alpha, const float* sm, const float* dctx, float* d_att, float* d_p_att,
                                   float* d_atth, float* d_w_rows, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(atth && p_att && att && masks && alpha_w && alpha && sm && dctx && d_att && d_p_att && d_atth && d_w_rows && S > 0 &&
                        len > 0 && len <= 64,
                    "subgc_attention_bwd: bad arguments");
    attention_bwd_kernel<<<dim3(S, kAttBwdParts), 256, (size_t)(2 * len + 512) * 4, ST>>>(atth, p_att, att, masks, alpha_w, alpha, sm, dctx, d_att,
                                                                                           d_p_att, d_atth, d_w_rows, len, H, AH, H);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_log_softmax_fwd(int rows, int V1, const float* logits, float* logp, size_t ld_out, subgc_stream_t stream);

extern "C" int subgc_log_softmax_bwd(int rows, int V1, const float* logp, const float* dlogp, size_t ld, float* dlogits, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(logp && dlogp && dlogits && rows > 0 && V1 > 0, "subgc_log_softmax_bwd: bad arguments");
    log_softmax_bwd_kernel<<<rows, 256, 0, ST>>>(logp, dlogp, ld, dlogits, V1);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_sgpn_pool_bwd(const subgc_dims* d, const subgc_subgraph_layout* lay, const float* x_obj, const int64_t* gpn_obj_ind,
                                   const int32_t* sub_len, const float* d_read_out, float* d_x_obj, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && lay && x_obj && gpn_obj_ind && sub_len && d_read_out && d_x_obj, "subgc_sgpn_pool_bwd: null argument");
    const int n_sub = subgraph_count(*lay);
    sgpn_pool_bwd_kernel<<<n_sub, 256, d->obj_num * sizeof(int), ST>>>(*lay, x_obj, reinterpret_cast<const long long*>(gpn_obj_ind), sub_len,
                                                                      d_read_out, d_x_obj, d->obj_num, d->gcn);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_gcn_node_train_fwd(int B, int N, int K, int L, const float* m_subj, const float* m_obj, const int64_t* rel_ind,
                                        const float* res, float* out, float* y0, float* y1, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(m_subj && m_obj && rel_ind && out && y0 && y1 && B > 0, "subgc_gcn_node_train_fwd: bad arguments");
    gcn_node_train_fwd_kernel<<<B * N, 256, 2 * K * sizeof(int), ST>>>(m_subj, m_obj, reinterpret_cast<const long long*>(rel_ind), res, out, y0, y1,
                                                                      N, K, L);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_gcn_node_bwd(int B, int N, int K, int L, const float* dx, const float* y0, const float* y1, const int64_t* rel_ind,
                                  float* dm_subj, float* dm_obj, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(dx && y0 && y1 && rel_ind && dm_subj && dm_obj && B > 0, "subgc_gcn_node_bwd: bad arguments");
    gcn_node_bwd_kernel<<<B * K, 256, 0, ST>>>(dx, y0, y1, reinterpret_cast<const long long*>(rel_ind), dm_subj, dm_obj, N, K, L);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_gcn_edge_fwd(int B, int N, int K, int L, const float* m_subj, const float* m_obj, const int64_t* rel_ind, const float* res,
                                  float* out, subgc_stream_t stream);

extern "C" int subgc_gcn_edge_bwd(int B, int N, int K, int L, const float* dp, const float* m_subj, const float* m_obj, const int64_t* rel_ind,
                                  float* dm_subj, float* dm_obj, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(dp && m_subj && m_obj && rel_ind && dm_subj && dm_obj && B > 0, "subgc_gcn_edge_bwd: bad arguments");
    gcn_edge_bwd_kernel<<<B * N, 256, 2 * K * sizeof(int), ST>>>(dp, m_subj, m_obj, reinterpret_cast<const long long*>(rel_ind), dm_subj, dm_obj, N, K,
                                                                L);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_bce_sigmoid_bwd(const subgc_subgraph_layout* lay, const float* score, float scale, float* dz, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(lay && score && dz, "subgc_bce_sigmoid_bwd: null argument");
    const int n_sub = subgraph_count(*lay);
    bce_sigmoid_bwd_kernel<<<(n_sub + 255) / 256, 256, 0, ST>>>(*lay, score, n_sub, scale, dz);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

namespace subgc {
// out[r, :] = act(src[idx[r], :])
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, int n_rows, int cols,
                                                          int ld_src, float* __restrict__ out, int relu) {
    const int r = blockIdx.x;
    if (r >= n_rows) return;
    const float* s = src + (size_t)idx[r] * ld_src;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        float v = s[c];
        out[(size_t)r * cols + c] = relu ? fmaxf(v, 0.f) : v;
    }
}
__global__ void __launch_bounds__(256) ew2_kernel(int op, size_t n, const float* __restrict__ a, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = a[i];
        out[i] = op == 0 ? fmaxf(v, 0.f) : sigmoidf_(v);
    }
}
}  // namespace subgc

extern "C" int subgc_gather_rows(int n_rows, int cols, const float* src, int ld_src, const int64_t* idx, float* out, int relu,
                                 subgc_stream_t stream) {
    SUBGC_CHECK_ARG(src && idx && out && cols > 0, "subgc_gather_rows: bad arguments");
    if (n_rows <= 0) return SUBGC_OK;
    gather_rows_kernel<<<n_rows, 256, 0, ST>>>(src, reinterpret_cast<const long long*>(idx), n_rows, cols, ld_src, out, relu);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

/* op 0: relu, 1: sigmoid */
extern "C" int subgc_unary(int op, size_t n, const float* a, float* out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(a && out && (op == 0 || op == 1), "subgc_unary: bad arguments");
    if (n == 0) return SUBGC_OK;
    ew2_kernel<<<ew_blocks(n), 256, 0, ST>>>(op, n, a, out);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

// =====================================================================================================================================
// Stage-level training entries: the teacher-forced decoder of AttModel._forward (models/AttModel.py:150-177) and its
// backward as ONE C call each.  Same arithmetic as the building blocks above, sequenced here instead of from Python, with everything
// that does not depend on the recurrence batched over the executed steps:
//   forward : per step ONE multi-segment contraction per LSTM ([h_lang | fc | x_t] . W_ih^T + h_att . W_hh^T, no concat buffer), the
//             logit contraction + log-softmax once over all T * R rows;
//   backward: d(logits) / d(h) of all steps in one contraction, per step only the recurrence (cells, attention, dX through transposed
//             weight blocks prepared once), and every weight / bias gradient as ONE contraction over K = T * R (dW = dY_all^T . X_all).
// =====================================================================================================================================
namespace subgc {

// x_t = relu(E[it]) * mask  (AttModel.py:332 embed = Embedding + ReLU + Dropout); one block per row
__global__ void __launch_bounds__(256) dec_embed_kernel(const float* __restrict__ embed, const long long* __restrict__ it, const float* __restrict__ mask,
                                                        float* __restrict__ xt, int E) {
    const int r = blockIdx.x;
    const float* e = embed + (size_t)it[r] * E;
    for (int j = threadIdx.x; j < E; j += blockDim.x) {
        float v = fmaxf(e[j], 0.f);
        if (mask) v *= mask[(size_t)r * E + j];
        xt[(size_t)r * E + j] = v;
    }
}

// out[r, j] = a[r * lda + j] (+ b[r * ldb + j]) (+ c[r * ldc + j]),  j < n  (column blocks of wider gradients)
__global__ void __launch_bounds__(256) add_cols_kernel(int R, int n, const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                                                       const float* __restrict__ c, int ldc, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)R * n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / n, j = i - r * n;
        float v = a[r * lda + j];
        if (b) v += b[r * ldb + j];
        if (c) v += c[r * ldc + j];
        out[i] = v;
    }
}

// tail of one backward step.  d_xa [R, ld] = dgates_att . [W_ih | W_hh]: columns [0,H) d(h_lang(t-1)), [H,2H) d(fc), [2H,2H+E) d(x_t),
// [2H+E, 3H+E) d(h_att(t-1)) (read in place by the next step).  d_fc += ; dh_lang_n = dh_lang_prev + ; embedding rows += d(x_t) through
// dropout mask and ReLU (x_t > 0 <=> relu'(E[it]) = 1 and the element was kept)
__global__ void __launch_bounds__(256) dec_bwd_tail_kernel(int R, int H, int E, const float* __restrict__ d_xa, int ld, const float* __restrict__ dh_lang_prev,
                                                           int ld_prev, const float* __restrict__ xt, const float* __restrict__ mask,
                                                           const long long* __restrict__ it, float* __restrict__ d_fc, float* __restrict__ dh_lang_n,
                                                           float* __restrict__ g_embed) {
    const int r = blockIdx.x;
    const float* row = d_xa + (size_t)r * ld;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        d_fc[(size_t)r * H + j] += row[H + j];
        dh_lang_n[(size_t)r * H + j] = dh_lang_prev[(size_t)r * ld_prev + j] + row[j];
    }
    float* ge = g_embed + (size_t)it[r] * E;
    for (int j = threadIdx.x; j < E; j += blockDim.x) {
        if (xt[(size_t)r * E + j] > 0.f) {
            float g = row[2 * H + j];
            if (mask) g *= mask[(size_t)r * E + j];
            atomicAdd(ge + j, g);
        }
    }
}

// out[i] = sum over t of in[t * n + i]
__global__ void __launch_bounds__(256) sum_steps_kernel(int T, size_t n, const float* __restrict__ in, float* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int t = 0; t < T; ++t) a += in[(size_t)t * n + i];
        out[i] = a;
    }
}

// Fused log-softmax + LanguageModelCriterion term of one row (misc/utils.py:115-124: -logp[target] * mask): the [rows, V1] log-probs are
// never written.  lse is kept for the backward; nll[r] = (lse - logit[target]) * mask[r].
__global__ void __launch_bounds__(256) nll_fwd_kernel(const float* __restrict__ logits, int V1, const long long* __restrict__ target,
                                                      const float* __restrict__ mask, float* __restrict__ lse, float* __restrict__ nll) {
    __shared__ float red[32];
    const size_t r = blockIdx.x;
    const float* x = logits + r * V1;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < V1; j += blockDim.x) m = fmaxf(m, x[j]);
    m = block_max(m, red);
    float s = 0.f;
    for (int j = threadIdx.x; j < V1; j += blockDim.x) s += expf(x[j] - m);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        const float lz = logf(s);
        lse[r] = m + lz;
        nll[r] = -((x[target[r]] - m) - lz) * mask[r];   // same rounding sequence as log_softmax_kernel: (x - max) - log(sum)
    }
}
// d(logits)[r, j] = coef[r] * (softmax[r, j] - [j == target[r]]),  coef[r] = mask[r] * d(loss) / sum(mask)
__global__ void __launch_bounds__(256) nll_bwd_kernel(const float* __restrict__ logits, int V1, const long long* __restrict__ target,
                                                      const float* __restrict__ lse, const float* __restrict__ coef, float* __restrict__ dlogits) {
    const size_t r = blockIdx.x;
    const float* x = logits + r * V1;
    float* dx = dlogits + r * V1;
    const float c = coef[r], l = lse[r];
    const int tg = (int)target[r];
    for (int j = threadIdx.x; j < V1; j += blockDim.x) dx[j] = c * (expf(x[j] - l) - (j == tg ? 1.f : 0.f));
}

struct StageWs {   // bump allocator over the caller's workspace; everything 256-byte aligned
    Workspace ws;
    StageWs(void* p, size_t n) : ws(p, n) {}
    float* f(size_t n) { return ws.take<float>(n); }
};

static size_t dec_gemm_ws_bytes(const subgc_dims* d, int R, int T) {
    const int H = d->rnn, E = d->enc, V1 = d->vocab1, AH = d->att_hid, TR = T * R;
    size_t m = 0;
    auto up = [&](int M, int N, int K) { const size_t b = gemm_workspace_bytes(M, N, K); if (b > m) m = b; };
    up(R, 4 * H, 3 * H + E); up(R, AH, H); up(TR, V1, H);                     // forward
    up(TR, H, V1); up(V1, H, TR); up(R, 3 * H, 4 * H); up(R, 3 * H + E, 4 * H); up(R, H, AH);
    up(4 * H, H, TR); up(4 * H, E, TR); up(4 * H, H, R); up(AH, H, TR);       // weight gradients
    return align_up(m, 256) + 1024;
}

static int gemm1(int M, int N, int K, const float* A, int lda, const float* W, int ldw, const float* bias, int accumulate, float* C, int ldc, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
    GemmProblem p;
    p.M = M; p.N = N; p.nseg = 1;
    p.seg[0] = make_seg(A, lda, W, ldw, K);
    p.epi.bias = bias; p.epi.accumulate = accumulate;
    p.C = C; p.ldc = ldc;
    return launch_gemm(p, ws, ws_bytes, st);
}
static void tr(const float* in, int rows, int cols, int ld_in, float* out, int ld_out, cudaStream_t st) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, 256, 0, st>>>(in, rows, cols, ld_in, out, ld_out);
}

}  // namespace subgc

extern "C" size_t subgc_decoder_train_workspace_bytes(const subgc_dims* d, int R, int len, int T) {
    if (!d || R < 1 || T < 1) return 0;
    const size_t H = d->rnn, E = d->enc, V1 = d->vocab1, AH = d->att_hid, TR = (size_t)T * R;
    size_t b = dec_gemm_ws_bytes(d, R, T);
    auto add = [&](size_t floats) { b += align_up(floats * 4, 256); };
    // forward: logits of all steps.  backward (the larger one): transposed weights, d(logits) + its transpose, per-step gate gradients ...
    add(H * V1); add(3 * H * 4 * H); add((3 * H + E) * 4 * H); add(H * AH);     // W^T blocks
    add(TR * V1); add(TR * V1);                                                 // dlogits_all, its transpose (forward: logits_all)
    add(TR * H); add(TR * 4 * H); add(TR * 4 * H); add(TR * AH); add(TR * AH);  // d_hd_all, dg2_all, dg1_all, d_atth_all, d_wrows_all
    add(4 * H * TR); add((H > E ? H : E) * TR);                                 // dY^T, X^T
    add((size_t)R * 3 * H); add((size_t)R * (3 * H + E)); add((size_t)R * H * 8); add((size_t)R * 4 * H);
    (void)len;
    return b + 4096;
}

/* Teacher-forced decoder, train mode (models/AttModel.py:150-177 with get_logprobs_state :328-341, TopDownCore :400-431, Attention :445-471).
 * tokens [T, R]: the token fed at step t (labels[:, t]).  Saved activations land in `b` (see include/subgc_b200.h). */
extern "C" int subgc_decoder_train_forward(const subgc_dims* d, const subgc_weights* w, int R, int len, int T, int T_total,
                                           const subgc_decoder_train_bufs* b, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && b && ws_ && R > 0 && T > 0 && T <= T_total && len > 0 && len <= 64, "subgc_decoder_train_forward: bad arguments");
    SUBGC_CHECK_ARG(ws_bytes >= subgc_decoder_train_workspace_bytes(d, R, len, T), "subgc_decoder_train_forward: workspace too small");
    cudaStream_t st = ST;
    const int H = d->rnn, E = d->enc, V1 = d->vocab1, AH = d->att_hid;
    const size_t RH = (size_t)R * H;
    StageWs sw(ws_, ws_bytes);
    const size_t gws_bytes = dec_gemm_ws_bytes(d, R, T);
    void* gws = sw.ws.take<char>(gws_bytes);
    float* logits = b->logits ? b->logits : sw.f((size_t)T * R * V1);
    SUBGC_CHECK_ARG(sw.ws.ok(), "subgc_decoder_train_forward: workspace too small");
    SUBGC_CHECK_ARG(b->outputs || (b->nll && b->lse && b->targets && b->tmask && b->logits),
                    "subgc_decoder_train_forward: either outputs or the fused-loss buffers (logits, targets, tmask, lse, nll) are needed");
    // init_hidden (AttModel.py:343-346): slot 0 of the state histories
    SUBGC_CUDA(cudaMemsetAsync(b->h_att, 0, RH * 4, st)); SUBGC_CUDA(cudaMemsetAsync(b->c_att, 0, RH * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(b->h_lang, 0, RH * 4, st)); SUBGC_CUDA(cudaMemsetAsync(b->c_lang, 0, RH * 4, st));
    for (int t = 0; t < T; ++t) {
        const float* h_att_p = b->h_att + (size_t)t * RH;   float* h_att_n = b->h_att + (size_t)(t + 1) * RH;
        const float* h_lang_p = b->h_lang + (size_t)t * RH; float* h_lang_n = b->h_lang + (size_t)(t + 1) * RH;
        float* xt = b->xt + (size_t)t * R * E;
        dec_embed_kernel<<<R, 256, 0, st>>>(w->embed, reinterpret_cast<const long long*>(b->tokens) + (size_t)t * R,
                                            b->m_x ? b->m_x + (size_t)t * R * E : nullptr, xt, E);
        SUBGC_LAUNCH_CHECK();
        // attention LSTM (AttModel.py:410-413): gates = W_ih [h_lang(t-1) | fc | x_t] + b_ih + W_hh h_att(t-1) + b_hh
        float* act1 = b->act1 + (size_t)t * R * 4 * H;
        {
            GemmProblem p;
            p.M = R; p.N = 4 * H; p.nseg = 4;
            p.seg[0] = make_seg(h_lang_p, H, w->att_w_ih, E + 2 * H, H);
            p.seg[1] = make_seg(b->fc, H, w->att_w_ih + H, E + 2 * H, H);
            p.seg[2] = make_seg(xt, E, w->att_w_ih + 2 * H, E + 2 * H, E);
            p.seg[3] = make_seg(h_att_p, H, w->att_w_hh, H, H);
            p.epi.bias = w->att_b_ih; p.epi.bias2 = w->att_b_hh;
            p.C = act1; p.ldc = 4 * H;
            SUBGC_TRY(launch_gemm(p, gws, gws_bytes, st));
        }
        SUBGC_TRY(subgc_lstm_cell_train_fwd(R, H, act1, b->c_att + (size_t)t * RH, h_att_n, b->c_att + (size_t)(t + 1) * RH, stream));
        // attention (AttModel.py:445-471)
        float* atth = b->atth + (size_t)t * R * AH;
        SUBGC_TRY(gemm1(R, AH, H, h_att_n, H, w->h2att.w, H, w->h2att.b, 0, atth, AH, gws, gws_bytes, st));
        float* ctx = b->ctx + (size_t)t * RH;
        SUBGC_TRY(subgc_attention_train_fwd(R, len, H, AH, atth, b->p_att, b->att, b->masks, w->alpha_net.w, w->alpha_net.b, ctx,
                                            b->alpha + (size_t)t * R * len, b->sm + (size_t)t * R * len, stream));
        // language LSTM (AttModel.py:423-426): gates = W_ih [ctx | h_att(t)] + b_ih + W_hh h_lang(t-1) + b_hh
        float* act2 = b->act2 + (size_t)t * R * 4 * H;
        {
            GemmProblem p;
            p.M = R; p.N = 4 * H; p.nseg = 3;
            p.seg[0] = make_seg(ctx, H, w->lang_w_ih, 2 * H, H);
            p.seg[1] = make_seg(h_att_n, H, w->lang_w_ih + H, 2 * H, H);
            p.seg[2] = make_seg(h_lang_p, H, w->lang_w_hh, H, H);
            p.epi.bias = w->lang_b_ih; p.epi.bias2 = w->lang_b_hh;
            p.C = act2; p.ldc = 4 * H;
            SUBGC_TRY(launch_gemm(p, gws, gws_bytes, st));
        }
        SUBGC_TRY(subgc_lstm_cell_train_fwd(R, H, act2, b->c_lang + (size_t)t * RH, h_lang_n, b->c_lang + (size_t)(t + 1) * RH, stream));
    }
    // output dropout, logit, log-softmax (AttModel.py:339-340,428-429) over all steps at once
    const float* hd = b->h_lang + RH;
    if (b->m_h) {
        SUBGC_TRY(subgc_ew(EW_MUL, (size_t)T * RH, b->h_lang + RH, b->m_h, b->hd, 0.f, stream));
        hd = b->hd;
    }
    SUBGC_TRY(gemm1(T * R, V1, H, hd, H, w->logit.w, H, w->logit.b, 0, logits, V1, gws, gws_bytes, st));
    if (b->outputs) {
        for (int t = 0; t < T; ++t)
            SUBGC_TRY(subgc_log_softmax_fwd(R, V1, logits + (size_t)t * R * V1, b->outputs + (size_t)t * V1, (size_t)T_total * V1, stream));
    }
    if (b->nll) {   // LossWrapper path: log-softmax + criterion fused, the log-probs are never materialised
        nll_fwd_kernel<<<T * R, 256, 0, st>>>(logits, V1, reinterpret_cast<const long long*>(b->targets), b->tmask, b->lse, b->nll);
        SUBGC_LAUNCH_CHECK();
    }
    return SUBGC_OK;
}

/* Backward of the above given d(outputs) [R, T_total, V1].  Gradients are ACCUMULATED into `g` (the caller zero-fills); d_fc [R, H],
 * d_att [R, len, H] and d_p_att [R, len, AH] are overwritten with the gradients of the decoder's inputs. */
extern "C" int subgc_decoder_train_backward(const subgc_dims* d, const subgc_weights* w, int R, int len, int T, int T_total,
                                            const subgc_decoder_train_bufs* b, const float* d_outputs, const subgc_decoder_grads* g, float* d_fc,
                                            float* d_att, float* d_p_att, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && b && g && d_fc && d_att && d_p_att && ws_ && R > 0 && T > 0 && T <= T_total && len > 0 && len <= 64,
                    "subgc_decoder_train_backward: bad arguments");
    SUBGC_CHECK_ARG(d_outputs || (b->coef && b->lse && b->targets && b->logits),
                    "subgc_decoder_train_backward: either d_outputs or the fused-loss buffers (coef, lse, targets, logits) are needed");
    SUBGC_CHECK_ARG(ws_bytes >= subgc_decoder_train_workspace_bytes(d, R, len, T), "subgc_decoder_train_backward: workspace too small");
    cudaStream_t st = ST;
    const int H = d->rnn, E = d->enc, V1 = d->vocab1, AH = d->att_hid, TR = T * R;
    const size_t RH = (size_t)R * H;
    const int XA = 3 * H + E;   // columns of d_xa: [h_lang | fc | x_t | h_att]
    StageWs sw(ws_, ws_bytes);
    const size_t gws_bytes = dec_gemm_ws_bytes(d, R, T);
    void* gws = sw.ws.take<char>(gws_bytes);
    float* wt_logit = sw.f((size_t)H * V1);            // [H, V1]
    float* wt_lang = sw.f((size_t)3 * H * 4 * H);      // [2H + H, 4H] = [W_ih^T ; W_hh^T]
    float* wt_att = sw.f((size_t)XA * 4 * H);          // [(2H + E) + H, 4H] = [W_ih^T ; W_hh^T]
    float* wt_h2att = sw.f((size_t)H * AH);            // [H, AH]
    float* dlogits = sw.f((size_t)TR * V1);
    float* big_t = sw.f((size_t)TR * V1);              // transposes of the dY operands (largest: [V1, TR])
    float* d_hd = sw.f((size_t)TR * H);
    float* dg2 = sw.f((size_t)TR * 4 * H);
    float* dg1 = sw.f((size_t)TR * 4 * H);
    float* d_atth = sw.f((size_t)TR * AH);
    float* d_wrows = sw.f((size_t)TR * AH);
    float* dyt = sw.f((size_t)4 * H * TR);
    float* xT = sw.f((size_t)(H > E ? H : E) * TR);
    float* d_xl = sw.f((size_t)R * 3 * H);             // [d_ctx | d_h_att | d_h_lang(t-1)]
    float* d_xa = sw.f((size_t)R * XA);
    float* tmp = sw.f((size_t)R * H * 8);
    float* dg1_sum = sw.f((size_t)R * 4 * H);
    SUBGC_CHECK_ARG(sw.ws.ok(), "subgc_decoder_train_backward: workspace too small");
    float* dh = tmp;                     // d(h_lang(t)) / d(h_att(t)) assembled for the cell backward
    float* dh_lang_n = tmp + RH;         // from step t+1
    float* dc_lang[2] = {tmp + 2 * RH, tmp + 3 * RH};
    float* dc_att[2] = {tmp + 4 * RH, tmp + 5 * RH};
    const int eb = ew_blocks(RH);

    // transposed weight blocks, once per backward pass (dX = dY . W needs W^T K-major)
    tr(w->logit.w, V1, H, H, wt_logit, V1, st);
    tr(w->lang_w_ih, 4 * H, 2 * H, 2 * H, wt_lang, 4 * H, st);
    tr(w->lang_w_hh, 4 * H, H, H, wt_lang + (size_t)2 * H * 4 * H, 4 * H, st);
    tr(w->att_w_ih, 4 * H, 2 * H + E, 2 * H + E, wt_att, 4 * H, st);
    tr(w->att_w_hh, 4 * H, H, H, wt_att + (size_t)(2 * H + E) * 4 * H, 4 * H, st);
    tr(w->h2att.w, AH, H, H, wt_h2att, AH, st);
    SUBGC_LAUNCH_CHECK();
    SUBGC_CUDA(cudaMemsetAsync(d_fc, 0, RH * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(d_att, 0, (size_t)R * len * H * 4, st));
    SUBGC_CUDA(cudaMemsetAsync(d_p_att, 0, (size_t)R * len * AH * 4, st));

    // ---- logit + log-softmax of all steps: d(logits), logit.weight / bias, d(h_lang after dropout)
    if (d_outputs) {
        for (int t = 0; t < T; ++t)
            SUBGC_TRY(subgc_log_softmax_bwd(R, V1, b->outputs + (size_t)t * V1, d_outputs + (size_t)t * V1, (size_t)T_total * V1,
                                            dlogits + (size_t)t * R * V1, stream));
    } else {
        nll_bwd_kernel<<<TR, 256, 0, st>>>(b->logits, V1, reinterpret_cast<const long long*>(b->targets), b->lse, b->coef, dlogits);
        SUBGC_LAUNCH_CHECK();
    }
    const float* hd = b->m_h ? b->hd : b->h_lang + RH;
    tr(dlogits, TR, V1, V1, big_t, TR, st);
    tr(hd, TR, H, H, xT, TR, st);
    SUBGC_LAUNCH_CHECK();
    SUBGC_TRY(gemm1(V1, H, TR, big_t, TR, xT, TR, nullptr, 1, g->logit_w, H, gws, gws_bytes, st));
    SUBGC_TRY(subgc_colsum(TR, V1, dlogits, V1, g->logit_b, 1, stream));
    SUBGC_TRY(gemm1(TR, H, V1, dlogits, V1, wt_logit, V1, nullptr, 0, d_hd, H, gws, gws_bytes, st));
    if (b->m_h) SUBGC_TRY(subgc_ew(EW_MUL, (size_t)TR * H, d_hd, b->m_h, d_hd, 0.f, stream));

    // ---- the recurrence, last step first
    for (int t = T - 1; t >= 0; --t) {
        const bool last = (t == T - 1);
        const int cur = t & 1, nxt = cur ^ 1;   // dc ping-pong: [nxt] came from step t+1, [cur] is written for step t-1
        // language LSTM cell: d(h_lang(t)) = d(hd(t)) + what step t+1 sent back
        const float* dh_l = d_hd + (size_t)t * RH;
        if (!last) {
            add_cols_kernel<<<eb, 256, 0, st>>>(R, H, d_hd + (size_t)t * RH, H, dh_lang_n, H, nullptr, 0, dh);
            SUBGC_LAUNCH_CHECK();
            dh_l = dh;
        }
        float* dg2_t = dg2 + (size_t)t * R * 4 * H;
        SUBGC_TRY(subgc_lstm_cell_bwd(R, H, b->act2 + (size_t)t * R * 4 * H, b->c_lang + (size_t)t * RH, b->c_lang + (size_t)(t + 1) * RH, dh_l,
                                      last ? nullptr : dc_lang[nxt], dg2_t, dc_lang[cur], stream));
        // d[ctx | h_att(t) | h_lang(t-1)] = dgates . [W_ih | W_hh]
        SUBGC_TRY(gemm1(R, 3 * H, 4 * H, dg2_t, 4 * H, wt_lang, 4 * H, nullptr, 0, d_xl, 3 * H, gws, gws_bytes, st));
        // attention backward (d_att / d_p_att accumulate over the steps)
        float* d_atth_t = d_atth + (size_t)t * R * AH;
        attention_bwd_kernel<<<dim3(R, kAttBwdParts), 256, (size_t)(2 * len + 512) * 4, st>>>(b->atth + (size_t)t * R * AH, b->p_att, b->att, b->masks, w->alpha_net.w,
                                                                         b->alpha + (size_t)t * R * len, b->sm + (size_t)t * R * len, d_xl, d_att, d_p_att,
                                                                         d_atth_t, d_wrows + (size_t)t * R * AH, len, H, AH, 3 * H);
        SUBGC_LAUNCH_CHECK();
        // d(h_att(t)) = language-LSTM input part + h2att part + what step t+1's attention LSTM sent back through W_hh
        add_cols_kernel<<<eb, 256, 0, st>>>(R, H, d_xl + H, 3 * H, last ? nullptr : d_xa + (2 * H + E), XA, nullptr, 0, dh);
        SUBGC_LAUNCH_CHECK();
        SUBGC_TRY(gemm1(R, H, AH, d_atth_t, AH, wt_h2att, AH, nullptr, 1, dh, H, gws, gws_bytes, st));
        float* dg1_t = dg1 + (size_t)t * R * 4 * H;
        SUBGC_TRY(subgc_lstm_cell_bwd(R, H, b->act1 + (size_t)t * R * 4 * H, b->c_att + (size_t)t * RH, b->c_att + (size_t)(t + 1) * RH, dh,
                                      last ? nullptr : dc_att[nxt], dg1_t, dc_att[cur], stream));
        // d[h_lang(t-1) | fc | x_t | h_att(t-1)] = dgates . [W_ih | W_hh]
        SUBGC_TRY(gemm1(R, XA, 4 * H, dg1_t, 4 * H, wt_att, 4 * H, nullptr, 0, d_xa, XA, gws, gws_bytes, st));
        dec_bwd_tail_kernel<<<R, 256, 0, st>>>(R, H, E, d_xa, XA, d_xl + 2 * H, 3 * H, b->xt + (size_t)t * R * E,
                                               b->m_x ? b->m_x + (size_t)t * R * E : nullptr,
                                               reinterpret_cast<const long long*>(b->tokens) + (size_t)t * R, d_fc, dh_lang_n, g->embed);
        SUBGC_LAUNCH_CHECK();
    }

    // ---- weight / bias gradients: one contraction over K = T * R per weight block
    auto dW = [&](const float* dy_t, int Mw, const float* x, int Kx, int ldx, int rowsK, float* G, int ldg) -> int {   // G[Mw, Kx] += dy^T . x
        tr(x, rowsK, Kx, ldx, xT, rowsK, st);
        SUBGC_LAUNCH_CHECK();
        return gemm1(Mw, Kx, rowsK, dy_t, rowsK, xT, rowsK, nullptr, 1, G, ldg, gws, gws_bytes, st);
    };
    // language LSTM
    tr(dg2, TR, 4 * H, 4 * H, dyt, TR, st);
    SUBGC_LAUNCH_CHECK();
    SUBGC_TRY(dW(dyt, 4 * H, b->ctx, H, H, TR, g->lang_w_ih, 2 * H));
    SUBGC_TRY(dW(dyt, 4 * H, b->h_att + RH, H, H, TR, g->lang_w_ih + H, 2 * H));
    SUBGC_TRY(dW(dyt, 4 * H, b->h_lang, H, H, TR, g->lang_w_hh, H));
    SUBGC_TRY(subgc_colsum(TR, 4 * H, dg2, 4 * H, g->lang_b_ih, 1, stream));
    SUBGC_TRY(subgc_colsum(TR, 4 * H, dg2, 4 * H, g->lang_b_hh, 1, stream));
    // attention LSTM
    tr(dg1, TR, 4 * H, 4 * H, dyt, TR, st);
    SUBGC_LAUNCH_CHECK();
    SUBGC_TRY(dW(dyt, 4 * H, b->h_lang, H, H, TR, g->att_w_ih, E + 2 * H));
    SUBGC_TRY(dW(dyt, 4 * H, b->xt, E, E, TR, g->att_w_ih + 2 * H, E + 2 * H));
    SUBGC_TRY(dW(dyt, 4 * H, b->h_att, H, H, TR, g->att_w_hh, H));
    SUBGC_TRY(subgc_colsum(TR, 4 * H, dg1, 4 * H, g->att_b_ih, 1, stream));
    SUBGC_TRY(subgc_colsum(TR, 4 * H, dg1, 4 * H, g->att_b_hh, 1, stream));
    // fc is the same at every step: contract the step sum of the gate gradients (K = R)
    sum_steps_kernel<<<ew_blocks((size_t)R * 4 * H), 256, 0, st>>>(T, (size_t)R * 4 * H, dg1, dg1_sum);
    tr(dg1_sum, R, 4 * H, 4 * H, dyt, R, st);
    SUBGC_LAUNCH_CHECK();
    SUBGC_TRY(dW(dyt, 4 * H, b->fc, H, H, R, g->att_w_ih + H, E + 2 * H));
    // h2att, alpha_net
    tr(d_atth, TR, AH, AH, dyt, TR, st);
    SUBGC_LAUNCH_CHECK();
    SUBGC_TRY(dW(dyt, AH, b->h_att + RH, H, H, TR, g->h2att_w, H));
    SUBGC_TRY(subgc_colsum(TR, AH, d_atth, AH, g->h2att_b, 1, stream));
    SUBGC_TRY(subgc_colsum(TR, AH, d_wrows, AH, g->alpha_w, 1, stream));
    return SUBGC_OK;
}

// =====================================================================================================================================
// Stage-level backward entries of the front-end: feature preparation, sGPN, GCN + fusion.  Each sequences the same
// building blocks as sub-gc_b200/subgc/train.py's reference orchestration (kept for the CPU emulation tests) in one C call.
// =====================================================================================================================================
namespace subgc {

struct BwdCtx {   // scratch shared by the dW / dX helpers of one call
    float* dyT; float* xT; float* wT; void* gws; size_t gws_bytes; cudaStream_t st;
    // G[Mw, Kx] (ld ldg) += dy^T . x        dy [rows, Mw], x [rows, Kx]
    int dW(float* G, int ldg, const float* dy, int Mw, const float* x, int Kx, int rows) const {
        tr(dy, rows, Mw, Mw, dyT, rows, st);
        tr(x, rows, Kx, Kx, xT, rows, st);
        return gemm1(Mw, Kx, rows, dyT, rows, xT, rows, nullptr, 1, G, ldg, gws, gws_bytes, st);
    }
    // out[rows, Kx] = dy . W                 W [Mw, Kx] (nn.Linear layout)
    int dX(float* out, const float* dy, int rows, int Mw, const float* W, int Kx) const {
        tr(W, Mw, Kx, Kx, wT, Mw, st);
        return gemm1(rows, Kx, Mw, dy, Mw, wT, Mw, nullptr, 0, out, Kx, gws, gws_bytes, st);
    }
};
static size_t bwd_scratch_floats(size_t rows_max, size_t dim_max) { return 2 * rows_max * dim_max + dim_max * dim_max; }
static size_t bwd_gemm_ws(size_t rows_max, int dim_max) {
    size_t a = gemm_workspace_bytes(dim_max, dim_max, (int)rows_max), b = gemm_workspace_bytes((int)rows_max, dim_max, dim_max);
    return align_up(a > b ? a : b, 256) + 1024;
}
static bool bwd_ctx(BwdCtx& c, Workspace& ws, size_t rows_max, int dim_max, cudaStream_t st) {
    c.gws_bytes = bwd_gemm_ws(rows_max, dim_max);
    c.gws = ws.take<char>(c.gws_bytes);
    c.dyT = ws.take<float>(rows_max * dim_max);
    c.xT = ws.take<float>(rows_max * dim_max);
    c.wT = ws.take<float>((size_t)dim_max * dim_max);
    c.st = st;
    return ws.ok();
}
static inline int imax(int a, int b) { return a > b ? a : b; }
static int ew(int op, size_t n, const float* a, const float* b, float* out, cudaStream_t st) {
    if (n == 0) return SUBGC_OK;
    ew_kernel<<<ew_blocks(n), 256, 0, st>>>(op, n, a, b, out, 0.f);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

}  // namespace subgc

extern "C" size_t subgc_frontend_backward_workspace_bytes(const subgc_dims* d, int n_images, int R, int len, int n_sub) {
    if (!d) return 0;
    const int dim = imax(imax(imax(2 * d->gcn, d->att_feat), imax(imax(d->fc_feat, d->rnn), imax(d->att_hid, d->embed))), d->low_rank);
    const size_t rows = (size_t)imax(imax(n_images * imax(d->obj_num, d->rel_num), R * len), imax(n_sub, R));
    size_t b = bwd_gemm_ws(rows, dim) + align_up(bwd_scratch_floats(rows, dim) * 4, 256) + 4096;
    b += 12 * align_up(rows * (size_t)dim * 4, 256);   // stage temporaries (gradients of intermediate activations)
    b += (size_t)(2 * d->gcn_layers + 2) * align_up((size_t)n_images * imax(d->obj_num, d->rel_num) * d->gcn * 4, 256);   // gx / gp per layer
    return b;
}

/* Backward of the feature preparation (AttModel.py:348-368 after gpn.py:79): fc_embed, read_out_proj (the read-out itself is detached,
 * gpn.py:78), att_embed (pack_wrapper: padded rows masked), ctx2att.  d_fc is modified in place.  d_x_obj [n_nodes, L] is overwritten. */
extern "C" int subgc_prepare_backward(const subgc_dims* d, const subgc_weights* w, int R, int len, int n_nodes, const subgc_prepare_train_saved* s,
                                      float* d_fc, const float* d_att, const float* d_p_att, const subgc_prepare_grads* g, float* d_x_obj,
                                      void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && s && g && d_fc && d_att && d_p_att && d_x_obj && ws_ && R > 0 && len > 0 && n_nodes > 0, "subgc_prepare_backward: bad arguments");
    cudaStream_t st = ST;
    const int H = d->rnn, FC = d->fc_feat, L = d->gcn, AH = d->att_hid, RL = R * len;
    Workspace ws(ws_, ws_bytes);
    BwdCtx c;
    const int dim = imax(imax(2 * L, FC), imax(H, AH));
    SUBGC_CHECK_ARG(bwd_ctx(c, ws, (size_t)imax(RL, R), dim, st), "subgc_prepare_backward: workspace too small");
    float* d_fcp = ws.take<float>((size_t)R * H);
    float* d_f1 = ws.take<float>((size_t)R * FC);
    float* d_gfc = ws.take<float>((size_t)R * 2 * L);
    float* d_hr = ws.take<float>((size_t)R * AH);
    float* d_att2 = ws.take<float>((size_t)RL * H);
    float* d_rows = ws.take<float>((size_t)RL * L);
    SUBGC_CHECK_ARG(ws.ok(), "subgc_prepare_backward: workspace too small");
    // fc path: dropout -> ReLU -> fc_embed.2 -> ReLU -> fc_embed.0 -> read_out_proj.1 -> read_out_proj.0
    if (s->m_fc) SUBGC_TRY(ew(EW_MUL, (size_t)R * H, d_fc, s->m_fc, d_fc, st));
    SUBGC_TRY(ew(EW_RELU_BWD, (size_t)R * H, s->fc_pre, d_fc, d_fcp, st));
    SUBGC_TRY(c.dW(g->fc2_w, FC, d_fcp, H, s->f1, FC, R)); SUBGC_TRY(subgc_colsum(R, H, d_fcp, H, g->fc2_b, 1, stream));
    SUBGC_TRY(c.dX(d_f1, d_fcp, R, H, w->fc_embed2.w, FC));
    SUBGC_TRY(ew(EW_RELU_BWD, (size_t)R * FC, s->f1, d_f1, d_f1, st));
    SUBGC_TRY(c.dW(g->fc0_w, 2 * L, d_f1, FC, s->g_fc, 2 * L, R)); SUBGC_TRY(subgc_colsum(R, FC, d_f1, FC, g->fc0_b, 1, stream));
    SUBGC_TRY(c.dX(d_gfc, d_f1, R, FC, w->fc_embed0.w, 2 * L));
    SUBGC_TRY(c.dW(g->ro1_w, AH, d_gfc, 2 * L, s->hr, AH, R)); SUBGC_TRY(subgc_colsum(R, 2 * L, d_gfc, 2 * L, g->ro1_b, 1, stream));
    SUBGC_TRY(c.dX(d_hr, d_gfc, R, 2 * L, w->read_out1.w, AH));
    SUBGC_TRY(c.dW(g->ro0_w, 2 * L, d_hr, AH, s->read_sel, 2 * L, R)); SUBGC_TRY(subgc_colsum(R, AH, d_hr, AH, g->ro0_b, 1, stream));
    // attention features: ctx2att, then att_embed through mask (dropout x valid rows) and ReLU, scattered back to the node rows
    SUBGC_TRY(c.dW(g->ctx2att_w, H, d_p_att, AH, s->att, H, RL)); SUBGC_TRY(subgc_colsum(RL, AH, d_p_att, AH, g->ctx2att_b, 1, stream));
    SUBGC_TRY(c.dX(d_att2, d_p_att, RL, AH, w->ctx2att.w, H));
    SUBGC_TRY(ew(EW_ADD, (size_t)RL * H, d_att2, d_att, d_att2, st));
    SUBGC_TRY(ew(EW_MUL, (size_t)RL * H, d_att2, s->m_att, d_att2, st));
    SUBGC_TRY(ew(EW_RELU_BWD, (size_t)RL * H, s->att_pre, d_att2, d_att2, st));
    SUBGC_TRY(c.dW(g->att_embed_w, L, d_att2, H, s->x_rows, L, RL)); SUBGC_TRY(subgc_colsum(RL, H, d_att2, H, g->att_embed_b, 1, stream));
    SUBGC_TRY(c.dX(d_rows, d_att2, RL, H, w->att_embed.w, L));
    SUBGC_CUDA(cudaMemsetAsync(d_x_obj, 0, (size_t)n_nodes * L * 4, st));
    SUBGC_TRY(subgc_scatter_add_rows(RL, L, d_rows, L, s->node_row, d_x_obj, L, stream));
    return SUBGC_OK;
}

/* Backward of the sGPN scorer (gpn.py:41-58): BCE(sigmoid) -> gpn_fc.3 -> Dropout(0.5) -> ReLU -> gpn_fc.0 -> max / mean pooling.
 * scale = d(gpn_loss) / n_sub.  d_x_obj [B, N, L] is ACCUMULATED into. */
extern "C" int subgc_sgpn_backward(const subgc_dims* d, const subgc_weights* w, const subgc_subgraph_layout* lay, const subgc_sgpn_train_saved* s,
                                   float scale, const float* x_obj, const int64_t* gpn_obj_ind, const subgc_sgpn_grads* g, float* d_x_obj, void* ws_,
                                   size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && lay && s && g && x_obj && gpn_obj_ind && d_x_obj && ws_, "subgc_sgpn_backward: bad arguments");
    cudaStream_t st = ST;
    const int n_sub = subgraph_count(*lay), L = d->gcn, AH = d->att_hid;
    Workspace ws(ws_, ws_bytes);
    BwdCtx c;
    SUBGC_CHECK_ARG(bwd_ctx(c, ws, (size_t)n_sub, imax(2 * L, AH), st), "subgc_sgpn_backward: workspace too small");
    float* dz = ws.take<float>(n_sub);
    float* d_hid = ws.take<float>((size_t)n_sub * AH);
    float* d_read = ws.take<float>((size_t)n_sub * 2 * L);
    SUBGC_CHECK_ARG(ws.ok(), "subgc_sgpn_backward: workspace too small");
    SUBGC_TRY(subgc_bce_sigmoid_bwd(lay, s->score, scale, dz, stream));
    SUBGC_TRY(c.dW(g->fc3_w, AH, dz, 1, s->hid_d, AH, n_sub)); SUBGC_TRY(subgc_colsum(n_sub, 1, dz, 1, g->fc3_b, 1, stream));
    SUBGC_TRY(c.dX(d_hid, dz, n_sub, 1, w->gpn_fc3.w, AH));
    if (s->m_gpn) SUBGC_TRY(ew(EW_MUL, (size_t)n_sub * AH, d_hid, s->m_gpn, d_hid, st));
    SUBGC_TRY(ew(EW_RELU_BWD, (size_t)n_sub * AH, s->hid, d_hid, d_hid, st));
    SUBGC_TRY(c.dW(g->fc0_w, 2 * L, d_hid, AH, s->read_out, 2 * L, n_sub)); SUBGC_TRY(subgc_colsum(n_sub, AH, d_hid, AH, g->fc0_b, 1, stream));
    SUBGC_TRY(c.dX(d_read, d_hid, n_sub, AH, w->gpn_fc0.w, 2 * L));
    return subgc_sgpn_pool_bwd(d, lay, x_obj, gpn_obj_ind, s->sub_len, d_read, d_x_obj, stream);
}

/* Backward of the GCN backbone (gcn_backbone.py:29-53, graph_conv.py:15-34, graph_conv_unit.py:28-36: live units only) and of the node
 * fusion (AttModel.py:370-378).  d_x_obj [B, N, L]: gradient of the encoder output (not modified). */
extern "C" int subgc_gcn_backward(const subgc_dims* d, const subgc_weights* w, int n_images, const subgc_gcn_train_saved* s, const float* d_x_obj,
                                  const subgc_gcn_grads* g, void* ws_, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && s && g && d_x_obj && ws_ && n_images > 0, "subgc_gcn_backward: bad arguments");
    cudaStream_t st = ST;
    const int B = n_images, N = d->obj_num, K = d->rel_num, L = d->gcn, Rk = d->low_rank, Ln = d->gcn_layers, Rs = d->gcn_residual;
    const int E = d->embed, A = d->att_feat;
    const size_t xn = (size_t)B * N * L, pn = (size_t)B * K * L;
    Workspace ws(ws_, ws_bytes);
    BwdCtx c;
    SUBGC_CHECK_ARG(bwd_ctx(c, ws, (size_t)B * imax(N, K), imax(imax(L, A), imax(E, Rk)), st), "subgc_gcn_backward: workspace too small");
    float* gx[SUBGC_MAX_GCN_LAYERS + 1];
    float* gp[SUBGC_MAX_GCN_LAYERS + 1];
    bool hx[SUBGC_MAX_GCN_LAYERS + 1], hp[SUBGC_MAX_GCN_LAYERS + 1];
    for (int l = 0; l <= Ln; ++l) { gx[l] = ws.take<float>(xn); gp[l] = ws.take<float>(pn); hx[l] = hp[l] = false; }
    const size_t big = (size_t)B * imax(N, K);
    float* dm_a = ws.take<float>(big * L);
    float* dm_b = ws.take<float>(big * L);
    float* dt = ws.take<float>(big * Rk);
    float* dsrc = ws.take<float>(big * L);
    float* d_emb = ws.take<float>((size_t)B * N * E);
    float* emb_rows = ws.take<float>((size_t)B * N * E);
    SUBGC_CHECK_ARG(ws.ok(), "subgc_gcn_backward: workspace too small");
    SUBGC_CUDA(cudaMemcpyAsync(gx[Ln], d_x_obj, xn * 4, cudaMemcpyDeviceToDevice, st));
    hx[Ln] = true;
    auto add_to = [&](float** lst, bool* has, int i, const float* v, size_t n) -> int {
        if (!has[i]) { SUBGC_CUDA(cudaMemcpyAsync(lst[i], v, n * 4, cudaMemcpyDeviceToDevice, st)); has[i] = true; return SUBGC_OK; }
        return ew(EW_ADD, n, lst[i], v, lst[i], st);
    };
    // backward of M = fc_rgt(fc_lft(src)) for one unit: gradients of both linears, d(src) accumulated into `acc` (first: overwrite)
    auto unit_bwd = [&](int l, int u, const float* src, const float* tmid, const float* dm, int rows, float* acc, bool first) -> int {
        SUBGC_TRY(c.dW(g->rgt_w[l][u], Rk, dm, L, tmid, Rk, rows)); SUBGC_TRY(subgc_colsum(rows, L, dm, L, g->rgt_b[l][u], 1, stream));
        SUBGC_TRY(c.dX(dt, dm, rows, L, w->gcn_rgt[l][u].w, Rk));
        SUBGC_TRY(c.dW(g->lft_w[l][u], L, dt, Rk, src, L, rows)); SUBGC_TRY(subgc_colsum(rows, Rk, dt, Rk, g->lft_b[l][u], 1, stream));
        if (first) return c.dX(acc, dt, rows, Rk, w->gcn_lft[l][u].w, L);
        SUBGC_TRY(c.dX(dm_a, dt, rows, Rk, w->gcn_lft[l][u].w, L));   // dm_a is free again by now
        return ew(EW_ADD, (size_t)rows * L, acc, dm_a, acc, st);
    };
    for (int l = Ln - 1; l >= 0; --l) {
        const subgc_gcn_layer_saved& rec = s->layer[l];
        const bool boundary = ((l + 1) % Rs == 0);
        if (hx[l + 1] && rec.y0) {   // units 0, 1 produced x(l+1) from the edge stream p(l)
            if (boundary) SUBGC_TRY(add_to(gx, hx, l + 1 - Rs, gx[l + 1], xn));
            SUBGC_TRY(subgc_gcn_node_bwd(B, N, K, L, gx[l + 1], rec.y0, rec.y1, s->rel_ind, dm_a, dm_b, stream));
            // unit 1 first (its d(src) lands in dsrc), then unit 0 adds to it: dm_a is consumed by unit 0's first contraction before reuse
            SUBGC_TRY(unit_bwd(l, 1, rec.p_in, rec.t1, dm_b, B * K, dsrc, true));
            {
                SUBGC_TRY(c.dW(g->rgt_w[l][0], Rk, dm_a, L, rec.t0, Rk, B * K)); SUBGC_TRY(subgc_colsum(B * K, L, dm_a, L, g->rgt_b[l][0], 1, stream));
                SUBGC_TRY(c.dX(dt, dm_a, B * K, L, w->gcn_rgt[l][0].w, Rk));
                SUBGC_TRY(c.dW(g->lft_w[l][0], L, dt, Rk, rec.p_in, L, B * K)); SUBGC_TRY(subgc_colsum(B * K, Rk, dt, Rk, g->lft_b[l][0], 1, stream));
                SUBGC_TRY(c.dX(dm_a, dt, B * K, Rk, w->gcn_lft[l][0].w, L));
                SUBGC_TRY(ew(EW_ADD, pn, dsrc, dm_a, dsrc, st));
            }
            SUBGC_TRY(add_to(gp, hp, l, dsrc, pn));
        }
        if (hp[l + 1] && rec.m2) {   // units 2, 3 produced p(l+1) from the node stream x(l)
            if (boundary) SUBGC_TRY(add_to(gp, hp, l + 1 - Rs, gp[l + 1], pn));
            SUBGC_TRY(subgc_gcn_edge_bwd(B, N, K, L, gp[l + 1], rec.m2, rec.m3, s->rel_ind, dm_a, dm_b, stream));
            SUBGC_TRY(unit_bwd(l, 3, rec.x_in, rec.t3, dm_b, B * N, dsrc, true));
            {
                SUBGC_TRY(c.dW(g->rgt_w[l][2], Rk, dm_a, L, rec.t2, Rk, B * N)); SUBGC_TRY(subgc_colsum(B * N, L, dm_a, L, g->rgt_b[l][2], 1, stream));
                SUBGC_TRY(c.dX(dt, dm_a, B * N, L, w->gcn_rgt[l][2].w, Rk));
                SUBGC_TRY(c.dW(g->lft_w[l][2], L, dt, Rk, rec.x_in, L, B * N)); SUBGC_TRY(subgc_colsum(B * N, Rk, dt, Rk, g->lft_b[l][2], 1, stream));
                SUBGC_TRY(c.dX(dm_a, dt, B * N, Rk, w->gcn_lft[l][2].w, L));
                SUBGC_TRY(ew(EW_ADD, xn, dsrc, dm_a, dsrc, st));
            }
            SUBGC_TRY(add_to(gx, hx, l, dsrc, xn));
        }
    }
    SUBGC_CHECK_ARG(hx[0], "subgc_gcn_backward: no gradient reached the fused node features");
    // fusion: x0 = relu(W_v att + b_v + W_e E[cls] + b_e)
    float* d_x0 = gx[0];
    SUBGC_TRY(ew(EW_RELU_BWD, xn, s->x0, d_x0, d_x0, st));
    SUBGC_TRY(c.dW(g->obj_v_w, A, d_x0, L, s->att_feats, A, B * N)); SUBGC_TRY(subgc_colsum(B * N, L, d_x0, L, g->obj_v_b, 1, stream));
    SUBGC_TRY(subgc_gather_rows(B * N, E, w->sg_obj_embed, E, s->cls, emb_rows, 0, stream));
    SUBGC_TRY(c.dW(g->obj_emb_w, E, d_x0, L, emb_rows, E, B * N)); SUBGC_TRY(subgc_colsum(B * N, L, d_x0, L, g->obj_emb_b, 1, stream));
    SUBGC_TRY(c.dX(d_emb, d_x0, B * N, L, w->obj_emb_proj.w, E));
    return subgc_scatter_add_rows(B * N, E, d_emb, E, s->cls, g->sg_obj_embed, E, stream);
}
