// Optimiser step of the Sub-GC training loop as two multi-tensor kernels:
//
//   utils.clip_gradient_norm(optimizer, 10.)   reference misc/utils.py:174-200: one global L2 norm over every gradient, then every
//                                              gradient scaled by clip / max(norm, clip)
//   optimizer.step()                           torch.optim.Adam as built by misc/utils.py:236 (train.py:107,163-164)
//
// The reference runs ~50 norm kernels, a host sync for the norm (`max(totalnorm, clip_norm)` on Python floats), ~50 scale kernels and
// Adam's multi-tensor sequence over 70 M parameters.  Here the norm is a deterministic two-level sum that stays on the device and the
// update reads p, g, m, v once and writes p, m, v once: 7 x 4 bytes per parameter, pure HBM streaming.
#include "common.cuh"

namespace subgc {

constexpr int kOptChunk = 16384;   // elements per block and table entry

struct OptChunk { float* p; const float* g; float* m; float* v; int n; int pad; };

// per-chunk sum of squares (fixed order inside a block), then one block adds the partials in chunk order: deterministic
__global__ void __launch_bounds__(256) grad_sqnorm_kernel(const OptChunk* __restrict__ chunks, float* __restrict__ partial) {
    __shared__ float red[32];
    const OptChunk c = chunks[blockIdx.x];
    float s = 0.f;
    const int n4 = c.n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(c.g);
    if ((reinterpret_cast<uintptr_t>(c.g) & 15) == 0) {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) {
            const float4 v = __ldg(g4 + i);
            s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (int i = 4 * n4 + threadIdx.x; i < c.n; i += blockDim.x) s += c.g[i] * c.g[i];
    } else {
        for (int i = threadIdx.x; i < c.n; i += blockDim.x) s += c.g[i] * c.g[i];
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
__global__ void __launch_bounds__(1024) grad_norm_finish_kernel(const float* __restrict__ partial, int n, float clip, float* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 32; ++i) t += red[i];
        const float norm = sqrtf((float)t);
        out[0] = norm;                                     // total norm (the reference's return value)
        out[1] = clip > 0.f ? clip / fmaxf(norm, clip) : 1.f;   // clip coefficient (misc/utils.py:196)
    }
}

struct AdamArgs { float lr, beta1, beta2, eps, weight_decay, bias1, bias2_sqrt; int write_grad; };

__device__ __forceinline__ void adam_elem(float& p, float& g, float& m, float& v, float coef, const AdamArgs& a) {
    g *= coef;                                           // p.grad.mul_(norm)
    float gg = g;
    if (a.weight_decay != 0.f) gg = fmaf(a.weight_decay, p, gg);
    m = m + (gg - m) * (1.f - a.beta1);                  // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(1.f - a.beta2, gg * gg, v * a.beta2);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    const float denom = sqrtf(v) / a.bias2_sqrt + a.eps;
    p = p - (a.lr / a.bias1) * (m / denom);              // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) clip_adam_kernel(const OptChunk* __restrict__ chunks, const float* __restrict__ norm_coef, const AdamArgs a) {
    const OptChunk c = chunks[blockIdx.x];
    const float coef = norm_coef ? norm_coef[1] : 1.f;
    const bool vec = ((reinterpret_cast<uintptr_t>(c.p) | reinterpret_cast<uintptr_t>(c.g) | reinterpret_cast<uintptr_t>(c.m) |
                       reinterpret_cast<uintptr_t>(c.v)) & 15) == 0;
    const int n4 = vec ? (c.n >> 2) : 0;
    float4* p4 = reinterpret_cast<float4*>(c.p);
    float4* m4 = reinterpret_cast<float4*>(c.m);
    float4* v4 = reinterpret_cast<float4*>(c.v);
    float4* g4 = reinterpret_cast<float4*>(const_cast<float*>(c.g));
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
        adam_elem(p.x, g.x, m.x, v.x, coef, a); adam_elem(p.y, g.y, m.y, v.y, coef, a);
        adam_elem(p.z, g.z, m.z, v.z, coef, a); adam_elem(p.w, g.w, m.w, v.w, coef, a);
        p4[i] = p; m4[i] = m; v4[i] = v;
        if (a.write_grad) g4[i] = g;
    }
    for (int i = 4 * n4 + threadIdx.x; i < c.n; i += blockDim.x) {
        float p = c.p[i], g = c.g[i], m = c.m[i], v = c.v[i];
        adam_elem(p, g, m, v, coef, a);
        c.p[i] = p; c.m[i] = m; c.v[i] = v;
        if (a.write_grad) const_cast<float*>(c.g)[i] = g;
    }
}

}  // namespace subgc

using namespace subgc;

extern "C" int subgc_opt_chunk_elems(void) { return kOptChunk; }

extern "C" int subgc_clip_adam_step(const void* chunks, int n_chunks, float clip_norm, float lr, float beta1, float beta2, float eps,
                                    float weight_decay, int step, int write_grad, float* partial, float* norm_out, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(chunks && n_chunks > 0 && partial && norm_out && step >= 1, "subgc_clip_adam_step: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const OptChunk* c = static_cast<const OptChunk*>(chunks);
    grad_sqnorm_kernel<<<n_chunks, 256, 0, st>>>(c, partial);
    SUBGC_LAUNCH_CHECK();
    grad_norm_finish_kernel<<<1, 1024, 0, st>>>(partial, n_chunks, clip_norm, norm_out);
    SUBGC_LAUNCH_CHECK();
    AdamArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
    a.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
    a.write_grad = write_grad;
    clip_adam_kernel<<<n_chunks, 256, 0, st>>>(c, norm_out, a);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}
