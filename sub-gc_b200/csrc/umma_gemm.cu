// Tensor-core version of the contraction block: C[M,N] = epi( sum_seg A_seg[M,K] . W_seg[N,K]^T ) on tcgen05 (sm_100a)
// with split-precision TF32 operands ("3xTF32"), fp32 accumulation in TMEM.
//
// Why split precision: token ids / beam indices must match the fp32 reference and the path's top-1/top-2 log-prob
// margin is ~1e-5, so a single TF32 pass (10-bit mantissa) is not admissible.  Each fp32 operand v is
// split as v = hi + lo with hi = rn_tf32(v), lo = rn_tf32(v - hi); the product is accumulated as
// hi*hi + lo*hi + hi*lo (the dropped lo*lo term is 2^-22 relative), which measures at the fp32 noise floor.
//
// Dataflow of one CTA (192 threads = 6 warps, 1 CTA / SM, persistent over its k-range):
//   warp 0      TMA producer: per k-block (16 fp32 = one 64-byte swizzled row) loads the raw fp32 weight tile
//               W[256 x 16] straight from the reference's un-repacked nn.Linear / LSTMCell tensor and the raw fp32
//               activation tile X[128 x 16] of the K segment it belongs to   (cp.async.bulk.tensor, mbarrier tx)
//   warps 2-5   transform: write lo = v - trunc_tf32(v) of both tiles to second buffers (the raw tiles serve as
//               "hi": the tensor core reads only their tf32 bits), fence.proxy.async, signal the MMA warp; after
//               the main loop they are the epilogue warps
//   warp 1      one elected thread issues 3 x 2 tcgen05.mma.kind::tf32 (M=128, N=256, K=8) per k-block into a
//               256-column TMEM accumulator, tcgen05.commit frees the stage / publishes the accumulator
//   epilogue    tcgen05.ld 32x32b.x32 -> registers -> per-warp smem transpose -> coalesced 128-byte row stores of
//               the split-K partial [z][M][N]; the reduction + epilogue kernel is shared with the SIMT block.
// K segments (the un-concatenated LSTM inputs) are separate tensor maps on both sides, so plain activations are read
// in place; only segments that need a row gather / shared rows / ReLU-on-load are materialised first (gather_seg_kernel).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace subgc {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 16;
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_W_BYTES = TC_BN * TC_BK * 4;                     // 16 KB
constexpr int TC_X_BYTES = TC_BM * TC_BK * 4;                     // 8 KB
constexpr int TC_STAGE_BYTES = 2 * TC_W_BYTES + 2 * TC_X_BYTES;   // W raw | W lo | X raw | X lo = 48 KB
constexpr int TC_TX_BYTES = TC_W_BYTES + TC_X_BYTES;              // bytes TMA lands per stage
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_MAX_SEG = 4;
constexpr bool TC_L2_PREFETCH = false;  // measured on B200: no gain in steady state (the loop is shared-memory-bandwidth
                                        // bound) and the burst of prefetches delays the first demand tile by ~3 us
constexpr int TC_PF_KB = 8;       // weights are requested into L2 in chunks of 8 k-blocks (128 columns = 512 contiguous bytes per row),
                                  // two chunks ahead of the smem ring: long row runs for the DRAM pages, latency cover without smem
constexpr int TC_MAX_CHAIN = 64;  // k-blocks (1024 columns) accumulated in TMEM before an fp32 round-to-nearest combine

struct TcParams {
    CUtensorMap tm_x[TC_MAX_SEG];
    CUtensorMap tm_w[TC_MAX_SEG];
    CUtensorMap tm_wpf[TC_MAX_SEG];  // same tensors, box 128 columns x 256 rows: L2 prefetch of 512-byte row runs (DRAM page locality)
    int seg_kb_end[TC_MAX_SEG];  // cumulative k-block count at the end of each segment
    int nseg;
    int M, N;
    int kb_total, kb_per_split;
    float* part;                 // [splits][M][N]
    int direct;                  // single split: apply the epilogue here and write C (no partial round trip)
    GemmEpilogue epi;
    float* C;
    int ldc;
    const int* active;
    int raw_hi;                  // 1 (default): the raw fp32 weight tile is the "hi" operand -- the tensor core reads only the tf32
                                 // bits, i.e. truncates (measured: same error as an explicit split) -- and only lo = w - trunc(w) is
                                 // written; 0 (SUBGC_TC_REWRITE_HI=1): hi = rn_tf32(w) is rewritten in place as well
    int collect;                 // A-operand collector reuse for the X-hi tile (SUBGC_TC_NOCOLLECT=1 disables)
    int epi_mode;                // debug (SUBGC_TC_EPI=1: TMEM loads only, 2: stores only)
    long long* trace;            // debug (SUBGC_TC_TRACE=1): per-role clock64 stamps of CTA (0,0,0); nullptr in normal operation
};
constexpr int TC_TRACE_SLOTS = 64;  // k-blocks traced per role

#define TC_STAMP(role, i)                                                                                      \
    do {                                                                                                       \
        if (p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (i) < TC_TRACE_SLOTS) \
            p.trace[(role) * TC_TRACE_SLOTS + (i)] = clock64();                                                \
    } while (0)

// ---- main kernel ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1) umma_gemm_kernel(const __grid_constant__ TcParams p) {
    if (p.active != nullptr && *p.active == 0) return;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar_base = base + TC_STAGES * TC_STAGE_BYTES;
    // barriers: full[s] @ +0, xf[s] @ +32, empty[s] @ +64, accum @ +96, tmem slot @ +104
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + TC_STAGES * TC_STAGE_BYTES + 104);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * TC_BN, m0 = blockIdx.y * TC_BM;
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_base + 8 * s, 1);        // full: producer arrive + tx bytes
            mbar_init(bar_base + 32 + 8 * s, 4);   // transformed: one arrive per transform warp
            mbar_init(bar_base + 64 + 8 * s, 1);   // empty: tcgen05.commit
        }
        mbar_init(bar_base + 96, 1);               // accumulator ready
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: two 256-column fp32 accumulators (main hi*hi chain | cross terms), i.e. all 512 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    if (threadIdx.x == 0) TC_STAMP(0, 0);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int sgi = 0; sgi < p.nseg; ++sgi) { prefetch_tensormap(&p.tm_w[sgi]); prefetch_tensormap(&p.tm_x[sgi]); }
            // weight k-block -> (segment, column) lookup; `pseg` trails the L2 prefetch cursor, `seg` the smem ring
            int seg = 0, pseg = 0;
            while (seg < p.nseg - 1 && kb_begin >= p.seg_kb_end[seg]) ++seg;
            pseg = seg;
            auto prefetch_chunk = [&](int kb0) {  // k-blocks [kb0, kb0 + TC_PF_KB), split at segment boundaries
                int kb = kb0;
                const int kend = min(kb0 + TC_PF_KB, kb_end);
                while (kb < kend) {
                    while (pseg < p.nseg - 1 && kb >= p.seg_kb_end[pseg]) ++pseg;
                    const int k0 = pseg == 0 ? 0 : p.seg_kb_end[pseg - 1];
                    tma_prefetch_l2_2d(&p.tm_wpf[pseg], (kb - k0) * TC_BK, n0);
                    kb = min(kend, p.seg_kb_end[pseg]);
                }
            };
            if (TC_L2_PREFETCH) {
                prefetch_chunk(kb_begin);
                prefetch_chunk(kb_begin + TC_PF_KB);
            }
            for (int i = 0; i < nkb; ++i) {
                const int kb = kb_begin + i;
                if (TC_L2_PREFETCH && (i % TC_PF_KB) == 0 && i + 2 * TC_PF_KB < nkb) prefetch_chunk(kb + 2 * TC_PF_KB);
                while (seg < p.nseg - 1 && kb >= p.seg_kb_end[seg]) ++seg;
                const int seg_kb0 = seg == 0 ? 0 : p.seg_kb_end[seg - 1];
                const int s = i % TC_STAGES;
                const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
                mbar_wait(bar_base + 64 + 8 * s, ph ^ 1u);
                TC_STAMP(1, i);
                const uint32_t st = base + s * TC_STAGE_BYTES;
                const uint32_t full = bar_base + 8 * s;
                mbar_arrive_expect_tx(full, TC_TX_BYTES);
                tma_load_2d(st, &p.tm_w[seg], full, (kb - seg_kb0) * TC_BK, n0);
                tma_load_2d(st + 2 * TC_W_BYTES, &p.tm_x[seg], full, (kb - seg_kb0) * TC_BK, m0);
                TC_STAMP(2, i);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D fp32, A/B tf32, both K-major, N=256, M=128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % TC_STAGES;
            const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
            mbar_wait(bar_base + 32 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                TC_STAMP(5, i);
                const uint32_t st = base + s * TC_STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    const uint64_t whi = umma_desc_sw64(st + ks * 32);
                    const uint64_t wlo = umma_desc_sw64(st + TC_W_BYTES + ks * 32);
                    const uint64_t xhi = umma_desc_sw64(st + 2 * TC_W_BYTES + ks * 32);
                    const uint64_t xlo = umma_desc_sw64(st + 2 * TC_W_BYTES + TC_X_BYTES + ks * 32);
                    // The tensor core truncates its fp32 accumulator on every add, a bias that grows with the chain length.
                    // The two cross terms (2^-11 of the main term) therefore get their own accumulator: the main chain
                    // sees one add per K-step instead of three, and the cross chain's truncation is 2^-11 smaller.
                    const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                    if (p.collect) {
                        umma_tf32_afill(tmem_d, xhi, whi, idesc, acc);
                        umma_tf32_alast(tmem_d + TC_BN, xhi, wlo, idesc, acc);
                        umma_tf32(tmem_d + TC_BN, xlo, whi, idesc, 1u);
                    } else {
                        umma_tf32(tmem_d, xhi, whi, idesc, acc);
                        umma_tf32(tmem_d + TC_BN, xlo, whi, idesc, acc);
                        umma_tf32(tmem_d + TC_BN, xhi, wlo, idesc, 1u);
                    }
                }
                umma_commit(bar_base + 64 + 8 * s);              // stage free once these MMAs have read it
                if (i == nkb - 1) umma_commit(bar_base + 96);    // accumulator complete
                TC_STAMP(6, i);
            }
            __syncwarp();
        }
    } else {
        // ===== transform warps (2..5), then epilogue =====
        const int t = threadIdx.x - 64;  // 0..127
        for (int i = 0; i < nkb; ++i) {
            const int s = i % TC_STAGES;
            const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
            mbar_wait(bar_base + 8 * s, ph);
            if (t == 0) TC_STAMP(3, i);
            float4* whi = reinterpret_cast<float4*>(gbase + s * TC_STAGE_BYTES);
            float4* wlo = reinterpret_cast<float4*>(gbase + s * TC_STAGE_BYTES + TC_W_BYTES);
#pragma unroll
            for (int j = 0; j < TC_W_BYTES / 16 / 128; ++j) {
                const int idx = t + 128 * j;
                const float4 v = whi[idx];
                // hi = v rounded to tf32 (integer round-half-up on the 13 dropped bits: 2 ALU ops, finite inputs), lo = v - hi is
                // exact in fp32; the tensor core reads only the tf32 bits of lo (a 2^-22 relative residual)
                float4 h, l;
                h.x = __uint_as_float((__float_as_uint(v.x) + 0x1000u) & 0xffffe000u);
                h.y = __uint_as_float((__float_as_uint(v.y) + 0x1000u) & 0xffffe000u);
                h.z = __uint_as_float((__float_as_uint(v.z) + 0x1000u) & 0xffffe000u);
                h.w = __uint_as_float((__float_as_uint(v.w) + 0x1000u) & 0xffffe000u);
                l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                if (p.raw_hi) {
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                } else {
                    whi[idx] = h;
                }
                wlo[idx] = l;
            }
            {
                const float4* xr = reinterpret_cast<const float4*>(gbase + s * TC_STAGE_BYTES + 2 * TC_W_BYTES);
                float4* xl = reinterpret_cast<float4*>(gbase + s * TC_STAGE_BYTES + 2 * TC_W_BYTES + TC_X_BYTES);
#pragma unroll
                for (int j = 0; j < TC_X_BYTES / 16 / 128; ++j) {
                    const int idx = t + 128 * j;
                    const float4 v = xr[idx];
                    float4 l;
                    l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                    l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                    l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                    l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                    xl[idx] = l;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 32 + 8 * s);
            if (t == 0) TC_STAMP(4, i);
        }
        // epilogue: TMEM lanes of this warp = 32 * (warp % 4)
        mbar_wait(bar_base + 96, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (t == 0) TC_STAMP(0, 1);
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;            // TMEM lane == output row
        float* out = p.part + (size_t)blockIdx.z * p.M * p.N + (size_t)row * p.N;
        const bool vec = ((p.N & 3) == 0);
#define TC_TMEM_LD32(R, ADDR)                                                                                                             \
    asm volatile(                                                                                                                        \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, " \
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                            \
        : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]), "=r"(R[9]),        \
          "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15]), "=r"(R[16]), "=r"(R[17]), "=r"(R[18]),           \
          "=r"(R[19]), "=r"(R[20]), "=r"(R[21]), "=r"(R[22]), "=r"(R[23]), "=r"(R[24]), "=r"(R[25]), "=r"(R[26]), "=r"(R[27]),           \
          "=r"(R[28]), "=r"(R[29]), "=r"(R[30]), "=r"(R[31])                                                                             \
        : "r"(ADDR))
        const uint32_t tlane = tmem_d + ((uint32_t)(q * 32) << 16);
        const int nchunks = min(TC_BN / 32, (p.N - n0 + 31) / 32);
        uint32_t ra[32], rb[32], na[32], nb[32];
        TC_TMEM_LD32(ra, tlane);
        TC_TMEM_LD32(rb, tlane + TC_BN);
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const int col0 = n0 + c * 32;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (t == 0) TC_STAMP(7, 2 * c);
            if (c + 1 < nchunks && p.epi_mode != 2) {  // next chunk's TMEM reads overlap this chunk's global stores
                TC_TMEM_LD32(na, tlane + (uint32_t)((c + 1) * 32));
                TC_TMEM_LD32(nb, tlane + TC_BN + (uint32_t)((c + 1) * 32));
            }
            if (row < p.M && p.direct) {
                // epilogue in place for the simple cases (+ bias + bias2, ReLU, += C); anything with a divisor, addend or
                // zero-padding goes through the partial buffer + reduce kernel.  Four warps do this for the whole tile: keep it lean.
                float* crow = p.C + (size_t)row * p.ldc;
                const bool vecc = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (col0 + 32 <= p.N);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    // the (row-invariant) bias / addend values of 8 columns are fetched first, as independent loads under uniform
                    // conditions: one dependent load->add chain per element would cost ~200 cycles each
                    float b1[8], b2[8], v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { b1[u] = 0.f; b2[u] = 0.f; }
                    if (p.epi.bias) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) b1[u] = __ldg(p.epi.bias + min(col0 + j + u, p.N - 1));
                    }
                    if (p.epi.bias2) {
#pragma unroll
                        for (int u = 0; u < 8; ++u) b2[u] = __ldg(p.epi.bias2 + min(col0 + j + u, p.N - 1));
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        float t = __uint_as_float(ra[j + u]) + __uint_as_float(rb[j + u]);
                        if (p.epi.bias) t += b1[u];
                        if (p.epi.bias2) t += b2[u];
                        v[u] = p.epi.relu ? fmaxf(t, 0.f) : t;
                    }
                    if (vecc) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float4* dst = reinterpret_cast<float4*>(crow + col0 + j + 4 * h);
                            float4 o = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
                            if (p.epi.accumulate) { const float4 t = *dst; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
                            *dst = o;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (col0 + j + u < p.N) crow[col0 + j + u] = p.epi.accumulate ? v[u] + crow[col0 + j + u] : v[u];
                    }
                }
            } else if (row < p.M && p.epi_mode != 1) {
                if (vec && col0 + 32 <= p.N) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v;
                        v.x = __uint_as_float(ra[j]) + __uint_as_float(rb[j]);
                        v.y = __uint_as_float(ra[j + 1]) + __uint_as_float(rb[j + 1]);
                        v.z = __uint_as_float(ra[j + 2]) + __uint_as_float(rb[j + 2]);
                        v.w = __uint_as_float(ra[j + 3]) + __uint_as_float(rb[j + 3]);
                        *reinterpret_cast<float4*>(out + col0 + j) = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N) out[col0 + j] = __uint_as_float(ra[j]) + __uint_as_float(rb[j]);
                }
            }
            if (t == 0) TC_STAMP(7, 2 * c + 1);
            if (c + 1 < nchunks) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) { ra[j] = na[j]; rb[j] = nb[j]; }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (t == 0) TC_STAMP(0, 2);
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
    }
}

// ---- materialisation of a K segment whose rows are gathered / shared / ReLU-ed on load --------------------------------------
__global__ void __launch_bounds__(256) gather_seg_kernel(const GemmSeg g, int M, int Kp, float* __restrict__ out, const int* __restrict__ active) {
    if (active != nullptr && *active == 0) return;
    const int kq = Kp >> 2;
    const size_t total = (size_t)M * kq;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx / kq), k = (int)(idx - (size_t)m * kq) << 2;
        const long long src = g.gather ? g.gather[m] : (g.gather32 ? (long long)g.gather32[m] : (long long)(m / g.a_row_div));
        const float* ptr = g.A + src * g.lda + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < g.K) v.x = __ldg(ptr);
        if (k + 1 < g.K) v.y = __ldg(ptr + 1);
        if (k + 2 < g.K) v.z = __ldg(ptr + 2);
        if (k + 3 < g.K) v.w = __ldg(ptr + 3);
        if (g.relu_a) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn();
void* tc_encode_fn() { return reinterpret_cast<void*>(encode_fn()); }
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &ptr, 12000, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    });
    return fn;
}

// 2-D fp32 tensor map: inner dim `cols` (contiguous), outer dim `rows` with `ld` floats between rows; box 16 x box_rows
static bool make_map(CUtensorMap* out, const float* base, int rows, int cols, long long ld, int box_rows, int box_cols = TC_BK) {
    struct Key {
        const void* base; int rows, cols; long long ld; int box, boxc;
        bool operator==(const Key& o) const {
            return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box && boxc == o.boxc;
        }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            size_t h = reinterpret_cast<size_t>(k.base);
            h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols; h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.box; h = h * 1000003u ^ (size_t)k.boxc;
            return h;
        }
    };
    static thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;  // encoding is pure: same key -> same descriptor
    Key key{base, rows, cols, ld, box_rows, box_cols};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const bool tile = box_cols == TC_BK;  // smem tiles are 64-byte swizzled; the wide box is only ever used for L2 prefetch
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    tile ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                    tile ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return false;
    if (cache.size() > 512) cache.clear();
    cache.emplace(key, *out);
    return true;
}

static int tc_mode() {  // SUBGC_GEMM=simt forces the fp32 FMA block (bisecting / A-B measurements)
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("SUBGC_GEMM");
        mode = (e && strcmp(e, "simt") == 0) ? 0 : 1;
    }
    return mode;
}

bool tc_eligible(const GemmProblem& p) {
    if (!tc_mode()) return false;
    if (p.M < 1 || p.N < 64 || p.nseg < 1 || p.nseg > TC_MAX_SEG) return false;
    long long ktot = 0;
    for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg& g = p.seg[s];
        if ((g.K & 3) || (g.ldw & 3) || (reinterpret_cast<uintptr_t>(g.W) & 15)) return false;
        ktot += g.K;
    }
    return ktot >= 64;
}

struct TcPlan { int m_tiles, n_tiles, kb_total, splits, kb_per_split, Mpad, Kpad; };

static TcPlan tc_plan(int M, int N, const int* segK, int nseg) {
    TcPlan pl;
    pl.m_tiles = (M + TC_BM - 1) / TC_BM;
    pl.n_tiles = (N + TC_BN - 1) / TC_BN;
    pl.kb_total = 0;
    for (int s = 0; s < nseg; ++s) pl.kb_total += (segK[s] + TC_BK - 1) / TC_BK;
    const int tiles = pl.m_tiles * pl.n_tiles;
    int splits = kNumSMs / tiles;
    const int by_k = pl.kb_total / 8;  // at least 8 k-blocks (128 columns) per split
    if (splits > by_k) splits = by_k;
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
    pl.kb_per_split = (pl.kb_total + splits - 1) / splits;
    if (pl.kb_per_split > TC_MAX_CHAIN) pl.kb_per_split = TC_MAX_CHAIN;  // bound the truncating accumulation chain
    pl.splits = (pl.kb_total + pl.kb_per_split - 1) / pl.kb_per_split;
    pl.Mpad = pl.m_tiles * TC_BM;
    pl.Kpad = pl.kb_total * TC_BK;
    return pl;
}

// upper bound of the workspace the tensor-core path needs for an (M, N, Ktotal) problem with <= 4 segments
size_t tc_workspace_bytes(int M, int N, int Ktotal) {
    const int m_tiles = (M + TC_BM - 1) / TC_BM, n_tiles = (N + TC_BN - 1) / TC_BN;
    int splits = kNumSMs / (m_tiles * n_tiles);
    if (splits > 32) splits = 32;
    if (splits < 1) splits = 1;
    const int by_chain = (Ktotal / TC_BK + TC_MAX_SEG + TC_MAX_CHAIN - 1) / TC_MAX_CHAIN + 1;
    if (splits < by_chain) splits = by_chain;
    const size_t Kpad = (size_t)Ktotal + TC_MAX_SEG * 8;   // per-segment padding of the materialised / split activation copies
    return align_up((size_t)M * Kpad * 4, 256) + 2 * TC_MAX_SEG * 256 + align_up((size_t)splits * M * N * 4, 256) + 1024;
}

void launch_splitk_reduce(const GemmProblem& p, const float* part, int splits, cudaStream_t stream, bool write_c16 = false);

// raw != nullptr: leave the split-K partials [splits][M][N] (no epilogue) for a fused consumer and report where they are
int launch_gemm_tc(const GemmProblem& p, void* ws_, size_t ws_bytes, cudaStream_t stream, RawPartials* raw) {
    int segK[TC_MAX_SEG];
    for (int s = 0; s < p.nseg; ++s) segK[s] = p.seg[s].K;
    const TcPlan pl = tc_plan(p.M, p.N, segK, p.nseg);
    Workspace ws(ws_, ws_bytes);
    const bool direct = (pl.splits == 1 && raw == nullptr && p.epi.div == 0.f && p.epi.addend == nullptr && p.epi.group == 0);
    float* part = direct ? nullptr : ws.take<float>((size_t)pl.splits * p.M * p.N);
    TcParams tp;
    int kb = 0;
    for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg& g = p.seg[s];
        const int nkb = (g.K + TC_BK - 1) / TC_BK;
        kb += nkb;
        tp.seg_kb_end[s] = kb;
        if (!make_map(&tp.tm_w[s], g.W, p.N, g.K, g.ldw, TC_BN)) {
            set_error("gemm(tc): cuTensorMapEncodeTiled failed for weight segment %d", s);
            return SUBGC_E_CUDA;
        }
        if (TC_L2_PREFETCH && !make_map(&tp.tm_wpf[s], g.W, p.N, g.K, g.ldw, TC_BN, TC_PF_KB * TC_BK)) {
            set_error("gemm(tc): cuTensorMapEncodeTiled failed for the prefetch map of weight segment %d", s);
            return SUBGC_E_CUDA;
        }
        const float* xa = g.A;
        long long xld = g.lda;
        const bool plain = g.gather == nullptr && g.gather32 == nullptr && g.a_row_div == 1 && g.relu_a == 0 && (g.lda & 3) == 0 &&
                           (reinterpret_cast<uintptr_t>(g.A) & 15) == 0;
        if (!plain) {  // gathered / shared / ReLU-ed rows: materialise this segment once, [M, K rounded up to 4]
            const int Kp = (g.K + 3) & ~3;
            float* tmp = ws.take<float>((size_t)p.M * Kp);
            if (!ws.ok()) break;
            size_t quads = (size_t)p.M * (Kp >> 2);
            int gb = (int)((quads + 255) / 256);
            if (gb > kNumSMs * 8) gb = kNumSMs * 8;
            gather_seg_kernel<<<gb, 256, 0, stream>>>(g, p.M, Kp, tmp, p.active);
            SUBGC_LAUNCH_CHECK();
            xa = tmp;
            xld = Kp;
        }
        if (!make_map(&tp.tm_x[s], xa, p.M, g.K, xld, TC_BM)) {
            set_error("gemm(tc): cuTensorMapEncodeTiled failed for activation segment %d", s);
            return SUBGC_E_CUDA;
        }
    }
    if (!ws.ok()) {
        set_error("gemm(tc): workspace too small (%zu bytes given)", ws_bytes);
        return SUBGC_E_WORKSPACE;
    }
    for (int s = p.nseg; s < TC_MAX_SEG; ++s) { tp.seg_kb_end[s] = kb; tp.tm_w[s] = tp.tm_w[0]; tp.tm_x[s] = tp.tm_x[0]; tp.tm_wpf[s] = tp.tm_w[0]; }
    if (!TC_L2_PREFETCH) for (int s = 0; s < p.nseg; ++s) tp.tm_wpf[s] = tp.tm_w[s];
    tp.nseg = p.nseg; tp.M = p.M; tp.N = p.N; tp.kb_total = pl.kb_total; tp.kb_per_split = pl.kb_per_split; tp.part = part; tp.active = p.active;
    tp.direct = direct ? 1 : 0; tp.epi = p.epi; tp.C = p.C; tp.ldc = p.ldc;
    static DeviceOnce once;
    SUBGC_CUDA(once.run([]() { return cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES); }));
    dim3 grid(pl.n_tiles, pl.m_tiles, pl.splits);
    static const bool trace_on = getenv("SUBGC_TC_TRACE") != nullptr;  // debugging aid only: allocates and synchronises
    static long long* trace_buf = nullptr;
    static const bool raw_hi = getenv("SUBGC_TC_REWRITE_HI") == nullptr;
    tp.raw_hi = raw_hi ? 1 : 0;
    static const int epi_mode = getenv("SUBGC_TC_EPI") ? atoi(getenv("SUBGC_TC_EPI")) : 0;
    tp.epi_mode = epi_mode;
    static const bool collect = getenv("SUBGC_TC_NOCOLLECT") == nullptr;
    tp.collect = collect ? 1 : 0;
    tp.trace = nullptr;
    if (trace_on) {
        if (!trace_buf) cudaMalloc(&trace_buf, 8 * TC_TRACE_SLOTS * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, 8 * TC_TRACE_SLOTS * sizeof(long long), stream);
        tp.trace = trace_buf;
    }
    umma_gemm_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tp);
    SUBGC_LAUNCH_CHECK();
    if (trace_on) {
        static int dumps = 0;
        long long h[8 * TC_TRACE_SLOTS];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
        if (dumps++ < 40) {
            const long long t0 = h[0];
            fprintf(stderr, "[tc-trace] M=%d N=%d kb=%d per_split=%d grid=(%d,%d,%d): acc_ready=%lld epi_done=%lld\n", p.M, p.N, pl.kb_total,
                    pl.kb_per_split, pl.n_tiles, pl.m_tiles, pl.splits, h[1] - t0, h[2] - t0);
            fprintf(stderr, "[tc-trace]  epilogue chunk (start, stored):");
            for (int c = 0; c < 8; ++c) fprintf(stderr, " (%lld, %lld)", h[7 * TC_TRACE_SLOTS + 2 * c] - t0, h[7 * TC_TRACE_SLOTS + 2 * c + 1] - t0);
            fprintf(stderr, "\n");
            for (int i = 0; i < pl.kb_per_split && i < TC_TRACE_SLOTS; ++i)
                fprintf(stderr, "[tc-trace]  kb %2d: prod_wait %6lld prod_issued %6lld | xf_start %6lld xf_done %6lld | mma_start %6lld mma_issued %6lld\n", i,
                        h[1 * TC_TRACE_SLOTS + i] - t0, h[2 * TC_TRACE_SLOTS + i] - t0, h[3 * TC_TRACE_SLOTS + i] - t0, h[4 * TC_TRACE_SLOTS + i] - t0,
                        h[5 * TC_TRACE_SLOTS + i] - t0, h[6 * TC_TRACE_SLOTS + i] - t0);
        }
    }
    if (direct) return SUBGC_OK;
    if (raw) {
        raw->part = part;
        raw->splits = pl.splits;
        return SUBGC_OK;
    }
    launch_splitk_reduce(p, part, pl.splits, stream);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

}  // namespace subgc
