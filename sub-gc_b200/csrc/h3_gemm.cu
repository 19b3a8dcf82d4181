// "h3": the contraction block on tcgen05 with split-fp16 operands, for weights that carry a packed copy (subgc_packed).
//
//   C[M,N] = epi( sum_seg X_seg[M,K] . W_seg[N,K]^T ),   v = hi + lo * 2^-11,  hi = rn_f16(v),  lo = rn_f16((v - hi) * 2^11)
//   X.W^T  = Xhi.Whi^T  +  2^-11 (Xhi.Wlo^T + Xlo.Whi^T)          (the dropped lo.lo term is 2^-22 relative)
//
// Why this form instead of the split-TF32 kernel (umma_gemm.cu): the decode step streams 136 MB of weights per token and is
// HBM-bound on paper, but splitting fp32 tiles into hi/lo inside the kernel moves every weight tile through shared memory six
// times and runs three TF32 MMAs, which measured at 31 % of the HBM roofline.  hi/lo in fp16 are exactly the 4 bytes per weight
// of the fp32 tensor (same algorithmic HBM bytes), need no in-kernel transform (TMA lands MMA-ready tiles) and kind::f16 runs
// at twice the TF32 rate; 22 mantissa bits per operand and fp32 accumulation keep the result at the fp32 noise floor
// (tools/gemm_check.py).  Activations are split by their producer kernel (decode loop) or by split_rows_kernel here.
//
// One CTA (128 threads = one warp per SM sub-partition, 1 CTA per SM) computes a 128 x 256 tile over its k-range:
//   warp 0     (lane 0) TMA producer: per k-block (32 fp16 = one 128-byte swizzled row) four tiles  Whi | Wlo [bn x 64], Xhi | Xlo [128 x 64]
//   warp 1     one thread issues 4 x 3 tcgen05.mma.kind::f16 (M=128, N=bn, K=16) per k-block: main chain -> TMEM columns [0,256),
//              the two cross terms -> columns [256,512) (separate chains: the accumulator truncation of the main chain is not
//              multiplied by three adds per step, and the cross sum is scaled by 2^-11 once, exactly, in the epilogue)
//   all 4      epilogue (after their main-loop role): tcgen05.ld -> main + cross * 2^-11 -> split-K partial [z][M][N] or the in-place epilogue (single split)
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace subgc {

constexpr int H3_BM = 128;
constexpr int H3_BN_MAX = 256;
constexpr int H3_BK = 64;   // fp16 elements per k-block: 128-byte rows (TMA issues one request per box row: 64-byte rows measured request-bound)
constexpr int H3_MAX_STAGES = 8;
constexpr int H3_THREADS = 128;   // 4 warps, one per SM sub-partition: leaves 12 K registers per sub-partition for co-resident kernels
constexpr int H3_X_BYTES = H3_BM * H3_BK * 2;                    // 16 KB
constexpr int H3_SMEM_EXTRA = 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int H3_SMEM_BUDGET = 196 * 1024;   // pipeline bytes: leaves ~30 KB of the SM's shared memory for a co-resident consumer kernel (PDL)
constexpr int H3_MAX_SEG = 4;
constexpr int H3_MAX_CHAIN = 40;   // k-blocks (128 MMA accumulation steps) per TMEM chain before the fp32 combine of split-K
constexpr float H3_LO_INV = kH3LoInv;

struct H3Params {
    CUtensorMap tm_xh[H3_MAX_SEG], tm_xl[H3_MAX_SEG], tm_wh[H3_MAX_SEG], tm_wl[H3_MAX_SEG];
    int seg_kb_end[H3_MAX_SEG];
    int nseg;
    int M, N;
    int bn;        // tile width (multiple of 32, <= 256): the weight tiles are [bn x 32]
    int stages;    // ring depth (<= 8)
    int kb_total, kb_per_split;
    float* part;   // [splits][M][N]
    int direct;    // single split: apply the epilogue here and write C
    GemmEpilogue epi;
    float* C;
    int ldc;
    const int* active;
    CUtensorMap tm_part;   // [splits, M, N] fp32 partials, box 32 x 32 x 1, 128-byte swizzle (valid when tma_part)
    int tma_part;          // 1: the partial tiles leave through shared memory + TMA stores (full 128-byte lines) instead of per-thread rows
    TraceSlot trace;
    CellEpilogue cell;     // kCell instantiation only
    int* overflow;         // fp16-range flag for the split copy of the result (epi.c16_*)
    int pf_dist;           // k-blocks the L2 prefetch of weight tiles runs ahead of the ring (0 = off)
    int dbg;               // timing experiments only (SUBGC_H3_DBG bit mask, results are wrong): 1 no activation loads after the first
                           // ring round, 2 no weight loads after it, 4 no MMAs
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
    return v;
}

// kCell: gate-grouped tiles (bn = 128 = 4 gates x 32 hidden units, weight rows q*H + j0 .. +32 per gate), the grid's z dimension is a
// cluster; the epilogue reduces the k-splits through distributed shared memory and applies the LSTM cell (CellEpilogue)
template <bool kCell>
__global__ void __launch_bounds__(H3_THREADS, 1) h3_gemm_kernel(const __grid_constant__ H3Params p) {
    extern __shared__ uint8_t smem_raw[];
    trace_begin(p.trace);
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const int bn = p.bn, stages = p.stages;
    const uint32_t w_bytes = (uint32_t)bn * H3_BK * 2;                 // one weight tile (hi or lo)
    const uint32_t stage_bytes = 2 * w_bytes + 2 * H3_X_BYTES;         // Whi | Wlo | Xhi | Xlo
    const uint32_t bar_base = base + (uint32_t)stages * stage_bytes;
    // barriers: full[s] @ +0, empty[s] @ +64, accum @ +128, tmem slot @ +136
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + (size_t)stages * stage_bytes + 136);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = kCell ? blockIdx.x * 32 : blockIdx.x * bn, m0 = blockIdx.y * H3_BM;   // kCell: n0 = first hidden unit of the tile
    const int kb_begin = blockIdx.z * p.kb_per_split;
    const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
    const int nkb = kb_end - kb_begin;
    const uint32_t tmem_cols = bn <= 128 ? 256u : 512u;   // main chain | cross chain
    const uint32_t cross = tmem_cols >> 1;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(bar_base + 8 * s, 2);        // full: two producer arrivals (weight tiles, activation tiles) + their tx bytes
            mbar_init(bar_base + 64 + 8 * s, 1);   // empty: tcgen05.commit
        }
        mbar_init(bar_base + 128, 1);              // accumulator ready
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    pdl_trigger();   // the next kernel of the stream may start its own prologue

    // ---- producer helpers: the weight tiles of a k-block do not depend on the preceding kernel, the activation tiles do
    int seg_w = 0, seg_x = 0;
    auto issue_w = [&](int i) {
        const int kb = kb_begin + i;
        while (seg_w < p.nseg - 1 && kb >= p.seg_kb_end[seg_w]) ++seg_w;
        const int kc = (kb - (seg_w == 0 ? 0 : p.seg_kb_end[seg_w - 1])) * H3_BK;
        const int s = i % stages;
        const uint32_t st = base + s * stage_bytes, full = bar_base + 8 * s;
        if ((p.dbg & 2) && i >= stages) { mbar_arrive(full); return; }
        mbar_arrive_expect_tx(full, 2 * w_bytes);
        if (kCell) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // 32 rows of each gate (boxes of 32 rows, 4 KB each)
                tma_load_3d(st + q * 4096, &p.tm_wh[seg_w], full, 0, q * p.cell.H + n0, kc / H3_BK);
                tma_load_3d(st + w_bytes + q * 4096, &p.tm_wl[seg_w], full, 0, q * p.cell.H + n0, kc / H3_BK);
            }
        } else {
            tma_load_3d(st, &p.tm_wh[seg_w], full, 0, n0, kc / H3_BK);           // packed weights are k-block-major: one contiguous run
            tma_load_3d(st + w_bytes, &p.tm_wl[seg_w], full, 0, n0, kc / H3_BK);
        }
    };
    // L2 prefetch of the weight tiles of k-block i: the ring holds only ~96 KB of weights per SM, which is less than HBM latency x
    // bandwidth; requests running `pf_dist` k-blocks ahead of the ring keep enough bytes in flight without shared memory
    int seg_p = 0;
    auto prefetch_w = [&](int i) {
        if (i >= nkb) return;
        const int kb = kb_begin + i;
        while (seg_p < p.nseg - 1 && kb >= p.seg_kb_end[seg_p]) ++seg_p;
        const int kbi = kb - (seg_p == 0 ? 0 : p.seg_kb_end[seg_p - 1]);
        if (kCell) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                tma_prefetch_l2_3d(&p.tm_wh[seg_p], 0, q * p.cell.H + n0, kbi);
                tma_prefetch_l2_3d(&p.tm_wl[seg_p], 0, q * p.cell.H + n0, kbi);
            }
        } else {
            tma_prefetch_l2_3d(&p.tm_wh[seg_p], 0, n0, kbi);
            tma_prefetch_l2_3d(&p.tm_wl[seg_p], 0, n0, kbi);
        }
    };
    auto issue_x = [&](int i) {
        const int kb = kb_begin + i;
        while (seg_x < p.nseg - 1 && kb >= p.seg_kb_end[seg_x]) ++seg_x;
        const int kc = (kb - (seg_x == 0 ? 0 : p.seg_kb_end[seg_x - 1])) * H3_BK;
        const int s = i % stages;
        const uint32_t st = base + s * stage_bytes, full = bar_base + 8 * s;
        if ((p.dbg & 1) && i >= stages) { mbar_arrive(full); return; }
        mbar_arrive_expect_tx(full, 2 * H3_X_BYTES);
        tma_load_2d(st + 2 * w_bytes, &p.tm_xh[seg_x], full, kc, m0);
        tma_load_2d(st + 2 * w_bytes + H3_X_BYTES, &p.tm_xl[seg_x], full, kc, m0);
    };
    const bool producer = (warp == 0 && lane == 0);
    const int pre = min(nkb, stages);
    if (producer) {   // fill the weight half of the ring while the predecessor kernel is still running
        for (int sgi = 0; sgi < p.nseg; ++sgi) { prefetch_tensormap(&p.tm_wh[sgi]); prefetch_tensormap(&p.tm_wl[sgi]); }
        for (int i = 0; i < pre; ++i) issue_w(i);
        for (int i = pre; i < pre + p.pf_dist; ++i) prefetch_w(i);
        for (int sgi = 0; sgi < p.nseg; ++sgi) { prefetch_tensormap(&p.tm_xh[sgi]); prefetch_tensormap(&p.tm_xl[sgi]); }
    }
    pdl_wait();      // from here on the activations / flags written by earlier kernels are visible
    trace_released(p.trace);
    // the activation tiles of the first ring round are requested before anything else (the `active` flag below costs an L2 round trip)
    if (producer)
        for (int i = 0; i < pre; ++i) issue_x(i);
    const bool act = (p.active == nullptr) || (*p.active != 0);
    // ---- kCell: operands of the fused LSTM cell (shared by the epilogue role and the reduction below)
    const CellEpilogue& ce = p.cell;
    const int H = ce.H;
    const int nsp = (int)gridDim.z, z = (int)blockIdx.z;   // cluster = the k-splits of this tile; rank == blockIdx.z
    const int nf4 = 8 / nsp;                               // 16-byte chunks (4 units) per gate that this CTA finishes (1, 2, 4 or 8)
    const int rl = (warp & 3) * 32 + lane;
    const int row = m0 + rl;
    constexpr int kMaxPre = 2;                             // chunks whose operands are prefetched (nf4 <= 2 for >= 4 splits)
    float4 pre_c[kMaxPre], pre_a[kMaxPre][4];
    const bool epi_thread = true;   // all four warps finish the tile
    const long long pr = (kCell && epi_thread && row < p.M && ce.parent) ? ce.parent[row] : row;
    auto load_operands = [&](int f, float4& cp, float4 (&ad)[4]) {
        const int j = n0 + 4 * (z * nf4 + f);
        if (!(row < p.M && j < H)) return;
        cp = *reinterpret_cast<const float4*>(ce.c_prev + (size_t)pr * H + j);
        if (ce.addend) {
            const float* a0 = ce.addend + (size_t)(row / ce.add_div) * 4 * H + j;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) ad[qq] = *reinterpret_cast<const float4*>(a0 + (size_t)qq * H);
        } else {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {   // (g + b_ih) + b_hh is the reference's association; biases are summed after the reduction
                ad[qq] = __ldg(reinterpret_cast<const float4*>(ce.b_ih + qq * H + j));
            }
        }
    };
    if (!act) {
        if (producer)   // drain the prefetched tiles before the CTA retires
            for (int i = 0; i < pre; ++i) mbar_wait(bar_base + 8 * i, 0);
    } else {
    if (warp == 0) {
        // ===== TMA producer (then epilogue warp 0) =====
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % stages;
                if (i >= pre) {
                    const uint32_t ph = (uint32_t)(i / stages) & 1u;
                    if (p.pf_dist > 0) prefetch_w(i + p.pf_dist);
                    mbar_wait(bar_base + 64 + 8 * s, ph ^ 1u);
                    issue_w(i);
                    issue_x(i);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (then epilogue warp 1) =====
        // instruction descriptor: D fp32, A/B fp16, both K-major, N=bn, M=128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(H3_BM >> 4) << 24);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % stages;
            const uint32_t ph = (uint32_t)(i / stages) & 1u;
            mbar_wait(bar_base + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                if (!kCell && i == 0) trace_mark(p.trace, 1);   // first stage landed
                if (!kCell && i == nkb - 1) trace_mark(p.trace, 2);   // last stage landed
                const uint32_t st = base + s * stage_bytes;
#pragma unroll
                for (int ks = 0; ks < H3_BK / 16; ++ks) {
                    const uint64_t whi = umma_desc_sw128(st + ks * 32);
                    const uint64_t wlo = umma_desc_sw128(st + w_bytes + ks * 32);
                    const uint64_t xhi = umma_desc_sw128(st + 2 * w_bytes + ks * 32);
                    const uint64_t xlo = umma_desc_sw128(st + 2 * w_bytes + H3_X_BYTES + ks * 32);
                    const uint32_t acc = (i > 0 || ks > 0) ? 1u : 0u;
                    if ((p.dbg & 4) && i > 0) continue;
                    umma_f16_afill(tmem_d, xhi, whi, idesc, acc);
                    umma_f16_alast(tmem_d + cross, xhi, wlo, idesc, acc);
                    umma_f16(tmem_d + cross, xlo, whi, idesc, 1u);
                }
                umma_commit(bar_base + 64 + 8 * s);              // stage free once these MMAs have read it
                if (i == nkb - 1) umma_commit(bar_base + 128);   // accumulator complete
            }
            __syncwarp();
        }
    }
    {
        // ===== epilogue, all four warps: TMEM lanes of this warp = 32 * warp =====
        if (kCell) {   // idle until the accumulator is complete: fetch the cell's reduction-independent operands meanwhile
#pragma unroll
            for (int f = 0; f < kMaxPre; ++f)
                if (f < nf4) load_operands(f, pre_c[f], pre_a[f]);
        }
        mbar_wait(bar_base + 128, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (threadIdx.x == 0) trace_mark(p.trace, 0);   // main loop done
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;            // TMEM lane == output row
        if (kCell) {
            // partial tile -> shared memory, [row][gate q: 32 units] fp32, 16-byte chunks XOR-swizzled by the row so that both this
            // (row-per-thread) write and the column-slice reads of the reduction are conflict-free
            const uint32_t tl = tmem_d + ((uint32_t)(q * 32) << 16);
            const int rl = q * 32 + lane;
            uint32_t ra[32], rb[32];
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                SUBGC_TMEM_LD32(ra, tl + (uint32_t)(c * 32));
                SUBGC_TMEM_LD32(rb, tl + cross + (uint32_t)(c * 32));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t addr = base + (uint32_t)rl * 512u + (uint32_t)(((8 * c + j) ^ (rl & 31)) << 4);
                    const float v0 = fmaf(__uint_as_float(rb[4 * j]), H3_LO_INV, __uint_as_float(ra[4 * j]));
                    const float v1 = fmaf(__uint_as_float(rb[4 * j + 1]), H3_LO_INV, __uint_as_float(ra[4 * j + 1]));
                    const float v2 = fmaf(__uint_as_float(rb[4 * j + 2]), H3_LO_INV, __uint_as_float(ra[4 * j + 2]));
                    const float v3 = fmaf(__uint_as_float(rb[4 * j + 3]), H3_LO_INV, __uint_as_float(ra[4 * j + 3]));
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        } else {
        float* out = p.part + (size_t)blockIdx.z * p.M * p.N + (size_t)row * p.N;
        const bool vec = ((p.N & 3) == 0);
        const uint32_t tlane = tmem_d + ((uint32_t)(q * 32) << 16);
        const int nchunks = min(bn / 32, (p.N - n0 + 31) / 32);
        uint32_t ra[32], rb[32];
        // staging tiles of the TMA-store path live in the pipeline stages (idle by now: every MMA has retired): 2 x 4 KB per warp
        const uint32_t sbuf = base + (uint32_t)q * 8192u;
        const bool tma_out = p.tma_part && (m0 + q * 32 < p.M);   // partial planes, or C itself (direct: epilogue applied on the way)
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            const int col0 = n0 + c * 32;
            SUBGC_TMEM_LD32(ra, tlane + (uint32_t)(c * 32));
            SUBGC_TMEM_LD32(rb, tlane + cross + (uint32_t)(c * 32));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (p.tma_part) {
                if (tma_out) {
                    bool keep = true;   // direct: rows beyond their group's length are written as exact zeros (pack_padded_sequence)
                    if (p.direct && p.epi.group && row < p.M) {
                        const int g = row / p.epi.group, jg = row - g * p.epi.group;
                        keep = jg < p.epi.group_len[p.epi.group_sel ? p.epi.group_sel[g] : g];
                    }
                    const uint32_t buf = sbuf + (uint32_t)(c & 1) * 4096u;
                    if (c >= 2) {  // the store that read this buffer two chunks ago must have drained it
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // 16-byte chunk j of row `lane` sits at chunk j ^ (lane & 7): conflict-free, and what SWIZZLE_128B expects
                        const uint32_t addr = buf + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4);
                        float v0 = fmaf(__uint_as_float(rb[4 * j]), H3_LO_INV, __uint_as_float(ra[4 * j]));
                        float v1 = fmaf(__uint_as_float(rb[4 * j + 1]), H3_LO_INV, __uint_as_float(ra[4 * j + 1]));
                        float v2 = fmaf(__uint_as_float(rb[4 * j + 2]), H3_LO_INV, __uint_as_float(ra[4 * j + 2]));
                        float v3 = fmaf(__uint_as_float(rb[4 * j + 3]), H3_LO_INV, __uint_as_float(ra[4 * j + 3]));
                        if (p.direct) {   // bias / ReLU / zero rows here; the tile leaves as full 128-byte lines instead of 16 bytes per row and store
                            const int cn = col0 + 4 * j, cl = p.N - 1;
                            if (p.epi.bias) {
                                v0 += __ldg(p.epi.bias + min(cn, cl)); v1 += __ldg(p.epi.bias + min(cn + 1, cl));
                                v2 += __ldg(p.epi.bias + min(cn + 2, cl)); v3 += __ldg(p.epi.bias + min(cn + 3, cl));
                            }
                            if (p.epi.bias2) {
                                v0 += __ldg(p.epi.bias2 + min(cn, cl)); v1 += __ldg(p.epi.bias2 + min(cn + 1, cl));
                                v2 += __ldg(p.epi.bias2 + min(cn + 2, cl)); v3 += __ldg(p.epi.bias2 + min(cn + 3, cl));
                            }
                            if (p.epi.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
                            if (!keep) { v0 = 0.f; v1 = 0.f; v2 = 0.f; v3 = 0.f; }
                            if (p.epi.c16_hi && row < p.M && cn + 4 <= p.N) {   // split-fp16 copy for the consumer contraction (8 bytes of hi and of lo)
                                unsigned short hh[4], hl[4];
                                int ovf = 0;
                                split_f16(v0, hh[0], hl[0], ovf); split_f16(v1, hh[1], hl[1], ovf); split_f16(v2, hh[2], hl[2], ovf); split_f16(v3, hh[3], hl[3], ovf);
                                const size_t o16 = (size_t)row * p.epi.ld16 + cn;
                                *reinterpret_cast<uint2*>(p.epi.c16_hi + o16) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
                                *reinterpret_cast<uint2*>(p.epi.c16_lo + o16) = make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16));
                                if (ovf && p.overflow) atomicOr(p.overflow, 1);
                            }
                        }
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&p.tm_part)),
                                     "r"(buf), "r"(col0), "r"(m0 + q * 32), "r"(p.direct ? 0 : (int)blockIdx.z)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else if (row < p.M && p.direct) {
                float* crow = p.C + (size_t)row * p.ldc;
                const bool vecc = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && (col0 + 32 <= p.N);
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
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
                        float tv = fmaf(__uint_as_float(rb[j + u]), H3_LO_INV, __uint_as_float(ra[j + u]));
                        if (p.epi.bias) tv += b1[u];
                        if (p.epi.bias2) tv += b2[u];
                        v[u] = p.epi.relu ? fmaxf(tv, 0.f) : tv;
                    }
                    if (p.epi.c16_hi && col0 + j + 8 <= p.N) {   // split-fp16 copy for the consumer contraction: 8 values = 16 bytes of hi and of lo
                        unsigned short hh[8], hl[8];
                        int ovf = 0;
#pragma unroll
                        for (int u = 0; u < 8; ++u) split_f16(v[u], hh[u], hl[u], ovf);
                        const size_t o16 = (size_t)row * p.epi.ld16 + col0 + j;
                        *reinterpret_cast<uint4*>(p.epi.c16_hi + o16) = make_uint4((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16),
                                                                                   (uint32_t)hh[4] | ((uint32_t)hh[5] << 16), (uint32_t)hh[6] | ((uint32_t)hh[7] << 16));
                        *reinterpret_cast<uint4*>(p.epi.c16_lo + o16) = make_uint4((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16),
                                                                                   (uint32_t)hl[4] | ((uint32_t)hl[5] << 16), (uint32_t)hl[6] | ((uint32_t)hl[7] << 16));
                        if (ovf && p.overflow) atomicOr(p.overflow, 1);
                    }
                    if (vecc) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float4* dst = reinterpret_cast<float4*>(crow + col0 + j + 4 * h);
                            float4 o = make_float4(v[4 * h], v[4 * h + 1], v[4 * h + 2], v[4 * h + 3]);
                            if (p.epi.accumulate) { const float4 tt = *dst; o.x += tt.x; o.y += tt.y; o.z += tt.z; o.w += tt.w; }
                            *dst = o;
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (col0 + j + u < p.N) crow[col0 + j + u] = p.epi.accumulate ? v[u] + crow[col0 + j + u] : v[u];
                    }
                }
            } else if (row < p.M) {
                if (vec && col0 + 32 <= p.N) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v;
                        v.x = fmaf(__uint_as_float(rb[j]), H3_LO_INV, __uint_as_float(ra[j]));
                        v.y = fmaf(__uint_as_float(rb[j + 1]), H3_LO_INV, __uint_as_float(ra[j + 1]));
                        v.z = fmaf(__uint_as_float(rb[j + 2]), H3_LO_INV, __uint_as_float(ra[j + 2]));
                        v.w = fmaf(__uint_as_float(rb[j + 3]), H3_LO_INV, __uint_as_float(ra[j + 3]));
                        *reinterpret_cast<float4*>(out + col0 + j) = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N) out[col0 + j] = fmaf(__uint_as_float(rb[j]), H3_LO_INV, __uint_as_float(ra[j]));
                }
            }
        }
        if (tma_out && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA retires
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
    }
    }
    if (kCell && act) {
        // ---- split-K reduction across the cluster (ranks in ascending order: deterministic) + LSTM cell on this CTA's share of units.
        // The epilogue is a chain of latencies (DSMEM ~0.2 us, global ~1 us), so everything that does not depend on the reduction --
        // previous cell state, hoisted gate term / biases -- is requested before the cluster barrier, and all remote reads of the
        // thread are in flight together.
        if (threadIdx.x == 0) trace_mark(p.trace, 1);   // tile staged
        cluster_sync_all();   // every CTA's partial tile is in its shared memory
        if (threadIdx.x == 0) trace_mark(p.trace, 2);   // cluster barrier passed
        if (epi_thread) {
            for (int f0 = 0; f0 < nf4; f0 += kMaxPre) {
                float4 g[kMaxPre][4];
#pragma unroll
                for (int ff = 0; ff < kMaxPre; ++ff) {
                    const int f = f0 + ff;
                    if (f >= nf4) continue;
                    const int u4 = z * nf4 + f;          // chunk of 4 units inside the tile's 32
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        const uint32_t addr = base + (uint32_t)rl * 512u + (uint32_t)(((8 * qq + u4) ^ (rl & 31)) << 4);
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int sidx = 0; sidx < nsp; ++sidx) {
                            const float4 v = ld_dsmem_f4(addr, (uint32_t)sidx);
                            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                        }
                        g[ff][qq] = acc;
                    }
                }
#pragma unroll
                for (int ff = 0; ff < kMaxPre; ++ff) {
                    const int f = f0 + ff;
                    if (f >= nf4) continue;
                    const int j = n0 + 4 * (z * nf4 + f);   // first hidden unit of the chunk
                    if (!(row < p.M && j < H)) continue;    // H % 4 == 0: a chunk is entirely valid or entirely padding
                    float4 cp, ad[4];
                    if (f0 == 0) {
                        cp = pre_c[ff];
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) ad[qq] = pre_a[ff][qq];
                    } else {
                        load_operands(f, cp, ad);
                    }
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        g[ff][qq].x += ad[qq].x; g[ff][qq].y += ad[qq].y; g[ff][qq].z += ad[qq].z; g[ff][qq].w += ad[qq].w;
                    }
                    if (!ce.addend) {
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            const float4 b2 = __ldg(reinterpret_cast<const float4*>(ce.b_hh + qq * H + j));
                            g[ff][qq].x += b2.x; g[ff][qq].y += b2.y; g[ff][qq].z += b2.z; g[ff][qq].w += b2.w;
                        }
                    }
                    const float gi[4] = {g[ff][0].x, g[ff][0].y, g[ff][0].z, g[ff][0].w}, gf[4] = {g[ff][1].x, g[ff][1].y, g[ff][1].z, g[ff][1].w};
                    const float gg[4] = {g[ff][2].x, g[ff][2].y, g[ff][2].z, g[ff][2].w}, go[4] = {g[ff][3].x, g[ff][3].y, g[ff][3].z, g[ff][3].w};
                    const float cpv[4] = {cp.x, cp.y, cp.z, cp.w};
                    float cn[4], hn[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        cn[e] = sigmoidf_(gf[e]) * cpv[e] + sigmoidf_(gi[e]) * tanhf(gg[e]);
                        hn[e] = sigmoidf_(go[e]) * tanhf(cn[e]);
                    }
                    *reinterpret_cast<float4*>(ce.c_out + (size_t)row * H + j) = make_float4(cn[0], cn[1], cn[2], cn[3]);
                    *reinterpret_cast<float4*>(ce.h_out + (size_t)row * H + j) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                    if (ce.h16_hi) {
                        unsigned short hh[4], hl[4];
                        int ovf = 0;
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_f16(hn[e], hh[e], hl[e], ovf);
                        const size_t o16 = (size_t)row * ce.Hp + j;   // Hp % 8 == 0 and j % 4 == 0: 8-byte aligned
                        *reinterpret_cast<uint2*>(ce.h16_hi + o16) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
                        *reinterpret_cast<uint2*>(ce.h16_lo + o16) = make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16));
                    }
                }
            }
        }
        if (threadIdx.x == 0) trace_mark(p.trace, 3);   // reduction + cell done
        cluster_sync_all();   // nobody retires while a peer may still read its tile
    }
    __syncthreads();
    trace_end(p.trace);
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
    }
}

// ---- fp32 -> (hi, lo) fp16 (split_f16, common.cuh) ------------------------------------------------------------------
// weights: [rows, cols] fp32 (ld ldw) -> hi / lo [kb][rows][32] (k-block-major, segments padded to 32 columns with zeros)
struct PackSegs { int n_seg; int col[SUBGC_PACK_MAX_SEG + 1]; int kb[SUBGC_PACK_MAX_SEG + 1]; };
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int rows, int ldw, const PackSegs sg,
                                                          unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int* __restrict__ overflow) {
    const size_t total = (size_t)sg.kb[sg.n_seg] * rows * H3_BK;
    int ovf = 0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx % H3_BK);
        const size_t t = idx / H3_BK;
        const int r = (int)(t % rows), kb = (int)(t / rows);
        int s = 0;
        while (s < sg.n_seg - 1 && kb >= sg.kb[s + 1]) ++s;
        const int c = sg.col[s] + (kb - sg.kb[s]) * H3_BK + j;
        unsigned short h = 0, l = 0;
        if (c < sg.col[s + 1]) split_f16(w[(size_t)r * ldw + c], h, l, ovf);
        hi[idx] = h;
        lo[idx] = l;
    }
    if (ovf && overflow) atomicOr(overflow, 1);
}
static bool make_pack_segs(int cols, int n_seg, const int32_t* seg_col, PackSegs& sg) {
    if (n_seg < 1 || n_seg > SUBGC_PACK_MAX_SEG || !seg_col || seg_col[0] != 0 || seg_col[n_seg] != cols) return false;
    sg.n_seg = n_seg;
    sg.kb[0] = 0;
    for (int s = 0; s < n_seg; ++s) {
        if (seg_col[s + 1] <= seg_col[s]) return false;
        sg.col[s] = seg_col[s];
        sg.kb[s + 1] = sg.kb[s] + (seg_col[s + 1] - seg_col[s] + H3_BK - 1) / H3_BK;
    }
    sg.col[n_seg] = cols;
    return true;
}

// activations of one K segment (row gather / shared rows / ReLU-on-load applied) -> hi / lo [M, Kp]
__global__ void __launch_bounds__(256) split_rows_kernel(const GemmSeg g, int M, int Kp, unsigned short* __restrict__ hi, unsigned short* __restrict__ lo,
                                                         const int* __restrict__ active, int* __restrict__ overflow) {
    pdl_trigger();
    pdl_wait();
    if (active != nullptr && *active == 0) return;
    const int kq = Kp >> 2;
    const size_t total = (size_t)M * kq;
    int ovf = 0;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx / kq), k = (int)(idx - (size_t)m * kq) << 2;
        const long long src = g.gather ? g.gather[m] : (g.gather32 ? (long long)g.gather32[m] : (long long)(m / g.a_row_div));
        const float* ptr = g.A + src * g.lda + k;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (k + 3 < g.K && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)) {
            const float4 t = *reinterpret_cast<const float4*>(ptr);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k + u < g.K) v[u] = ptr[u];
        }
        unsigned short h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g.relu_a) v[u] = fmaxf(v[u], 0.f);
            split_f16(v[u], h[u], l[u], ovf);
        }
        reinterpret_cast<uint2*>(hi)[idx] = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        reinterpret_cast<uint2*>(lo)[idx] = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
    if (ovf && overflow) atomicOr(overflow, 1);
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D fp16 tensor map: inner dim `cols` (contiguous), outer dim `rows`, `ld` elements between rows; box 32 x box_rows, 64-byte swizzle
// 3-D fp16 map of a packed weight segment [nkb][plane_rows][32]: box 32 x box_rows x 1, 64-byte swizzle; `rows` valid rows from `base`
static bool make_map_w16(CUtensorMap* out, const unsigned short* base, int rows, int plane_rows, int nkb, int box_rows) {
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tc_encode_fn());
    if (!fn) return false;
    cuuint64_t gdim[3] = {(cuuint64_t)H3_BK, (cuuint64_t)rows, (cuuint64_t)nkb};
    cuuint64_t gstride[2] = {(cuuint64_t)H3_BK * 2, (cuuint64_t)plane_rows * H3_BK * 2};
    cuuint32_t box[3] = {(cuuint32_t)H3_BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<unsigned short*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static bool make_map16(CUtensorMap* out, const unsigned short* base, int rows, int cols, long long ld, int box_rows) {
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tc_encode_fn());
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)H3_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<unsigned short*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// 3-D fp32 map of the split-K partials [splits, M, N]: box 32 columns x 32 rows x 1 split, 128-byte swizzle
// ld: floats between consecutive rows (N for the partial planes, ldc for the in-place epilogue writing C itself)
static bool make_map_part(CUtensorMap* out, float* part, int splits, int M, int N, long long ld = 0) {
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(tc_encode_fn());
    if (!fn) return false;
    if (ld <= 0) ld = N;
    cuuint64_t gdim[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)splits};
    cuuint64_t gstride[2] = {(cuuint64_t)ld * 4, (cuuint64_t)M * (cuuint64_t)ld * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, part, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

static int h3_mode() {  // SUBGC_H3=0 ignores packed weights (A/B runs against the split-TF32 path)
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("SUBGC_H3");
        mode = (e && e[0] == '0') ? 0 : 1;
    }
    return mode;
}

bool h3_eligible(const GemmProblem& p) {
    if (!h3_mode()) return false;
    if (p.M < 1 || p.N < 64 || p.nseg < 1 || p.nseg > H3_MAX_SEG) return false;
    long long ktot = 0;
    for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg& g = p.seg[s];
        if (!g.W16_hi || !g.W16_lo || g.w16_rows < p.N || (reinterpret_cast<uintptr_t>(g.W16_hi) & 15) || (reinterpret_cast<uintptr_t>(g.W16_lo) & 15))
            return false;
        if (g.A16_hi && ((g.lda16 & 7) || (reinterpret_cast<uintptr_t>(g.A16_hi) & 15) || (reinterpret_cast<uintptr_t>(g.A16_lo) & 15) || !g.A16_lo))
            return false;
        ktot += g.K;
    }
    return ktot >= 64;
}

void resolve_packs(GemmProblem& p, const subgc_weights* w) {
    if (!w || !w->packs || w->n_packs <= 0 || !h3_mode()) return;
    p.overflow = w->h3_overflow;
    for (int s = 0; s < p.nseg; ++s) {
        GemmSeg& g = p.seg[s];
        if (g.W16_hi) continue;
        for (int i = 0; i < w->n_packs; ++i) {
            const subgc_packed& pk = w->packs[i];
            if (!pk.w || g.W < pk.w || g.W >= pk.w + (size_t)pk.rows * pk.cols || g.ldw != pk.cols) continue;
            const size_t off = (size_t)(g.W - pk.w);
            const size_t r = off / pk.cols, c = off - r * pk.cols;
            int kb0 = 0, found = -1;   // the operand must be exactly one packed K segment (from its first column, over its full width)
            for (int q = 0; q < pk.n_seg; ++q) {
                if ((size_t)pk.seg_col[q] == c && pk.seg_col[q + 1] - pk.seg_col[q] == g.K) { found = q; break; }
                kb0 += (pk.seg_col[q + 1] - pk.seg_col[q] + H3_BK - 1) / H3_BK;
            }
            if (found < 0 || r + (size_t)p.N > (size_t)pk.rows) break;
            g.W16_hi = pk.hi + ((size_t)kb0 * pk.rows + r) * H3_BK;
            g.W16_lo = pk.lo + ((size_t)kb0 * pk.rows + r) * H3_BK;
            g.w16_rows = pk.rows;
            break;
        }
    }
}

bool h3_presplit(const float* A, int M, int K, int lda, const subgc_weights* w, Workspace& ws, cudaStream_t stream, const unsigned short** hi,
                 const unsigned short** lo, int* ld16) {
    if (!w || !w->packs || w->n_packs <= 0 || !h3_mode() || M <= 0 || K <= 0) return false;
    const int Kp = (K + 7) & ~7;
    if (ws.remaining() < 2 * align_up((size_t)M * Kp * 2, 256) + 512) return false;
    unsigned short* th = ws.take<unsigned short>((size_t)M * Kp);
    unsigned short* tl = ws.take<unsigned short>((size_t)M * Kp);
    if (!ws.ok()) return false;
    GemmSeg g = make_seg(A, lda, nullptr, 0, K);
    const size_t quads = (size_t)M * (Kp >> 2);
    int gb = (int)((quads + 255) / 256);
    if (gb > kNumSMs * 8) gb = kNumSMs * 8;
    if (launch_pdl(split_rows_kernel, dim3(gb), dim3(256), (size_t)0, stream, g, M, Kp, th, tl, (const int*)nullptr, w->h3_overflow) != cudaSuccess) return false;
    count_launch();
    *hi = th; *lo = tl; *ld16 = Kp;
    return true;
}

struct H3Plan { int m_tiles, n_tiles, kb_total, splits, kb_per_split, bn, stages; };

// Tile width and split count.  One CTA streams bn weight rows over its k-range; narrow tiles leave more n-tiles, i.e. fewer
// k-splits and proportionally less split-K partial traffic (the consumers' read volume), but re-read the activation tile more
// often.  A small cost model (cycles per CTA: HBM share vs shared-memory traffic per k-block, plus the partial-tile epilogue,
// times the number of waves) picks among {256, 224, 192, 160, 128}.
static H3Plan h3_plan(int M, int N, const int* segK, int nseg) {
    static const int forced_bn = getenv("SUBGC_H3_BN") ? atoi(getenv("SUBGC_H3_BN")) : 0;
    H3Plan best{};
    double best_cost = 1e30;
    int kb_total = 0;
    for (int s = 0; s < nseg; ++s) kb_total += (segK[s] + H3_BK - 1) / H3_BK;
    const int cands[5] = {256, 224, 192, 160, 128};
    for (int ci = 0; ci < 5; ++ci) {
        const int bn = cands[ci];
        if (forced_bn && bn != forced_bn) continue;
        H3Plan pl;
        pl.bn = bn;
        pl.m_tiles = (M + H3_BM - 1) / H3_BM;
        pl.n_tiles = (N + bn - 1) / bn;
        pl.kb_total = kb_total;
        const int tiles = pl.m_tiles * pl.n_tiles;
        int splits = kNumSMs / tiles;
        const int by_k = kb_total / 2;  // at least 2 k-blocks (128 columns) per split
        if (splits > by_k) splits = by_k;
        if (splits > 32) splits = 32;
        if (splits < 1) splits = 1;
        pl.kb_per_split = (kb_total + splits - 1) / splits;
        if (pl.kb_per_split > H3_MAX_CHAIN) pl.kb_per_split = H3_MAX_CHAIN;
        pl.splits = (kb_total + pl.kb_per_split - 1) / pl.kb_per_split;
        const int stage_bytes = 2 * bn * H3_BK * 2 + 2 * H3_X_BYTES;
        pl.stages = H3_SMEM_BUDGET / stage_bytes;
        if (pl.stages > H3_MAX_STAGES) pl.stages = H3_MAX_STAGES;
        const long long ctas = (long long)tiles * pl.splits;
        const long long waves = (ctas + kNumSMs - 1) / kNumSMs;
        const double per_wave = (double)(ctas < kNumSMs ? ctas : kNumSMs);
        const double hbm = (2.0 * bn * 128) / (3400.0 / per_wave < 64.0 ? 3400.0 / per_wave : 64.0);       // bytes / (bytes per cycle per CTA)
        const double smem = ((2.0 * bn * 128 + 2.0 * H3_X_BYTES) + 3.0 * (bn * 128 + H3_X_BYTES) - H3_X_BYTES) / 128.0;
        const double mma = 12.0 * bn / 2.0;
        double per_kb = hbm > smem ? hbm : smem;
        if (mma > per_kb) per_kb = mma;
        // a ring stage turns around in (TMA latency + its MMAs + barrier hand-offs); with few stages this, not bandwidth, bounds the loop
        // (measured: 3 stages of 64 KB -> ~0.75 us per k-block whatever is switched off)
        const double turn = (2300.0 + mma + 600.0) / pl.stages;
        if (turn > per_kb) per_kb = turn;
        static const double epi_w = getenv("SUBGC_H3_EPIW") ? atof(getenv("SUBGC_H3_EPIW")) : 1.0;
        const double epi = (pl.splits > 1 ? 1.0 + epi_w * pl.splits / 2.0 : 1.0) * bn * 128.0 * 4.0 / 64.0;   // partial tile out + the consumer's read of all splits
        const double cost = waves * (pl.kb_per_split * per_kb + epi + 4000.0);
        if (cost < best_cost) { best_cost = cost; best = pl; }
    }
    return best;
}

void launch_splitk_reduce(const GemmProblem& p, const float* part, int splits, cudaStream_t stream, bool write_c16 = false);

// Tensor maps of every K segment (weights: boxes of box_w rows; activations: pre-split copies or split here into `ws`)
static int h3_segments(const GemmProblem& p, int box_w, Workspace& ws, cudaStream_t stream, H3Params& hp) {
    int kb = 0;
    for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg& g = p.seg[s];
        const int nkb_seg = (g.K + H3_BK - 1) / H3_BK;
        kb += nkb_seg;
        hp.seg_kb_end[s] = kb;
        if (!make_map_w16(&hp.tm_wh[s], g.W16_hi, p.N, g.w16_rows, nkb_seg, box_w) || !make_map_w16(&hp.tm_wl[s], g.W16_lo, p.N, g.w16_rows, nkb_seg, box_w)) {
            set_error("gemm(h3): cuTensorMapEncodeTiled failed for weight segment %d", s);
            return SUBGC_E_CUDA;
        }
        const unsigned short *xh = g.A16_hi, *xl = g.A16_lo;
        long long xld = g.lda16;
        if (!xh) {  // no pre-split copy from the producer: split (and gather / ReLU) this segment here, [M, K rounded up to 8]
            const int Kp = (g.K + 7) & ~7;
            unsigned short* th = ws.take<unsigned short>((size_t)p.M * Kp);
            unsigned short* tl = ws.take<unsigned short>((size_t)p.M * Kp);
            if (!ws.ok()) {
                set_error("gemm(h3): workspace too small");
                return SUBGC_E_WORKSPACE;
            }
            const size_t quads = (size_t)p.M * (Kp >> 2);
            int gb = (int)((quads + 255) / 256);
            if (gb > kNumSMs * 8) gb = kNumSMs * 8;
            SUBGC_CUDA(launch_pdl(split_rows_kernel, dim3(gb), dim3(256), (size_t)0, stream, g, p.M, Kp, th, tl, p.active, p.overflow));
            SUBGC_LAUNCH_CHECK();
            xh = th; xl = tl; xld = Kp;
        }
        if (!make_map16(&hp.tm_xh[s], xh, p.M, g.K, xld, H3_BM) || !make_map16(&hp.tm_xl[s], xl, p.M, g.K, xld, H3_BM)) {
            set_error("gemm(h3): cuTensorMapEncodeTiled failed for activation segment %d", s);
            return SUBGC_E_CUDA;
        }
    }
    for (int s = p.nseg; s < H3_MAX_SEG; ++s) {
        hp.seg_kb_end[s] = kb; hp.tm_wh[s] = hp.tm_wh[0]; hp.tm_wl[s] = hp.tm_wl[0]; hp.tm_xh[s] = hp.tm_xh[0]; hp.tm_xl[s] = hp.tm_xl[0];
    }
    hp.nseg = p.nseg; hp.M = p.M; hp.N = p.N; hp.kb_total = kb; hp.active = p.active;
    return SUBGC_OK;
}

static int h3_set_smem_attr() {
    static DeviceOnce once;
    SUBGC_CUDA(once.run([]() -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(h3_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM_BUDGET + H3_SMEM_EXTRA);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(h3_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_SMEM_BUDGET + H3_SMEM_EXTRA);
        return e;
    }));
    return SUBGC_OK;
}

// Workspace: within tc_workspace_bytes(M, N, Ktotal) (split activations take M * Kp * 4 bytes per segment like the fp32 copies there,
// and the split count is never larger because a k-block covers four times the columns).
int launch_split_rows(const float* A, int M, int K, int lda, unsigned short* hi, unsigned short* lo, int ld16, int* overflow, cudaStream_t stream) {
    GemmSeg g = make_seg(A, lda, nullptr, 0, K);
    const size_t quads = (size_t)M * (ld16 >> 2);
    int gb = (int)((quads + 255) / 256);
    if (gb > kNumSMs * 8) gb = kNumSMs * 8;
    if (gb < 1) gb = 1;
    SUBGC_CUDA(launch_pdl(split_rows_kernel, dim3(gb), dim3(256), (size_t)0, stream, g, M, ld16, hi, lo, (const int*)nullptr, overflow));
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

int launch_gemm_h3(const GemmProblem& p, void* ws_, size_t ws_bytes, cudaStream_t stream, RawPartials* raw, bool* wrote_c16) {
    int segK[H3_MAX_SEG];
    for (int s = 0; s < p.nseg; ++s) segK[s] = p.seg[s].K;
    const H3Plan pl = h3_plan(p.M, p.N, segK, p.nseg);
    Workspace ws(ws_, ws_bytes);
    // in-place epilogue of single-split problems; zero-padded row groups need the TMA-store variant (decided below)
    const bool c_tma_ok = (p.N & 3) == 0 && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && !p.epi.accumulate &&
                          getenv("SUBGC_H3_NO_TMA_STORE") == nullptr && getenv("SUBGC_H3_NO_TMA_C") == nullptr;
    const bool direct = (pl.splits == 1 && raw == nullptr && p.epi.div == 0.f && p.epi.addend == nullptr && (p.epi.group == 0 || c_tma_ok));
    float* part = direct ? nullptr : ws.take<float>((size_t)pl.splits * p.M * p.N);
    if (!ws.ok()) {
        set_error("gemm(h3): workspace too small (%zu bytes given)", ws_bytes);
        return SUBGC_E_WORKSPACE;
    }
    H3Params hp;
    SUBGC_TRY(h3_segments(p, pl.bn, ws, stream, hp));
    hp.bn = pl.bn; hp.stages = pl.stages; hp.kb_per_split = pl.kb_per_split; hp.part = part;
    hp.direct = direct ? 1 : 0; hp.epi = p.epi; hp.C = p.C; hp.ldc = p.ldc; hp.overflow = p.overflow;
    // the in-kernel split copy needs whole 8-column groups and 16-byte aligned rows; anything else is split by the caller afterwards
    const bool c16_here = direct && p.epi.c16_hi && p.epi.c16_lo && (p.N & 7) == 0 && (p.epi.ld16 & 7) == 0 && ((p.ldc & 3) == 0) &&
                          (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && !p.epi.accumulate &&
                          (reinterpret_cast<uintptr_t>(p.epi.c16_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.epi.c16_lo) & 15) == 0;
    if (!c16_here) { hp.epi.c16_hi = nullptr; hp.epi.c16_lo = nullptr; }
    if (wrote_c16) *wrote_c16 = c16_here;
    static const bool tma_store = getenv("SUBGC_H3_NO_TMA_STORE") == nullptr;
    hp.tma_part = 0;
    hp.tm_part = hp.tm_wh[0];
    if (!direct && tma_store && (p.N & 3) == 0 && make_map_part(&hp.tm_part, part, pl.splits, p.M, p.N)) hp.tma_part = 1;
    if (direct && c_tma_ok && make_map_part(&hp.tm_part, p.C, 1, p.M, p.N, p.ldc)) hp.tma_part = 1;   // C leaves through shared memory + TMA stores
    if (direct && p.epi.group && !hp.tma_part) { set_error("gemm(h3): tensor map of C could not be encoded"); return SUBGC_E_CUDA; }
    SUBGC_TRY(h3_set_smem_attr());
    hp.trace = next_trace_slot(1);
    static const int dbg = getenv("SUBGC_H3_DBG") ? atoi(getenv("SUBGC_H3_DBG")) : 0;
    hp.dbg = dbg;
    static const int pf = getenv("SUBGC_H3_PF") ? atoi(getenv("SUBGC_H3_PF")) : 0;   // measured: no gain (the loop is bound by stage turn-around, not by HBM latency), costs TMA issue slots
    hp.pf_dist = pf;
    const size_t smem_bytes = (size_t)pl.stages * (2 * pl.bn * H3_BK * 2 + 2 * H3_X_BYTES) + H3_SMEM_EXTRA;
    dim3 grid(pl.n_tiles, pl.m_tiles, pl.splits);
    SUBGC_CUDA(launch_pdl(h3_gemm_kernel<false>, grid, dim3(H3_THREADS), smem_bytes, stream, hp));
    SUBGC_LAUNCH_CHECK();
    if (direct) return SUBGC_OK;
    if (raw) {
        raw->part = part;
        raw->splits = pl.splits;
        return SUBGC_OK;
    }
    // the reduction pass writes the split-fp16 copy of the result itself when its vector path applies
    const bool red16 = p.epi.c16_hi && p.epi.c16_lo && (p.N & 3) == 0 && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 &&
                       (p.epi.ld16 & 3) == 0 && (reinterpret_cast<uintptr_t>(p.epi.c16_hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(p.epi.c16_lo) & 7) == 0;
    launch_splitk_reduce(p, part, pl.splits, stream, red16);
    SUBGC_LAUNCH_CHECK();
    if (wrote_c16) *wrote_c16 = red16;
    return SUBGC_OK;
}

// gates -> LSTM cell inside the contraction (CellEpilogue): gate-grouped tiles of 32 hidden units, the k-splits of a tile are one
// thread-block cluster (1, 1, splits), splits in {8, 4, 2, 1}
int launch_gemm_cell(const GemmProblem& p0, const CellEpilogue& cell, void* ws_, size_t ws_bytes, cudaStream_t stream, bool* fused) {
    *fused = false;
    // opt-in (SUBGC_FUSED_CELL=1): correct, but the DSMEM reduction + cell on 128 threads per SM measured ~3 us slower per LSTM than
    // split-K partials + the (all-resident, PDL-released) cell kernel -- see DESIGN.md
    static const bool off = !(getenv("SUBGC_FUSED_CELL") != nullptr && getenv("SUBGC_FUSED_CELL")[0] == '1');
    int clus = 0;
    if (!off) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&clus, cudaDevAttrClusterLaunch, dev);
    }
    GemmProblem p = p0;
    if (p.wts) { resolve_packs(p, p.wts); p.wts = nullptr; }
    const int H = cell.H;
    if (off || !clus || H <= 0 || (H & 3) || p.N != 4 * H || !h3_eligible(p)) return SUBGC_OK;
    if (cell.h16_hi && (cell.Hp & 7)) return SUBGC_OK;
    for (int s = 0; s < p.nseg; ++s)
        if (p.seg[s].w16_rows < 4 * H) return SUBGC_OK;
    int kb_total = 0;
    for (int s = 0; s < p.nseg; ++s) kb_total += (p.seg[s].K + H3_BK - 1) / H3_BK;
    const int n_tiles = (H + 31) / 32, m_tiles = (p.M + H3_BM - 1) / H3_BM;
    const long long tiles = (long long)n_tiles * m_tiles;
    auto valid = [&](int sp) {   // every split non-empty, at least 2 k-blocks each (unless unsplit), TMEM chain within its limit
        const int k = (kb_total + sp - 1) / sp;
        return (sp == 1 || k >= 2) && (kb_total + k - 1) / k == sp && k <= H3_MAX_CHAIN;
    };
    int splits = 0;
    const int desc[4] = {8, 4, 2, 1};
    for (int i = 0; i < 4 && !splits; ++i)
        if (valid(desc[i]) && tiles * desc[i] <= kNumSMs) splits = desc[i];      // one wave: as many splits as fit
    for (int i = 3; i >= 0 && !splits; --i)
        if (valid(desc[i])) splits = desc[i];                                    // several waves: as few splits as the chain limit allows
    if (!splits) return SUBGC_OK;
    const int kps = (kb_total + splits - 1) / splits;
    Workspace ws(ws_, ws_bytes);
    H3Params hp;
    SUBGC_TRY(h3_segments(p, 32, ws, stream, hp));
    hp.bn = 128; hp.kb_per_split = kps; hp.part = nullptr; hp.direct = 0; hp.epi = GemmEpilogue(); hp.C = nullptr; hp.ldc = 0; hp.overflow = nullptr;
    hp.tma_part = 0; hp.tm_part = hp.tm_wh[0];
    const int stage_bytes = 2 * 128 * H3_BK * 2 + 2 * H3_X_BYTES;
    hp.stages = H3_SMEM_BUDGET / stage_bytes;
    hp.cell = cell;
    SUBGC_TRY(h3_set_smem_attr());
    hp.trace = next_trace_slot(6);
    hp.dbg = 0;
    hp.pf_dist = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tiles, m_tiles, splits);
    cfg.blockDim = dim3(H3_THREADS);
    cfg.dynamicSmemBytes = (size_t)hp.stages * stage_bytes + H3_SMEM_EXTRA;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = splits;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    SUBGC_CUDA(cudaLaunchKernelEx(&cfg, h3_gemm_kernel<true>, hp));
    SUBGC_LAUNCH_CHECK();
    *fused = true;
    return SUBGC_OK;
}

}  // namespace subgc

using namespace subgc;

extern "C" size_t subgc_pack_elems(int rows, int n_seg, const int32_t* seg_col) {
    PackSegs sg;
    if (rows <= 0 || !seg_col || n_seg < 1 || n_seg > SUBGC_PACK_MAX_SEG || !make_pack_segs(seg_col[n_seg], n_seg, seg_col, sg)) return 0;
    return (size_t)sg.kb[n_seg] * rows * H3_BK;
}

extern "C" int subgc_pack_weight(int rows, int cols, const float* w, int ldw, int n_seg, const int32_t* seg_col, uint16_t* hi, uint16_t* lo,
                                 int32_t* overflow, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(rows > 0 && cols > 0 && w && hi && lo && ldw >= cols, "subgc_pack_weight: bad arguments");
    PackSegs sg;
    SUBGC_CHECK_ARG(make_pack_segs(cols, n_seg, seg_col, sg), "subgc_pack_weight: bad column segments");
    const size_t total = (size_t)sg.kb[n_seg] * rows * H3_BK;
    int blocks = (int)((total + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    pack_weight_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, rows, ldw, sg, hi, lo, overflow);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

extern "C" int subgc_linear_packed_forward(int M, int N, int K, const float* A, int lda, const int64_t* a_gather, const subgc_packed* pk,
                                           const float* bias, int relu, float* C, int ldc, void* ws, size_t ws_bytes, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(A && pk && pk->w && pk->hi && pk->lo && C && M >= 0 && N > 0 && K > 0, "subgc_linear_packed_forward: bad arguments");
    SUBGC_CHECK_ARG(pk->rows == N && pk->cols == K && pk->n_seg == 1, "subgc_linear_packed_forward: pack does not match [N, K] (one segment)");
    GemmProblem p;
    p.M = M; p.N = N; p.nseg = 1;
    p.seg[0] = make_seg(A, lda, pk->w, K, K);
    p.seg[0].gather = reinterpret_cast<const long long*>(a_gather);
    p.seg[0].W16_hi = pk->hi; p.seg[0].W16_lo = pk->lo; p.seg[0].w16_rows = pk->rows;
    p.epi.bias = bias;
    p.epi.relu = relu;
    p.C = C; p.ldc = ldc;
    return launch_gemm(p, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
