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
//               W[256 x 16] straight from the reference's un-repacked nn.Linear / LSTMCell tensor, plus the
//               pre-split activation tiles Xhi / Xlo [128 x 16]            (cp.async.bulk.tensor, mbarrier tx)
//   warps 2-5   transform: split the weight tile in shared memory (hi in place, lo to a second buffer),
//               fence.proxy.async, signal the MMA warp; after the main loop they are the epilogue warps
//   warp 1      one elected thread issues 3 x 2 tcgen05.mma.kind::tf32 (M=128, N=256, K=8) per k-block into a
//               256-column TMEM accumulator, tcgen05.commit frees the stage / publishes the accumulator
//   epilogue    tcgen05.ld 32x32b.x32 -> registers -> per-warp smem transpose -> coalesced 128-byte row stores of
//               the split-K partial [z][M][N]; the reduction + epilogue kernel is shared with the SIMT block.
// Activations are packed once per GEMM by pack_split_kernel (gather / shared rows / ReLU-on-load / concat of the K
// segments, zero padding to 16 columns per segment and 128 rows) so that one 2-D tensor map describes them.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace subgc {

constexpr int TC_BM = 128;
constexpr int TC_BN = 256;
constexpr int TC_BK = 16;
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 192;
constexpr int TC_W_BYTES = TC_BN * TC_BK * 4;                     // 16 KB
constexpr int TC_X_BYTES = TC_BM * TC_BK * 4;                     // 8 KB
constexpr int TC_STAGE_BYTES = 2 * TC_W_BYTES + 2 * TC_X_BYTES;   // Whi | Wlo | Xhi | Xlo = 48 KB
constexpr int TC_TX_BYTES = TC_W_BYTES + 2 * TC_X_BYTES;          // bytes TMA lands per stage
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TC_MAX_SEG = 4;

struct TcParams {
    CUtensorMap tm_xhi, tm_xlo;
    CUtensorMap tm_w[TC_MAX_SEG];
    int seg_kb_end[TC_MAX_SEG];  // cumulative k-block count at the end of each segment
    int nseg;
    int M, N;
    int kb_total, kb_per_split;
    float* part;                 // [splits][M][N]
    const int* active;
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
// K-major, 64-byte swizzle, rows of 64 bytes: 8-row groups are 512 bytes apart (SBO), LBO unused for swizzled K-major
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                          // leading byte offset (16-byte units), bits [16,30)
    d |= (uint64_t)(512 >> 4) << 32;                 // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)4 << 61;                          // layout type SWIZZLE_64B, bits [61,64)
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

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
    if (warp == 0) {  // TMEM: 256 fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int seg = 0;
            while (seg < p.nseg - 1 && kb_begin >= p.seg_kb_end[seg]) ++seg;
            for (int i = 0; i < nkb; ++i) {
                const int kb = kb_begin + i;
                while (seg < p.nseg - 1 && kb >= p.seg_kb_end[seg]) ++seg;
                const int seg_kb0 = seg == 0 ? 0 : p.seg_kb_end[seg - 1];
                const int s = i % TC_STAGES;
                const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
                mbar_wait(bar_base + 64 + 8 * s, ph ^ 1u);
                const uint32_t st = base + s * TC_STAGE_BYTES;
                const uint32_t full = bar_base + 8 * s;
                mbar_arrive_expect_tx(full, TC_TX_BYTES);
                tma_load_2d(st, &p.tm_w[seg], full, (kb - seg_kb0) * TC_BK, n0);
                tma_load_2d(st + 2 * TC_W_BYTES, &p.tm_xhi, full, kb * TC_BK, m0);
                tma_load_2d(st + 2 * TC_W_BYTES + TC_X_BYTES, &p.tm_xlo, full, kb * TC_BK, m0);
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
                const uint32_t st = base + s * TC_STAGE_BYTES;
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    const uint64_t whi = umma_desc_sw64(st + ks * 32);
                    const uint64_t wlo = umma_desc_sw64(st + TC_W_BYTES + ks * 32);
                    const uint64_t xhi = umma_desc_sw64(st + 2 * TC_W_BYTES + ks * 32);
                    const uint64_t xlo = umma_desc_sw64(st + 2 * TC_W_BYTES + TC_X_BYTES + ks * 32);
                    umma_tf32(tmem_d, xhi, whi, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                    umma_tf32(tmem_d, xlo, whi, idesc, 1u);
                    umma_tf32(tmem_d, xhi, wlo, idesc, 1u);
                }
                umma_commit(bar_base + 64 + 8 * s);              // stage free once these MMAs have read it
                if (i == nkb - 1) umma_commit(bar_base + 96);    // accumulator complete
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
            float4* whi = reinterpret_cast<float4*>(gbase + s * TC_STAGE_BYTES);
            float4* wlo = reinterpret_cast<float4*>(gbase + s * TC_STAGE_BYTES + TC_W_BYTES);
#pragma unroll
            for (int j = 0; j < TC_W_BYTES / 16 / 128; ++j) {
                const int idx = t + 128 * j;
                float4 v = whi[idx];
                float4 h, l;
                h.x = __uint_as_float(to_tf32(v.x)); h.y = __uint_as_float(to_tf32(v.y));
                h.z = __uint_as_float(to_tf32(v.z)); h.w = __uint_as_float(to_tf32(v.w));
                l.x = __uint_as_float(to_tf32(v.x - h.x)); l.y = __uint_as_float(to_tf32(v.y - h.y));
                l.z = __uint_as_float(to_tf32(v.z - h.z)); l.w = __uint_as_float(to_tf32(v.w - h.w));
                whi[idx] = h;
                wlo[idx] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_base + 32 + 8 * s);
        }
        // epilogue: TMEM lanes of this warp = 32 * (warp % 4)
        mbar_wait(bar_base + 96, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3;
        float* scratch = reinterpret_cast<float*>(gbase) + (warp - 2) * (32 * 33);  // stage memory is idle now
        float* out = p.part + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll 1
        for (int c = 0; c < TC_BN / 32; ++c) {
            if (n0 + c * 32 >= p.N) break;
            uint32_t r[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = __uint_as_float(r[j]);
            __syncwarp();
            const int col = n0 + c * 32 + lane;
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
                const int row = m0 + q * 32 + rr;
                if (row < p.M && col < p.N) out[(size_t)row * p.N + col] = scratch[rr * 33 + lane];
            }
            __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(256) : "memory");
    }
}

// ---- activation packing: concat segments, gather, ReLU-on-load, zero pad, hi/lo split ----------------------------------
struct PackArgs {
    GemmSeg seg[TC_MAX_SEG];
    int seg_col0[TC_MAX_SEG + 1];  // first packed column of each segment (multiples of 16); [nseg] = Kpad
    int nseg, M, Mpad, Kpad;
    float* xhi;
    float* xlo;
    const int* active;
};

__global__ void __launch_bounds__(256) pack_split_kernel(const PackArgs a) {
    if (a.active != nullptr && *a.active == 0) return;
    const int kq = a.Kpad >> 2;
    const size_t total = (size_t)a.Mpad * kq;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(idx / kq), c = (int)(idx - (size_t)m * kq) << 2;
        int s = 0;
        while (s < a.nseg - 1 && c >= a.seg_col0[s + 1]) ++s;
        const GemmSeg& g = a.seg[s];
        const int k = c - a.seg_col0[s];
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (m < a.M && k < g.K) {
            long long src = g.gather ? g.gather[m] : (g.gather32 ? (long long)g.gather32[m] : (long long)(m / g.a_row_div));
            const float* ptr = g.A + src * g.lda + k;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k + j < g.K) v[j] = g.relu_a ? fmaxf(__ldg(ptr + j), 0.f) : __ldg(ptr + j);
        }
        float4 h, l;
        h.x = __uint_as_float(to_tf32(v[0])); h.y = __uint_as_float(to_tf32(v[1]));
        h.z = __uint_as_float(to_tf32(v[2])); h.w = __uint_as_float(to_tf32(v[3]));
        l.x = __uint_as_float(to_tf32(v[0] - h.x)); l.y = __uint_as_float(to_tf32(v[1] - h.y));
        l.z = __uint_as_float(to_tf32(v[2] - h.z)); l.w = __uint_as_float(to_tf32(v[3] - h.w));
        reinterpret_cast<float4*>(a.xhi)[idx] = h;
        reinterpret_cast<float4*>(a.xlo)[idx] = l;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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
static bool make_map(CUtensorMap* out, const float* base, int rows, int cols, long long ld, int box_rows) {
    struct Key {
        const void* base; int rows, cols; long long ld; int box;
        bool operator==(const Key& o) const { return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box == o.box; }
    };
    struct Hash {
        size_t operator()(const Key& k) const {
            size_t h = reinterpret_cast<size_t>(k.base);
            h = h * 1000003u ^ (size_t)k.rows; h = h * 1000003u ^ (size_t)k.cols; h = h * 1000003u ^ (size_t)k.ld; h = h * 1000003u ^ (size_t)k.box;
            return h;
        }
    };
    static thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;  // encoding is pure: same key -> same descriptor
    Key key{base, rows, cols, ld, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return true; }
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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
    const size_t Kpad = (size_t)Ktotal + TC_MAX_SEG * TC_BK;
    return 2 * align_up((size_t)m_tiles * TC_BM * Kpad * 4, 256) + align_up((size_t)splits * M * N * 4, 256) + 512;
}

void launch_splitk_reduce(const GemmProblem& p, const float* part, int splits, cudaStream_t stream);

int launch_gemm_tc(const GemmProblem& p, void* ws_, size_t ws_bytes, cudaStream_t stream) {
    int segK[TC_MAX_SEG];
    for (int s = 0; s < p.nseg; ++s) segK[s] = p.seg[s].K;
    const TcPlan pl = tc_plan(p.M, p.N, segK, p.nseg);
    Workspace ws(ws_, ws_bytes);
    float* xhi = ws.take<float>((size_t)pl.Mpad * pl.Kpad);
    float* xlo = ws.take<float>((size_t)pl.Mpad * pl.Kpad);
    float* part = ws.take<float>((size_t)pl.splits * p.M * p.N);
    if (!ws.ok()) {
        set_error("gemm(tc): workspace too small (%zu bytes given)", ws_bytes);
        return SUBGC_E_WORKSPACE;
    }
    PackArgs pa;
    pa.nseg = p.nseg; pa.M = p.M; pa.Mpad = pl.Mpad; pa.Kpad = pl.Kpad; pa.xhi = xhi; pa.xlo = xlo; pa.active = p.active;
    TcParams tp;
    int col = 0, kb = 0;
    for (int s = 0; s < p.nseg; ++s) {
        pa.seg[s] = p.seg[s];
        pa.seg_col0[s] = col;
        const int nkb = (p.seg[s].K + TC_BK - 1) / TC_BK;
        col += nkb * TC_BK;
        kb += nkb;
        tp.seg_kb_end[s] = kb;
        if (!make_map(&tp.tm_w[s], p.seg[s].W, p.N, p.seg[s].K, p.seg[s].ldw, TC_BN)) {
            set_error("gemm(tc): cuTensorMapEncodeTiled failed for weight segment %d", s);
            return SUBGC_E_CUDA;
        }
    }
    for (int s = p.nseg; s < TC_MAX_SEG; ++s) { tp.seg_kb_end[s] = kb; tp.tm_w[s] = tp.tm_w[0]; }
    pa.seg_col0[p.nseg] = col;
    if (!make_map(&tp.tm_xhi, xhi, pl.Mpad, pl.Kpad, pl.Kpad, TC_BM) || !make_map(&tp.tm_xlo, xlo, pl.Mpad, pl.Kpad, pl.Kpad, TC_BM)) {
        set_error("gemm(tc): cuTensorMapEncodeTiled failed for the activation tiles");
        return SUBGC_E_CUDA;
    }
    tp.nseg = p.nseg; tp.M = p.M; tp.N = p.N; tp.kb_total = pl.kb_total; tp.kb_per_split = pl.kb_per_split; tp.part = part; tp.active = p.active;
    size_t quads = (size_t)pl.Mpad * (pl.Kpad >> 2);
    int pblocks = (int)((quads + 255) / 256);
    if (pblocks > kNumSMs * 8) pblocks = kNumSMs * 8;
    pack_split_kernel<<<pblocks, 256, 0, stream>>>(pa);
    SUBGC_LAUNCH_CHECK();
    static bool attr_set = false;
    if (!attr_set) {
        SUBGC_CUDA(cudaFuncSetAttribute(umma_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(pl.n_tiles, pl.m_tiles, pl.splits);
    umma_gemm_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(tp);
    SUBGC_LAUNCH_CHECK();
    launch_splitk_reduce(p, part, pl.splits, stream);
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

}  // namespace subgc
