// fp32 "TN" contraction  C[M,N] = epi( sum_seg A_seg[M,K] . W_seg[N,K]^T )  — the nn.Linear / nn.LSTMCell product
// that every stage of the Sub-GC path is built from (reference: nn.Linear calls in models/AttModel.py:72-120,
// models/lib/graph_conv_unit.py:29-30, models/lib/gpn.py:25-36 and nn.LSTMCell in models/AttModel.py:397-398).
//
// fp32 FMA accumulation on purpose: greedy/beam token ids must match the fp32 reference bit-for-bit and the
// measured top-1/top-2 log-prob margin of the path is ~1e-5 (SURVEY §7 hard part 1), so a single-pass TF32/BF16
// tensor-core product is not admissible.  Layout notes:
//   * both operands are K-contiguous; a BMxBK / BNxBK tile is read with 16-byte loads along K and stored
//     transposed in shared memory so the inner product reads conflict-free float4 rows,
//   * K is walked segment by segment (GemmSeg) with zero-filled tails, which removes every concat / repack,
//   * skinny problems (decode: M = rows in flight) are split along K so that >= 1 wave of CTAs streams the
//     weights; partials are reduced in a fixed order (deterministic) by the consumer.
#include "common.cuh"

namespace subgc {

constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int kTargetCtas = 2 * kNumSMs;  // two resident CTAs per SM

struct GemmKernelArgs {
    GemmProblem p;
    int total_tiles;      // k-tiles over all segments
    int tiles_per_split;
    int splits;
    float* part;          // != nullptr: raw partial sums [splits][M][N]
};

template <int ROWS, bool VEC>
__device__ __forceinline__ void load_tile(float4 (&reg)[ROWS / 64], const float* base, int ld, int K, int k0, int row0,
                                          int row_limit, const long long* gather, const int* gather32, int row_div, bool relu, int tid) {
#pragma unroll
    for (int i = 0; i < ROWS / 64; ++i) {
        int f = tid + i * GEMM_THREADS;
        int r = f >> 2, kq = (f & 3) << 2;
        int row = row0 + r;
        int k = k0 + kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < row_limit && k < K) {
            long long src = gather ? gather[row] : (gather32 ? (long long)gather32[row] : (long long)(row / row_div));
            const float* p = base + src * ld + k;
            if (VEC && k + 3 < K) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                v.x = __ldg(p);
                if (k + 1 < K) v.y = __ldg(p + 1);
                if (k + 2 < K) v.z = __ldg(p + 2);
                if (k + 3 < K) v.w = __ldg(p + 3);
            }
            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        }
        reg[i] = v;
    }
}

template <int ROWS>
__device__ __forceinline__ void store_tile(float (*s)[ROWS + 4], const float4 (&reg)[ROWS / 64], int tid) {
#pragma unroll
    for (int i = 0; i < ROWS / 64; ++i) {
        int f = tid + i * GEMM_THREADS;
        int r = f >> 2, kq = (f & 3) << 2;
        s[kq + 0][r] = reg[i].x;
        s[kq + 1][r] = reg[i].y;
        s[kq + 2][r] = reg[i].z;
        s[kq + 3][r] = reg[i].w;
    }
}

template <int BM, int BN, bool VEC>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_tn_kernel(const GemmKernelArgs a) {
    constexpr int TM = BM / 16, TN = BN / 16;
    const GemmProblem& p = a.p;
    if (p.active != nullptr && *p.active == 0) return;

    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kt_begin = blockIdx.z * a.tiles_per_split;
    const int kt_end = min(a.total_tiles, kt_begin + a.tiles_per_split);

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // locate (segment, k0) of the first tile
    int seg = 0, seg_tile0 = 0;
    while (seg < p.nseg - 1 && kt_begin >= seg_tile0 + (p.seg[seg].K + BK - 1) / BK) {
        seg_tile0 += (p.seg[seg].K + BK - 1) / BK;
        ++seg;
    }

    float4 ra[BM / 64], rb[BN / 64];
    auto fetch = [&](int kt) {
        while (seg < p.nseg - 1 && kt >= seg_tile0 + (p.seg[seg].K + BK - 1) / BK) {
            seg_tile0 += (p.seg[seg].K + BK - 1) / BK;
            ++seg;
        }
        const GemmSeg& s = p.seg[seg];
        int k0 = (kt - seg_tile0) * BK;
        load_tile<BM, VEC>(ra, s.A, s.lda, s.K, k0, m0, p.M, s.gather, s.gather32, s.a_row_div, s.relu_a != 0, tid);
        load_tile<BN, VEC>(rb, s.W, s.ldw, s.K, k0, n0, p.N, nullptr, nullptr, 1, false, tid);
    };

    if (kt_begin < kt_end) {
        fetch(kt_begin);
        store_tile<BM>(As[0], ra, tid);
        store_tile<BN>(Bs[0], rb, tid);
    }
    __syncthreads();

    for (int kt = kt_begin; kt < kt_end; ++kt) {
        const int buf = (kt - kt_begin) & 1;
        const bool has_next = kt + 1 < kt_end;
        if (has_next) fetch(kt + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int c = 0; c < TM / 4; ++c) {
                float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][c * 64 + ty * 4]);
                av[c * 4 + 0] = t.x; av[c * 4 + 1] = t.y; av[c * 4 + 2] = t.z; av[c * 4 + 3] = t.w;
            }
#pragma unroll
            for (int c = 0; c < TN / 4; ++c) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][c * 64 + tx * 4]);
                bv[c * 4 + 0] = t.x; bv[c * 4 + 1] = t.y; bv[c * 4 + 2] = t.z; bv[c * 4 + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (has_next) {
            store_tile<BM>(As[buf ^ 1], ra, tid);
            store_tile<BN>(Bs[buf ^ 1], rb, tid);
        }
        __syncthreads();
    }

    // ---- epilogue
    const bool raw = a.part != nullptr;
    float* out = raw ? a.part + (size_t)blockIdx.z * p.M * p.N : p.C;
    const int ldo = raw ? p.N : p.ldc;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
        if (m >= p.M) continue;
#pragma unroll
        for (int c = 0; c < TN / 4; ++c) {
            int n = n0 + c * 64 + tx * 4;
            if (n >= p.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = acc[i][c * 4 + j];
            if (!raw) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) v[j] = epilogue_apply(p.epi, v[j], m, n + j);
            }
            float* dst = out + (size_t)m * ldo + n;
            if (!raw && p.epi.accumulate) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) v[j] += dst[j];
            }
            if (n + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) dst[j] = v[j];
            }
        }
    }
}

// Fixed-order reduction of split-K partials followed by the epilogue (z ascending: deterministic).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const GemmProblem p, const float* __restrict__ part, int splits) {
    if (p.active != nullptr && *p.active == 0) return;
    const size_t total = (size_t)p.M * p.N;
    if ((p.N & 3) == 0 && (p.ldc & 3) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 15) == 0) {
        const int n4 = p.N >> 2;
        const size_t quads = total >> 2;
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (size_t)gridDim.x * blockDim.x) {
            const int m = (int)(q / n4), n = (int)(q - (size_t)m * n4) << 2;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int z = 0; z < splits; ++z) {
                const float4 t = *reinterpret_cast<const float4*>(part + (size_t)z * total + (q << 2));
                v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
            }
            v.x = epilogue_apply(p.epi, v.x, m, n); v.y = epilogue_apply(p.epi, v.y, m, n + 1);
            v.z = epilogue_apply(p.epi, v.z, m, n + 2); v.w = epilogue_apply(p.epi, v.w, m, n + 3);
            float4* dst = reinterpret_cast<float4*>(p.C + (size_t)m * p.ldc + n);
            if (p.epi.accumulate) { const float4 o = *dst; v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w; }
            *dst = v;
            if (p.epi.c16_hi) {   // split-fp16 copy for the consumer contraction (ld16 % 4 == 0, 8-byte aligned: checked by the launcher)
                unsigned short hh[4], hl[4];
                int ovf = 0;
                split_f16(v.x, hh[0], hl[0], ovf); split_f16(v.y, hh[1], hl[1], ovf); split_f16(v.z, hh[2], hl[2], ovf); split_f16(v.w, hh[3], hl[3], ovf);
                const size_t o16 = (size_t)m * p.epi.ld16 + n;
                *reinterpret_cast<uint2*>(p.epi.c16_hi + o16) = make_uint2((uint32_t)hh[0] | ((uint32_t)hh[1] << 16), (uint32_t)hh[2] | ((uint32_t)hh[3] << 16));
                *reinterpret_cast<uint2*>(p.epi.c16_lo + o16) = make_uint2((uint32_t)hl[0] | ((uint32_t)hl[1] << 16), (uint32_t)hl[2] | ((uint32_t)hl[3] << 16));
                if (ovf && p.overflow) atomicOr(p.overflow, 1);
            }
        }
        return;
    }
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int m = (int)(idx / p.N), n = (int)(idx - (size_t)m * p.N);
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += part[(size_t)z * total + idx];
        v = epilogue_apply(p.epi, v, m, n);
        if (p.epi.accumulate) v += p.C[(size_t)m * p.ldc + n];
        p.C[(size_t)m * p.ldc + n] = v;
    }
}

struct GemmPlan {
    int bm, bn, splits, tiles_per_split, total_tiles, grid_m, grid_n;
};

static GemmPlan plan_gemm(int M, int N, int total_tiles) {
    GemmPlan pl;
    long long t128 = (long long)((M + 127) / 128) * ((N + 127) / 128);
    bool small = (M <= 64) || (t128 * 2 <= kNumSMs && (long long)M * N <= 128LL * 1024);
    pl.bm = pl.bn = small ? 64 : 128;
    pl.grid_m = (M + pl.bm - 1) / pl.bm;
    pl.grid_n = (N + pl.bn - 1) / pl.bn;
    long long tiles = (long long)pl.grid_m * pl.grid_n;
    int splits = 1;
    if (tiles < kTargetCtas) {
        splits = (int)(kTargetCtas / tiles);
        int max_by_k = total_tiles / 4;  // at least 4 k-tiles (64 columns) per split
        if (splits > max_by_k) splits = max_by_k;
        if (splits < 1) splits = 1;
        if (splits > 32) splits = 32;
    }
    pl.tiles_per_split = (total_tiles + splits - 1) / splits;
    pl.splits = (total_tiles + pl.tiles_per_split - 1) / pl.tiles_per_split;
    if (pl.splits < 1) pl.splits = 1;
    pl.total_tiles = total_tiles;
    return pl;
}

static int count_tiles(const GemmProblem& p) {
    int t = 0;
    for (int s = 0; s < p.nseg; ++s) t += (p.seg[s].K + BK - 1) / BK;
    return t;
}

// Upper bounds independent of how K is cut into segments (each segment rounds its tile count up separately).
static int max_splits(int M, int N) {
    long long tiles64 = (long long)((M + 63) / 64) * ((N + 63) / 64);
    long long tiles128 = (long long)((M + 127) / 128) * ((N + 127) / 128);
    long long tiles = tiles64 < tiles128 ? tiles64 : tiles128;  // whichever tile shape the plan picks, it has >= this many CTAs
    if (tiles >= kTargetCtas) return 1;
    long long s = kTargetCtas / tiles;
    return (int)(s > 32 ? 32 : s);
}

size_t gemm_workspace_bytes(int M, int N, int Ktotal) {
    int s = max_splits(M, N);
    size_t simt = s > 1 ? align_up((size_t)s * M * N * sizeof(float), 256) : 0;
    size_t tc = tc_workspace_bytes(M, N, Ktotal);
    return simt > tc ? simt : tc;
}

size_t gemm_partial_elems(int M, int N, int Ktotal) {
    (void)Ktotal;
    return (size_t)max_splits(M, N) * M * N;
}

static bool vec_ok(const GemmProblem& p) {
    for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg& g = p.seg[s];
        if ((reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(g.W) & 15) || (g.lda & 3) || (g.ldw & 3))
            return false;
    }
    return true;
}

template <int BM, int BN>
static void launch_cfg(const GemmKernelArgs& a, const GemmPlan& pl, bool vec, cudaStream_t stream) {
    dim3 grid(pl.grid_n, pl.grid_m, pl.splits);
    if (vec)
        gemm_tn_kernel<BM, BN, true><<<grid, GEMM_THREADS, 0, stream>>>(a);
    else
        gemm_tn_kernel<BM, BN, false><<<grid, GEMM_THREADS, 0, stream>>>(a);
}

// raw_part != nullptr: always leave the per-split partial sums there (consumer reduces) and report the split
// count through *out_splits; otherwise reduce + epilogue into p.C using ws for the partials when needed.
int launch_gemm_ex(const GemmProblem& p, float* raw_part, size_t raw_part_elems, int* out_splits, void* ws, size_t ws_bytes,
                   cudaStream_t stream) {
    SUBGC_CHECK_ARG(p.M >= 0 && p.N > 0 && p.nseg >= 1 && p.nseg <= 4, "gemm: bad shape M=%d N=%d nseg=%d", p.M, p.N, p.nseg);
    if (p.M == 0) {
        if (out_splits) *out_splits = 1;
        return SUBGC_OK;
    }
    for (int s = 0; s < p.nseg; ++s)
        SUBGC_CHECK_ARG(p.seg[s].A && p.seg[s].W && p.seg[s].K > 0 && p.seg[s].a_row_div >= 1, "gemm: bad segment %d", s);
    GemmKernelArgs a;
    a.p = p;
    GemmPlan pl = plan_gemm(p.M, p.N, count_tiles(p));
    a.total_tiles = pl.total_tiles;
    a.tiles_per_split = pl.tiles_per_split;
    a.splits = pl.splits;
    a.part = nullptr;
    if (raw_part) {
        SUBGC_CHECK_ARG((size_t)pl.splits * p.M * p.N <= raw_part_elems, "gemm: partial buffer too small");
        a.part = raw_part;
        *out_splits = pl.splits;
    } else if (pl.splits > 1) {
        size_t need = (size_t)pl.splits * p.M * p.N * sizeof(float);
        if (ws == nullptr || ws_bytes < need) {
            set_error("gemm: workspace too small (%zu < %zu)", ws_bytes, need);
            return SUBGC_E_WORKSPACE;
        }
        a.part = static_cast<float*>(ws);
    } else {
        SUBGC_CHECK_ARG(p.C != nullptr && p.ldc >= p.N, "gemm: bad output");
    }
    bool vec = vec_ok(p);
    if (pl.bm == 128)
        launch_cfg<128, 128>(a, pl, vec, stream);
    else
        launch_cfg<64, 64>(a, pl, vec, stream);
    SUBGC_LAUNCH_CHECK();
    if (!raw_part && pl.splits > 1) {
        size_t total = (size_t)p.M * p.N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
        GemmProblem pr = p;   // the split-fp16 copy (epi.c16_*) is written by launch_gemm's split pass on this path
        pr.epi.c16_hi = nullptr; pr.epi.c16_lo = nullptr;
        splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(pr, a.part, pl.splits);
        SUBGC_LAUNCH_CHECK();
    }
    return SUBGC_OK;
}

// write_c16: the caller has checked the alignment the vector path's split-fp16 stores need; otherwise epi.c16_* is ignored here
void launch_splitk_reduce(const GemmProblem& p0, const float* part, int splits, cudaStream_t stream, bool write_c16) {
    GemmProblem p = p0;
    if (!write_c16) { p.epi.c16_hi = nullptr; p.epi.c16_lo = nullptr; }
    size_t total = (size_t)p.M * p.N;
    size_t work = ((p.N & 3) == 0) ? total / 4 : total;
    int blocks = (int)((work + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (blocks < 1) blocks = 1;
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(p, part, splits);
}

int launch_gemm_raw(const GemmProblem& p, void* ws, size_t ws_bytes, cudaStream_t stream, RawPartials* raw) {
    SUBGC_CHECK_ARG(p.M > 0 && raw != nullptr, "gemm(raw): bad arguments");
    if (p.wts) {
        GemmProblem q = p;
        q.wts = nullptr;
        resolve_packs(q, p.wts);
        return launch_gemm_raw(q, ws, ws_bytes, stream, raw);
    }
    if (h3_eligible(p)) return launch_gemm_h3(p, ws, ws_bytes, stream, raw);
    if (tc_eligible(p)) return launch_gemm_tc(p, ws, ws_bytes, stream, raw);
    int splits = 1;
    SUBGC_TRY(launch_gemm_ex(p, static_cast<float*>(ws), ws_bytes / sizeof(float), &splits, nullptr, 0, stream));
    raw->part = static_cast<const float*>(ws);
    raw->splits = splits;
    return SUBGC_OK;
}

int launch_gemm(const GemmProblem& p, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (p.wts) {
        GemmProblem q = p;
        q.wts = nullptr;
        resolve_packs(q, p.wts);
        return launch_gemm(q, ws, ws_bytes, stream);
    }
    bool wrote16 = false;
    int rc;
    if (p.M > 0 && h3_eligible(p)) {
        SUBGC_CHECK_ARG(p.N > 0 && p.C != nullptr && p.ldc >= p.N, "gemm: bad output");
        rc = launch_gemm_h3(p, ws, ws_bytes, stream, nullptr, &wrote16);
    } else if (p.M > 0 && tc_eligible(p)) {
        SUBGC_CHECK_ARG(p.N > 0 && p.C != nullptr && p.ldc >= p.N, "gemm: bad output");
        rc = launch_gemm_tc(p, ws, ws_bytes, stream);
    } else {
        rc = launch_gemm_ex(p, nullptr, 0, nullptr, ws, ws_bytes, stream);
    }
    if (rc == SUBGC_OK && p.M > 0 && p.epi.c16_hi && p.epi.c16_lo && !wrote16)   // the path taken did not split its result itself
        rc = launch_split_rows(p.C, p.M, p.N, p.ldc, p.epi.c16_hi, p.epi.c16_lo, p.epi.ld16, p.overflow, stream);
    return rc;
}

}  // namespace subgc

using namespace subgc;

extern "C" size_t subgc_linear_workspace_bytes(int M, int N, int K) { return gemm_workspace_bytes(M, N, K); }

extern "C" int subgc_linear_forward(int M, int N, int K, const float* A, int lda, const int64_t* a_gather, const float* W, int ldw,
                                    const float* bias, int relu, float* C, int ldc, void* ws, size_t ws_bytes,
                                    subgc_stream_t stream) {
    SUBGC_CHECK_ARG(A && W && C && M >= 0 && N > 0 && K > 0, "subgc_linear_forward: bad arguments");
    GemmProblem p;
    p.M = M; p.N = N; p.nseg = 1;
    p.seg[0] = make_seg(A, lda, W, ldw, K);
    p.seg[0].gather = reinterpret_cast<const long long*>(a_gather);
    p.epi.bias = bias;
    p.epi.relu = relu;
    p.C = C; p.ldc = ldc;
    return launch_gemm(p, ws, ws_bytes, static_cast<cudaStream_t>(stream));
}
