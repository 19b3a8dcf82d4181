// Persistent decode kernel: the whole greedy / top-k sampling loop of AttModel._sample (reference models/AttModel.py:278-326) as ONE
// cooperative launch -- every step = get_logprobs_state (AttModel.py:328-341) = TopDownCore (:400-431) + Attention (:445-471) + logit.
//
// Why: with one kernel per stage (decode.cu) a step is 8 dependent launches and the weight stream (152 MB per token) stops at every
// kernel boundary; measured 0.33 of the HBM roofline.  tools/ubench_ingest.cu showed that the SMs can ingest the weight stream at
// 7.0-7.3 TB/s next to an equally large L2-resident activation stream, as long as >= 96 KB per SM stay in flight, and that a grid-wide
// hand-off costs 1.2-2.4 us.  So here one CTA per SM owns a fixed slice of every contraction of the step and
//   * a producer thread streams that slice (its own contiguous, pre-swizzled region of the "stream pack", cp.async.bulk, L2 evict-first)
//     into a 3-slot ring and never waits for anything but a free slot: the HBM pipe keeps running across phase hand-offs;
//   * a second thread loads the activation tiles (split-fp16, pre-swizzled by their producers) once their dependency counter is up;
//   * one thread issues the tcgen05.mma chains (split-fp16: main + cross accumulators in TMEM, two accumulator sets);
//   * 16 worker warps drain accumulators (TMEM -> split-K partial in L2), and run the LSTM cells, the attention of "their" row and the
//     token selection of "their" row (12 warps; the other 4 drain the next step's attention-LSTM accumulator meanwhile), handing results
//     over through global counters (release / acquire), not kernel boundaries.  A hand-off is polled by ONE thread per CTA (or one per
//     row for the tokens) and every polled word has a 32-byte sector of its own: many pollers on one line keep an L2 slice busy.
// Work that does not depend on the newest result (the h_att / h_lang segments of the NEXT contraction) is issued while a hand-off is
// in flight.  The schedule (which CTA contracts which weight tile in which order) is a table built on the host (mega_plan).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace subgc {

// ---- geometry ------------------------------------------------------------------------------------------------------------
constexpr int MG_THREADS = 640;          // warps 0-3: W producer, X loader, MMA issuer, TMEM owner; warps 4-19: workers
constexpr int MG_NW = 512;               // worker threads
constexpr int MG_NSEL = 384;             // threads of the 12 selection warps of a row CTA (the other 4 worker warps drain A(t+1) meanwhile)
constexpr int MG_WSLOTS = 3, MG_WSLOT_BYTES = 36864;   // weight ring: tiles of <= 144 rows x 64 k (hi | lo)
constexpr int MG_XSLOTS = 3, MG_XTILE_BYTES = 32768;   // activation ring: [128 rows x 64 k] hi | lo
constexpr int MG_SCR_BYTES = 16384;
constexpr int MG_TOK_STRIDE = 8;         // ints per token slot: one 32-byte sector each, so that the polls of a step spread over the L2 slices
constexpr int MG_OFF_X = MG_WSLOTS * MG_WSLOT_BYTES;            // 110592
constexpr int MG_OFF_SCR = MG_OFF_X + MG_XSLOTS * MG_XTILE_BYTES;   // 208896
constexpr int MG_OFF_BAR = MG_OFF_SCR + MG_SCR_BYTES;           // 225280
constexpr int MG_SMEM = MG_OFF_BAR + 2048 + 1024;               // + control block + alignment slack
// TMEM columns of the two accumulator sets (set 0: tiles <= 112 wide, set 1: <= 144) and the offset of the cross-term accumulator
__host__ __device__ constexpr int mg_set_col(int set) { return set ? 224 : 0; }
__host__ __device__ constexpr int mg_set_cross(int set) { return set ? 144 : 112; }
constexpr int MG_MAX_TASKS = 8;
constexpr int MG_TRACE_EVENTS = 48;   // 0-16 workers, 20 + 2k / 21 + 2k MMA task k (operands landed / issued), 40 producer (step issued)
constexpr int MG_MAX_TILES = 192;

enum { MG_X_CTX = 1, MG_X_HATT = 2, MG_X_HLANG_PREV = 3, MG_X_HLANG = 4 };
enum { MG_F_START = 1, MG_F_COMMIT = 2, MG_F_NEXT = 4 };
enum { MG_JOB_A = 0, MG_JOB_B = 1, MG_JOB_C = 2, MG_JOB_D = 3 };
// sync counters (uint32 indices into MgParams::sync)
// (one 128-byte line per global counter, one 32-byte sector per tile counter: pollers of different counters do not meet on a line)
constexpr int MG_TILE_STRIDE = 8;
enum { MG_C_XT = 0, MG_C_HATT = 32, MG_C_HLANG = 64, MG_C_CTX = 96, MG_C_B = 128, MG_C_D = 160, MG_C_ABORT = 192, MG_C_TILE_A = 256,
       MG_C_TILE_C = 256 + MG_MAX_TILES * MG_TILE_STRIDE, MG_C_TOTAL = 256 + 2 * MG_MAX_TILES * MG_TILE_STRIDE };

struct MgTask { int n_blk, n_rows, x_src, x_kb0, acc_set, flags, rot, pad1; };   // block b covers k-block x_kb0 + (b + rot) % n_blk
struct MgJob {
    int present, set, n_rows;
    int part_off;      // floats: A/C: this job's [col][128] tile; B/D: first column inside the plane
    int plane;         // split index (B/D: plane of the row-major partials)
    int tile, n_split; // A/C: tile id (sync counter) and partials per tile
    int tile_u0, tile_units;   // A/C: hidden units of the tile
    int u_lo, u_n;     // A/C: this CTA's share of the tile's units in the cell
    int part_tile_off; // A/C: floats, partial of split 0 of the tile (split z at + z * 16384)
};
struct MgCta {
    MgTask task[MG_MAX_TASKS];
    MgJob job[4];
    int n_task;
    int step_bytes;
    unsigned long long w_off;
};

struct MgParams {
    const MgCta* ctas;
    const uint8_t* wstream;
    int n_cta, S, T, len_max, H, AH, V1, kbH;
    int nL, nB, nD, zB, zD, ldB, ldD;
    const float* fc_pre; const float* att; const float* p_att; const float* masks;
    const float* h2att_b; const float* alpha_w; const float* alpha_b; const float* logit_b;
    const float* xt_table;       // [V1, H, 4] (gates i,f,g,o of a unit adjacent): W_ih[:, 2H:2H+E] relu(E[v]) for every token v (subgc_mega_pack)
    const float* lang_b;         // [H, 4]: b_ih + b_hh of the language LSTM, gates of a unit adjacent (built next to fc_pre at launch)
    uint8_t* x_ctx; uint8_t* x_hatt[2]; uint8_t* x_hlang[2];
    float* partA; float* partC; float* partB; float* partD;
    unsigned* sync;
    int* tok_slots;              // [T][128][MG_TOK_STRIDE], zero at launch: token of (step, row) + 1 once selected (the cells of the next step spin on it)
    long long* seq; float* seq_lp; int* steps_done; int* overflow;
    int mode; float temp; int top_k; unsigned long long seed, offset; const float* uniforms;
    const int* counts;           // nullable DEVICE int32[2]: (rows, longest sub-graph) decided by the NMS kernel earlier in the stream; S / len_max
                                 // are then upper bounds (buffer strides) and the real values are read here: no host round trip
    int len_stride;              // nodes per row in att / p_att / masks
    unsigned long long* trace;   // debug (SUBGC_MEGA_TRACE=1): [cta][step][MG_TRACE_EVENTS] globaltimer stamps, else nullptr
    unsigned long long timeout_ns;   // a wait longer than this aborts the kernel (SUBGC_MEGA_TIMEOUT_S, default 4 s)
};

// ---- PTX helpers local to this kernel ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int ld_relaxed_s32(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_s32(int* p, int v) { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.global;" ::: "memory"); }   // generic <-> async proxy, global memory
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void sel_bar() { asm volatile("bar.sync 2, 384;" ::: "memory"); }     // the 12 selection warps
__device__ __forceinline__ void drain_bar() { asm volatile("bar.sync 3, 128;" ::: "memory"); }   // the 4 drain warps
#define MG_TMEM_LD16(R, ADDR)                                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(R[0]), "=r"(R[1]), "=r"(R[2]), "=r"(R[3]), "=r"(R[4]), "=r"(R[5]), "=r"(R[6]), "=r"(R[7]), "=r"(R[8]), "=r"(R[9]),  \
                   "=r"(R[10]), "=r"(R[11]), "=r"(R[12]), "=r"(R[13]), "=r"(R[14]), "=r"(R[15])                                          \
                 : "r"(ADDR))

// control block in shared memory (after the rings and the worker scratch)
struct MgCtl {
    unsigned long long w_full[MG_WSLOTS], w_empty[MG_WSLOTS], x_full[MG_XSLOTS], x_empty[MG_XSLOTS], acc_full[2], acc_free[2], mma_done;
    uint32_t tmem_slot;
    volatile int stop;      // 1: all rows finished (normal early exit), 2: aborted (time-out; MG_C_ABORT holds the site)
    volatile int fail;      // sticky: a worker thread gave up a spin wait (the kernel is stopping)
    volatile int flag[4];   // worker broadcast of wait results (two per round, rounds alternate); [2], [3]: the drain group's
    // 16-byte aligned: the compiler reads these arrays with 16-byte loads; unaligned, such a load also covered flag[3], which the drain
    // group writes while the selection group reduces (compute-sanitizer racecheck: a WAR hazard on a value nobody uses)
    alignas(16) int red_i[16];
    float red_f[16];
    float red_s[16];
    float red_s2[16];       // top-k selection (k <= 3): per-warp sum and the two best elements other than the row maximum
    float cand_v[2][16];
    int cand_i[2][16];
    float topv[16];
    int topi[16];
    MgCta cta;
};
static_assert(sizeof(MgCtl) <= 2048, "control block must fit its 2 KB");
static_assert(MG_SMEM <= 232448, "shared memory of a CTA");

// a wait longer than MgParams::timeout_ns aborts the kernel (guard: never hang the GPU).  Default 4 s; SUBGC_MEGA_TIMEOUT_S raises it
// for runs under compute-sanitizer, where the kernel is ~100x slower.

struct MgWait {
    MgCtl* ctl;
    const MgParams& p;
    unsigned long long t0;
    __device__ __forceinline__ bool expired(int site) {
        if (ld_acquire(p.sync + MG_C_ABORT) != 0) { ctl->stop = 2; return true; }
        if (globaltimer_ns() - t0 > p.timeout_ns) {
            atomicCAS(p.sync + MG_C_ABORT, 0u, (unsigned)(site * 1000 + (int)blockIdx.x + 1));
            ctl->stop = 2;
            return true;
        }
        return false;
    }
    // false: the kernel is stopping (early exit or abort)
    __device__ __forceinline__ bool mbar(unsigned long long* bar, uint32_t parity, int site) {
        const uint32_t b = smem_u32(bar);
        for (uint32_t it = 1;; ++it) {
            if (mbar_try(b, parity)) return true;
            if (ctl->stop) return false;
            if ((it & 255u) == 0 && expired(site)) return false;
        }
    }
    __device__ __forceinline__ bool counter_masked(const unsigned* c, unsigned target, unsigned mask, int site, unsigned& out) {
        for (uint32_t it = 1;; ++it) {
            const unsigned v = ld_acquire(c);
            if ((v & mask) >= target) { out = v; return true; }
            if (ctl->stop) return false;
            if ((it & 255u) == 0 && expired(site)) return false;
        }
    }
    __device__ __forceinline__ bool counter(const unsigned* c, unsigned target, int site) {
        for (uint32_t it = 1;; ++it) {
            if (ld_acquire(c) >= target) return true;
            if (ctl->stop) return false;
            if ((it & 255u) == 0 && expired(site)) return false;
        }
    }
};

// element (row, col) of a split-fp16 activation tensor in tile layout: [kb][hi | lo][128 rows][64 cols], 128-byte swizzle
__device__ __forceinline__ size_t x_tile_off(int row, int col) {
    const int kb = col >> 6, c = col & 63;
    return (size_t)kb * MG_XTILE_BYTES + (size_t)row * 128 + (size_t)((((c >> 3) ^ (row & 7)) << 4) + ((c & 7) << 1));
}
__device__ __forceinline__ void x_store(uint8_t* buf, int row, int col, float v, int& ovf) {
    unsigned short h, l;
    split_f16(v, h, l, ovf);
    const size_t o = x_tile_off(row, col);
    *reinterpret_cast<unsigned short*>(buf + o) = h;
    *reinterpret_cast<unsigned short*>(buf + o + MG_XTILE_BYTES / 2) = l;
}

// LSTM cell non-linearities from the ex2 / rcp units: sigmoid and tanh to ~2e-7 absolute (libdevice's expf / tanhf cost ~5x the
// instructions, and the cells are two of the hand-offs every step waits for); |h| < 1, so this is far below the 2e-5 bar
__device__ __forceinline__ float mg_sigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float mg_tanh(float x) {
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
__device__ __forceinline__ float mg_tanh_score(float x) {   // same form as decode.cu: ex2-based, |err| <= 2e-7
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

// reductions over the 384 selection threads (wt < MG_NSEL); every selection thread must call
__device__ __forceinline__ float workers_sum(float v, MgCtl* ctl, int wt) {
    v = warp_sum(v);
    sel_bar();
    if ((wt & 31) == 0) ctl->red_f[wt >> 5] = v;
    sel_bar();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < MG_NSEL / 32; ++i) t += ctl->red_f[i];
    return t;
}
__device__ __forceinline__ void workers_argmax(float& v, int& i, MgCtl* ctl, int wt) {
    warp_argmax(v, i);
    sel_bar();
    if ((wt & 31) == 0) { ctl->red_f[wt >> 5] = v; ctl->red_i[wt >> 5] = i; }
    sel_bar();
    v = ctl->red_f[0]; i = ctl->red_i[0];
#pragma unroll
    for (int w = 1; w < MG_NSEL / 32; ++w) argmax_combine(v, i, ctl->red_f[w], ctl->red_i[w]);
}

struct MgPhilox {
    static __device__ __forceinline__ void round(unsigned (&c)[4], unsigned k0, unsigned k1) {
        unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static __device__ float uniform(unsigned long long seed, unsigned long long offset, unsigned t, unsigned r) {   // identical to decode.cu's stream
        unsigned c[4] = {(unsigned)offset, (unsigned)(offset >> 32), t, r};
        unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
        for (int i = 0; i < 10; ++i) { round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        return (float)(c[0] >> 8) * (1.0f / 16777216.0f);
    }
};

constexpr int MG_SELQ = 7;        // logit column quads per selection thread: thread wt owns quads wt + i * MG_NSEL, V1 <= 10752
constexpr int MG_SELVALS = 4 * MG_SELQ;
#define MG_SELCOL(WT_, I_) (4 * ((WT_) + ((I_) >> 2) * MG_NSEL) + ((I_) & 3))   // column of value slot I_ of selection thread WT_
constexpr int MG_ATT_ITEMS = 10; // (node, 32-quad slice) items per worker warp and pass

// kTrace: the instantiation with the time stamps (SUBGC_MEGA_TRACE=1); the production one carries none of that code (every step runs
// through ~50 stamp sites once, and straight-line code that is executed once per step comes from a cold instruction cache)
template <bool kTrace>
__global__ void __launch_bounds__(MG_THREADS, 1) mega_decode_kernel(const __grid_constant__ MgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    MgCtl* ctl = reinterpret_cast<MgCtl*>(gbase + MG_OFF_BAR);
    float* scr = reinterpret_cast<float*>(gbase + MG_OFF_SCR);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta_id = blockIdx.x;
    const int S = p.counts ? min(p.S, __ldg(p.counts)) : p.S, T = p.T, H = p.H;
    const int len_rt = p.counts ? min(p.len_max, __ldg(p.counts + 1)) : p.len_max;   // attention length (<= len_stride)

    // ---- set-up: schedule of this CTA, barriers, tensor memory
    {
        const int* src = reinterpret_cast<const int*>(p.ctas + cta_id);
        int* dst = reinterpret_cast<int*>(&ctl->cta);
        for (int i = tid; i < (int)(sizeof(MgCta) / 4); i += MG_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int s = 0; s < MG_WSLOTS; ++s) { mbar_init(smem_u32(&ctl->w_full[s]), 1); mbar_init(smem_u32(&ctl->w_empty[s]), 1); }
        for (int s = 0; s < MG_XSLOTS; ++s) { mbar_init(smem_u32(&ctl->x_full[s]), 1); mbar_init(smem_u32(&ctl->x_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&ctl->acc_full[s]), 1); mbar_init(smem_u32(&ctl->acc_free[s]), 16); }
        mbar_init(smem_u32(&ctl->mma_done), 1);
        ctl->stop = 0; ctl->fail = 0; ctl->flag[0] = 0; ctl->flag[1] = 0; ctl->flag[2] = 0; ctl->flag[3] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 3) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ctl->tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = ctl->tmem_slot;
    const MgCta& cta = ctl->cta;
    MgWait wt_{ctl, p, globaltimer_ns()};
#define MG_STAMP(T_, EV_) do { if (kTrace && p.trace) p.trace[((size_t)cta_id * T + (T_)) * MG_TRACE_EVENTS + (EV_)] = globaltimer_ns(); } while (0)

    // The kernel launches with 96 registers per thread (640 threads = 61440).  The four single-thread roles need fewer: their warpgroup
    // hands registers over to the 16 worker warps (4 warpgroups: 128 x 64 + 512 x 104 = 61440), whose selection phase keeps 56 partial
    // values in flight.  A spill is expensive here: the 227 KB of shared memory leave almost no L1 behind local memory.
#define MG_REG_DEC() asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory")
#define MG_REG_INC() asm volatile("setmaxnreg.inc.sync.aligned.u32 104;" ::: "memory")
    if (warp == 0) {
        MG_REG_DEC();
        // ===================================================== weight stream producer =====================================================
        if (lane == 0) {
            uint64_t pol;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
            uint32_t cnt = 0;
            bool alive = true;
            for (int t = 0; t < T && alive; ++t) {
                const uint8_t* src = p.wstream + cta.w_off;
                for (int k = 0; k < cta.n_task && alive; ++k) {
                    const MgTask& tk = cta.task[k];
                    const uint32_t bytes = (uint32_t)tk.n_rows * 256u;
                    if ((tk.flags & MG_F_NEXT) && t == T - 1) { src += (size_t)tk.n_blk * bytes; continue; }
                    for (int b = 0; b < tk.n_blk; ++b) {
                        const uint32_t s = cnt % MG_WSLOTS;
                        if (cnt >= MG_WSLOTS && !wt_.mbar(&ctl->w_empty[s], ((cnt / MG_WSLOTS) & 1u) ^ 1u, 1)) { alive = false; break; }
                        const uint32_t full = smem_u32(&ctl->w_full[s]);
                        mbar_arrive_expect_tx(full, bytes);
                        bulk_load_hint(base + s * MG_WSLOT_BYTES, src, bytes, full, pol);
                        src += bytes;
                        ++cnt;
                    }
                }
                MG_STAMP(t, 40);
            }
            // tiles still in flight must land before the CTA may retire (they were requested ahead of a stop)
            if (!alive && ctl->stop == 1)
                for (uint32_t i = (cnt > MG_WSLOTS ? cnt - MG_WSLOTS : 0); i < cnt; ++i)
                    while (!mbar_try(smem_u32(&ctl->w_full[i % MG_WSLOTS]), (i / MG_WSLOTS) & 1u)) {}
        }
    } else if (warp == 1) {
        MG_REG_DEC();
        // ===================================================== activation tile loader =====================================================
        if (lane == 0) {
            uint32_t cnt = 0;
            bool alive = true;
            int steps = T + 1;   // executed steps as the reference counts them (AttModel.py:312-314)
            // last value seen of every counter: a dependency that is known to be satisfied costs no further L2 round trip.
            // MG_C_XT packs two counts: bits 0-15 rows whose token is selected, bits 16+ rows still unfinished, summed over the steps
            unsigned seen_xt = 0, seen_ctx = 0, seen_hatt = 0, seen_hlang = 0, unf_before = 0;
            auto need = [&](const unsigned* c, unsigned& seen, unsigned target, unsigned mask, int site) -> bool {
                while ((seen & mask) < target) {
                    unsigned v = 0;
                    if (!wt_.counter_masked(c, target, mask, site, v)) return false;
                    seen = v;
                }
                return true;
            };
            for (int t = 0; t < T && alive; ++t) {
                if (t >= 1) {   // every row's selection of step t-1 is done: the all-finished early exit is decided here
                    if (!need(p.sync + MG_C_XT, seen_xt, (unsigned)S * (unsigned)t, 0xffffu, 2)) { alive = false; break; }
                    const unsigned unf = seen_xt >> 16;
                    if (unf == unf_before) { ctl->stop = 1; steps = t; alive = false; break; }
                    unf_before = unf;
                }
                for (int k = 0; k < cta.n_task && alive; ++k) {
                    const MgTask& tk = cta.task[k];
                    if ((tk.flags & MG_F_NEXT) && t == T - 1) continue;
                    const uint8_t* xb;
                    bool ok = true;
                    switch (tk.x_src) {
                        case MG_X_CTX: xb = p.x_ctx; ok = need(p.sync + MG_C_CTX, seen_ctx, (unsigned)S * (unsigned)(t + 1), ~0u, 3); break;
                        case MG_X_HATT: xb = p.x_hatt[t & 1]; ok = need(p.sync + MG_C_HATT, seen_hatt, (unsigned)p.nL * (unsigned)(t + 1), ~0u, 3); break;
                        case MG_X_HLANG_PREV: xb = p.x_hlang[(t + 1) & 1]; ok = need(p.sync + MG_C_HLANG, seen_hlang, (unsigned)p.nL * (unsigned)t, ~0u, 3); break;
                        default: xb = p.x_hlang[t & 1]; ok = need(p.sync + MG_C_HLANG, seen_hlang, (unsigned)p.nL * (unsigned)(t + 1), ~0u, 3); break;
                    }
                    if (!ok) { alive = false; break; }
                    fence_proxy_async_all();   // the tiles were written with generic stores by other SMs; the bulk copies below read them
                    for (int b = 0; b < tk.n_blk; ++b) {
                        const uint32_t s = cnt % MG_XSLOTS;
                        if (cnt >= MG_XSLOTS && !wt_.mbar(&ctl->x_empty[s], ((cnt / MG_XSLOTS) & 1u) ^ 1u, 4)) { alive = false; break; }
                        const uint32_t full = smem_u32(&ctl->x_full[s]);
                        mbar_arrive_expect_tx(full, MG_XTILE_BYTES);
                        bulk_load(base + MG_OFF_X + s * MG_XTILE_BYTES, xb + (size_t)(tk.x_kb0 + (b + tk.rot) % tk.n_blk) * MG_XTILE_BYTES, MG_XTILE_BYTES, full);
                        ++cnt;
                    }
                }
            }
            if (cta_id == 0 && ctl->stop != 2) {
                if (alive && need(p.sync + MG_C_XT, seen_xt, (unsigned)S * (unsigned)T, 0xffffu, 13) && (seen_xt >> 16) == unf_before) steps = T;
                if (ctl->stop != 2) p.steps_done[0] = steps;
            }
        }
    } else if (warp == 2) {
        MG_REG_DEC();
        // ===================================================== MMA issuer =====================================================
        if (lane == 0) {
            uint32_t wcnt = 0, xcnt = 0, jobs[2] = {0, 0};
            int drain_set = -1;   // accumulator set whose epilogue is in progress
            bool alive = true;
            for (int t = 0; t < T && alive; ++t) {
                for (int k = 0; k < cta.n_task && alive; ++k) {
                    const MgTask& tk = cta.task[k];
                    if ((tk.flags & MG_F_NEXT) && t == T - 1) continue;
                    const int set = tk.acc_set;
                    if (drain_set >= 0) {
                        // tcgen05.ld of an epilogue runs ~2.5x slower while MMAs write tensor memory (measured): the work issued next is
                        // never what the hand-off waits for, so it starts once the accumulator has been drained
                        if (!wt_.mbar(&ctl->acc_free[drain_set], (jobs[drain_set] - 1) & 1u, 14)) { alive = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        drain_set = -1;
                    }
                    const bool fresh = (tk.flags & MG_F_START) != 0;
                    if (fresh && jobs[set] > 0) {   // the workers must have drained the previous accumulator of this set
                        if (!wt_.mbar(&ctl->acc_free[set], (jobs[set] - 1) & 1u, 5)) { alive = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const uint32_t n = (uint32_t)tk.n_rows;
                    const uint32_t idesc = (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);   // D fp32, A/B fp16 K-major, N = n, M = 128
                    const uint32_t d_main = tmem + (uint32_t)mg_set_col(set), d_cross = d_main + (uint32_t)mg_set_cross(set);
                    for (int b = 0; b < tk.n_blk; ++b) {
                        const uint32_t ws = wcnt % MG_WSLOTS, xs = xcnt % MG_XSLOTS;
                        if (!wt_.mbar(&ctl->w_full[ws], (wcnt / MG_WSLOTS) & 1u, 6)) { alive = false; break; }
                        if (!wt_.mbar(&ctl->x_full[xs], (xcnt / MG_XSLOTS) & 1u, 7)) { alive = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (b == 0) MG_STAMP(t, 20 + 2 * k);
                        const uint32_t wa = base + ws * MG_WSLOT_BYTES, xa = base + MG_OFF_X + xs * MG_XTILE_BYTES;
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t whi = umma_desc_sw128(wa + ks * 32);
                            const uint64_t wlo = umma_desc_sw128(wa + n * 128u + ks * 32);
                            const uint64_t xhi = umma_desc_sw128(xa + ks * 32);
                            const uint64_t xlo = umma_desc_sw128(xa + MG_XTILE_BYTES / 2 + ks * 32);
                            const uint32_t acc = (fresh && b == 0 && ks == 0) ? 0u : 1u;
                            umma_f16_afill(d_main, xhi, whi, idesc, acc);
                            umma_f16_alast(d_cross, xhi, wlo, idesc, acc);
                            umma_f16(d_cross, xlo, whi, idesc, 1u);
                        }
                        umma_commit(smem_u32(&ctl->w_empty[ws]));
                        umma_commit(smem_u32(&ctl->x_empty[xs]));
                        ++wcnt; ++xcnt;
                    }
                    if (alive && (tk.flags & MG_F_COMMIT)) { umma_commit(smem_u32(&ctl->acc_full[set])); ++jobs[set]; drain_set = set; }
                    MG_STAMP(t, 21 + 2 * k);
                }
            }
            // every MMA issued so far (the early exit leaves next-step work behind) has retired before tensor memory is released
            if (ctl->stop != 2) {
                umma_commit(smem_u32(&ctl->mma_done));
                for (uint32_t it = 0; it < (1u << 24) && !mbar_try(smem_u32(&ctl->mma_done), 0); ++it) {}
            }
        }
    } else if (warp == 3) {
        MG_REG_DEC();
    } else {
        MG_REG_INC();
        // ===================================================== workers =====================================================
        const int wt = tid - 128, ww = wt >> 5;
        // parities in one register: bit 0 worker-wide waits, bit 1 drain-group waits, bits 2 / 3 epilogues done on accumulator set 0 / 1
        uint32_t par = 0;
        bool alive = true;
        const bool has_row = cta_id < S;
        const int row = cta_id;
        // broadcast wait helpers: worker thread 0 waits, everybody learns the outcome
        auto w_mbar = [&](unsigned long long* bar, uint32_t parity, int site) -> bool {
            if (wt == 0) ctl->flag[par & 1] = wt_.mbar(bar, parity, site) ? 1 : 0;
            worker_bar();
            const bool ok = ctl->flag[par & 1] != 0;
            par ^= 1u;
            return ok;
        };
        auto w_counter = [&](const unsigned* c, unsigned target, int site) -> bool {
            if (wt == 0) ctl->flag[par & 1] = wt_.counter(c, target, site) ? 1 : 0;
            worker_bar();
            const bool ok = ctl->flag[par & 1] != 0;
            par ^= 1u;
            return ok;
        };
        // everything this CTA's workers wrote becomes visible to whoever acquires the counter afterwards
        auto w_signal = [&](unsigned* c, bool tiles = false, unsigned inc = 1u) {
            if (tiles) fence_proxy_async_all();   // activation tiles are read through the async proxy (bulk copies) on the other side
            worker_bar();
            if (wt == 0) red_release(c, inc);   // release at gpu scope: cumulative over what the barrier made visible to this thread
        };
        // accumulator -> split-K partial.  col_major: LSTM tile as [unit][128 rows] float4 (i, f, g, o), else row-major plane [128][ld]
        // sub: only the 4 drain warps (ww >= 12, one per lane quarter) call, while the other 12 run the token selection
        // bias: added to the partial here (the plane-0 tiles of the logit contraction: the selection then sums planes only)
        auto epilogue = [&](const MgJob& jb, float* dst, int ld, bool col_major, int t, int ev, bool sub = false, const float* bias = nullptr,
                            int bias_n = 0) -> bool {
            const int set = jb.set;
            // the bias of this tile's columns (<= 144 floats) -> L1 while the accumulator is still being computed: the broadcast loads
            // below would otherwise pay an L2 round trip per 16-column chunk on the way to the hand-off
            if (bias && wt < 8 && wt * 32 < jb.n_rows) asm volatile("prefetch.global.L1 [%0];" ::"l"(bias + wt * 32));
            if (!sub) {
                if (!w_mbar(&ctl->acc_full[set], (par >> (2 + set)) & 1u, 8)) return false;
            } else {
                if (wt == MG_NSEL) ctl->flag[2 + ((par >> 1) & 1)] = wt_.mbar(&ctl->acc_full[set], (par >> (2 + set)) & 1u, 8) ? 1 : 0;
                drain_bar();
                const bool ok = ctl->flag[2 + ((par >> 1) & 1)] != 0;
                par ^= 2u;
                if (!ok) return false;
            }
            par ^= 4u << set;
            if (wt == (sub ? MG_NSEL : 0)) MG_STAMP(t, ev);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int q = ww & 3, g = sub ? 0 : ww >> 2, gs = sub ? 1 : 4, ncc = sub ? 9 : 3;
            const int r = q * 32 + lane;
            const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)mg_set_col(set);
#pragma unroll 1
            for (int cc = 0; cc < ncc; ++cc) {
                const int chunk = g + gs * cc;   // 16-column chunks, interleaved over the warps of a lane quarter (<= 9 chunks)
                if (chunk * 16 >= jb.n_rows) break;
                uint32_t ra[16], rb[16];
                MG_TMEM_LD16(ra, tl + (uint32_t)(chunk * 16));
                MG_TMEM_LD16(rb, tl + (uint32_t)mg_set_cross(set) + (uint32_t)(chunk * 16));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaf(__uint_as_float(rb[j]), kH3LoInv, __uint_as_float(ra[j]));
                if (bias) {   // the same address in every lane: one broadcast transaction per column
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] += (chunk * 16 + j < bias_n) ? __ldg(bias + chunk * 16 + j) : 0.f;
                }
                if (col_major) {   // LSTM tile: columns are [unit][gate]; slot (unit, row) is one float4 of the four gates
                    float4* o = reinterpret_cast<float4*>(dst) + (size_t)(chunk * 4) * 128 + r;
#pragma unroll
                    for (int uu = 0; uu < 4; ++uu) __stcg(o + (size_t)uu * 128, make_float4(v[4 * uu], v[4 * uu + 1], v[4 * uu + 2], v[4 * uu + 3]));
                } else {
                    // A lane holds 64 contiguous bytes of its row.  Lane pairs swap halves so that every store instruction writes
                    // 32 contiguous bytes per pair (full sectors) instead of 16 per lane (half sectors: measured ~1.5 us slower per tile)
                    const bool odd = lane & 1;
                    float sa[4], sb[4], ra_[4], rb_[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) { sa[e] = odd ? v[e] : v[4 + e]; sb[e] = odd ? v[8 + e] : v[12 + e]; }
#pragma unroll
                    for (int e = 0; e < 4; ++e) { ra_[e] = __shfl_xor_sync(0xffffffffu, sa[e], 1); rb_[e] = __shfl_xor_sync(0xffffffffu, sb[e], 1); }
                    const int r_even = r & ~1, r_odd = r | 1;
                    float* oe = dst + (size_t)r_even * ld + chunk * 16 + (odd ? 4 : 0);
                    float* oo = dst + (size_t)r_odd * ld + chunk * 16 + (odd ? 4 : 0);
                    // even row: pieces 0,1 then 2,3; odd row: pieces 0,1 then 2,3 (piece = 4 floats; the even lane writes pieces 0 / 2)
                    const float4 e0 = odd ? make_float4(ra_[0], ra_[1], ra_[2], ra_[3]) : make_float4(v[0], v[1], v[2], v[3]);
                    const float4 e1 = odd ? make_float4(rb_[0], rb_[1], rb_[2], rb_[3]) : make_float4(v[8], v[9], v[10], v[11]);
                    const float4 o0 = odd ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(ra_[0], ra_[1], ra_[2], ra_[3]);
                    const float4 o1 = odd ? make_float4(v[12], v[13], v[14], v[15]) : make_float4(rb_[0], rb_[1], rb_[2], rb_[3]);
                    __stcg(reinterpret_cast<float4*>(oe), e0);
                    __stcg(reinterpret_cast<float4*>(oe + 8), e1);
                    __stcg(reinterpret_cast<float4*>(oo), o0);
                    __stcg(reinterpret_cast<float4*>(oo + 8), o1);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
                for (int i = 0; i < (sub ? 4 : 1); ++i) mbar_arrive(smem_u32(&ctl->acc_free[set]));   // the barrier counts 16 warps
            return true;
        };

        const MgJob& jA = cta.job[MG_JOB_A];
        const MgJob& jB = cta.job[MG_JOB_B];
        const MgJob& jC = cta.job[MG_JOB_C];
        const MgJob& jD = cta.job[MG_JOB_D];
        // cell state of this thread's (row, unit) elements lives in registers for the whole loop (init_hidden: zeros)
        float cA[2] = {0.f, 0.f}, cC[2] = {0.f, 0.f};
        int unfinished = 1;   // worker thread 0 of a row CTA
        int ovf = 0;
        // scratch: [AH] atth | [AH] alpha_net weight | [64] scores | [64] mask | [256] score partials | [2][H] context partials | [128] tokens
        float* s_h = scr;
        float* s_w = scr + p.AH;
        float* s_e = scr + 2 * p.AH;
        float* s_mask = s_e + 64;
        float* s_sp = s_mask + 64;
        float* s_c = s_sp + 256;
        int* s_tok = reinterpret_cast<int*>(s_c + 2 * p.H);   // [128] tokens of the previous step, one poller per row
        if (has_row) {
            for (int j = wt; j < p.AH; j += MG_NW) s_w[j] = __ldg(p.alpha_w + j);
            if (wt < len_rt) s_mask[wt] = __ldg(p.masks + (size_t)row * p.len_stride + wt);
        }

        // LSTM cell on this CTA's share of a tile: gates = sum of the tile's split-K partials (split order) + addend.
        // Attention LSTM (is_att): the partials hold W_hh h_att(t-1) + W_ih[:, :H] h_lang(t-1) (contracted and drained during step t-1,
        // nothing at t = 0: init_hidden is zero); the addends are fc_pre (fc segment + both biases) and the token's row of xt_table
        // (= W_ih[:, 2H:] relu(E[it]), AttModel.py:332,410-413; it = <bos> = 0 at t = 0, AttModel.py:283-284), so the embedding segment
        // is a 16 KB row gather instead of a contraction that has to wait for the selection.
        auto cell = [&](const MgJob& jb, const float* part, unsigned* tile_cnt, int t, float (&cst)[2], bool is_att, uint8_t* xout, int ev) -> bool {
            const int nel = 128 * jb.u_n;
            const bool have_part = !is_att || t > 0;
            // nothing is requested before the polls: an acquire load is not served before the thread's earlier loads have landed
            // (measured: operands requested ahead of the wait delayed the poll by ~2 us)
            // nothing is requested before the polls: an acquire load is not served before the thread's earlier loads have landed
            // (measured: operands requested ahead of the wait delayed the poll by ~2 us)
            int tok[2] = {0, 0};
            if (is_att && t > 0) {
                // One worker thread per row spins on that row's token slot (the poll returns the token itself) and hands it to the other
                // threads through shared memory; worker thread 0 polls the tile's partial counter in the same loop.  (All 512 threads
                // polling, 148 CTAs on the same few lines, kept one L2 slice busy for microseconds: measured as late releases.)
                if (wt < 128) {
                    const int* slot = p.tok_slots + ((size_t)(t - 1) * 128 + wt) * MG_TOK_STRIDE;
                    const unsigned tile_target = (unsigned)jb.n_split * (unsigned)t;
                    int v = wt < S ? 0 : 1;
                    bool tile_ok = wt != 0;
                    for (uint32_t it = 1; v == 0 || !tile_ok; ++it) {
                        if (v == 0) v = ld_relaxed_s32(slot);
                        if (!tile_ok) tile_ok = ld_acquire(tile_cnt) >= tile_target;
                        if (v != 0 && tile_ok) break;
                        if (ctl->stop || ((it & 1023u) == 0 && wt_.expired(15))) { ctl->fail = 1; break; }
                    }
                    s_tok[wt] = v - 1;
                }
                worker_bar();
                if (ctl->fail) return false;
                tok[0] = tok[1] = s_tok[wt & 127];
            } else if (!is_att) {
                if (!w_counter(tile_cnt, (unsigned)jb.n_split * (unsigned)(t + 1), 9)) return false;
            }
            if (wt == 0) MG_STAMP(t, ev);
            // fc_pre (attention LSTM: fc segment + both biases), the language LSTM's bias (b_ih + b_hh), xt_table and the partials keep the
            // four gates of a unit adjacent: one 16-byte load per element (and split)
            float4 add[2], tab[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int e = wt + k * MG_NW;
                const int r = e & 127, u = jb.tile_u0 + jb.u_lo + (e >> 7);
                const bool live = e < nel && r < S;
                const float4* ap = is_att ? reinterpret_cast<const float4*>(p.fc_pre) + (size_t)r * H + u : reinterpret_cast<const float4*>(p.lang_b) + u;
                add[k] = live ? __ldg(ap) : make_float4(0.f, 0.f, 0.f, 0.f);
                tab[k] = (live && is_att) ? __ldg(reinterpret_cast<const float4*>(p.xt_table) + (size_t)tok[k] * H + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // every partial of both elements of this thread is requested before the first one is used (one L2 round trip);
            // slot (split z, unit ul, row r) is the float4 at z * 4096 + ul * 128 + r of the tile; n_split <= 4 (mega_plan)
            float acc2[2][4];
            {
                const float4* g4 = reinterpret_cast<const float4*>(part + jb.part_tile_off);
                float4 pz[2][4];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int e = wt + k * MG_NW;
                    const bool live = e < nel && (e & 127) < S && have_part;
                    const float4* g = g4 + (size_t)(jb.u_lo + (e >> 7)) * 128 + (e & 127);
#pragma unroll
                    for (int z = 0; z < 4; ++z) pz[k][z] = (live && z < jb.n_split) ? __ldcg(g + (size_t)z * 4096) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {   // split order
                    acc2[k][0] = ((pz[k][0].x + pz[k][1].x) + pz[k][2].x) + pz[k][3].x;
                    acc2[k][1] = ((pz[k][0].y + pz[k][1].y) + pz[k][2].y) + pz[k][3].y;
                    acc2[k][2] = ((pz[k][0].z + pz[k][1].z) + pz[k][2].z) + pz[k][3].z;
                    acc2[k][3] = ((pz[k][0].w + pz[k][1].w) + pz[k][2].w) + pz[k][3].w;
                }
            }
            if (kTrace && p.trace && wt == 0) {   // the stamp must not be taken before the operands have landed
                asm volatile("" ::"f"(acc2[0][0]), "f"(acc2[1][3]), "f"(tab[0].x), "f"(tab[1].w), "f"(add[0].x), "f"(add[1].w) : "memory");
                MG_STAMP(t, is_att ? 36 : 37);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int e = wt + k * MG_NW;
                if (e >= nel) continue;
                const int r = e & 127, ul = jb.u_lo + (e >> 7), u = jb.tile_u0 + ul;
                if (r >= S) continue;
                float acc[4] = {acc2[k][0], acc2[k][1], acc2[k][2], acc2[k][3]};
                const float a4[4] = {add[k].x, add[k].y, add[k].z, add[k].w}, t4[4] = {tab[k].x, tab[k].y, tab[k].z, tab[k].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = is_att ? (acc[q] + a4[q]) + t4[q] : acc[q] + a4[q];
                const float c = mg_sigmoid(acc[1]) * cst[k] + mg_sigmoid(acc[0]) * mg_tanh(acc[2]);
                cst[k] = c;
                x_store(xout, r, u, mg_sigmoid(acc[3]) * mg_tanh(c), ovf);
            }
            if (wt == 0) MG_STAMP(t, ev == 3 ? 17 : 19);
            return true;
        };

#define MG_WSTAMP(EV_) do { if (wt == 0) MG_STAMP(t, EV_); } while (0)
        for (int t = 0; t < T && alive; ++t) {
            MG_WSTAMP(0);
            // ---------------- attention LSTM: gates -> partial -> cell -> h_att(t)
            if (jA.present) {
                if (!cell(jA, p.partA, p.sync + MG_C_TILE_A + jA.tile * MG_TILE_STRIDE, t, cA, true, p.x_hatt[t & 1], 3)) break;
                w_signal(p.sync + MG_C_HATT, true);
                MG_WSTAMP(4);
            }
            // ---------------- h2att partial
            if (jB.present) {
                if (!epilogue(jB, p.partB + (size_t)jB.plane * 128 * p.ldB + jB.part_off, p.ldB, false, t, 5)) break;
                w_signal(p.sync + MG_C_B);
                MG_WSTAMP(6);
            }
            // ---------------- attention of this CTA's row (AttModel.py:445-471) -> ctx(t)
            if (has_row) {
                const int AH = p.AH, len_max = len_rt, lst = p.len_stride;
                const int AH4 = AH >> 2, Q = AH4 >> 5;          // AH % 128 == 0 (checked on the host)
                const int items = len_max * Q;                  // (node, 32-quad slice)
                const float4* pa4 = reinterpret_cast<const float4*>(p.p_att + (size_t)row * lst * AH);
                {   // att row -> L2 (the weight stream may have evicted it)
                    const char* af = reinterpret_cast<const char*>(p.att + (size_t)row * lst * H);
                    const size_t nb_a = (size_t)len_max * H * 4;
                    for (size_t o = (size_t)wt * 128; o < nb_a; o += (size_t)MG_NW * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(af + o));
                }
                for (int i0 = 0; i0 < items && alive; i0 += MG_ATT_ITEMS * 16) {
                    // item = (node, slice) = (item / Q, item % Q); with Q = 1, 2 or 4 slices per node the 16-item stride of a warp is a whole
                    // number of nodes: one division per pass instead of two per item (they were ~30 % of the scoring instructions)
                    const int it0 = i0 + ww, n0 = it0 / Q, q0 = it0 - n0 * Q, nstep = 16 / Q;
                    const bool whole = nstep * Q == 16;
                    float4 pv[MG_ATT_ITEMS];
#pragma unroll
                    for (int k = 0; k < MG_ATT_ITEMS; ++k) {   // p_att does not depend on the step: in flight during the wait below
                        const int item = it0 + k * 16;
                        const int node = whole ? n0 + k * nstep : item / Q, sl = whole ? q0 : item % Q;
                        pv[k] = item < items ? __ldg(pa4 + (size_t)node * AH4 + sl * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (i0 == 0) {
                        if (!w_counter(p.sync + MG_C_B, (unsigned)p.nB * (unsigned)(t + 1), 10)) { alive = false; break; }
                        MG_WSTAMP(7);
                        for (int j = wt; j < AH; j += MG_NW) {
                            float pz[8];
#pragma unroll
                            for (int z = 0; z < 8; ++z) pz[z] = z < p.zB ? __ldcg(p.partB + ((size_t)z * 128 + row) * p.ldB + j) : 0.f;
                            float a = 0.f;
#pragma unroll
                            for (int z = 0; z < 8; ++z) a += pz[z];
                            s_h[j] = a + __ldg(p.h2att_b + j);
                        }
                        worker_bar();
                        MG_WSTAMP(41);
                    }
                    const float4* h4 = reinterpret_cast<const float4*>(s_h);
                    const float4* w4 = reinterpret_cast<const float4*>(s_w);
#pragma unroll
                    for (int k = 0; k < MG_ATT_ITEMS; ++k) {
                        const int item = it0 + k * 16;
                        if (item >= items) continue;
                        const int j4 = (whole ? q0 : item % Q) * 32 + lane;
                        const float4 v = pv[k], hh = h4[j4], wv = w4[j4];
                        float a = wv.x * mg_tanh_score(v.x + hh.x);
                        a = fmaf(wv.y, mg_tanh_score(v.y + hh.y), a);
                        a = fmaf(wv.z, mg_tanh_score(v.z + hh.z), a);
                        a = fmaf(wv.w, mg_tanh_score(v.w + hh.w), a);
                        a = warp_sum(a);
                        if (lane == 0) s_sp[item] = a;
                    }
                }
                if (!alive) break;
                worker_bar();
                MG_WSTAMP(42);
                // the att rows of the context sum do not depend on the weights: the first nodes of this thread's column quad are requested
                // now and land during the softmax
                constexpr int kPre = 12;
                const float* af = p.att + (size_t)row * lst * H;
                const int grp = wt >> 8, tg = wt & 255, H4 = H >> 2;
                float4 pre[kPre];
#pragma unroll
                for (int i = 0; i < kPre; ++i) {
                    const int n = grp + 2 * i;
                    pre[i] = (tg < H4 && n < len_max) ? __ldg(reinterpret_cast<const float4*>(af + (size_t)n * H) + tg) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                // Every warp runs the softmax itself (lane l owns nodes l and l + 32; len_max <= 64) and keeps the weights in registers:
                // no barrier between scores and context, and no warp waits for another one's softmax (one warp ran it while fifteen
                // waited: 2.3 us).  softmax, then mask, then renormalise: two-stage as the reference (AttModel.py:462-466); ex2 / rcp units.
                float wgt[2];
                {
                    float e[2];
                    const float ab = __ldg(p.alpha_b);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int n = lane + 32 * i;
                        float acc = 0.f;
                        if (n < len_max)
                            for (int q = 0; q < Q; ++q) acc += s_sp[n * Q + q];
                        e[i] = acc + ab;
                    }
                    float m = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (lane + 32 * i < len_max) m = fmaxf(m, e[i]);
                    m = warp_max(m);
                    float ex[2] = {0.f, 0.f}, sum = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (lane + 32 * i < len_max) { ex[i] = __expf(e[i] - m); sum += ex[i]; }
                    sum = warp_sum(sum);
                    float msum = 0.f;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (lane + 32 * i < len_max) { ex[i] = __fdividef(ex[i], sum) * s_mask[lane + 32 * i]; msum += ex[i]; }
                    msum = warp_sum(msum);
#pragma unroll
                    for (int i = 0; i < 2; ++i) wgt[i] = (lane + 32 * i < len_max) ? __fdividef(ex[i], msum) : 0.f;
                }
                MG_WSTAMP(43);
                {   // context: two thread groups take interleaved node subsets, partials combined in fixed order.  The node index is the
                    // same in every lane of a warp: its weight comes from the owning lane by shuffle (outside any lane-dependent branch)
                    {
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < kPre; ++i) {
                            const int n = grp + 2 * i;   // < 32 + grp for every i < kPre <= 16
                            const float wv = __shfl_sync(0xffffffffu, n < 32 ? wgt[0] : wgt[1], n & 31);
                            if (n < len_max) { a.x = fmaf(wv, pre[i].x, a.x); a.y = fmaf(wv, pre[i].y, a.y); a.z = fmaf(wv, pre[i].z, a.z); a.w = fmaf(wv, pre[i].w, a.w); }
                        }
#pragma unroll 12
                        for (int n = grp + 2 * kPre; n < len_max; n += 2) {
                            const float wv = __shfl_sync(0xffffffffu, n < 32 ? wgt[0] : wgt[1], n & 31);
                            const float4 v = tg < H4 ? __ldg(reinterpret_cast<const float4*>(af + (size_t)n * H) + tg) : make_float4(0.f, 0.f, 0.f, 0.f);
                            a.x = fmaf(wv, v.x, a.x); a.y = fmaf(wv, v.y, a.y); a.z = fmaf(wv, v.z, a.z); a.w = fmaf(wv, v.w, a.w);
                        }
                        if (tg < H4) reinterpret_cast<float4*>(s_c + (size_t)grp * H)[tg] = a;
                    }
                    for (int j4b = 256; j4b < H4; j4b += 256) {   // H > 1024 only
                        const int j4 = j4b + tg;
                        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int n = grp; n < len_max; n += 2) {
                            const float wv = __shfl_sync(0xffffffffu, n < 32 ? wgt[0] : wgt[1], n & 31);
                            const float4 v = j4 < H4 ? __ldg(reinterpret_cast<const float4*>(af + (size_t)n * H) + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
                            a.x = fmaf(wv, v.x, a.x); a.y = fmaf(wv, v.y, a.y); a.z = fmaf(wv, v.z, a.z); a.w = fmaf(wv, v.w, a.w);
                        }
                        if (j4 < H4) reinterpret_cast<float4*>(s_c + (size_t)grp * H)[j4] = a;
                    }
                    worker_bar();
                    MG_WSTAMP(44);
                    for (int j = wt; j < H; j += MG_NW) x_store(p.x_ctx, row, j, s_c[j] + s_c[H + j], ovf);
                    MG_WSTAMP(18);
                }
                w_signal(p.sync + MG_C_CTX, true);
                MG_WSTAMP(8);
            }
            // ---------------- language LSTM -> h_lang(t)
            if (jC.present) {
                if (!epilogue(jC, p.partC + jC.part_off, 128, true, t, 9)) break;
                w_signal(p.sync + MG_C_TILE_C + jC.tile * MG_TILE_STRIDE);
                MG_WSTAMP(10);
                if (!cell(jC, p.partC, p.sync + MG_C_TILE_C + jC.tile * MG_TILE_STRIDE, t, cC, false, p.x_hlang[t & 1], 11)) break;
                w_signal(p.sync + MG_C_HLANG, true);
                MG_WSTAMP(12);
            }
            // ---------------- logit partial
            if (jD.present) {
                if (!epilogue(jD, p.partD + (size_t)jD.plane * 128 * p.ldD + jD.part_off, p.ldD, false, t, 13, false,
                              jD.plane == 0 ? p.logit_b + jD.part_off : nullptr, p.V1 - jD.part_off)) break;
                w_signal(p.sync + MG_C_D);
                MG_WSTAMP(14);
            }
            // ---------------- token selection of this CTA's row (AttModel.py:292-318) -> xt(t+1), on 12 of the 16 worker warps.
            // The other 4 (one per tensor-memory lane quarter) drain the attention LSTM accumulator of the NEXT step meanwhile (its
            // h_att(t) / h_lang(t) segments are contracted right after the logit tile): those partials are what cell A(t+1) would
            // otherwise wait for after the token is out (measured: published ~2 us after the token when all 16 warps drained it afterwards).
            const bool next_a = jA.present && t + 1 < T;
            if (has_row && wt < MG_NSEL) {
                // top-k: the uniform of (step, row) does not depend on the logits; every selection thread draws it ahead of the wait
                float u_row = 0.f;
                if (p.mode != 0) u_row = p.uniforms ? __ldg(p.uniforms + (size_t)t * p.S + row) : MgPhilox::uniform(p.seed, p.offset, (unsigned)t, (unsigned)row);
                if (wt == 0) ctl->flag[par & 1] = wt_.counter(p.sync + MG_C_D, (unsigned)p.nD * (unsigned)(t + 1), 11) ? 1 : 0;
                sel_bar();
                const bool seen = ctl->flag[par & 1] != 0;
                if (seen) {
                MG_WSTAMP(15);
                // logit partials of this row: thread wt owns the column quads wt + i * MG_NSEL; both planes (zD <= 2: mega_plan) are
                // requested together, 16 bytes per load.  Plane 0 carries the bias (added by its epilogue).
                const int V1 = p.V1;
                const float4* pl = reinterpret_cast<const float4*>(p.partD + (size_t)row * p.ldD);
                const size_t plane = (size_t)32 * p.ldD;   // float4 units
                const int nquad = p.ldD >> 2;   // ldD is a multiple of 16; columns V1 .. ldD - 1 are written (zero weights) but not selectable
                float v[MG_SELVALS];
                {
                    float4 a0[MG_SELQ], a1[MG_SELQ];
#pragma unroll
                    for (int i = 0; i < MG_SELQ; ++i) a0[i] = (wt + i * MG_NSEL < nquad) ? __ldcg(pl + wt + i * MG_NSEL) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < MG_SELQ; ++i)
                        a1[i] = (p.zD > 1 && wt + i * MG_NSEL < nquad) ? __ldcg(pl + plane + wt + i * MG_NSEL) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < MG_SELQ; ++i) {   // split order; a1 = +0 with one plane
                        v[4 * i] = a0[i].x + a1[i].x; v[4 * i + 1] = a0[i].y + a1[i].y; v[4 * i + 2] = a0[i].z + a1[i].z; v[4 * i + 3] = a0[i].w + a1[i].w;
                    }
                }
                MG_WSTAMP(45);
                // arg-max and log-sum-exp in ONE block reduction: every thread sums exp(v - its own max), the partial sums are rescaled
                // to the warp's and then to the row's maximum (the sum differs from the two-pass one by rounding only)
                float bv = -INFINITY;
                int bi = 0x7fffffff;
#pragma unroll
                for (int i = 0; i < MG_SELVALS; ++i) {
                    const int j = MG_SELCOL(wt, i);
                    if (j < V1 && (v[i] > bv || bi == 0x7fffffff)) { bv = v[i]; bi = j; }
                }
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < MG_SELVALS; ++i)
                    if (MG_SELCOL(wt, i) < V1) s += __expf(v[i] - bv);
                {
                    const float own = bv;
                    warp_argmax(bv, bi);
                    s = warp_sum(bi == 0x7fffffff || own == -INFINITY ? 0.f : s * __expf(own - bv));
                    if ((wt & 31) == 0) { ctl->red_f[wt >> 5] = bv; ctl->red_i[wt >> 5] = bi; ctl->red_s[wt >> 5] = s; }
                    sel_bar();
                    bv = ctl->red_f[0]; bi = ctl->red_i[0];
#pragma unroll
                    for (int w = 1; w < MG_NSEL / 32; ++w) argmax_combine(bv, bi, ctl->red_f[w], ctl->red_i[w]);
                    s = 0.f;
#pragma unroll
                    for (int w = 0; w < MG_NSEL / 32; ++w) s += ctl->red_i[w] == 0x7fffffff ? 0.f : ctl->red_s[w] * __expf(ctl->red_f[w] - bv);
                }
                const float m = bv;
                const float lz = logf(s);
                MG_WSTAMP(46);
                int tok;
                float lp;
                if (p.mode == 0) {
                    tok = bi;
                    lp = (m - m) - lz;
                } else {   // top-k sampling: q = log_softmax(logp / temp), keep the k best, Categorical over them (AttModel.py:296-303)
                    // the scaled log-probs are evaluated ONCE per element: v[i] <- (logp / temp) - max, then q = that - log(sum).  The
                    // division is a multiplication by 1 / temp and the exponential comes from the ex2 unit (<= 2 ulp each, far below
                    // the 2e-5 bar): the IEEE division + libdevice expf were ~1000 instructions per thread and step, straight-line code
                    // that ran from a cold instruction cache (measured 3.8 us for this loop)
                    const float rtemp = 1.f / p.temp;
                    const float ym = ((m - m) - lz) * rtemp;
                    float s2 = 0.f;
#pragma unroll
                    for (int i = 0; i < MG_SELVALS; ++i) {
                        v[i] = ((v[i] - m) - lz) * rtemp - ym;
                        if (MG_SELCOL(wt, i) < V1) s2 += __expf(v[i]);
                    }
                    MG_WSTAMP(38);
                    if (p.top_k <= 3) {
                        // The best of q is the row maximum found above (q is a monotone map of the logit; first index on ties): q = 0 - lz2.
                        // The next two come from ONE more block reduction, shared with the sum: every thread keeps its two best other
                        // elements, a warp merges them by shuffles, and every thread merges the 12 warps' pairs itself (comparisons
                        // before or after the - lz2 shift are the same).  Was: a sum and three arg-max reductions, nine barriers.
                        float c1v = -INFINITY, c2v = -INFINITY;
                        int c1i = 0x7fffffff, c2i = 0x7fffffff;
                        auto better = [](float av, int ai, float bv_, int bi_) { return av > bv_ || (av == bv_ && ai < bi_); };
                        auto offer = [&](float xv, int xi) {   // insert into the sorted pair
                            if (better(xv, xi, c1v, c1i)) { c2v = c1v; c2i = c1i; c1v = xv; c1i = xi; }
                            else if (better(xv, xi, c2v, c2i)) { c2v = xv; c2i = xi; }
                        };
#pragma unroll
                        for (int i = 0; i < MG_SELVALS; ++i) {
                            const int j = MG_SELCOL(wt, i);
                            if (j < V1 && j != bi) offer(v[i], j);
                        }
                        s2 = warp_sum(s2);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const float o1v = __shfl_xor_sync(0xffffffffu, c1v, o), o2v = __shfl_xor_sync(0xffffffffu, c2v, o);
                            const int o1i = __shfl_xor_sync(0xffffffffu, c1i, o), o2i = __shfl_xor_sync(0xffffffffu, c2i, o);
                            offer(o1v, o1i);
                            offer(o2v, o2i);
                        }
                        if ((wt & 31) == 0) {
                            const int w = wt >> 5;
                            ctl->red_s2[w] = s2; ctl->cand_v[0][w] = c1v; ctl->cand_i[0][w] = c1i; ctl->cand_v[1][w] = c2v; ctl->cand_i[1][w] = c2i;
                        }
                        sel_bar();
                        MG_WSTAMP(39);
                        s2 = 0.f;
#pragma unroll
                        for (int w = 0; w < MG_NSEL / 32; ++w) s2 += ctl->red_s2[w];
                        const float lz2 = logf(s2);
                        c1v = c2v = -INFINITY; c1i = c2i = 0x7fffffff;
#pragma unroll
                        for (int w = 0; w < MG_NSEL / 32; ++w) { offer(ctl->cand_v[0][w], ctl->cand_i[0][w]); offer(ctl->cand_v[1][w], ctl->cand_i[1][w]); }
                        const float tv[3] = {0.f - lz2, c1i == 0x7fffffff ? -INFINITY : c1v - lz2, c2i == 0x7fffffff ? -INFINITY : c2v - lz2};
                        const int ti[3] = {bi, c1i, c2i};
                        const int k = p.top_k;
                        const float u = u_row;
                        float den = 0.f;
#pragma unroll
                        for (int c = 0; c < 3; ++c) den += c < k ? expf(tv[c] - tv[0]) : 0.f;
                        float cdf = 0.f;
                        int pos = 0;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            if (c < k) {
                                cdf += expf(tv[c] - tv[0]) / den;
                                if (u >= cdf) pos = c + 1;
                            }
                        }
                        if (pos > k - 1) pos = k - 1;
                        tok = pos == 0 ? ti[0] : (pos == 1 ? ti[1] : ti[2]);
                        lp = pos == 0 ? tv[0] : (pos == 1 ? tv[1] : tv[2]);
                    } else {
                    s2 = workers_sum(s2, ctl, wt);
                    const float lz2 = logf(s2);
#pragma unroll
                    for (int i = 0; i < MG_SELVALS; ++i) v[i] -= lz2;
                    unsigned taken = 0;
                    for (int c = 0; c < p.top_k; ++c) {
                        float cv = -INFINITY;
                        int ci = 0x7fffffff;
#pragma unroll
                        for (int i = 0; i < MG_SELVALS; ++i) {
                            const int j = MG_SELCOL(wt, i);
                            if (j >= V1 || ((taken >> i) & 1u)) continue;
                            const float qv = v[i];
                            if (qv > cv || ci == 0x7fffffff) { cv = qv; ci = j; }
                        }
                        workers_argmax(cv, ci, ctl, wt);
                        if (ci != 0x7fffffff && ((ci >> 2) % MG_NSEL) == wt) taken |= 1u << ((((ci >> 2) / MG_NSEL) << 2) | (ci & 3));
                        if (wt == 0) { ctl->topv[c] = cv; ctl->topi[c] = ci; }
                    }
                    sel_bar();
                    const int k = p.top_k;
                    const float u = u_row;
                    float den = 0.f;
                    for (int c = 0; c < k; ++c) den += expf(ctl->topv[c] - ctl->topv[0]);
                    float cdf = 0.f;
                    int pos = 0;
                    for (int c = 0; c < k; ++c) {
                        cdf += expf(ctl->topv[c] - ctl->topv[0]) / den;
                        if (u >= cdf) pos = c + 1;
                    }
                    if (pos > k - 1) pos = k - 1;
                    tok = ctl->topi[pos];
                    lp = ctl->topv[pos];
                    }
                }
                if (wt == 0) {
                    const int unf = (t == 0 ? 1 : unfinished) && (tok > 0);
                    unfinished = unf;
                    const long long it = unf ? tok : 0;
                    // the cells of step t + 1 spin on this; the slot carries the token itself, nothing else has to be visible with it
                    st_relaxed_s32(p.tok_slots + ((size_t)t * 128 + row) * MG_TOK_STRIDE, (int)it + 1);
                    p.seq[(size_t)row * T + t] = it;
                    p.seq_lp[(size_t)row * T + t] = lp;
                }
                {   // the cells gather this token's xt_table row next: on its way into L2 while the token is published and polled for
                    const size_t row_bytes = (size_t)4 * H * sizeof(float);
                    const char* tr = reinterpret_cast<const char*>(p.xt_table) + (size_t)tok * row_bytes;
                    if (wt == 32) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(tr), "r"((uint32_t)row_bytes) : "memory");
                }
                MG_WSTAMP(47);
                sel_bar();
                if (wt == 0) red_release(p.sync + MG_C_XT, 1u + (unfinished ? 0x10000u : 0u));
                MG_WSTAMP(16);
                } else if (wt == 0) ctl->fail = 1;
            }
            // next-step accumulator: the 4 drain warps of a row CTA (beside the selection), all 16 worker warps elsewhere
            if (next_a && (!has_row || wt >= MG_NSEL)) {
                if (epilogue(jA, p.partA + jA.part_off, 128, true, t, 1, has_row)) {
                    if (has_row) drain_bar(); else worker_bar();
                    if (wt == (has_row ? MG_NSEL : 0)) { red_release(p.sync + MG_C_TILE_A + jA.tile * MG_TILE_STRIDE, 1u); MG_STAMP(t, 2); }
                } else if (wt == (has_row ? MG_NSEL : 0)) ctl->fail = 1;
            } else if (next_a) par ^= 4u << jA.set;
            if (has_row) par ^= 1u;   // the selection group used flag[par & 1]
            worker_bar();   // the groups meet again; a give-up of either ends the loop for all
            if (ctl->fail) break;
        }
        if (ovf && p.overflow) atomicOr(p.overflow, 1);
    }
    __syncthreads();
    if (tid == 0) {   // a time-out anywhere is reported as a negative step count (site * 1000 + CTA + 1)
        const unsigned ab = ld_acquire(p.sync + MG_C_ABORT);
        if (ab != 0) p.steps_done[0] = -(int)ab;
    }
    if (warp == 3) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// [rows, 4H] gate-major (row q * H + u of an LSTM weight) -> [rows, H, 4]: the cell reads the four gates of a unit with one 16-byte load
// b_a / b_b (nullable): two gate-major bias vectors [4H]; their sum goes, interleaved the same way, to dst row `rows` (H more float4)
__global__ void __launch_bounds__(256) gate_interleave_kernel(const float* __restrict__ src, float4* __restrict__ dst, int rows, int H,
                                                              const float* __restrict__ b_a, const float* __restrict__ b_b) {
    const size_t n = (size_t)rows * H + (b_a ? H : 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / H;
        const int u = (int)(i - r * H);
        if (r < (size_t)rows) {
            const float* s = src + r * 4 * H + u;
            dst[i] = make_float4(s[0], s[H], s[2 * (size_t)H], s[3 * (size_t)H]);
        } else {
            dst[i] = make_float4(b_a[u] + b_b[u], b_a[H + u] + b_b[H + u], b_a[2 * H + u] + b_b[2 * H + u], b_a[3 * H + u] + b_b[3 * H + u]);
        }
    }
}
static void launch_gate_interleave(const float* src, float* dst, int rows, int H, cudaStream_t st, const float* b_a = nullptr, const float* b_b = nullptr) {
    const size_t n = (size_t)rows * H + (b_a ? H : 0);
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)kNumSMs * 8);
    gate_interleave_kernel<<<blocks, 256, 0, st>>>(src, reinterpret_cast<float4*>(dst), rows, H, b_a, b_b);
}

// ---- stream pack ---------------------------------------------------------------------------------------------------------------
// One block = the [n_rows x 64] weight tile of one k-block as the tensor core wants it in shared memory: hi tile then lo tile, rows of
// 128 bytes, 16-byte chunks XOR-swizzled by (row & 7).  The blocks of a CTA follow each other in the order the CTA contracts them.
struct MgBlockDesc {
    const float* src; int ld;
    int gate;          // 1: LSTM tile, the four gates of a unit adjacent (row r = unit u0 + r / 4, gate r % 4 -> source row gate * H + unit), 0: row r -> r0 + r
    int r0, U, Hrows;  // plain: first source row and number of source rows; gate: first unit, units in the tile, H
    int n_rows;
    int col0, ncols;   // source columns [col0, col0 + ncols) are valid (the rest of the 64 is zero)
    unsigned long long dst;
};
__global__ void __launch_bounds__(256) mega_pack_kernel(const MgBlockDesc* __restrict__ descs, uint8_t* __restrict__ out, int* __restrict__ overflow) {
    const MgBlockDesc d = descs[blockIdx.x];
    uint8_t* hi = out + d.dst;
    uint8_t* lo = hi + (size_t)d.n_rows * 128;
    int ovf = 0;
    for (int idx = threadIdx.x; idx < d.n_rows * 64; idx += blockDim.x) {
        const int r = idx >> 6, c = idx & 63;
        long long srow;
        bool valid;
        if (d.gate) {
            const int q = r & 3, u = d.r0 + (r >> 2);
            valid = (r >> 2) < d.U && u < d.Hrows;
            srow = (long long)q * d.Hrows + u;
        } else {
            valid = d.r0 + r < d.Hrows;
            srow = d.r0 + r;
        }
        valid = valid && c < d.ncols;
        unsigned short h = 0, l = 0;
        if (valid) split_f16(d.src[srow * d.ld + d.col0 + c], h, l, ovf);
        const size_t o = (size_t)r * 128 + (size_t)((((c >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
        *reinterpret_cast<unsigned short*>(hi + o) = h;
        *reinterpret_cast<unsigned short*>(lo + o) = l;
    }
    if (ovf && overflow) atomicOr(overflow, 1);
}

// ---- schedule ------------------------------------------------------------------------------------------------------------------
struct MgPlan {
    bool ok = false;
    int n_cta = 0, kbH = 0, H = 0, T = 0;
    int zL = 0, tilesL = 0, nL = 0;
    int zB = 0, tilesB = 0, nB = 0, ldB = 0;
    int zD = 0, tilesD = 0, nD = 0, ldD = 0;
    std::vector<MgCta> ctas;
    std::vector<MgBlockDesc> blocks;   // src pointers filled by mega_fill_sources
    std::vector<int> block_src;        // source matrix id per block
    unsigned long long w_bytes = 0;
};
enum { MG_SRC_ATT_IH = 0, MG_SRC_ATT_HH, MG_SRC_LANG_IH, MG_SRC_LANG_HH, MG_SRC_H2ATT, MG_SRC_LOGIT };

static void split_groups(int groups, int tiles, std::vector<int>& g0, std::vector<int>& gn) {   // as even as possible
    g0.resize(tiles); gn.resize(tiles);
    const int b = groups / tiles, rem = groups % tiles;
    int at = 0;
    for (int i = 0; i < tiles; ++i) { gn[i] = b + (i < rem ? 1 : 0); g0[i] = at; at += gn[i]; }
}

static MgPlan mega_plan(const subgc_dims* d, int n_cta) {
    MgPlan pl;
    const int H = d->rnn, E = d->enc, AH = d->att_hid, V1 = d->vocab1;
    if (n_cta < 1 || n_cta > 1024 || H < 4 || (H & 3) || E < 1 || AH < 128 || (AH & 127) || AH > 512 || V1 < 2 || V1 > MG_SELVALS * MG_NSEL) return pl;
    if ((size_t)(2 * AH + 64 + 64 + 256 + 2 * H + 128) * 4 > MG_SCR_BYTES) return pl;
    pl.n_cta = n_cta;
    const int kbH = (H + 63) / 64;
    pl.kbH = kbH; pl.H = H; pl.T = d->seq_length;
    // LSTM contractions: tiles of <= 7 groups of 4 hidden units (x 4 gates = <= 112 weight rows), zL k-splits per tile
    const int G = H / 4;
    int zL = 0, tilesL = 0;
    for (int z = 4; z >= 1; z >>= 1) {   // <= 4: the cell requests every partial of its elements in one round trip
        if (z > kbH) continue;
        const int tl = std::min(G, n_cta / z);
        if (tl < 1 || (G + tl - 1) / tl > 7) continue;
        const int umax = 4 * ((G + tl - 1) / tl);
        if ((umax + z - 1) / z > 8) continue;   // the cell keeps <= 2 elements per worker thread: <= 8 units x 128 rows per CTA
        zL = z; tilesL = tl;
        break;
    }
    if (!zL || tilesL > MG_MAX_TILES) return pl;
    pl.zL = zL; pl.tilesL = tilesL; pl.nL = zL * tilesL;
    // h2att: tiles of <= 7 groups of 16 rows
    const int GB = (AH + 15) / 16;
    int zB = 1;
    while (zB * 2 <= 8 && zB * 2 <= kbH) zB *= 2;
    int tilesB = (GB + 6) / 7;
    while (zB > 1 && tilesB * zB > n_cta) zB >>= 1;
    if (tilesB * zB > n_cta) return pl;
    pl.zB = zB; pl.tilesB = tilesB; pl.nB = zB * tilesB; pl.ldB = GB * 16;
    // logit: tiles of <= 9 groups of 16 rows
    const int GD = (V1 + 15) / 16;
    int zD = 0, tilesD = 0;
    for (int z = 2; z >= 1; z >>= 1) {   // <= 2: the selection requests both planes of its row in one round trip
        if (z > kbH) continue;
        const int td = std::min(GD, n_cta / z);
        if (td < 1 || (GD + td - 1) / td > 9) continue;
        zD = z; tilesD = td;
        break;
    }
    if (!zD) return pl;
    pl.zD = zD; pl.tilesD = tilesD; pl.nD = zD * tilesD; pl.ldD = GD * 16;

    std::vector<int> gL0, gLn, gB0, gBn, gD0, gDn;
    split_groups(G, tilesL, gL0, gLn);
    split_groups(GB, tilesB, gB0, gBn);
    split_groups(GD, tilesD, gD0, gDn);
    pl.ctas.assign(n_cta, MgCta());
    for (auto& c : pl.ctas) memset(&c, 0, sizeof(MgCta));
    auto kb_range = [](int kb, int z, int Z, int& a, int& n) { a = z * kb / Z; n = (z + 1) * kb / Z - a; };
    unsigned long long w_at = 0;
    for (int c = 0; c < n_cta; ++c) {
        MgCta& ct = pl.ctas[c];
        ct.w_off = w_at;
        int nt = 0;
        // CTAs that contract the same k-blocks start at different ones (rot): they would otherwise all ask L2 for the same activation
        // tile at the same moment
        auto add_task = [&](int src_id, int rot_seed, int gate, int r0, int U, int Hrows, int n_rows, int seg_col0, int seg_cols, int kb0, int nkb, int x_src,
                            int set, int flags) {
            if (nkb <= 0) return;
            MgTask& tk = ct.task[nt++];
            tk.n_blk = nkb; tk.n_rows = n_rows; tk.x_src = x_src; tk.x_kb0 = kb0; tk.acc_set = set; tk.flags = flags;
            tk.rot = rot_seed % nkb;
            for (int b = 0; b < nkb; ++b) {
                MgBlockDesc bd;
                memset(&bd, 0, sizeof(bd));
                bd.gate = gate; bd.r0 = r0; bd.U = U; bd.Hrows = Hrows; bd.n_rows = n_rows;
                const int k0 = (kb0 + (b + tk.rot) % nkb) * 64;
                bd.col0 = seg_col0 + k0;
                bd.ncols = std::max(0, std::min(64, seg_cols - k0));
                bd.dst = w_at;
                pl.blocks.push_back(bd);
                pl.block_src.push_back(src_id);
                w_at += (unsigned long long)n_rows * 256;
            }
        };
        const bool hasL = c < pl.nL;
        const int tileL = hasL ? c / zL : 0, zl = hasL ? c % zL : 0;
        const int bidx = c - (n_cta - pl.nB);
        const bool hasB = bidx >= 0;
        const bool hasD = c < pl.nD;
        int U = 0, u0 = 0, nrL = 0, ha = 0, hn = 0;
        if (hasL) {
            U = 4 * gLn[tileL]; u0 = 4 * gL0[tileL]; nrL = 4 * U;
            kb_range(kbH, zl, zL, ha, hn);
        }
        // issue order of a step (see the header): C_hlang(t) | B(t) | C_hatt(t) | A_hatt(t+1) | C_ctx(t) | D(t) | A_hlang(t+1)
        if (hasL) add_task(MG_SRC_LANG_HH, tileL, 1, u0, U, H, nrL, 0, H, ha, hn, MG_X_HLANG_PREV, 1, MG_F_START);
        if (hasB) {
            const int tb = bidx / zB, zb = bidx % zB;
            int a, n;
            kb_range(kbH, zb, zB, a, n);
            add_task(MG_SRC_H2ATT, tb, 0, 16 * gB0[tb], 0, AH, 16 * gBn[tb], 0, H, a, n, MG_X_HATT, 0, MG_F_START | MG_F_COMMIT);
            MgJob& jb = ct.job[MG_JOB_B];
            jb.present = n > 0; jb.set = 0; jb.n_rows = 16 * gBn[tb]; jb.part_off = 16 * gB0[tb]; jb.plane = zb;
        }
        if (hasL) add_task(MG_SRC_LANG_IH, tileL, 1, u0, U, H, nrL, H, H, ha, hn, MG_X_HATT, 1, 0);
        if (hasL) add_task(MG_SRC_ATT_HH, tileL, 1, u0, U, H, nrL, 0, H, ha, hn, MG_X_HATT, 0, MG_F_START | MG_F_NEXT);
        if (hasL) add_task(MG_SRC_LANG_IH, tileL, 1, u0, U, H, nrL, 0, H, ha, hn, MG_X_CTX, 1, MG_F_COMMIT);
        if (hasD) {
            const int td = c / zD, zd = c % zD;
            int a, n;
            kb_range(kbH, zd, zD, a, n);
            add_task(MG_SRC_LOGIT, td, 0, 16 * gD0[td], 0, V1, 16 * gDn[td], 0, H, a, n, MG_X_HLANG, 1, MG_F_START | MG_F_COMMIT);
            MgJob& jb = ct.job[MG_JOB_D];
            jb.present = n > 0; jb.set = 1; jb.n_rows = 16 * gDn[td]; jb.part_off = 16 * gD0[td]; jb.plane = zd;
        }
        if (hasL) add_task(MG_SRC_ATT_IH, tileL, 1, u0, U, H, nrL, 0, H, ha, hn, MG_X_HLANG, 0, MG_F_NEXT | MG_F_COMMIT);
        ct.n_task = nt;
        ct.step_bytes = (int)(w_at - ct.w_off);
        if (hasL) {
            for (int which = 0; which < 2; ++which) {
                MgJob& jb = ct.job[which == 0 ? MG_JOB_A : MG_JOB_C];
                jb.present = 1; jb.set = which; jb.n_rows = nrL;
                jb.part_off = (tileL * zL + zl) * 16384; jb.plane = zl;
                jb.tile = tileL; jb.n_split = zL; jb.tile_u0 = u0; jb.tile_units = U;
                jb.u_lo = zl * U / zL; jb.u_n = (zl + 1) * U / zL - jb.u_lo;
                jb.part_tile_off = tileL * zL * 16384;
            }
        }
    }
    pl.w_bytes = w_at;
    pl.ok = true;
    return pl;
}

static size_t mega_table_bytes(const MgPlan& pl) {
    return align_up(pl.ctas.size() * sizeof(MgCta), 1024) + align_up(pl.blocks.size() * sizeof(MgBlockDesc), 1024);
}
// the pack buffer: schedule tables | weight stream | xt_table [V1, 4H] fp32 | workspace of the contraction that builds the table
constexpr int MG_XT_CHUNK = 1024;   // token rows per contraction when xt_table is built
static size_t mega_table_off(const MgPlan& pl) { return align_up((size_t)pl.w_bytes, 1024); }   // from the start of the weight stream
static size_t mega_xt_bytes(const subgc_dims* d) { return align_up((size_t)d->vocab1 * 4 * d->rnn * sizeof(float), 1024); }
static size_t mega_xt_tmp_bytes(const subgc_dims* d) { return align_up((size_t)MG_XT_CHUNK * 4 * d->rnn * sizeof(float), 1024); }   // gate-major chunk
static size_t mega_xt_ws_bytes(const subgc_dims* d) { return mega_xt_tmp_bytes(d) + align_up(gemm_workspace_bytes(MG_XT_CHUNK, 4 * d->rnn, d->enc), 1024) + 1024; }
static size_t mega_pack_total(const subgc_dims* d, const MgPlan& pl) {
    return mega_table_bytes(pl) + mega_table_off(pl) + mega_xt_bytes(d) + mega_xt_ws_bytes(d);
}

struct MgScratch {   // per-call device scratch (inside the decode workspace)
    uint8_t *x_ctx, *x_hatt[2], *x_hlang[2];
    int* tok_slots;
    float *partA, *partC, *partB, *partD, *fc_il;
    unsigned* sync;
    size_t zero_bytes;   // bytes from x_ctx that must be zero at launch (activation tiles + counters)
};
static size_t mega_scratch_bytes(const MgPlan& pl) {
    size_t b = 0;
    b += 5 * align_up((size_t)pl.kbH * MG_XTILE_BYTES, 1024) + MG_C_TOTAL * 4 + (size_t)pl.T * 128 * MG_TOK_STRIDE * 4;
    b += 2 * align_up((size_t)pl.nL * 16384 * 4, 1024);
    b += align_up((size_t)pl.zB * 128 * pl.ldB * 4, 1024) + align_up((size_t)pl.zD * 128 * pl.ldD * 4, 1024);
    b += align_up((size_t)129 * 4 * pl.H * 4, 1024);   // gate-interleaved copy of fc_pre + the interleaved language-LSTM bias
    return b + 2048;
}
static bool mega_take_scratch(const MgPlan& pl, Workspace& ws, MgScratch& sc) {
    const size_t xh = align_up((size_t)pl.kbH * MG_XTILE_BYTES, 1024);
    uint8_t* z = ws.take<uint8_t>(5 * xh + MG_C_TOTAL * 4 + (size_t)pl.T * 128 * MG_TOK_STRIDE * 4 + 1024);
    if (!z) return false;
    z = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(z), 1024));
    sc.x_ctx = z; sc.x_hatt[0] = sc.x_ctx + xh; sc.x_hatt[1] = sc.x_hatt[0] + xh; sc.x_hlang[0] = sc.x_hatt[1] + xh;
    sc.x_hlang[1] = sc.x_hlang[0] + xh;
    sc.sync = reinterpret_cast<unsigned*>(sc.x_hlang[1] + xh);
    sc.tok_slots = reinterpret_cast<int*>(sc.sync + MG_C_TOTAL);
    sc.zero_bytes = 5 * xh + MG_C_TOTAL * 4 + (size_t)pl.T * 128 * MG_TOK_STRIDE * 4;
    sc.partA = ws.take<float>((size_t)pl.nL * 16384);
    sc.partC = ws.take<float>((size_t)pl.nL * 16384);
    sc.partB = ws.take<float>((size_t)pl.zB * 128 * pl.ldB);
    sc.partD = ws.take<float>((size_t)pl.zD * 128 * pl.ldD);
    sc.fc_il = ws.take<float>((size_t)129 * 4 * pl.H);
    return ws.ok();
}

static const MgPlan& cached_plan(const subgc_dims* d, int n_cta) {   // the schedule is a pure function of (dims, CTA count)
    struct Key { int H, E, AH, V1, n, T; };
    static thread_local Key key{0, 0, 0, 0, 0, 0};
    static thread_local MgPlan plan;
    if (!(key.H == d->rnn && key.E == d->enc && key.AH == d->att_hid && key.V1 == d->vocab1 && key.n == n_cta && key.T == d->seq_length)) {
        plan = mega_plan(d, n_cta);
        key = Key{d->rnn, d->enc, d->att_hid, d->vocab1, n_cta, d->seq_length};
    }
    return plan;
}

// debug (SUBGC_MEGA_TRACE=1): per-CTA, per-step time stamps of the most recent launch, read back by subgc_debug_mega_trace
static unsigned long long* mega_trace_buffer(size_t* elems) {
    static unsigned long long* buf = nullptr;
    static int on = -1;
    static const size_t n = (size_t)256 * 32 * MG_TRACE_EVENTS;
    if (on < 0) {
        on = getenv("SUBGC_MEGA_TRACE") != nullptr ? 1 : 0;
        if (on) { cudaMalloc(&buf, n * sizeof(unsigned long long)); cudaMemset(buf, 0, n * sizeof(unsigned long long)); }
    }
    if (elems) *elems = n;
    return buf;
}

struct MgTiming { bool on = false, pending = false; cudaEvent_t e0 = nullptr, e1 = nullptr; };
static MgTiming& mega_timing() { static thread_local MgTiming t; return t; }

bool mega_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SUBGC_MEGA"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}

size_t mega_decode_scratch_bytes(const subgc_dims* d, const subgc_weights* w) {
    if (!w || !w->mega || w->mega_ctas <= 0) return 0;
    const MgPlan& pl = cached_plan(d, w->mega_ctas);
    return pl.ok ? mega_scratch_bytes(pl) : 0;
}
// worst case over CTA counts, for workspace queries that have no weights at hand
size_t mega_decode_scratch_bytes_max(const subgc_dims* d) {
    const MgPlan& pl = cached_plan(d, kNumSMs);
    return pl.ok ? mega_scratch_bytes(pl) + (1u << 20) : 0;
}

bool mega_decode_eligible(const subgc_dims* d, const subgc_weights* w, int S, int len_max, const float* att_weights) {
    if (!mega_enabled() || !w || !w->mega || w->mega_ctas <= 0 || att_weights != nullptr) return false;
    if (S < 1 || S > 128 || S > w->mega_ctas || len_max < 1 || len_max > 64) return false;
    const MgPlan& pl = cached_plan(d, w->mega_ctas);
    return pl.ok && w->mega_bytes >= mega_pack_total(d, pl);
}

// The loop of subgc_decode_sample as one cooperative launch.  fc_pre: [S, 4H] hoisted fc segment + both biases (launch_fc_pre).
int launch_mega_decode(const subgc_dims* d, const subgc_weights* w, int S, int len_max, int mode, float temp, int top_k, uint64_t seed, uint64_t offset,
                       const float* uniforms, const float* fc_pre, const float* att, const float* p_att, const float* masks, int64_t* seq,
                       float* seq_lp, int32_t* steps_done, Workspace& ws, cudaStream_t st, const int32_t* counts) {
    const MgPlan& pl = cached_plan(d, w->mega_ctas);
    MgScratch sc;
    if (!mega_take_scratch(pl, ws, sc)) { set_error("subgc_decode_sample: workspace too small for the persistent decode kernel"); return SUBGC_E_WORKSPACE; }
    SUBGC_CUDA(cudaMemsetAsync(sc.x_ctx, 0, sc.zero_bytes, st));
    MgParams p;
    memset(&p, 0, sizeof(p));
    const uint8_t* mb = static_cast<const uint8_t*>(w->mega);
    p.ctas = reinterpret_cast<const MgCta*>(mb);
    p.wstream = mb + mega_table_bytes(pl);
    p.n_cta = pl.n_cta; p.S = S; p.T = d->seq_length; p.len_max = len_max; p.H = d->rnn; p.AH = d->att_hid; p.V1 = d->vocab1;
    p.kbH = pl.kbH; p.nL = pl.nL; p.nB = pl.nB; p.nD = pl.nD; p.zB = pl.zB; p.zD = pl.zD; p.ldB = pl.ldB; p.ldD = pl.ldD;
    // S is the row capacity when `counts` decides the real count; row S of the copy is the language LSTM's bias (b_ih + b_hh)
    launch_gate_interleave(fc_pre, sc.fc_il, S, d->rnn, st, w->lang_b_ih, w->lang_b_hh);
    SUBGC_LAUNCH_CHECK();
    p.fc_pre = sc.fc_il; p.att = att; p.p_att = p_att; p.masks = masks;
    p.h2att_b = w->h2att.b; p.alpha_w = w->alpha_net.w; p.alpha_b = w->alpha_net.b; p.logit_b = w->logit.b;
    p.xt_table = reinterpret_cast<const float*>(p.wstream + mega_table_off(pl));
    p.lang_b = sc.fc_il + (size_t)S * 4 * d->rnn;
    p.x_ctx = sc.x_ctx; p.x_hatt[0] = sc.x_hatt[0]; p.x_hatt[1] = sc.x_hatt[1]; p.x_hlang[0] = sc.x_hlang[0]; p.x_hlang[1] = sc.x_hlang[1];
    p.partA = sc.partA; p.partC = sc.partC; p.partB = sc.partB; p.partD = sc.partD; p.sync = sc.sync; p.tok_slots = sc.tok_slots;
    p.seq = reinterpret_cast<long long*>(seq); p.seq_lp = seq_lp; p.steps_done = steps_done; p.overflow = w->h3_overflow;
    p.mode = mode; p.temp = temp; p.top_k = top_k; p.seed = seed; p.offset = offset; p.uniforms = uniforms;
    p.counts = counts; p.len_stride = len_max;
    p.trace = (pl.n_cta <= 256 && d->seq_length <= 32) ? mega_trace_buffer(nullptr) : nullptr;
    static const unsigned long long timeout_ns = [] {
        const char* e = getenv("SUBGC_MEGA_TIMEOUT_S");
        const double s = e ? atof(e) : 0.0;
        return (unsigned long long)((s > 0.0 ? s : 4.0) * 1e9);
    }();
    p.timeout_ns = timeout_ns;
    static DeviceOnce once;
    SUBGC_CUDA(once.run([]() {
        cudaError_t e = cudaFuncSetAttribute(mega_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
        return e != cudaSuccess ? e : cudaFuncSetAttribute(mega_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM);
    }));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.n_cta); cfg.blockDim = dim3(MG_THREADS); cfg.dynamicSmemBytes = MG_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;   // every CTA must be resident: they wait for each other
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // bench / profiling (subgc_mega_timing): CUDA events around this launch alone, on the launching stream (not while capturing a graph)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    MgTiming& tm = mega_timing();
    const bool timed = tm.on && cap == cudaStreamCaptureStatusNone;
    if (timed) {
        if (!tm.e0) { cudaEventCreate(&tm.e0); cudaEventCreate(&tm.e1); }
        cudaEventRecord(tm.e0, st);
    }
    if (p.trace) SUBGC_CUDA(cudaLaunchKernelEx(&cfg, mega_decode_kernel<true>, p));
    else SUBGC_CUDA(cudaLaunchKernelEx(&cfg, mega_decode_kernel<false>, p));
    if (timed) { cudaEventRecord(tm.e1, st); tm.pending = true; }
    SUBGC_LAUNCH_CHECK();
    return SUBGC_OK;
}

}  // namespace subgc

using namespace subgc;

extern "C" int subgc_debug_mega_trace(unsigned long long* host_out, int n_cta, int n_steps) {
    size_t n = 0;
    unsigned long long* buf = mega_trace_buffer(&n);
    if (!buf || !host_out || (size_t)n_cta * n_steps * MG_TRACE_EVENTS > n) return SUBGC_E_INVALID;
    cudaDeviceSynchronize();
    return cudaMemcpy(host_out, buf, (size_t)n_cta * n_steps * MG_TRACE_EVENTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? SUBGC_OK : SUBGC_E_CUDA;
}

extern "C" int subgc_mega_timing(int on, float* last_ms) {
    MgTiming& tm = mega_timing();
    if (last_ms) {
        *last_ms = -1.f;
        if (tm.pending && tm.e0 && tm.e1 && cudaEventSynchronize(tm.e1) == cudaSuccess) cudaEventElapsedTime(last_ms, tm.e0, tm.e1);
        tm.pending = false;
    }
    tm.on = on != 0;
    return SUBGC_OK;
}

extern "C" size_t subgc_mega_pack_bytes(const subgc_dims* d, int n_cta) {
    if (!d) return 0;
    const MgPlan& pl = cached_plan(d, n_cta);
    return pl.ok ? mega_pack_total(d, pl) : 0;
}

extern "C" int subgc_mega_pack(const subgc_dims* d, const subgc_weights* w, int n_cta, void* buf, size_t bytes, int32_t* overflow, subgc_stream_t stream) {
    SUBGC_CHECK_ARG(d && w && buf, "subgc_mega_pack: null argument");
    const MgPlan& pl = cached_plan(d, n_cta);
    SUBGC_CHECK_ARG(pl.ok, "subgc_mega_pack: these dimensions are not supported by the persistent decode kernel");
    const size_t tb = mega_table_bytes(pl);
    SUBGC_CHECK_ARG(bytes >= mega_pack_total(d, pl), "subgc_mega_pack: buffer too small (%zu < %zu)", bytes, mega_pack_total(d, pl));
    SUBGC_CHECK_ARG((reinterpret_cast<uintptr_t>(buf) & 1023) == 0, "subgc_mega_pack: buffer must be 1024-byte aligned");
    const int H = d->rnn, E = d->enc;
    const float* srcs[6] = {w->att_w_ih, w->att_w_hh, w->lang_w_ih, w->lang_w_hh, w->h2att.w, w->logit.w};
    const int lds[6] = {E + 2 * H, H, 2 * H, H, H, H};
    std::vector<MgBlockDesc> blocks = pl.blocks;
    for (size_t i = 0; i < blocks.size(); ++i) { blocks[i].src = srcs[pl.block_src[i]]; blocks[i].ld = lds[pl.block_src[i]]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* b = static_cast<uint8_t*>(buf);
    const size_t cta_bytes = align_up(pl.ctas.size() * sizeof(MgCta), 1024);
    // tables: plain (pageable) host memory, so the copies are staged before these calls return
    SUBGC_CUDA(cudaMemcpyAsync(b, pl.ctas.data(), pl.ctas.size() * sizeof(MgCta), cudaMemcpyHostToDevice, st));
    SUBGC_CUDA(cudaMemcpyAsync(b + cta_bytes, blocks.data(), blocks.size() * sizeof(MgBlockDesc), cudaMemcpyHostToDevice, st));
    mega_pack_kernel<<<(unsigned)blocks.size(), 256, 0, st>>>(reinterpret_cast<const MgBlockDesc*>(b + cta_bytes), b + tb, overflow);
    SUBGC_LAUNCH_CHECK();
    // xt_table[v] = W_ih[:, 2H:2H+E] relu(E[v]) (AttModel.py:332,410): the embedding segment of the attention LSTM for every token, so
    // that the decode loop gathers a row instead of contracting the segment after every selection
    float* table = reinterpret_cast<float*>(b + tb + mega_table_off(pl));
    float* tmp = reinterpret_cast<float*>(b + tb + mega_table_off(pl) + mega_xt_bytes(d));
    void* gws = reinterpret_cast<uint8_t*>(tmp) + mega_xt_tmp_bytes(d);
    const size_t gws_bytes = mega_xt_ws_bytes(d) - mega_xt_tmp_bytes(d) - 1024;
    for (int v0 = 0; v0 < d->vocab1; v0 += MG_XT_CHUNK) {
        GemmProblem gp;
        gp.wts = w;
        gp.M = std::min(MG_XT_CHUNK, d->vocab1 - v0); gp.N = 4 * H; gp.nseg = 1;
        gp.seg[0] = make_seg(w->embed + (size_t)v0 * E, E, w->att_w_ih + 2 * H, E + 2 * H, E);
        gp.seg[0].relu_a = 1;
        gp.C = tmp; gp.ldc = 4 * H;
        gp.overflow = overflow;
        SUBGC_TRY(launch_gemm(gp, gws, gws_bytes, st));
        launch_gate_interleave(tmp, table + (size_t)v0 * 4 * H, gp.M, H, st);
        SUBGC_LAUNCH_CHECK();
    }
    return SUBGC_OK;
}
