// Internal helpers shared by the subgc_b200 translation units (not part of the C ABI).
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "subgc_b200.h"

namespace subgc {

void set_error(const char* fmt, ...);
void count_launch();  // process-wide tally of kernels launched through this library (subgc_launch_count)

#define SUBGC_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            subgc::set_error(__VA_ARGS__);         \
            return SUBGC_E_INVALID;                \
        }                                          \
    } while (0)

#define SUBGC_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            subgc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SUBGC_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define SUBGC_LAUNCH_CHECK()                                                                  \
    do {                                                                                      \
        subgc::count_launch();                                                                \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            subgc::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return SUBGC_E_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define SUBGC_TRY(expr)            \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != SUBGC_OK) return rc__; \
    } while (0)

constexpr int kNumSMs = 148;  // B200

// Function attributes (dynamic shared-memory limit, carve-out) are per device.  DeviceOnce::run(f) runs f() once per device for its
// call site and remembers the device only when f succeeded; threads driving the same device (DataParallel-style) wait for the one
// that sets the attributes instead of launching before they are applied.
struct DeviceOnce {
    std::mutex mu;
    std::atomic<bool> done[64];
    DeviceOnce() { for (auto& d : done) d.store(false); }
    template <typename F>
    cudaError_t run(F&& f) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return f();
        if (done[dev].load(std::memory_order_acquire)) return cudaSuccess;
        std::lock_guard<std::mutex> lock(mu);
        if (done[dev].load(std::memory_order_relaxed)) return cudaSuccess;
        const cudaError_t e = f();
        if (e == cudaSuccess) done[dev].store(true, std::memory_order_release);
        return e;
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller-provided workspace.
struct Workspace {
    char* base;
    size_t size;
    size_t used;
    Workspace(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t off = align_up(used, 256);
        size_t bytes = count * sizeof(T);
        if (base == nullptr || off + bytes > size) {
            used = size + 1;  // poison
            return nullptr;
        }
        used = off + bytes;
        return reinterpret_cast<T*>(base + off);
    }
    bool ok() const { return used <= size; }
    size_t remaining() const { size_t off = align_up(used, 256); return off >= size ? 0 : size - off; }
    void* cursor() { return base + align_up(used, 256); }
};

// ---- GEMM building block -----------------------------------------------------------------------------
// C[M,N] = epilogue( sum over segments  A_seg[M,K_seg] . W_seg[N,K_seg]^T ).  Segments let the decoder contract
// the concatenated LSTM input [h_lang | fc | relu(E[it])] (+ h_att through weight_hh) straight from the caller's
// tensors and the reference's un-repacked weight_ih / weight_hh, with no concat buffer.
struct GemmSeg {
    const float* A;          // [rows, K] row-major, leading dim lda
    const float* W;          // [N, K] row-major (nn.Linear layout), leading dim ldw
    const long long* gather; // optional: A row index per output row (e.g. token ids into the embedding table)
    int lda, ldw, K;
    int a_row_div;           // A row = m / a_row_div (rows shared by consecutive outputs; 1 = identity)
    int relu_a;              // apply ReLU to A elements on load (embed + ReLU)
    const int* gather32;     // optional 32-bit variant of `gather` (sub-graph selections)
    // split-fp16 operands of the h3 tensor-core path (h3_gemm.cu): v = hi + lo * 2^-11, both fp16.  W16_* come from a packed copy
    // of the weight tensor (subgc_pack_weight, resolved by resolve_packs); A16_* are optional pre-split activations written by
    // the producer kernel (rows 1:1 with the output rows), otherwise the launcher splits A itself.
    const unsigned short* W16_hi = nullptr;   // first k-block plane of this segment, [kb][w16_rows][32] (k-block-major)
    const unsigned short* W16_lo = nullptr;
    int w16_rows = 0;                         // rows per k-block plane of the packed tensor
    const unsigned short* A16_hi = nullptr;
    const unsigned short* A16_lo = nullptr;
    int lda16 = 0;
};

inline GemmSeg make_seg(const float* A, int lda, const float* W, int ldw, int K) {
    GemmSeg s;
    s.A = A; s.W = W; s.gather = nullptr; s.lda = lda; s.ldw = ldw; s.K = K; s.a_row_div = 1; s.relu_a = 0; s.gather32 = nullptr;
    return s;  // the split-fp16 fields default to null
}

struct GemmEpilogue {
    const float* bias = nullptr;    // [N]
    const float* bias2 = nullptr;   // [N] second bias (LSTM b_ih + b_hh)
    const float* addend = nullptr;  // optional [*, ld_add] rows added before activation
    const long long* add_gather = nullptr;  // addend row index per output row (nullptr: row m)
    const int* add_gather32 = nullptr;
    int ld_add = 0;
    float div = 0.f;                // != 0: divide by it (GCN mean with count 1: 1 + 1e-7 in fp32)
    int relu = 0;
    int accumulate = 0;             // C += result (weight-gradient accumulation); applied last
    // rows are grouped in blocks of `group` rows; row (g, j) is valid iff j < group_len[g]; invalid rows are
    // written as exact zeros (pack_padded_sequence semantics).  group_sel maps block -> index into group_len.
    const int* group_len = nullptr;
    const int* group_sel = nullptr;
    int group = 0;
    // optional split-fp16 copy of the result (v = hi + lo * 2^-11), [M, ld16], for a consumer contraction that reads it as its
    // activation operand (GemmSeg::A16_*): written by the h3 contraction's own epilogue when it finishes the tile itself, otherwise
    // by one split pass over C after the contraction (launch_gemm guarantees it is there either way)
    unsigned short* c16_hi = nullptr;
    unsigned short* c16_lo = nullptr;
    int ld16 = 0;
};

struct GemmProblem {
    int M = 0, N = 0;
    int nseg = 0;
    GemmSeg seg[4];
    GemmEpilogue epi;
    float* C = nullptr;
    int ldc = 0;
    const int* active = nullptr;  // optional device flag: kernel exits immediately when *active == 0
    const subgc_weights* wts = nullptr;  // when set, segments whose weight has a packed copy there take the split-fp16 path
    int* overflow = nullptr;             // device flag raised when an activation saturates in the fp16 split (from wts->h3_overflow)
};

size_t gemm_workspace_bytes(int M, int N, int Ktotal);
// Launches the contraction (+ split-K reduction when used).  ws may be null when gemm_workspace_bytes == 0.
int launch_gemm(const GemmProblem& p, void* ws, size_t ws_bytes, cudaStream_t stream);
// raw_part != nullptr: leave the per-split partial sums [splits][M][N] there (the consumer reduces them in z order)
// and report the split count through *out_splits; no epilogue is applied.
int launch_gemm_ex(const GemmProblem& p, float* raw_part, size_t raw_part_elems, int* out_splits, void* ws, size_t ws_bytes,
                   cudaStream_t stream);
// tensor-core (tcgen05, split-precision TF32) variant of the same contraction, umma_gemm.cu
bool tc_eligible(const GemmProblem& p);
size_t tc_workspace_bytes(int M, int N, int Ktotal);
struct RawPartials { const float* part; int splits; };  // [splits][M][N] partial sums, to be added in z order
int launch_gemm_tc(const GemmProblem& p, void* ws, size_t ws_bytes, cudaStream_t stream, RawPartials* raw = nullptr);
// Fused LSTM-cell epilogue of the h3 contraction (gates [S, 4H] never reach memory): the CTA tile is gate-grouped (4 x 32 weight
// rows: gates i, f, g, o of 32 hidden units), the k-splits of a tile form a thread-block cluster that reduces its partial tiles
// through distributed shared memory in split order, and each CTA applies the cell to its share of the units.
struct CellEpilogue {
    int H = 0;
    const float* c_prev = nullptr;        // [*, H]
    const long long* parent = nullptr;    // nullable: previous-state row per output row (beam re-ordering)
    const float* addend = nullptr;        // nullable [S / add_div, 4H]: pre-computed gate term incl. both biases
    int add_div = 1;
    const float* b_ih = nullptr;          // used when addend == nullptr
    const float* b_hh = nullptr;
    float* h_out = nullptr;               // [S, H]
    float* c_out = nullptr;               // [S, H]
    unsigned short* h16_hi = nullptr;     // nullable split-fp16 copy of h', [S, Hp]
    unsigned short* h16_lo = nullptr;
    int Hp = 0;
};
// gates = p (N must be 4H, no epilogue fields) -> LSTM cell.  *fused = false when the problem does not qualify (caller falls back)
int launch_gemm_cell(const GemmProblem& p, const CellEpilogue& cell, void* ws, size_t ws_bytes, cudaStream_t stream, bool* fused);
// tensor-core (tcgen05, split-fp16 "h3") variant for weights that have a packed copy, h3_gemm.cu
bool h3_eligible(const GemmProblem& p);
int launch_gemm_h3(const GemmProblem& p, void* ws, size_t ws_bytes, cudaStream_t stream, RawPartials* raw = nullptr, bool* wrote_c16 = nullptr);
// fp32 rows [M, K] (leading dim lda) -> split-fp16 hi / lo [M, ld16] (ld16 % 8 == 0, >= K)
int launch_split_rows(const float* A, int M, int K, int lda, unsigned short* hi, unsigned short* lo, int ld16, int* overflow, cudaStream_t stream);
// fills W16_hi / W16_lo of every segment whose weight pointer lies inside a packed tensor of `w` (no-op without packs)
void resolve_packs(GemmProblem& p, const subgc_weights* w);
// Splits plain activation rows A [M, K] once into hi / lo [M, K rounded up to 8] (taken from `ws`) so that several contractions can
// share the copy through GemmSeg::A16_*; returns false (and takes nothing) when the split-fp16 path is not in use for `w`
bool h3_presplit(const float* A, int M, int K, int lda, const subgc_weights* w, Workspace& ws, cudaStream_t stream, const unsigned short** hi,
                 const unsigned short** lo, int* ld16);
void* tc_encode_fn();  // cuTensorMapEncodeTiled through the runtime's driver entry point (nullptr when unavailable)
// contraction without epilogue: the partial sums stay in `ws` for a consumer kernel that reduces them itself
int launch_gemm_raw(const GemmProblem& p, void* ws, size_t ws_bytes, cudaStream_t stream, RawPartials* raw);
// upper bound of splits * M * N floats for launch_gemm_ex raw partials
size_t gemm_partial_elems(int M, int N, int Ktotal);

// ---- persistent decode kernel (mega_decode.cu): the greedy / top-k loop as one cooperative launch ---------------------------
bool mega_decode_eligible(const subgc_dims* d, const subgc_weights* w, int S, int len_max, const float* att_weights);
size_t mega_decode_scratch_bytes_max(const subgc_dims* d);
int launch_mega_decode(const subgc_dims* d, const subgc_weights* w, int S, int len_max, int mode, float temp, int top_k, uint64_t seed, uint64_t offset,
                       const float* uniforms, const float* fc_pre, const float* att, const float* p_att, const float* masks, int64_t* seq,
                       float* seq_lp, int32_t* steps_done, Workspace& ws, cudaStream_t st, const int32_t* counts = nullptr);

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// The decode loop is a chain of ~6 dependent kernels per token, each of them short: launch latency, grid drain and kernel
// prologues are a third of the step.  Kernels launched through launch_pdl may start while their predecessor in the stream is
// still running (as soon as every block of it has passed pdl_trigger()); they do their input-independent prologue (barrier /
// TMEM set-up, weight-tile prefetch) and then block in pdl_wait() until the predecessor has completed and its writes are
// visible.  RULE: a kernel launched with launch_pdl must call pdl_wait() before it reads anything another kernel wrote and
// before it writes anything another kernel may still read.  Without the launch attribute both instructions are no-ops.
bool pdl_enabled();  // SUBGC_NO_PDL=1 switches the attribute off (plain stream order)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- timeline tracing (debug, SUBGC_TRACE=1): [first block start, last block end] of every launch of the decode loop ----------
struct TraceSlot { unsigned long long* buf; int seq; };   // buf == nullptr in normal operation
TraceSlot next_trace_slot(int kernel_id);                  // host: hands out consecutive slots (api.cu)
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_begin(const TraceSlot& t) {
    if (t.buf != nullptr && threadIdx.x == 0) {
        const unsigned long long now = globaltimer_ns();
        atomicMin(&t.buf[8 * t.seq], now);
        atomicMax(&t.buf[8 * t.seq + 7], now);   // last block to start
    }
}
__device__ __forceinline__ void trace_end(const TraceSlot& t) {
    if (t.buf != nullptr && threadIdx.x == 0) atomicMax(&t.buf[8 * t.seq + 1], globaltimer_ns());
}
__device__ __forceinline__ void trace_mark(const TraceSlot& t, int i) {   // extra per-kernel marks 0..3: last block to reach the point
    if (t.buf != nullptr) atomicMax(&t.buf[8 * t.seq + 4 + i], globaltimer_ns());
}
__device__ __forceinline__ void trace_released(const TraceSlot& t) {   // right after pdl_wait(): [first, last] block released
    if (t.buf != nullptr && threadIdx.x == 0) {
        const unsigned long long now = globaltimer_ns();
        atomicMin(&t.buf[8 * t.seq + 2], now);
        atomicMax(&t.buf[8 * t.seq + 3], now);
    }
}
#endif

// ---- device helpers ------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float block_sum(float v, float* red /*[32]*/) {
    // all threads of the block must call; returns the total to every thread (fixed reduction order)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red /*[32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float t = red[0];
    for (int i = 1; i < nw; ++i) t = fmaxf(t, red[i]);
    return t;
}
// (value, index) arg-max with first-index tie-break, the semantics of torch.max(dim) the reference relies on.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, i, o);
        argmax_combine(v, i, ov, oi);
    }
}
__device__ __forceinline__ void block_argmax(float& v, int& i, float* redv /*[32]*/, int* redi /*[32]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    warp_argmax(v, i);
    __syncthreads();
    if (lane == 0) { redv[wid] = v; redi[wid] = i; }
    __syncthreads();
    v = redv[0]; i = redi[0];
    for (int w = 1; w < nw; ++w) argmax_combine(v, i, redv[w], redi[w]);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// fp32 -> split fp16 (h3 operands): v = hi + lo * 2^-11.  Values beyond the fp16 range saturate and raise ovf.
constexpr float kH3LoScale = 2048.f;
constexpr float kH3LoInv = 1.f / 2048.f;
__device__ __forceinline__ void split_f16(float v, unsigned short& hi, unsigned short& lo, int& ovf) {
    const float c = fminf(fmaxf(v, -65504.f), 65504.f);
    if (!(c == v)) ovf = 1;                      // out of range or NaN
    const __half h = __float2half_rn(c);
    const float r = (c - __half2float(h)) * kH3LoScale;   // exact in fp32
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(r));
}
__device__ __forceinline__ void split_f16_store(float v, unsigned short* hi, unsigned short* lo, size_t idx, int* overflow = nullptr) {
    int ovf = 0;
    unsigned short h, l;
    split_f16(v, h, l, ovf);
    hi[idx] = h;
    lo[idx] = l;
    if (ovf && overflow) atomicOr(overflow, 1);
}
// bias / addend / divide / ReLU / zero-padding part of the contraction epilogue (accumulate-into-C is applied by the caller)
__device__ __forceinline__ float epilogue_apply(const GemmEpilogue& e, float v, int m, int n) {
    if (e.bias) v += __ldg(e.bias + n);
    if (e.bias2) v += __ldg(e.bias2 + n);
    if (e.addend) {
        long long r = e.add_gather ? e.add_gather[m] : (e.add_gather32 ? (long long)e.add_gather32[m] : (long long)m);
        v += __ldg(e.addend + r * e.ld_add + n);
    }
    // division by 1 is exact; never hand the compiler a speculative x / 0 (it if-converts the branch and every element
    // then takes the slow-path division subroutine: measured 12x slower epilogues)
    v = v / (e.div != 0.f ? e.div : 1.f);
    if (e.relu) v = fmaxf(v, 0.f);
    if (e.group) {
        int g = m / e.group, j = m - g * e.group;
        int len = e.group_len[e.group_sel ? e.group_sel[g] : g];
        if (j >= len) v = 0.f;
    }
    return v;
}

#endif

// locate sub-graph s in the loader tensors [rows, 2, per_half, N]: returns the flat (row*2+half)*per_half+g.
__host__ __device__ inline int subgraph_slot(const subgc_subgraph_layout& l, int s, int* image) {
    int half, row, g;
    if (l.order == 0) {
        g = s % l.per_half;
        int t = s / l.per_half;
        row = t % l.rows;
        half = t / l.rows;
    } else {
        g = s % l.per_half;
        int t = s / l.per_half;
        half = t % 2;
        row = (t / 2) * l.seq_per_img;
    }
    if (image) *image = row / l.seq_per_img;
    return (row * 2 + half) * l.per_half + g;
}
inline int subgraph_count(const subgc_subgraph_layout& l) {
    return l.order == 0 ? 2 * l.rows * l.per_half : 2 * (l.rows / l.seq_per_img) * l.per_half;
}

}  // namespace subgc
