// Error reporting and version of the subgc_b200 C ABI.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace subgc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// process-wide: the backward pass of a training step runs on autograd's worker thread
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SUBGC_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    return on == 1;
}
constexpr int kTraceSlots = 4096;
static unsigned long long* g_trace_buf = nullptr;
static int g_trace_ids[kTraceSlots];
static int g_trace_n = 0;
TraceSlot next_trace_slot(int kernel_id) {
    static int on = -1;
    if (on < 0) {
        on = getenv("SUBGC_TRACE") != nullptr ? 1 : 0;
        if (on) cudaMalloc(&g_trace_buf, kTraceSlots * 8 * sizeof(unsigned long long));
    }
    if (!on || g_trace_buf == nullptr) return TraceSlot{nullptr, 0};
    const int seq = g_trace_n % kTraceSlots;
    g_trace_ids[seq] = kernel_id;
    ++g_trace_n;
    return TraceSlot{g_trace_buf, seq};
}
}  // namespace subgc

extern "C" const char* subgc_last_error(void) { return subgc::g_err; }
extern "C" int subgc_version(void) { return SUBGC_ABI_VERSION; }
extern "C" unsigned long long subgc_launch_count(void) { return subgc::g_launches.load(std::memory_order_relaxed); }

/* debugging aid (SUBGC_TRACE=1): op 0 = restart slot numbering (call before the launch sequence to be traced / captured),
 * op 1 = reset the time stamps (before a replay), op 2 = copy out [n][8] stamps and the kernel ids; returns slots in use */
extern "C" int subgc_debug_trace(int op, unsigned long long* stamps, int* ids, int n) {
    using namespace subgc;
    if (op == 0) { g_trace_n = 0; return 0; }
    if (g_trace_buf == nullptr) return -1;
    if (op == 1) {
        static unsigned long long init[kTraceSlots * 8];
        for (int i = 0; i < kTraceSlots; ++i) {
            for (int j = 0; j < 8; ++j) init[8 * i + j] = 0;
            init[8 * i] = ~0ull; init[8 * i + 2] = ~0ull;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(g_trace_buf, init, sizeof(init), cudaMemcpyHostToDevice);
        return 0;
    }
    cudaDeviceSynchronize();
    const int m = g_trace_n < n ? g_trace_n : n;
    cudaMemcpy(stamps, g_trace_buf, (size_t)m * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    for (int i = 0; i < m; ++i) ids[i] = g_trace_ids[i];
    return m;
}
