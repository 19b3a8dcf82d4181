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
static thread_local unsigned long long g_launches = 0;
void count_launch() { ++g_launches; }
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SUBGC_NO_PDL"); on = (e && e[0] == '1') ? 0 : 1; }
    return on == 1;
}
}  // namespace subgc

extern "C" const char* subgc_last_error(void) { return subgc::g_err; }
extern "C" int subgc_version(void) { return SUBGC_ABI_VERSION; }
extern "C" unsigned long long subgc_launch_count(void) { return subgc::g_launches; }
