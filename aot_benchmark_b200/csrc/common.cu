// Error reporting + version for libaotb200.so.
#include "common.cuh"
#include <cstdarg>
#include <cstring>

namespace aotb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launches(int n) { g_launches += (unsigned long long)n; }
static bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
}  // namespace aotb

// enable / disable programmatic dependent launch for every kernel of the library (default off)
extern "C" void aotb_set_pdl(int on) { aotb::g_pdl = on != 0; }

// number of kernels this library has launched in this process (bench.py's gpu_launches)
extern "C" unsigned long long aotb_launch_count(void) { return aotb::g_launches; }

extern "C" const char* aotb_last_error_string(void) { return aotb::g_err; }
extern "C" int aotb_version(void) { return 100; }  // 0.1.0
extern "C" const char* aotb_arch(void) { return "sm_100a"; }
