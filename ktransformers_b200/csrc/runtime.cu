// Host-side plumbing shared by all translation units: error string, launch counter, SM count.
#include <cstdarg>
#include <cstdio>
#include <atomic>

#include "common.cuh"

namespace ktb {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int num_sms(int device) {
    static int cached[64] = {0};
    if (device < 0 || device >= 64) device = 0;
    if (!cached[device]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
        cached[device] = n;
    }
    return cached[device];
}

}  // namespace ktb

extern "C" {
const char* ktb200_last_error(void) { return ktb::g_err; }
const char* ktb200_version(void) { return "ktb200 0.1 (sm_100a)"; }
long ktb200_type_size(int t) { return ktb::type_size(t); }
long ktb200_blck_size(int t) { return ktb::blck_size(t); }
unsigned long long ktb200_launch_count(void) { return ktb::g_launches.load(); }
}
