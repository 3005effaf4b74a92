// api.cu -- library-level entry points of the C ABI (include/rcg.h): version, error string,
// device discovery, launch accounting.
#include <atomic>
#include <cstring>

#include "rcg_host.h"

namespace rcg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        set_error("no CUDA device available (%s); librcg_b200 has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return RCG_ENODEV;
    }
    return 0;
}

}  // namespace rcg

extern "C" {

int rcg_version(void) { return RCG_VERSION; }

const char *rcg_last_error_string(void) { return rcg::g_err; }

int rcg_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        rcg::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return RCG_ENODEV;
    }
    return n;
}

int rcg_dim_state(int32_t sys_id) { return rcg::sys_n(sys_id); }
int rcg_dim_input(int32_t sys_id) { return rcg::sys_m(sys_id); }
int rcg_dim_critic(int32_t cs, int32_t n, int32_t m) { return rcg::dim_critic_c(cs, n, m); }

int64_t rcg_launch_count(void) { return rcg::g_launches.load(); }
void    rcg_reset_launch_count(void) { rcg::g_launches.store(0); }

}  // extern "C"
