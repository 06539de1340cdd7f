// Error reporting, launch accounting and device queries shared by every entry point of libsdfg.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace sdfg {

static thread_local char g_last_error[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs backward on its own threads

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(SDFG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return SDFG_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        cached = p.multiProcessorCount;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace sdfg

extern "C" {
const char* sdfg_last_error(void) { return sdfg::g_last_error; }
int sdfg_version(void) { return 100; }
int64_t sdfg_launch_count(void) { return sdfg::g_launches.load(); }
void sdfg_launch_count_reset(void) { sdfg::g_launches.store(0); }
}
