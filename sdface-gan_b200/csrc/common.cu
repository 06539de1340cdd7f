// Error reporting, launch accounting and device queries shared by every entry point of libsdfg.so.
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>
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

struct ProfRec { cudaEvent_t a, b; };
static std::mutex g_prof_mu;
static std::vector<ProfRec*> g_prof_recs;
static std::atomic<int> g_prof_on{0};
static char g_prof_filter[64] = "";

ProfScope::ProfScope(const char* tag, cudaStream_t s) : rec(nullptr), st(s) {
    if (!g_prof_on.load(std::memory_order_relaxed) || !strstr(tag, g_prof_filter)) return;
    ProfRec* r = new ProfRec;
    if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->a, st);
    rec = r;
}
ProfScope::~ProfScope() {
    if (!rec) return;
    ProfRec* r = (ProfRec*)rec;
    cudaEventRecord(r->b, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back(r);
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

void sdfg_prof_enable(int on, const char* tag_substring) {
    using namespace sdfg;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    snprintf(g_prof_filter, sizeof(g_prof_filter), "%s", tag_substring ? tag_substring : "");
    g_prof_on.store(on ? 1 : 0);
}

int sdfg_prof_collect(double* total_ms, int64_t* launches) {
    using namespace sdfg;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double tot = 0;
    int64_t n = 0;
    for (ProfRec* r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r->b) == cudaSuccess && cudaEventElapsedTime(&ms, r->a, r->b) == cudaSuccess) { tot += ms; n++; }
        cudaEventDestroy(r->a);
        cudaEventDestroy(r->b);
        delete r;
    }
    g_prof_recs.clear();
    if (total_ms) *total_ms = tot;
    if (launches) *launches = n;
    return SDFG_OK;
}
}
