// lib.cu -- library-wide state of libtgn_b200.so: error text, launch counter, device facts.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"
#include "tgn_b200.h"

namespace tgn {
namespace {
thread_local char g_err[512] = "";
std::atomic<int> g_launches{0};
}  // namespace

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what)
{
    g_launches.fetch_add(1, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return TGN_ERR_CUDA;
    }
    return TGN_OK;
}

// Stream-ordered scratch (cudaMallocAsync) is used by the bucket FPS and the ball query; keep freed
// blocks in the device's default pool instead of returning them to the OS at every synchronisation.
void keep_async_pool()
{
    static bool done = false;
    if (done) return;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    (void)cudaGetLastError();
    done = true;
}

int sm_count()
{
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;   // B200
    }
    return cached;
}

}  // namespace tgn

extern "C" {
int tgn_version(void) { return 100; }
const char* tgn_last_error(void) { return tgn::g_err; }
int tgn_launch_count(void) { return tgn::g_launches.load(std::memory_order_relaxed); }
}
