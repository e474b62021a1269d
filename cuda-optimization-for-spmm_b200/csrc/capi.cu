// capi.cu -- error string, launch counter, device queries of the C ABI.
#include "common.cuh"

#include <map>
#include <mutex>
#include <stdarg.h>
#include <utility>
#include <string.h>

namespace cuspmm_b200 {

static thread_local char g_err[512] = "";
static thread_local unsigned long long g_launches = 0;

int set_error(int status, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

void count_launch(unsigned n) { g_launches += n; }

struct DevProps { int sms = 0; size_t l2 = 0; };

static DevProps props() {      // cached per device ordinal (any ordinal: a map, not a fixed table)
    static std::mutex mu;
    static std::map<int, DevProps> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(dev);
    if (it != cache.end()) return it->second;
    DevProps p;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) p.sms = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev) == cudaSuccess) p.l2 = (size_t)v;
    if (p.sms <= 0) p.sms = 148;
    cache[dev] = p;
    return p;
}
int sm_count() { return props().sms; }

int set_smem_once_impl(const void *kern, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    std::lock_guard<std::mutex> lock(mu);
    auto it = done.find({kern, dev});
    if (it != done.end() && it->second >= bytes) return (int)cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done[{kern, dev}] = bytes;
    return (int)e;
}
size_t l2_bytes() { return props().l2; }

} // namespace cuspmm_b200

using namespace cuspmm_b200;

extern "C" int cuspmm_version(void) { return CUSPMM_B200_VERSION; }
extern "C" const char *cuspmm_last_error(void) { return g_err; }
extern "C" unsigned long long cuspmm_launch_count(void) { return g_launches; }
extern "C" void cuspmm_reset_launch_count(void) { g_launches = 0; }

extern "C" int cuspmm_device_count(int *count) {
    CUSPMM_REQUIRE(count, "null pointer");
    *count = 0;
    CUSPMM_CUDA(cudaGetDeviceCount(count));
    return CUSPMM_OK;
}

extern "C" int cuspmm_device_info(int device, int *sms, int *major, int *minor, size_t *l2, size_t *mem) {
    cudaDeviceProp p;
    CUSPMM_CUDA(cudaGetDeviceProperties(&p, device));
    if (sms) *sms = p.multiProcessorCount;
    if (major) *major = p.major;
    if (minor) *minor = p.minor;
    if (l2) *l2 = (size_t)p.l2CacheSize;
    if (mem) *mem = p.totalGlobalMem;
    return CUSPMM_OK;
}

extern "C" int cuspmm_host_alloc(void **ptr, size_t bytes) {
    CUSPMM_REQUIRE(ptr, "null pointer");
    CUSPMM_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return CUSPMM_OK;
}
extern "C" int cuspmm_host_free(void *ptr) {
    if (ptr) CUSPMM_CUDA(cudaFreeHost(ptr));
    return CUSPMM_OK;
}
