// common.cu -- error state, device info (host side of libkvq_sm100.so).
#include <cstdarg>
#include <cstdio>

#include "kvq_common.cuh"

namespace kvq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(KVQ_ERR_CUDA, "%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return KVQ_OK;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached = n;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace kvq

extern "C" int kvq_version(void) { return KVQ_VERSION; }

extern "C" const char* kvq_last_error(void) { return kvq::g_err; }

extern "C" int kvq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0, n = 0, ma = 0, mi = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return kvq::fail(KVQ_ERR_NO_DEVICE, "no CUDA device");
    KVQ_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    KVQ_CUDA(cudaDeviceGetAttribute(&ma, cudaDevAttrComputeCapabilityMajor, dev));
    KVQ_CUDA(cudaDeviceGetAttribute(&mi, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = ma;
    if (cc_minor) *cc_minor = mi;
    if (ma != 10) return kvq::fail(KVQ_ERR_NO_DEVICE, "libkvq_sm100 is built for sm_100a only; device is sm_%d%d", ma, mi);
    return KVQ_OK;
}
