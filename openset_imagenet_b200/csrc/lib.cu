// Library-level entry points: version, error string, device checks.
#include "osb_common.cuh"
#include <string.h>

namespace osb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct DevInfo { int checked; int rc; int sms; };
static DevInfo g_dev[64];

static int probe(int dev) {
    DevInfo& d = g_dev[dev];
    if (d.checked) return d.rc;
    int major = 0, minor = 0, sms = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        set_error("cannot query CUDA device %d", dev);
        d.rc = OSB_ERR_ARCH;
    } else if (major != 10) {
        set_error("libosb is built for sm_100a (B200) only; device %d is sm_%d%d and there is no fallback path",
                  dev, major, minor);
        d.rc = OSB_ERR_ARCH;
    } else {
        d.rc = OSB_OK;
    }
    d.sms = sms;
    d.checked = 1;
    return d.rc;
}

int check_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        set_error("no usable CUDA device (%s); libosb has no CPU fallback", cudaGetErrorName(e));
        return OSB_ERR_ARCH;
    }
    // Bind the device's primary context to THIS thread once: driver entry points called without any runtime call
    // before them (cuTensorMapEncodeTiled from an autograd worker thread whose allocations all came from PyTorch's
    // cache) otherwise fail with CUDA_ERROR_INVALID_CONTEXT.
    static thread_local int bound_dev = -1;
    if (bound_dev != dev) {
        cudaFree(nullptr);
        bound_dev = dev;
    }
    return probe(dev);
}

int cta_limit(int max_ctas) {
    const int sms = sm_count();
    return (max_ctas > 0 && max_ctas < sms) ? max_ctas : sms;
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    probe(dev);
    return g_dev[dev].sms;
}

}  // namespace osb

extern "C" int osb_abi_version(void) { return OSB_ABI_VERSION; }
extern "C" const char* osb_last_error(void) { return osb::g_err; }
extern "C" int osb_device_check(void) { return osb::check_device(); }
extern "C" int osb_sm_count(void) { return osb::sm_count(); }
