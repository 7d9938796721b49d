// common.cu — see common.cuh.
#include "common.cuh"

namespace xrb {

std::atomic<uint64_t> g_launches{0};
static thread_local std::string t_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
}

int select_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error("no CUDA device visible (%s); xrsfm_b200 has no CPU fallback",
                  e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
        cudaGetLastError();
        return XRB_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device %d out of range (0..%d)", device, n - 1);
        return XRB_ERR_INVALID;
    }
    cudaDeviceProp prop;
    XRB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library ships sm_100a code only", device,
                  prop.major, prop.minor);
        return XRB_ERR_NO_DEVICE;
    }
    XRB_CUDA(cudaSetDevice(device));
    return XRB_OK;
}

int DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return XRB_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    XRB_CUDA(cudaMalloc(&p, bytes));
    cap = bytes;
    return XRB_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

}  // namespace xrb

extern "C" {
int xrb_abi_version(void) { return XRB_ABI_VERSION; }
const char *xrb_last_error(void) { return xrb::t_error.c_str(); }
uint64_t xrb_kernel_launch_count(void) { return xrb::g_launches.load(); }
}
