// common.cuh — error plumbing, launch accounting and small device helpers shared by the
// matcher and bundle-adjustment translation units of libxrsfm_b200.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/xrsfm_b200.h"

namespace xrb {

extern std::atomic<uint64_t> g_launches;
void set_error(const char *fmt, ...);

#define XRB_CUDA(expr)                                                                    \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ::xrb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                             __FILE__, __LINE__);                                         \
            return XRB_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define XRB_LAUNCHED() (::xrb::g_launches.fetch_add(1, std::memory_order_relaxed))

// Select `device` and verify it is a Blackwell sm_100 part; no CPU fallback exists.
int select_device(int device);

struct DevBuf {  // grow-only device buffer
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes);
    void release();
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

}  // namespace xrb
