// SiftMatchGPU_b200.h — header-compatible stand-in for the SiftMatchGPU façade of
// 3rdparty/SiftGPU/SiftGPU.h:277-372, backed by libxrsfm_b200.so.
//
// A maintainer drops this header in place of the SiftGPU one for the MATCHER only (SIFT
// extraction keeps using SiftGPU): src/feature/feature_processing.cc compiles unchanged —
// CreateSiftGPUMatcher (:53-88) and SiftMatch (:118-154) use exactly the members below.
// Semantics per method are those of SiftMatchCU (3rdparty/SiftGPU/SiftMatchCU.cpp).
#pragma once
#include <cstdint>

#include "xrsfm_b200.h"

class SiftMatchGPU {
  public:
    enum SIFTMATCH_LANGUAGE {
        SIFTMATCH_SAME_AS_SIFTGPU = 0,
        SIFTMATCH_GLSL = 2,
        SIFTMATCH_CUDA = 3,
        SIFTMATCH_CUDA_DEVICE0 = 3
    };

    int gpu_index = 0;

    explicit SiftMatchGPU(int max_sift = 4096) : max_sift_(max_sift) {}
    // the reference copy-assigns a temporary (feature_processing.cc:65); handles are not shared
    SiftMatchGPU(const SiftMatchGPU &o) : gpu_index(o.gpu_index), max_sift_(o.max_sift_), device_(o.device_) {}
    SiftMatchGPU &operator=(const SiftMatchGPU &o) {
        if (this != &o) {
            Release();
            gpu_index = o.gpu_index, max_sift_ = o.max_sift_, device_ = o.device_;
        }
        return *this;
    }
    virtual ~SiftMatchGPU() { Release(); }

    virtual void SetLanguage(int gpu_language) {
        if (gpu_language >= SIFTMATCH_CUDA_DEVICE0) device_ = gpu_language - SIFTMATCH_CUDA_DEVICE0;
    }
    virtual void SetDeviceParam(int, char **) {}
    int CreateContextGL() { return VerifyContextGL(); }
    int VerifyContextGL() { return Ensure() ? 1 : 0; }  // SiftMatch.cpp:598-642 picks CUDA

    // SiftMatchCU::Allocate (SiftMatchCU.cpp:55-82): worst-case buffers for max_sift features
    virtual bool Allocate(int max_sift, int /*mbm*/) {
        max_sift_ = max_sift;
        Release();
        return Ensure();
    }
    virtual void SetMaxSift(int max_sift) { Allocate(max_sift, 1); }
    virtual int GetMaxSift() const { return h_ ? xrb_match_max_features(h_) : max_sift_; }

    virtual void SetDescriptors(int index, int num, const unsigned char *descriptors, int id = -1) {
        if (Ensure()) xrb_match_set_descriptors(h_, index, num, descriptors, id);
    }
    // float descriptors (SiftGPU.h:328-330) are not on XRSfM's path (feature_processing.cc:90-116
    // is only reached with FeatureDescriptors); quantise like sift_extractor.h:22-34 upstream.

    virtual int GetSiftMatch(int max_match, uint32_t match_buffer[][2], float distmax = 0.7f,
                             float ratiomax = 0.8f, int mutual_best_match = 1) {
        if (!Ensure()) return 0;  // SiftMatchCU.cpp:177-178: not initialised -> 0
        return xrb_match_get(h_, max_match, match_buffer, distmax, ratiomax, mutual_best_match);
    }

    xrb_matcher *handle() { return Ensure() ? h_ : nullptr; }  // batched entry points

  private:
    bool Ensure() {
        if (!h_) h_ = xrb_match_create(max_sift_, device_);
        return h_ != nullptr;
    }
    void Release() {
        if (h_) xrb_match_destroy(h_);
        h_ = nullptr;
    }
    xrb_matcher *h_ = nullptr;
    int max_sift_ = 4096;
    int device_ = 0;
};
