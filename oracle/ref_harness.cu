/*
 * oracle/ref_harness.cu — host harness that drives the REFERENCE's own matcher kernels.
 *
 * TEST INFRASTRUCTURE ONLY (never linked into the product).  The kernels
 * MultiplyDescriptor_Kernel / RowMatch_Kernel / ColMatch_Kernel are compiled verbatim from
 * /root/reference/3rdparty/SiftGPU/ProgramCU.cu (see oracle/Makefile: target _ref); no
 * reference source is copied into this repository.  ProgramCU.cu expects its CuTexImage
 * device-buffer class (declared in the reference's CuTexImage.h, implemented in
 * CuTexImage.cpp, which needs OpenGL and cannot be built here) and the GlobalParam statics;
 * this file supplies just enough of both, then replays the call sequence of
 * SiftMatchCU::SetDescriptors / GetSiftMatch / GetBestMatch (SiftMatchCU.cpp:100-118,
 * 175-215) with the same blocking copies.
 */
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <cstring>
#include <vector>

#include "GL/glew.h"
#include "CuTexImage.h"
#include "GlobalUtil.h"
#include "ProgramCU.h"

// ---- GlobalParam statics referenced by ProgramCU.cu (values: GlobalUtil.cpp defaults
// are irrelevant to the matcher kernels; zero/neutral here) -------------------------------
int GlobalParam::_MemCapGPU = 0;
int GlobalParam::_texMaxDimGL = 16384;
int GlobalParam::_MaxOrientation = 2;
int GlobalParam::_NormalizedSIFT = 1;
int GlobalParam::_FixedOrientation = 0;
int GlobalParam::_KeepExtremumSign = 0;
float GlobalParam::_FilterWidthFactor = 4.0f;
int GlobalParam::_UseDynamicIndexing = 0;
int GlobalParam::_SubpixelLocalization = 0;
float GlobalParam::_DescriptorWindowFactor = 3.0f;
float GlobalParam::_OrientationWindowFactor = 2.0f;
float GlobalParam::_OrientationGaussianFactor = 1.5f;
int GlobalParam::_verbose = 0;

// ---- minimal CuTexImage: linear device buffers only -------------------------------------
CuTexImage::CuTexObj::~CuTexObj() { cudaDestroyTextureObject(handle); }

CuTexImage::CuTexImage()
    : _cuData(nullptr), _cuData2D(nullptr), _numChannel(0), _numBytes(0), _imgWidth(0),
      _imgHeight(0), _texWidth(0), _texHeight(0), _fromPBO(0) {}
CuTexImage::CuTexImage(int, int, int, GLuint) : CuTexImage() {}
CuTexImage::~CuTexImage() {
    if (_cuData) cudaFree(_cuData);
}
void CuTexImage::SetImageSize(int w, int h) { _imgWidth = w, _imgHeight = h; }

bool CuTexImage::InitTexture(int w, int h, int nchannel) {
    // same sizing rule as the reference (CuTexImage.cpp:150-183): channels clamped to 1..4,
    // 4-byte elements, grow-only.
    _imgWidth = w, _imgHeight = h;
    _numChannel = nchannel < 1 ? 1 : (nchannel > 4 ? 4 : nchannel);
    const size_t need = (size_t)w * h * _numChannel * sizeof(float);
    if (need >= (size_t)INT_MAX * sizeof(float)) return false;
    if (need <= _numBytes) return true;
    if (_cuData) cudaFree(_cuData);
    _cuData = nullptr;
    if (cudaMalloc(&_cuData, need) != cudaSuccess) {
        _numBytes = 0;
        return false;
    }
    _numBytes = need;
    return true;
}

CuTexImage::CuTexObj CuTexImage::BindTexture(const cudaTextureDesc &td,
                                             const cudaChannelFormatDesc &fmt) {
    CuTexObj obj;
    cudaResourceDesc rd;
    memset(&rd, 0, sizeof rd);
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = _cuData;
    rd.res.linear.desc = fmt;
    rd.res.linear.sizeInBytes = _numBytes;
    cudaCreateTextureObject(&obj.handle, &rd, &td, nullptr);
    return obj;
}
CuTexImage::CuTexObj CuTexImage::BindTexture2D(const cudaTextureDesc &td,
                                               const cudaChannelFormatDesc &fmt) {
    return BindTexture(td, fmt);  // never reached by the matcher
}
void CuTexImage::CopyFromHost(const void *buf) {
    if (_cuData)
        cudaMemcpy(_cuData, buf, (size_t)_imgWidth * _imgHeight * _numChannel * 4,
                   cudaMemcpyHostToDevice);
}
void CuTexImage::CopyToHost(void *buf) {
    if (_cuData)
        cudaMemcpy(buf, _cuData, (size_t)_imgWidth * _imgHeight * _numChannel * 4,
                   cudaMemcpyDeviceToHost);
}
void CuTexImage::CopyToHost(void *buf, int) { CopyToHost(buf); }
int CuTexImage::CopyToPBO(GLuint) { return 0; }
void CuTexImage::CopyFromPBO(int, int, GLuint) {}

// ---- the SiftMatchCU call sequence -------------------------------------------------------
namespace {
struct RefMatcher {
    CuTexImage des[2], dot, crt, match[2];
    int num[2] = {0, 0};
    std::vector<int> buf;
};
RefMatcher *g_ref = nullptr;
}  // namespace

extern "C" {

/* SiftMatchCU::SetDescriptors(u8) + GetSiftMatch + GetBestMatch for one pair.
 * Returns #matches, or -1 on CUDA error (SiftMatchCU.cpp:209-212). */
int xrref_match_pair(int n1, const uint8_t *d1, int n2, const uint8_t *d2, float distmax,
                     float ratiomax, int mbm, int max_match, uint32_t (*out)[2],
                     int32_t *m12_out, int32_t *m21_out) {
    if (!g_ref) g_ref = new RefMatcher();
    RefMatcher &r = *g_ref;
    if (n1 <= 0 || n2 <= 0) return 0;
    const uint8_t *d[2] = {d1, d2};
    const int n[2] = {n1, n2};
    for (int k = 0; k < 2; ++k) {  // SiftMatchCU.cpp:115-117
        r.num[k] = n[k];
        if (!r.des[k].InitTexture(8 * n[k], 1, 4)) return -1;
        r.des[k].CopyFromHost(d[k]);
    }
    ProgramCU::MultiplyDescriptor(r.des, r.des + 1, &r.dot, mbm ? &r.crt : nullptr);
    r.buf.resize((size_t)n1 + n2);
    int *b1 = r.buf.data(), *b2 = r.buf.data() + n1;
    r.match[0].InitTexture(n1, 1);
    ProgramCU::GetRowMatch(&r.dot, r.match, distmax, ratiomax);
    r.match[0].CopyToHost(b1);
    if (mbm) {
        r.match[1].InitTexture(n2, 1);
        ProgramCU::GetColMatch(&r.crt, r.match + 1, distmax, ratiomax);
        r.match[1].CopyToHost(b2);
    }
    int nmatch = 0;
    for (int i = 0; i < n1 && nmatch < max_match; ++i) {
        int j = b1[i];
        if (j >= 0 && (!mbm || b2[j] == i)) {
            out[nmatch][0] = (uint32_t)i;
            out[nmatch][1] = (uint32_t)j;
            ++nmatch;
        }
    }
    if (m12_out) memcpy(m12_out, b1, (size_t)n1 * 4);
    if (m21_out && mbm) memcpy(m21_out, b2, (size_t)n2 * 4);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return nmatch;
}

void xrref_release(void) {
    delete g_ref;
    g_ref = nullptr;
}
}
