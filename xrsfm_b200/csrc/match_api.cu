// match_api.cu — C ABI of path M (see include/xrsfm_b200.h for the reference lines each
// entry point replaces).  Host logic only: buffer ownership, batching, copies.
#include <algorithm>
#include <cstring>
#include <vector>

#include "io_formats.cuh"
#include "match_kernels.cuh"

using namespace xrb;

struct xrb_matcher {
    int device = 0;
    int max_features = 4096;
    int variant = 1;  // generation in force
    cudaStream_t stream = nullptr;

    // resident image set
    DevBuf images;                  // owned descriptor block (when uploaded from host)
    const uint8_t *block = nullptr; // device pointer actually used (owned or attached)
    std::vector<int64_t> offsets;   // n_images + 1 row offsets (host copy)
    DevBuf offsets_dev;
    int n_images = 0;

    // per-pair compatibility slots (SiftMatchCU::_texDes[2], SiftMatchCU.cpp:100-118)
    DevBuf slot[2];
    int slot_n[2] = {0, 0};
    int slot_id[2] = {0, 0};  // SiftMatchCU.cpp:50 initialises _id_sift to 0

    // batch scratch
    int chunk_pairs = 0, state_stride = 0;
    DevBuf rows_best, rows_second, cols_best, cols_second;
    DevBuf pairdesc, counts, strided, packed, pack_offsets, vlow, pair_idx;
    DevBuf dist_tab;  // float(acos(double(min(v*2^-18,1)))) for v = 0..2^18 (generation 3)
    int strided_stride = 0;
    // host-output pipeline of xrb_match_pairs: chunk c's match lists travel to the caller's
    // buffer on copy_stream while chunk c+1 is being scored
    DevBuf packed2, pack_offsets2;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_off[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    int64_t *h_off[2] = {nullptr, nullptr};
    size_t h_off_cap = 0;
};

namespace {

// A device pointer handed in by the caller must live on the matcher's GPU: the kernels would fault
// on a pointer of another device (no peer mapping is set up), long after the call returned.
int check_on_device(const xrb_matcher *m, const void *p, const char *what) {
    if (!p) return XRB_OK;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: not a CUDA pointer", what);
        return XRB_ERR_INVALID;
    }
    if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) {
        set_error("%s: expected device memory", what);
        return XRB_ERR_INVALID;
    }
    if (a.type == cudaMemoryTypeDevice && a.device != m->device) {
        set_error("%s lives on GPU %d but the matcher was created on GPU %d", what, a.device, m->device);
        return XRB_ERR_INVALID;
    }
    return XRB_OK;
}

int ensure_scratch(xrb_matcher *m, int n_pairs_hint) {
    // Size the batch so the top-2 state stays below ~512 MiB.
    const int stride = m->max_features;
    const size_t per_pair = (size_t)stride * (8 + 4) * 2;
    int chunk = (int)std::max<size_t>(1, std::min<size_t>(2048, (512ull << 20) / per_pair));
    chunk = std::min(chunk, std::max(1, n_pairs_hint));
    if (chunk <= m->chunk_pairs && stride == m->state_stride) return XRB_OK;
    chunk = std::max(chunk, m->chunk_pairs);
    const size_t n = (size_t)chunk * stride;
    int rc;
    if ((rc = m->rows_best.reserve(n * 8))) return rc;
    if ((rc = m->cols_best.reserve(n * 8))) return rc;
    if ((rc = m->rows_second.reserve(n * 4))) return rc;
    if ((rc = m->cols_second.reserve(n * 4))) return rc;
    XRB_CUDA(cudaMemsetAsync(m->rows_best.p, 0, n * 8, m->stream));
    XRB_CUDA(cudaMemsetAsync(m->cols_best.p, 0, n * 8, m->stream));
    XRB_CUDA(cudaMemsetAsync(m->rows_second.p, 0, n * 4, m->stream));
    XRB_CUDA(cudaMemsetAsync(m->cols_second.p, 0, n * 4, m->stream));
    if ((rc = m->pairdesc.reserve((size_t)chunk * sizeof(PairDesc)))) return rc;
    if ((rc = m->counts.reserve((size_t)chunk * 4))) return rc;
    if ((rc = m->pack_offsets.reserve((size_t)(chunk + 1) * 8))) return rc;
    if ((rc = m->vlow.reserve(16))) return rc;
    m->chunk_pairs = chunk;
    m->state_stride = stride;
    return XRB_OK;
}

int score(xrb_matcher *m, const PairDesc *pd_dev, int n, int max_n1, int max_n2,
          cudaStream_t st, bool slots = false) {
    Top2State rows{m->rows_best.as<unsigned long long>(), m->rows_second.as<unsigned int>()};
    Top2State cols{m->cols_best.as<unsigned long long>(), m->cols_second.as<unsigned int>()};
    if (m->variant >= 2) {
        if (slots)
            return launch_score_tc(pd_dev, n, m->slot[0].as<uint8_t>(), (uint64_t)m->max_features,
                                   m->slot[1].as<uint8_t>(), (uint64_t)m->max_features,
                                   m->state_stride, rows, cols, m->vlow.as<int>(), st);
        const uint64_t total = (uint64_t)m->offsets[m->n_images];
        return launch_score_tc(pd_dev, n, m->block, total, m->block, total, m->state_stride, rows,
                               cols, m->vlow.as<int>(), st);
    }
    return launch_score_dp4a(pd_dev, n, max_n1, max_n2, m->state_stride, rows, cols,
                             m->vlow.as<int>(), st);
}

// generation 3 applies when every image of the call fits the fused kernel's shared-memory state
bool use_fused(const xrb_matcher *m, int max_feat) {
    return m->variant == 3 && max_feat <= fused_max_features() && m->dist_tab.p != nullptr;
}

int match_fused(xrb_matcher *m, const PairDesc *pd_dev, int n, bool slots, float distmax, float ratiomax, int mbm,
                int max_match, int32_t *counts_dev, uint32_t (*out_dev)[2], int out_stride, cudaStream_t st) {
    const uint8_t *bA, *bB;
    uint64_t rA, rB;
    if (slots) {
        bA = m->slot[0].as<uint8_t>(), bB = m->slot[1].as<uint8_t>(), rA = rB = (uint64_t)m->max_features;
    } else {
        bA = bB = m->block, rA = rB = (uint64_t)m->offsets[m->n_images];
    }
    return launch_match_fused(pd_dev, n, bA, rA, bB, rB, m->vlow.as<int>(), m->dist_tab.as<float>(), distmax,
                              ratiomax, mbm, max_match, counts_dev, out_dev, out_stride, st);
}

int finalize(xrb_matcher *m, const PairDesc *pd_dev, int n, float distmax, float ratiomax,
             int mbm, int max_match, int32_t *counts_dev, uint32_t (*out_dev)[2],
             int out_stride, cudaStream_t st) {
    Top2State rows{m->rows_best.as<unsigned long long>(), m->rows_second.as<unsigned int>()};
    Top2State cols{m->cols_best.as<unsigned long long>(), m->cols_second.as<unsigned int>()};
    return launch_finalize(pd_dev, n, m->state_stride, rows, cols, distmax, ratiomax, mbm,
                           max_match, counts_dev, out_dev, out_stride, st);
}

}  // namespace

extern "C" {

xrb_matcher *xrb_match_create(int max_features, int device) {
    if (select_device(device) != XRB_OK) return nullptr;
    xrb_matcher *m = new xrb_matcher();
    m->device = device;
    // SiftMatchCU ctor / SetMaxSift round up to a multiple of 32 (SiftMatchCU.cpp:47-53,84-87)
    m->max_features = max_features <= 0 ? 4096 : ((max_features + 31) / 32) * 32;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        delete m;
        return nullptr;
    }
    m->variant = score_tc_available() ? 3 : 1;
    if (m->dist_tab.reserve(262145 * sizeof(float)) != XRB_OK ||
        launch_dist_table(m->dist_tab.as<float>(), 262145, m->stream) != XRB_OK ||
        cudaStreamSynchronize(m->stream) != cudaSuccess) {
        set_error("matcher: building the distance table failed");
        xrb_match_destroy(m);
        return nullptr;
    }
    return m;
}

void xrb_match_destroy(xrb_matcher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    DevBuf *bufs[] = {&m->images, &m->offsets_dev, &m->slot[0], &m->slot[1], &m->rows_best,
                      &m->rows_second, &m->cols_best, &m->cols_second, &m->pairdesc,
                      &m->counts, &m->strided, &m->packed, &m->pack_offsets, &m->vlow,
                      &m->pair_idx, &m->dist_tab};
    for (DevBuf *b : bufs) b->release();
    m->packed2.release(), m->pack_offsets2.release();
    for (int b = 0; b < 2; ++b) {
        if (m->ev_off[b]) cudaEventDestroy(m->ev_off[b]);
        if (m->ev_copied[b]) cudaEventDestroy(m->ev_copied[b]);
        if (m->h_off[b]) cudaFreeHost(m->h_off[b]);
    }
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    cudaStreamDestroy(m->stream);
    delete m;
}

int xrb_match_max_features(const xrb_matcher *m) { return m ? m->max_features : 0; }

int xrb_match_set_variant(xrb_matcher *m, int variant) {
    if (!m) return XRB_ERR_INVALID;
    if (variant == 0) variant = score_tc_available() ? 3 : 1;
    if (variant >= 2 && !score_tc_available()) variant = 1;
    if (variant < 1 || variant > 3) {
        set_error("unknown matcher variant %d", variant);
        return XRB_ERR_INVALID;
    }
    m->variant = variant;
    return variant;
}

int xrb_match_set_descriptors(xrb_matcher *m, int index, int num, const uint8_t *desc,
                              int id) {
    if (!m) return XRB_ERR_INVALID;
    XRB_CUDA(cudaSetDevice(m->device));
    index = index > 1 ? 1 : (index < 0 ? 0 : index);  // SiftMatchCU.cpp:104-107
    if (id != -1 && id == m->slot_id[index]) return XRB_OK;  // :110-111
    m->slot_id[index] = id;
    if (num > m->max_features) num = m->max_features;  // :113-114
    if (num < 0) num = 0;
    m->slot_n[index] = num;
    if (num == 0) return XRB_OK;
    if (!desc) {
        set_error("set_descriptors: null descriptor pointer");
        return XRB_ERR_INVALID;
    }
    int rc = m->slot[index].reserve((size_t)m->max_features * kDim);
    if (rc) return rc;
    XRB_CUDA(cudaMemcpyAsync(m->slot[index].p, desc, (size_t)num * kDim,
                             cudaMemcpyHostToDevice, m->stream));
    return XRB_OK;
}

int xrb_match_get(xrb_matcher *m, int max_match, uint32_t (*match_buffer)[2], float distmax,
                  float ratiomax, int mutual_best_match) {
    if (!m) return 0;
    if (m->slot_n[0] <= 0 || m->slot_n[1] <= 0) return 0;  // SiftMatchCU.cpp:179-180
    if (max_match <= 0) return 0;
    if (cudaSetDevice(m->device) != cudaSuccess) return -1;
    if (ensure_scratch(m, 1) != XRB_OK) return -1;
    const int n1 = m->slot_n[0], n2 = m->slot_n[1];
    const int stride = std::min(max_match, n1);
    if (m->strided.reserve((size_t)stride * 8) != XRB_OK) return -1;
    PairDesc pd{m->slot[0].as<uint8_t>(), m->slot[1].as<uint8_t>(), n1, n2};
    cudaStream_t st = m->stream;
    if (cudaMemcpyAsync(m->pairdesc.p, &pd, sizeof pd, cudaMemcpyHostToDevice, st) !=
        cudaSuccess)
        return -1;
    if (launch_vlow(distmax, ratiomax, m->vlow.as<int>(), st)) return -1;
    {
        // ONE pair: the persistent tcgen05 kernels give a pair to one CTA (1/148 of the device, ~1.1 ms for
        // 4096 x 4096); the tiled dp4a kernel spreads the same pair over (n1/128) x (n2/128) CTAs.  Same
        // integer dot products, same top-2 state, same finalize: bit-identical lists.
        Top2State rows{m->rows_best.as<unsigned long long>(), m->rows_second.as<unsigned int>()};
        Top2State cols{m->cols_best.as<unsigned long long>(), m->cols_second.as<unsigned int>()};
        if (launch_score_dp4a(m->pairdesc.as<PairDesc>(), 1, n1, n2, m->state_stride, rows, cols, m->vlow.as<int>(), st))
            return -1;
        if (finalize(m, m->pairdesc.as<PairDesc>(), 1, distmax, ratiomax, mutual_best_match,
                     max_match, m->counts.as<int32_t>(), m->strided.as<uint32_t[2]>(), stride, st))
            return -1;
    }
    int32_t n = 0;
    if (cudaMemcpyAsync(&n, m->counts.p, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        return -1;
    if (cudaStreamSynchronize(st) != cudaSuccess) {  // SiftMatchCU.cpp:209-212
        set_error("matcher: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    if (n > 0 &&
        cudaMemcpy(match_buffer, m->strided.p, (size_t)n * 8, cudaMemcpyDeviceToHost) !=
            cudaSuccess)
        return -1;
    return n;
}

static int set_offsets(xrb_matcher *m, int n_images, const int64_t *row_offsets) {
    m->offsets.assign(row_offsets, row_offsets + n_images + 1);
    m->n_images = n_images;
    int rc = m->offsets_dev.reserve((size_t)(n_images + 1) * 8);
    if (rc) return rc;
    XRB_CUDA(cudaMemcpyAsync(m->offsets_dev.p, m->offsets.data(), (size_t)(n_images + 1) * 8,
                             cudaMemcpyHostToDevice, m->stream));
    XRB_CUDA(cudaStreamSynchronize(m->stream));
    return XRB_OK;
}

int xrb_match_upload_images(xrb_matcher *m, int n_images, const int32_t *counts,
                            const uint8_t *const *descs) {
    if (!m || n_images < 0 || (n_images && (!counts || !descs))) {
        set_error("upload_images: bad arguments");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(m->device));
    std::vector<int64_t> off(n_images + 1, 0);
    for (int i = 0; i < n_images; ++i) {
        if (counts[i] < 0) {
            set_error("upload_images: negative count for image %d", i);
            return XRB_ERR_INVALID;
        }
        off[i + 1] = off[i] + counts[i];
    }
    int rc = m->images.reserve(std::max<size_t>(16, (size_t)off[n_images] * kDim));
    if (rc) return rc;
    for (int i = 0; i < n_images; ++i)
        if (counts[i])
            XRB_CUDA(cudaMemcpyAsync(m->images.as<uint8_t>() + off[i] * kDim, descs[i],
                                     (size_t)counts[i] * kDim, cudaMemcpyHostToDevice,
                                     m->stream));
    m->block = m->images.as<uint8_t>();
    return set_offsets(m, n_images, off.data());
}

int xrb_match_upload_packed(xrb_matcher *m, int n_images, const int64_t *row_offsets,
                            const uint8_t *desc_block) {
    if (!m || n_images < 0 || !row_offsets || (row_offsets[n_images] && !desc_block)) {
        set_error("upload_packed: bad arguments");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(m->device));
    const size_t bytes = (size_t)row_offsets[n_images] * kDim;
    int rc = m->images.reserve(std::max<size_t>(16, bytes));
    if (rc) return rc;
    if (bytes)
        XRB_CUDA(cudaMemcpyAsync(m->images.p, desc_block, bytes, cudaMemcpyHostToDevice,
                                 m->stream));
    m->block = m->images.as<uint8_t>();
    return set_offsets(m, n_images, row_offsets);
}

int xrb_match_upload_ftr(xrb_matcher *m, const char *path) {
    if (!m || !path) {
        set_error("upload_ftr: bad arguments");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(m->device));
    // pass 1: frame sizes (the file is the reference's ftr.bin, io_feature.hpp:76-100)
    std::vector<int64_t> off(1, 0);
    int64_t max_points = 0;
    int rc;
    {
        FtrReader r;
        if ((rc = r.open(path))) return rc;
        std::string name;
        for (int i = 0; i < r.n_frames(); ++i) {
            int32_t np = 0;
            if ((rc = r.header(&name, &np)) || (rc = r.keypoints(nullptr)) || (rc = r.descriptors(nullptr))) return rc;
            off.push_back(off.back() + np);
            max_points = std::max<int64_t>(max_points, np);
        }
    }
    const int n_images = (int)off.size() - 1;
    if ((rc = m->images.reserve(std::max<size_t>(16, (size_t)off[n_images] * kDim)))) return rc;
    // pass 2: descriptors file -> pinned ring -> HBM; the copy of frame i overlaps the read of i + 1
    uint8_t *stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    const size_t stage_bytes = std::max<size_t>(16, (size_t)max_points * kDim);
    auto cleanup = [&]() {
        for (int b = 0; b < 2; ++b) {
            if (done[b]) cudaEventDestroy(done[b]);
            if (stage[b]) cudaFreeHost(stage[b]);
        }
    };
    for (int b = 0; b < 2; ++b) {
        if (cudaMallocHost(&stage[b], stage_bytes) != cudaSuccess ||
            cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming) != cudaSuccess) {
            cleanup();
            set_error("upload_ftr: pinned staging allocation failed");
            return XRB_ERR_CUDA;
        }
    }
    FtrReader r;
    if ((rc = r.open(path))) {
        cleanup();
        return rc;
    }
    std::string name;
    for (int i = 0; i < n_images && rc == XRB_OK; ++i) {
        const int b = i & 1;
        int32_t np = 0;
        if (i >= 2 && cudaEventSynchronize(done[b]) != cudaSuccess) rc = XRB_ERR_CUDA;
        if (rc == XRB_OK) rc = r.header(&name, &np);
        if (rc == XRB_OK && np != off[i + 1] - off[i]) {
            set_error("%s changed while it was being read", path);
            rc = XRB_ERR_INVALID;
        }
        if (rc == XRB_OK) rc = r.keypoints(nullptr);
        if (rc == XRB_OK) rc = r.descriptors(np ? stage[b] : nullptr);
        if (rc == XRB_OK && np &&
            (cudaMemcpyAsync(m->images.as<uint8_t>() + off[i] * kDim, stage[b], (size_t)np * kDim, cudaMemcpyHostToDevice,
                             m->stream) != cudaSuccess ||
             cudaEventRecord(done[b], m->stream) != cudaSuccess)) {
            set_error("upload_ftr: host-to-device copy failed");
            rc = XRB_ERR_CUDA;
        }
    }
    cudaStreamSynchronize(m->stream);
    cleanup();
    if (rc) return rc;
    m->block = m->images.as<uint8_t>();
    return set_offsets(m, n_images, off.data());
}

int xrb_match_attach_device(xrb_matcher *m, int n_images, const int64_t *row_offsets_host,
                            const uint8_t *desc_block_device) {
    if (!m || n_images < 0 || !row_offsets_host || !desc_block_device) {
        set_error("attach_device: bad arguments");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(m->device));
    if (int rc = check_on_device(m, desc_block_device, "attach_device: descriptor block")) return rc;
    m->block = desc_block_device;
    return set_offsets(m, n_images, row_offsets_host);
}

int xrb_match_pairs_device(xrb_matcher *m, int n_pairs, const int32_t (*pairs_dev)[2],
                           float distmax, float ratiomax, int mutual_best_match,
                           int max_match, int32_t *counts_dev, uint32_t (*out_dev)[2],
                           int out_stride, void *stream) {
    if (!m || n_pairs < 0 || !m->block) {
        set_error("pairs_device: no images resident / bad arguments");
        return XRB_ERR_INVALID;
    }
    if (n_pairs == 0) return XRB_OK;
    XRB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if ((rc = check_on_device(m, pairs_dev, "pairs_device: pair list"))) return rc;
    if ((rc = check_on_device(m, counts_dev, "pairs_device: counts"))) return rc;
    if ((rc = check_on_device(m, out_dev, "pairs_device: output"))) return rc;
    if ((rc = ensure_scratch(m, n_pairs))) return rc;
    if (m->stream != st) XRB_CUDA(cudaStreamSynchronize(m->stream));  // scratch memsets
    if (out_stride < std::min(max_match, m->max_features)) {
        // every pair may legitimately produce min(max_match, n1) matches
        int64_t worst = 0;
        for (int i = 0; i < m->n_images; ++i)
            worst = std::max<int64_t>(worst, m->offsets[i + 1] - m->offsets[i]);
        worst = std::min<int64_t>(std::min<int64_t>(worst, m->max_features), max_match);
        if (out_stride < worst) {
            set_error("pairs_device: out_stride %d < worst-case matches %lld", out_stride,
                      (long long)worst);
            return XRB_ERR_CAPACITY;
        }
    }
    if (launch_vlow(distmax, ratiomax, m->vlow.as<int>(), st)) return XRB_ERR_CUDA;
    int64_t max_n = 0;
    for (int i = 0; i < m->n_images; ++i)
        max_n = std::max<int64_t>(max_n, m->offsets[i + 1] - m->offsets[i]);
    const int max_feat = (int)std::min<int64_t>(max_n, m->max_features);
    for (int p0 = 0; p0 < n_pairs; p0 += m->chunk_pairs) {
        const int n = std::min(m->chunk_pairs, n_pairs - p0);
        PairDesc *pd = m->pairdesc.as<PairDesc>();
        if ((rc = launch_build_pairs(pairs_dev + p0, n, m->n_images, m->offsets_dev.as<int64_t>(), m->block,
                                     m->max_features, pd, st)))
            return rc;
        if (use_fused(m, max_feat)) {
            if ((rc = match_fused(m, pd, n, false, distmax, ratiomax, mutual_best_match, max_match, counts_dev + p0,
                                  out_dev + (size_t)p0 * out_stride, out_stride, st)))
                return rc;
            continue;
        }
        if ((rc = score(m, pd, n, max_feat, max_feat, st))) return rc;
        if ((rc = finalize(m, pd, n, distmax, ratiomax, mutual_best_match, max_match,
                           counts_dev + p0, out_dev + (size_t)p0 * out_stride, out_stride, st)))
            return rc;
    }
    return XRB_OK;
}

int xrb_match_pairs(xrb_matcher *m, int n_pairs, const int32_t (*pairs)[2], float distmax,
                    float ratiomax, int mutual_best_match, int max_match,
                    int64_t *out_offsets, uint32_t (*out)[2], int64_t out_capacity) {
    if (!m || n_pairs < 0 || !out_offsets || (n_pairs && !pairs) || !m->block) {
        set_error("match_pairs: no images resident / bad arguments");
        return XRB_ERR_INVALID;
    }
    out_offsets[0] = 0;
    if (n_pairs == 0) return XRB_OK;
    for (int p = 0; p < n_pairs; ++p)
        if (pairs[p][0] < 0 || pairs[p][0] >= m->n_images || pairs[p][1] < 0 ||
            pairs[p][1] >= m->n_images) {
            set_error("match_pairs: pair %d references image outside [0,%d)", p, m->n_images);
            return XRB_ERR_INVALID;
        }
    XRB_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = m->stream;
    int rc = ensure_scratch(m, n_pairs);
    if (rc) return rc;
    int64_t max_n = 0;
    for (int i = 0; i < m->n_images; ++i)
        max_n = std::max<int64_t>(max_n, m->offsets[i + 1] - m->offsets[i]);
    const int max_feat = (int)std::min<int64_t>(max_n, m->max_features);
    const int stride = std::max(1, std::min(max_match, max_feat));
    // Pipeline in chunks of <= 512 pairs: while the host drains chunk c (offsets, then the packed
    // match list into the caller's buffer) the device is already scoring chunk c + 1.
    const int chunk = std::min(m->chunk_pairs, 512);
    if ((rc = m->strided.reserve((size_t)chunk * stride * 8))) return rc;
    if ((rc = m->packed.reserve((size_t)chunk * stride * 8))) return rc;
    if ((rc = m->packed2.reserve((size_t)chunk * stride * 8))) return rc;
    if ((rc = m->pack_offsets2.reserve((size_t)(m->chunk_pairs + 1) * 8))) return rc;
    if ((rc = m->pair_idx.reserve((size_t)chunk * 8))) return rc;
    if (!m->copy_stream) {
        XRB_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        for (int b = 0; b < 2; ++b) {
            XRB_CUDA(cudaEventCreateWithFlags(&m->ev_off[b], cudaEventDisableTiming));
            XRB_CUDA(cudaEventCreateWithFlags(&m->ev_copied[b], cudaEventDisableTiming));
        }
    }
    if (m->h_off_cap < (size_t)chunk + 1) {
        for (int b = 0; b < 2; ++b) {
            if (m->h_off[b]) cudaFreeHost(m->h_off[b]);
            m->h_off[b] = nullptr;
            XRB_CUDA(cudaMallocHost(&m->h_off[b], ((size_t)chunk + 1) * 8));
        }
        m->h_off_cap = (size_t)chunk + 1;
    }
    if (launch_vlow(distmax, ratiomax, m->vlow.as<int>(), st)) return XRB_ERR_CUDA;

    int64_t total = 0;
    bool overflow = false;
    DevBuf *packed[2] = {&m->packed, &m->packed2}, *poff[2] = {&m->pack_offsets, &m->pack_offsets2};
    auto drain = [&](int c) -> int {  // host side of chunk c: offsets, then the D2H of its matches
        const int b = c & 1, p0 = c * chunk, n = std::min(chunk, n_pairs - p0);
        XRB_CUDA(cudaEventSynchronize(m->ev_off[b]));
        const int64_t got = m->h_off[b][n];
        for (int k = 0; k < n; ++k) out_offsets[p0 + k + 1] = total + m->h_off[b][k + 1];
        if (total + got <= out_capacity && out) {
            if (got)
                XRB_CUDA(cudaMemcpyAsync(out + total, packed[b]->p, (size_t)got * 8, cudaMemcpyDeviceToHost,
                                         m->copy_stream));
        } else {
            overflow = true;
        }
        XRB_CUDA(cudaEventRecord(m->ev_copied[b], m->copy_stream));
        total += got;
        return XRB_OK;
    };
    const int n_chunks = (n_pairs + chunk - 1) / chunk;
    for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1, p0 = c * chunk, n = std::min(chunk, n_pairs - p0);
        if (c >= 2) XRB_CUDA(cudaStreamWaitEvent(st, m->ev_copied[b], 0));  // packed[b] is free again
        XRB_CUDA(cudaMemcpyAsync(m->pair_idx.p, pairs + p0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
        PairDesc *pd = m->pairdesc.as<PairDesc>();
        if ((rc = launch_build_pairs(m->pair_idx.as<int32_t[2]>(), n, m->n_images, m->offsets_dev.as<int64_t>(), m->block,
                                     m->max_features, pd, st)))
            return rc;
        if (use_fused(m, max_feat)) {
            if ((rc = match_fused(m, pd, n, false, distmax, ratiomax, mutual_best_match, max_match,
                                  m->counts.as<int32_t>(), m->strided.as<uint32_t[2]>(), stride, st)))
                return rc;
        } else {
            if ((rc = score(m, pd, n, max_feat, max_feat, st))) return rc;
            if ((rc = finalize(m, pd, n, distmax, ratiomax, mutual_best_match, max_match,
                               m->counts.as<int32_t>(), m->strided.as<uint32_t[2]>(), stride, st)))
                return rc;
        }
        if ((rc = launch_pack(m->counts.as<int32_t>(), n, m->strided.as<uint32_t[2]>(), stride,
                              poff[b]->as<int64_t>(), packed[b]->as<uint32_t[2]>(), 0, (int64_t)chunk * stride, st)))
            return rc;
        XRB_CUDA(cudaMemcpyAsync(m->h_off[b], poff[b]->p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, st));
        XRB_CUDA(cudaEventRecord(m->ev_off[b], st));
        if (c >= 1 && (rc = drain(c - 1))) return rc;
    }
    if ((rc = drain(n_chunks - 1))) return rc;
    XRB_CUDA(cudaStreamSynchronize(m->copy_stream));
    XRB_CUDA(cudaStreamSynchronize(st));
    if (overflow) {
        set_error("match_pairs: out_capacity %lld < %lld matches", (long long)out_capacity,
                  (long long)total);
        return XRB_ERR_CAPACITY;
    }
    return XRB_OK;
}

/* test hook (not part of the reference surface): CUDA's float(acos(double)) table */
int xrb_match_debug_dist_table(float *out_host, int n) {
    float *d = nullptr;
    XRB_CUDA(cudaMalloc(&d, (size_t)n * 4));
    int rc = launch_dist_table(d, n, nullptr);
    if (rc == XRB_OK) {
        cudaError_t e = cudaMemcpy(out_host, d, (size_t)n * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = XRB_ERR_CUDA;
    }
    cudaFree(d);
    return rc;
}

}  // extern "C"
