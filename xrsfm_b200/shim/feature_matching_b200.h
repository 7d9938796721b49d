// feature_matching_b200.h — compiled replacement of xrsfm::FeatureMatching
// (src/feature/feature_processing.cc:222-308) on the batched C ABI of libxrsfm_b200.so.
//
// Same contract: every candidate pair is matched (descriptor ratio test + mutual check, the
// thresholds of SiftMatch :118-154: distmax 0.7, ratio 0.8, at most `max_match` matches), pairs with
// at least 15 matches are verified geometrically, pairs that keep max(15, 25 %) inliers survive with
// their matches filtered to the inliers.  What changes is the data movement: all descriptors go to HBM
// once (xrb_match_upload_images) and all pairs are matched by one call (xrb_match_pairs) instead of
// two H2D copies + one blocking read-back per pair.
//
// The geometric verification is a callable `verify(points1, points2, frame_pair)` with the contract of
// SolveFundamnetalCOLMAP (feature_processing.cc:256-296 / epipolar_geometry.hpp:10-27): it fills
// frame_pair.inlier_num and frame_pair.inlier_mask.  Template code: compiles against the reference's
// Frame / FramePair / Match (src/base/map.h, types.h) and against tests/mock/xrsfm_mock.h.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include "xrsfm_b200.h"

namespace xrsfm_b200 {

template <class FrameT, class FramePairT, class VerifyFn>
int FeatureMatching(const std::vector<FrameT> &frames, const std::vector<std::pair<int, int>> &candidate_pairs,
                    std::vector<FramePairT> &frame_pairs, bool b_use_fundamental, VerifyFn verify, int device = 0,
                    int max_match = 16384) {
    constexpr int min_num_matches = 15;
    constexpr int min_num_inlier = 15;
    constexpr double min_ratio_inlier = 0.25;
    using MatchT = typename std::decay<decltype(frame_pairs[0].matches[0])>::type;
    using PointT = typename std::decay<decltype(frames[0].points[0])>::type;

    static xrb_matcher *matcher = nullptr;  // CreateSiftGPUMatcher (:53-88): one matcher per process
    if (!matcher) matcher = xrb_match_create(max_match, device);
    if (!matcher) {
        std::fprintf(stderr, "ERROR: SiftMatchGPU not fully supported: %s\n", xrb_last_error());
        return XRB_ERR_NO_DEVICE;
    }
    // descriptors of every frame -> HBM, once
    std::vector<int32_t> counts(frames.size());
    std::vector<const uint8_t *> descs(frames.size());
    for (size_t i = 0; i < frames.size(); ++i) {
        counts[i] = (int32_t)frames[i].uint_descs_.rows();
        descs[i] = frames[i].uint_descs_.data();
    }
    int rc = xrb_match_upload_images(matcher, (int)frames.size(), counts.data(), descs.data());
    if (rc != XRB_OK) {
        std::fprintf(stderr, "xrb_match_upload_images failed (%d): %s\n", rc, xrb_last_error());
        return rc;
    }
    // all pairs in one call
    const int n_pairs = (int)candidate_pairs.size();
    std::vector<int32_t> pairs(2 * (size_t)n_pairs);
    int64_t cap = 0;
    for (int p = 0; p < n_pairs; ++p) {
        pairs[2 * p] = candidate_pairs[p].first, pairs[2 * p + 1] = candidate_pairs[p].second;
        cap += std::min(max_match, std::min(counts[candidate_pairs[p].first], counts[candidate_pairs[p].second]));
    }
    std::vector<int64_t> offsets((size_t)n_pairs + 1, 0);
    std::vector<uint32_t> out(2 * (size_t)std::max<int64_t>(cap, 1));
    rc = xrb_match_pairs(matcher, n_pairs, reinterpret_cast<const int32_t(*)[2]>(pairs.data()), 0.7f, 0.8f, 1, max_match,
                         offsets.data(), reinterpret_cast<uint32_t(*)[2]>(out.data()), cap);
    if (rc != XRB_OK) {
        std::fprintf(stderr, "xrb_match_pairs failed (%d): %s\n", rc, xrb_last_error());
        return rc;
    }
    for (int p = 0; p < n_pairs; ++p) {
        FramePairT frame_pair;
        frame_pair.id1 = candidate_pairs[p].first, frame_pair.id2 = candidate_pairs[p].second;
        frame_pair.inlier_num = 0;
        for (int64_t k = offsets[p]; k < offsets[p + 1]; ++k)
            frame_pair.matches.emplace_back(MatchT((int)out[2 * k], (int)out[2 * k + 1]));
        frame_pairs.emplace_back(frame_pair);
    }
    // geometric verification, thresholds and filtering as in :259-297
    int count_inlier_pairs = 0;
#pragma omp parallel for schedule(static, 8)
    for (int i = 0; i < (int)frame_pairs.size(); ++i) {
        auto &frame_pair = frame_pairs[i];
        if ((int)frame_pair.matches.size() < min_num_matches) continue;
        const auto &frame1 = frames[frame_pair.id1];
        const auto &frame2 = frames[frame_pair.id2];
        if (!b_use_fundamental) continue;  // the reference CHECK(false)s here (:273)
        std::vector<PointT> points1, points2;
        for (const auto &match : frame_pair.matches) {
            points1.push_back(frame1.points[match.id1]);
            points2.push_back(frame2.points[match.id2]);
        }
        verify(points1, points2, frame_pair);
        const int inlier_threshold = std::max(min_num_inlier, (int)(min_ratio_inlier * frame_pair.matches.size()));
        if (frame_pair.inlier_num < inlier_threshold) {
            frame_pair.inlier_num = 0;
            continue;
        }
        std::vector<MatchT> inlier_matches;  // ExtractInlierMatches (:310-321)
        for (size_t k = 0; k < frame_pair.matches.size(); ++k)
            if (frame_pair.inlier_mask[k]) inlier_matches.push_back(frame_pair.matches[k]);
        frame_pair.inlier_mask.assign(inlier_matches.size(), true);
        frame_pair.matches.swap(inlier_matches);
    }
    std::vector<FramePairT> kept;
    for (auto &frame_pair : frame_pairs) {
        if (frame_pair.inlier_num == 0) continue;
        kept.emplace_back(frame_pair);
        count_inlier_pairs++;
    }
    frame_pairs = std::move(kept);
    std::printf("matched image pairs: %d/%zu\n", count_inlier_pairs, candidate_pairs.size());
    return XRB_OK;
}

// SolveFundamnetalCOLMAP for EVERY pair of the batch in one launch (xrb_fm_loransac_batch, one CTA per pair): the
// verification callable of the overload above is replaced by a gather of the matched points, one call, and a
// scatter of inlier_num / inlier_mask / F (row-major into frame_pair.F(r, c) when the pair type has it).
namespace detail {
template <class FP>
auto set_F(FP &fp, const double *F, int) -> decltype(fp.F(0, 0), void()) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) fp.F(r, c) = F[r * 3 + c];
}
template <class FP>
void set_F(FP &, const double *, long) {}
}  // namespace detail

template <class FrameT, class FramePairT>
int FeatureMatching(const std::vector<FrameT> &frames, const std::vector<std::pair<int, int>> &candidate_pairs,
                    std::vector<FramePairT> &frame_pairs, bool b_use_fundamental, int device = 0, int max_match = 16384) {
    // pass 1: matching only (the stand-in verification marks nothing, so every pair is returned untouched)
    std::vector<FramePairT> all;
    {
        auto keep_all = [](const auto &p1, const auto &, FramePairT &fp) {
            fp.inlier_num = (int)p1.size();
            fp.inlier_mask.assign(p1.size(), 1);
        };
        // thresholds would drop pairs below 15 matches: those are dropped by the reference as well (:260-262)
        const int rc = FeatureMatching(frames, candidate_pairs, all, true, keep_all, device, max_match);
        if (rc != XRB_OK) return rc;
    }
    if (!b_use_fundamental) return XRB_ERR_INVALID;  // the reference CHECK(false)s on this branch (:273)
    std::vector<int64_t> offsets(all.size() + 1, 0);
    for (size_t i = 0; i < all.size(); ++i) offsets[i + 1] = offsets[i] + (int64_t)all[i].matches.size();
    std::vector<double> pts1(2 * (size_t)offsets.back()), pts2(2 * (size_t)offsets.back());
    for (size_t i = 0; i < all.size(); ++i) {
        const auto &f1 = frames[all[i].id1];
        const auto &f2 = frames[all[i].id2];
        int64_t w = offsets[i];
        for (const auto &m : all[i].matches) {
            pts1[2 * w] = f1.points[m.id1](0), pts1[2 * w + 1] = f1.points[m.id1](1);
            pts2[2 * w] = f2.points[m.id2](0), pts2[2 * w + 1] = f2.points[m.id2](1);
            ++w;
        }
    }
    xrb_fm_options opt;
    xrb_fm_default_options(&opt);  // epipolar_geometry.hpp:13-18
    std::vector<xrb_fm_report> reports(all.size());
    std::vector<char> mask((size_t)std::max<int64_t>(1, offsets.back()));
    const int rc = xrb_fm_loransac_batch(device, (int)all.size(), offsets.data(), pts1.data(), pts2.data(), &opt, reports.data(),
                                         mask.data());
    if (rc != XRB_OK) {
        std::fprintf(stderr, "xrb_fm_loransac_batch failed (%d): %s\n", rc, xrb_last_error());
        return rc;
    }
    int count_inlier_pairs = 0;
    for (size_t i = 0; i < all.size(); ++i) {
        auto &fp = all[i];
        const int n = (int)fp.matches.size();
        fp.inlier_num = (int)reports[i].num_inliers;
        detail::set_F(fp, reports[i].F, 0);
        const int inlier_threshold = std::max(15, (int)(0.25 * n));  // :283-285
        if (!reports[i].success || fp.inlier_num < inlier_threshold) continue;
        using MatchT = typename std::decay<decltype(fp.matches[0])>::type;
        std::vector<MatchT> inliers;
        for (int k = 0; k < n; ++k)
            if (mask[offsets[i] + k]) inliers.push_back(fp.matches[k]);
        fp.inlier_mask.assign(inliers.size(), true);
        fp.matches.swap(inliers);
        frame_pairs.emplace_back(fp);
        count_inlier_pairs++;
    }
    std::printf("matched image pairs: %d/%zu\n", count_inlier_pairs, candidate_pairs.size());
    return XRB_OK;
}

}  // namespace xrsfm_b200
