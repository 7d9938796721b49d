// pnp_b200.h — drop-in for the pose-refinement block of xrsfm::RegisterImage (src/geometry/pnp.cc:38-71) on top of
// libxrsfm_b200.so.  The reference builds a ceres::Problem per frame (ReProjectionCost + HuberLoss(5.99) over the
// inlier correspondences, quaternion parameterisation, points and intrinsics constant) and runs ten iterations.
// Here the frames are queued and refined together: one kernel launch, one CTA per frame (xrb_pose_refine_batch).
// Header-only template code: it reads frame.points[p2d_id], frame.Tcw.{q,t}, camera.model_id_ / params_ and the
// vectors RegisterImage already holds (id_pair_vec, points3ds, inlier_mask), so it compiles against the reference's
// types unchanged (and against tests/mock/xrsfm_mock.h).
//
//   reference (one frame)                         replacement
//   ------------------------------------------    ---------------------------------------------------------------
//   SolvePnP_colmap(..., frame.Tcw, inlier_mask)   unchanged (CPU P3P LORANSAC)
//   ceres::Problem ... ceres::Solve (:38-71)       PoseRefiner r; r.Add(frame, camera, id_pair_vec, points3ds,
//                                                  inlier_mask); ... (more frames) ...; r.Run();
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <vector>

#include "xrsfm_b200.h"

namespace xrsfm_b200 {

template <class FrameT>
class PoseRefiner {
  public:
    explicit PoseRefiner(int device = 0) : device_(device) { xrb_pose_default_options(&options); }

    xrb_ba_options options;  // ceres::Solver::Options defaults, max_num_iterations = 10 (pnp.cc:56-57)
    std::vector<xrb_pose_summary> summaries;
    int last_status = XRB_OK;

    // Queue one frame: the arguments RegisterImage holds when it reaches pnp.cc:38.
    template <class CameraT, class Vec3T>
    void Add(FrameT &frame, const CameraT &camera, const std::vector<std::pair<int, int>> &id_pair_vec,
             const std::vector<Vec3T> &points3ds, const std::vector<char> &inlier_mask) {
        frames_.push_back(&frame);
        for (size_t id = 0; id < id_pair_vec.size(); ++id) {
            const auto &uv = frame.points[id_pair_vec[id].first];
            uv_.push_back(uv(0)), uv_.push_back(uv(1));
            const double *X = points3ds[id].data();
            xyz_.insert(xyz_.end(), X, X + 3);
            inlier_.push_back(id < inlier_mask.size() && inlier_mask[id] ? 1 : 0);
        }
        offsets_.push_back((int64_t)inlier_.size());
        for (int k = 0; k < 8; ++k) intr_.push_back(k < (int)camera.params_.size() ? camera.params_[k] : 0.0);
        model_.push_back((int32_t)camera.model_id_);
        const double *q = frame.Tcw.q.coeffs().data(), *t = frame.Tcw.t.data();
        q_.insert(q_.end(), q, q + 4), t_.insert(t_.end(), t, t + 3);
    }

    size_t size() const { return frames_.size(); }

    // Refine every queued frame and write the poses back; prints what pnp.cc:60-67 prints per frame.
    int Run(bool verbose = true) {
        const int n = (int)frames_.size();
        summaries.assign(n, xrb_pose_summary{});
        last_status = xrb_pose_refine_batch(device_, n, offsets_.data(), uv_.data(), xyz_.data(), inlier_.data(), intr_.data(),
                                            model_.data(), q_.data(), t_.data(), &options, summaries.data());
        if (last_status != XRB_OK) {
            std::fprintf(stderr, "PoseRefiner: %s\n", xrb_last_error());
            return last_status;
        }
        for (int i = 0; i < n; ++i) {
            double *q = frames_[i]->Tcw.q.coeffs().data(), *t = frames_[i]->Tcw.t.data();
            for (int k = 0; k < 4; ++k) q[k] = q_[4 * (size_t)i + k];
            for (int k = 0; k < 3; ++k) t[k] = t_[3 * (size_t)i + k];
            if (verbose && summaries[i].num_residuals > 0) {
                std::printf("Initial cost : %.6g [px]\n", std::sqrt(summaries[i].initial_cost / summaries[i].num_residuals));
                std::printf("Final cost : %.6g [px]\n", std::sqrt(summaries[i].final_cost / summaries[i].num_residuals));
            }
        }
        Clear();
        return XRB_OK;
    }

    void Clear() {
        frames_.clear(), uv_.clear(), xyz_.clear(), inlier_.clear(), intr_.clear(), model_.clear(), q_.clear(), t_.clear();
        offsets_.assign(1, 0);
    }

  private:
    int device_;
    std::vector<FrameT *> frames_;
    std::vector<int64_t> offsets_{0};
    std::vector<double> uv_, xyz_, intr_, q_, t_;
    std::vector<uint8_t> inlier_;
    std::vector<int32_t> model_;
};

}  // namespace xrsfm_b200
