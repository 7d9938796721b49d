// ba_solver_b200.h — drop-in bodies for xrsfm::BASolver::{GBA,KGBA} (src/optimization/
// ba_solver.cc:594-678) on top of libxrsfm_b200.so.  Header-only template code: it touches
// Map/Frame/Track/Camera only through the members BASolver::SetUp reads
// (ba_solver.cc:330-356), so it compiles against the reference's src/base/map.h unchanged
// (and against tests/mock/xrsfm_mock.h, which is how this repository compile-checks it
// without Eigen).
//
// Flattening rules:
//   cameras   = registered frames (GBA :598-601) / registered key-frames (KGBA :647-654)
//   obs       = every i with frame.track_ids_[i] != -1 -> (frame, track, frame.points[i])
//   points    = tracks referenced above
//   constants = intrinsics always (:602-606); translations of map.init_id1/2 (:611-614);
//               every pose when fix_all_frames (:616-621)
//   q order   = Eigen coeffs x,y,z,w as stored, not re-normalised
#pragma once
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <map>
#include <unordered_map>
#include <set>
#include <vector>

#include "xrsfm_b200.h"

namespace xrsfm_b200 {

struct FlatBA {
    std::vector<double> cam_q, cam_t, pts, intr, obs_uv;
    std::vector<int32_t> intr_model, cam_intr, obs_cam, obs_pt;
    std::vector<uint8_t> cam_q_fixed, cam_t_fixed, pt_fixed;
    std::vector<int> frame_of_cam, track_of_pt;
    xrb_ba_problem problem() {
        xrb_ba_problem p;
        p.n_cams = (int32_t)frame_of_cam.size(), p.n_pts = (int32_t)track_of_pt.size();
        p.n_obs = (int32_t)obs_cam.size(), p.n_intr = (int32_t)intr_model.size();
        p.cam_q = cam_q.data(), p.cam_t = cam_t.data(), p.pts = pts.data();
        p.intr = intr.data(), p.intr_model = intr_model.data(), p.cam_intr = cam_intr.data();
        p.obs_cam = obs_cam.data(), p.obs_pt = obs_pt.data(), p.obs_uv = obs_uv.data();
        p.cam_q_fixed = cam_q_fixed.data(), p.cam_t_fixed = cam_t_fixed.data(), p.pt_fixed = pt_fixed.data();
        return p;
    }
};

template <class MapT>
FlatBA Flatten(MapT &map, bool keyframes_only, bool fix_all_frames) {
    FlatBA f;
    std::unordered_map<int, int> intr_of_camera, pt_of_track;
    for (auto &frame : map.frames_) {
        if (!frame.registered) continue;
        if (keyframes_only && !frame.is_keyframe) continue;
        int n_mea = 0;
        for (size_t i = 0; i < frame.track_ids_.size(); ++i) n_mea += frame.track_ids_[i] != -1;
        if (n_mea == 0) {
            std::fprintf(stderr, "BA: NO Measurement In Frame %d\n", frame.id);  // ba_solver.cc:350-351
            continue;
        }
        const int cam = (int)f.frame_of_cam.size();
        f.frame_of_cam.push_back(frame.id);
        const double *q = frame.Tcw.q.coeffs().data(), *t = frame.Tcw.t.data();
        f.cam_q.insert(f.cam_q.end(), q, q + 4);
        f.cam_t.insert(f.cam_t.end(), t, t + 3);
        auto it = intr_of_camera.find(frame.camera_id);
        if (it == intr_of_camera.end()) {
            auto &camera = map.Camera(frame.camera_id);
            it = intr_of_camera.emplace(frame.camera_id, (int)f.intr_model.size()).first;
            f.intr_model.push_back(camera.model_id_);
            for (int k = 0; k < 8; ++k) f.intr.push_back(k < (int)camera.params_.size() ? camera.params_[k] : 0.0);
        }
        f.cam_intr.push_back(it->second);
        const bool gauge = frame.id == map.init_id1 || frame.id == map.init_id2;
        f.cam_q_fixed.push_back(fix_all_frames ? 1 : 0);
        f.cam_t_fixed.push_back((fix_all_frames || gauge) ? 1 : 0);
        for (size_t i = 0; i < frame.track_ids_.size(); ++i) {
            const int tid = frame.track_ids_[i];
            if (tid == -1) continue;
            auto pit = pt_of_track.find(tid);
            if (pit == pt_of_track.end()) {
                pit = pt_of_track.emplace(tid, (int)f.track_of_pt.size()).first;
                f.track_of_pt.push_back(tid);
                const double *X = map.tracks_[tid].point3d_.data();
                f.pts.insert(f.pts.end(), X, X + 3);
                f.pt_fixed.push_back(0);
            }
            f.obs_cam.push_back(cam), f.obs_pt.push_back(pit->second);
            f.obs_uv.push_back(frame.points[i](0)), f.obs_uv.push_back(frame.points[i](1));
        }
    }
    return f;
}

template <class MapT>
void Scatter(const FlatBA &f, MapT &map) {
    for (size_t c = 0; c < f.frame_of_cam.size(); ++c) {
        auto &frame = map.frames_[f.frame_of_cam[c]];
        double *q = frame.Tcw.q.coeffs().data(), *t = frame.Tcw.t.data();
        for (int k = 0; k < 4; ++k) q[k] = f.cam_q[4 * c + k];
        for (int k = 0; k < 3; ++k) t[k] = f.cam_t[3 * c + k];
    }
    for (size_t p = 0; p < f.track_of_pt.size(); ++p) {
        double *X = map.tracks_[f.track_of_pt[p]].point3d_.data();
        for (int k = 0; k < 3; ++k) X[k] = f.pts[3 * p + k];
    }
}

// PrintSolverSummary (ba_solver.cc:14-68): same lines, same fields.
inline void PrintSolverSummary(const xrb_ba_summary &s) {
    const double n = s.num_residuals_reduced > 0 ? s.num_residuals_reduced : 1;
    static const char *term[] = {"Convergence", "No convergence", "Failure"};
    std::printf("%16s%d\n%16s%d\n%16s%d\n%16s%g [s]\n%16s%.6g [px]\n%16s%.6g [px]\n%16s%s\n\n", "Residuals : ",
                s.num_residuals_reduced, "Parameters : ", s.num_effective_parameters_reduced, "Iterations : ",
                s.num_successful_steps + s.num_unsuccessful_steps, "Time : ", s.total_time_in_seconds,
                "Initial cost : ", std::sqrt(s.initial_cost / n), "Final cost : ", std::sqrt(s.final_cost / n),
                "Termination : ", s.termination_type >= 0 && s.termination_type <= 2 ? term[s.termination_type] : "Unknown");
}

// One engine per device, created on first use (a handle owns its streams and buffers).
inline xrb_ba_solver *Engine(int device = 0) {
    static std::map<int, xrb_ba_solver *> engines;
    auto it = engines.find(device);
    if (it == engines.end()) it = engines.emplace(device, xrb_ba_create(device)).first;
    return it->second;
}

// Body of BASolver::GBA (ba_solver.cc:594-638).
template <class MapT>
int GBA(MapT &map, bool accurate = true, bool fix_all_frames = false, int device = 0) {
    FlatBA f = Flatten(map, /*keyframes_only=*/false, fix_all_frames);
    xrb_ba_options o;
    xrb_ba_default_options(&o);
    o.verbose = 1;  // minimizer_progress_to_stdout (:625)
    o.max_iterations = accurate ? 50 : 20;
    o.function_tolerance = accurate ? 1e-5 : 1e-4;
    o.parameter_tolerance = accurate ? 1e-6 : 1e-5;
    xrb_ba_problem p = f.problem();
    xrb_ba_summary s;
    const int rc = xrb_ba_solve(Engine(device), &p, &o, &s);
    if (rc != XRB_OK) {
        std::fprintf(stderr, "xrb_ba_solve failed (%d): %s\n", rc, xrb_last_error());
        return rc;
    }
    Scatter(f, map);
    PrintSolverSummary(s);
    return rc;
}

// Solver part of BASolver::KGBA (ba_solver.cc:645-675); the caller keeps
// KeyFrameSelection(map, ...) before and UpdateByRefFrame(map) after (:641,:677).
template <class MapT>
int KGBA_Solve(MapT &map, int device = 0) {
    FlatBA f = Flatten(map, /*keyframes_only=*/true, false);
    xrb_ba_options o;
    xrb_ba_default_options(&o);
    o.verbose = 1;
    o.initial_radius = 1e6, o.max_iterations = 20, o.function_tolerance = 1e-4, o.parameter_tolerance = 1e-5;
    xrb_ba_problem p = f.problem();
    xrb_ba_summary s;
    const int rc = xrb_ba_solve(Engine(device), &p, &o, &s);
    if (rc != XRB_OK) {
        std::fprintf(stderr, "xrb_ba_solve failed (%d): %s\n", rc, xrb_last_error());
        return rc;
    }
    Scatter(f, map);
    PrintSolverSummary(s);
    return rc;
}

// Flat problem of BASolver::LBA (ba_solver.cc:523-584 with SetUpLBA :358-391) for the window the
// caller selected: local_frame_ids1 = CovisibilityNeibors(frame_id, map), local_frame_ids2 =
// FindLocalBundle(frame_id, map) (both stay on the reference side).  Frames are visited in the
// order of the std::set the reference builds; a point is constant when track.angle_ > 5 or the
// new frame does not observe it (:380-382); translations: the gauge frames if they are in the
// window, else the last two of ids2, else the last two of ids1, else the new frame (:552-584).
template <class MapT>
FlatBA FlattenLBA(MapT &map, int frame_id, const std::vector<int> &local_frame_ids1,
                  const std::vector<int> &local_frame_ids2) {
    std::set<int> local(local_frame_ids1.begin(), local_frame_ids1.end());
    local.insert(local_frame_ids2.begin(), local_frame_ids2.end());
    std::set<int> fixed_t;
    if (local.count(map.init_id1)) fixed_t.insert(map.init_id1);
    if (local.count(map.init_id2)) fixed_t.insert(map.init_id2);
    if (fixed_t.empty()) {
        const std::vector<int> *src = local_frame_ids2.size() >= 2 ? &local_frame_ids2
                                      : local_frame_ids1.size() >= 2 ? &local_frame_ids1 : nullptr;
        if (src) {
            fixed_t.insert((*src)[src->size() - 1]), fixed_t.insert((*src)[src->size() - 2]);
        } else {
            std::printf("!!!LBA only one frame\n");
            fixed_t.insert(frame_id);
        }
    }
    FlatBA f;
    std::unordered_map<int, int> intr_of_camera, pt_of_track;
    for (const int id : local) {
        auto &frame = map.frames_.at(id);
        int n_mea = 0;
        for (size_t i = 0; i < frame.track_ids_.size(); ++i) n_mea += frame.track_ids_[i] != -1;
        if (n_mea == 0) {
            std::fprintf(stderr, "LBA: NO Measurement In Frame %d\n", frame.id);  // ba_solver.cc:384-385
            continue;
        }
        const int cam = (int)f.frame_of_cam.size();
        f.frame_of_cam.push_back(frame.id);
        const double *q = frame.Tcw.q.coeffs().data(), *t = frame.Tcw.t.data();
        f.cam_q.insert(f.cam_q.end(), q, q + 4);
        f.cam_t.insert(f.cam_t.end(), t, t + 3);
        auto it = intr_of_camera.find(frame.camera_id);
        if (it == intr_of_camera.end()) {
            auto &camera = map.Camera(frame.camera_id);
            it = intr_of_camera.emplace(frame.camera_id, (int)f.intr_model.size()).first;
            f.intr_model.push_back(camera.model_id_);
            for (int k = 0; k < 8; ++k) f.intr.push_back(k < (int)camera.params_.size() ? camera.params_[k] : 0.0);
        }
        f.cam_intr.push_back(it->second);
        f.cam_q_fixed.push_back(0);
        f.cam_t_fixed.push_back(fixed_t.count(frame.id) ? 1 : 0);
        for (size_t i = 0; i < frame.track_ids_.size(); ++i) {
            const int tid = frame.track_ids_[i];
            if (tid == -1) continue;
            auto pit = pt_of_track.find(tid);
            if (pit == pt_of_track.end()) {
                pit = pt_of_track.emplace(tid, (int)f.track_of_pt.size()).first;
                f.track_of_pt.push_back(tid);
                auto &track = map.tracks_[tid];
                const double *X = track.point3d_.data();
                f.pts.insert(f.pts.end(), X, X + 3);
                f.pt_fixed.push_back((track.angle_ > 5 || track.observations_.count(frame_id) == 0) ? 1 : 0);
            }
            f.obs_cam.push_back(cam), f.obs_pt.push_back(pit->second);
            f.obs_uv.push_back(frame.points[i](0)), f.obs_uv.push_back(frame.points[i](1));
        }
    }
    return f;
}

// Solver part of BASolver::LBA (ba_solver.cc:586-591): 5 iterations, 1e-4 / 1e-5, no summary print.
template <class MapT>
int LBA_Solve(MapT &map, int frame_id, const std::vector<int> &local_frame_ids1,
              const std::vector<int> &local_frame_ids2, int device = 0) {
    FlatBA f = FlattenLBA(map, frame_id, local_frame_ids1, local_frame_ids2);
    xrb_ba_options o;
    xrb_ba_default_options(&o);
    o.max_iterations = 5, o.function_tolerance = 1e-4, o.parameter_tolerance = 1e-5;
    xrb_ba_problem p = f.problem();
    xrb_ba_summary s;
    const int rc = xrb_ba_solve(Engine(device), &p, &o, &s);
    if (rc != XRB_OK) {
        std::fprintf(stderr, "xrb_ba_solve failed (%d): %s\n", rc, xrb_last_error());
        return rc;
    }
    Scatter(f, map);
    return rc;
}

// class BASolver of src/optimization/ba_solver.h:14-30 with the same four public methods.  The map
// walks that are not bundle adjustment stay what they are in the reference and are found by
// argument-dependent lookup in MapT's namespace (namespace xrsfm in the reference build):
//   KeyFrameSelection(map, ids, is_sequential)  UpdateByRefFrame(map)      src/base/map.h:213-216
//   CovisibilityNeibors(frame_id, map)  FindLocalBundle(frame_id, map)     ba_solver.cc:393-521
//   ScalePoseGraphUnorder                                                  ba_solver.cc:79-328 (pose graph,
//       Ceres on the host: SURVEY.md row B14 keeps it on the reference; `pose_graph` forwards to it)
template <class MapT, class LoopInfoT = int>
class BASolverT {
  public:
    explicit BASolverT(int device = 0) : device_(device) {}

    void (*pose_graph)(const LoopInfoT &, MapT &, bool) = nullptr;
    void ScalePoseGraphUnorder(const LoopInfoT &loop_info, MapT &map, bool use_key = false) {
        if (pose_graph)
            pose_graph(loop_info, map, use_key);
        else
            std::fprintf(stderr, "ScalePoseGraphUnorder: not part of the B200 path, set BASolverT::pose_graph\n");
    }
    void KGBA(MapT &map, const std::vector<int> fix_key_frame_ids, const bool is_sequential_data) {
        KeyFrameSelection(map, fix_key_frame_ids, is_sequential_data);  // ba_solver.cc:641
        int num_rf = 0, num_kf = 0;
        for (auto &frame : map.frames_) {
            if (!frame.registered) continue;
            num_rf++;
            num_kf += frame.is_keyframe ? 1 : 0;
        }
        last_status = KGBA_Solve(map, device_);
        std::printf("kf: %d/%d\n", num_kf, num_rf);  // :676
        UpdateByRefFrame(map);                       // :677
    }
    void GBA(MapT &map, bool accurate = true, bool fix_all_frames = false) {
        last_status = xrsfm_b200::GBA(map, accurate, fix_all_frames, device_);
    }
    void LBA(int frame_id, MapT &map) {
        const std::vector<int> ids1 = CovisibilityNeibors(frame_id, map);  // ba_solver.cc:525
        const std::vector<int> ids2 = FindLocalBundle(frame_id, map);      // :526
        std::set<int> local(ids1.begin(), ids1.end());
        local.insert(ids2.begin(), ids2.end());
        std::printf("LBA: ");
        for (const int id : local) std::printf(" %d", id);
        std::printf("\n");
        last_status = LBA_Solve(map, frame_id, ids1, ids2, device_);
    }
    int last_status = XRB_OK;

  private:
    int device_;
};

}  // namespace xrsfm_b200
