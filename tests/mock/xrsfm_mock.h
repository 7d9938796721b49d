// Minimal stand-ins for the reference types the shims touch (src/base/map.h:12-195,
// src/base/types.h:32-61, src/base/camera.hpp:10-76) so the shim headers can be
// compile-checked in an image without Eigen.  Test infrastructure.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <map>
#include <utility>
#include <vector>

namespace mock {
struct Vec2 { double v[2]; double operator()(int i) const { return v[i]; } };
struct Vec3 { double v[3]; double *data() { return v; } const double *data() const { return v; } };
struct Coeffs { double v[4]; double *data() { return v; } const double *data() const { return v; } };
struct Quat { Coeffs c; Coeffs &coeffs() { return c; } const Coeffs &coeffs() const { return c; } };
struct Pose { Quat q; Vec3 t; };
struct CameraT { int model_id_ = 2; std::vector<double> params_; };
struct Track { Vec3 point3d_; bool outlier = false; double angle_ = 0; std::map<int, int> observations_; };
struct U8Descs {  // Eigen::Matrix<uint8_t, Dynamic, 128, RowMajor> as far as the matcher reads it
    std::vector<uint8_t> v;
    long rows() const { return (long)(v.size() / 128); }
    const uint8_t *data() const { return v.data(); }
};
struct Match {  // src/base/types.h:14-21
    Match(int _id1 = 0, int _id2 = 0, double _dist = 0) : id1(_id1), id2(_id2), distance(_dist) {}
    int id1, id2;
    double distance;
};
struct FramePair {  // src/base/map.h:81-99 (the members FeatureMatching touches)
    int id1 = 0, id2 = 0;
    std::vector<Match> matches;
    int inlier_num = 0;
    std::vector<char> inlier_mask;
};
struct Frame {
    int id = 0, camera_id = 0;
    bool registered = false, is_keyframe = false;
    Pose Tcw;
    std::vector<Vec2> points;
    std::vector<int> track_ids_;
    U8Descs uint_descs_;
};
struct Map {
    std::vector<Frame> frames_;
    std::vector<Track> tracks_;
    std::map<int, CameraT> cameras_;
    int init_id1 = 0, init_id2 = 1;
    CameraT &Camera(int id) { return cameras_[id]; }
};
// Stand-ins for the reference's map walks that stay on the host (found by ADL from the shim's BASolverT).
inline void KeyFrameSelection(Map &map, std::vector<int>, bool) {
    for (auto &f : map.frames_) f.is_keyframe = f.registered;
}
inline void UpdateByRefFrame(Map &) {}
inline std::vector<int> CovisibilityNeibors(int frame_id, Map &map, size_t num_images = 4) {
    std::map<int, int> cov;
    for (int tid : map.frames_[frame_id].track_ids_)
        if (tid != -1)
            for (auto &o : map.tracks_[tid].observations_) cov[o.first] += 1;
    std::vector<std::pair<int, int>> v(cov.begin(), cov.end());
    std::stable_sort(v.begin(), v.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });
    std::vector<int> ids;
    for (auto &e : v) {
        ids.push_back(e.first);
        if (ids.size() == num_images) break;
    }
    return ids;
}
inline std::vector<int> FindLocalBundle(int frame_id, Map &map, size_t num_images = 4) {
    std::vector<int> ids{frame_id};
    for (int id : CovisibilityNeibors(frame_id, map, num_images + 1))
        if (id != frame_id && ids.size() < num_images) ids.push_back(id);
    return ids;
}
}  // namespace mock
