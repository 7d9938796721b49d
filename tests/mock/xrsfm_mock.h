// Minimal stand-ins for the reference types the shims touch (src/base/map.h:12-195,
// src/base/types.h:32-61, src/base/camera.hpp:10-76) so the shim headers can be
// compile-checked in an image without Eigen.  Test infrastructure.
#pragma once
#include <array>
#include <map>
#include <vector>

namespace mock {
struct Vec2 { double v[2]; double operator()(int i) const { return v[i]; } };
struct Vec3 { double v[3]; double *data() { return v; } const double *data() const { return v; } };
struct Coeffs { double v[4]; double *data() { return v; } const double *data() const { return v; } };
struct Quat { Coeffs c; Coeffs &coeffs() { return c; } const Coeffs &coeffs() const { return c; } };
struct Pose { Quat q; Vec3 t; };
struct CameraT { int model_id_ = 2; std::vector<double> params_; };
struct Track { Vec3 point3d_; bool outlier = false; double angle_ = 0; std::map<int, int> observations_; };
struct Frame {
    int id = 0, camera_id = 0;
    bool registered = false, is_keyframe = false;
    Pose Tcw;
    std::vector<Vec2> points;
    std::vector<int> track_ids_;
};
struct Map {
    std::vector<Frame> frames_;
    std::vector<Track> tracks_;
    std::map<int, CameraT> cameras_;
    int init_id1 = 0, init_id2 = 1;
    CameraT &Camera(int id) { return cameras_[id]; }
};
}  // namespace mock
