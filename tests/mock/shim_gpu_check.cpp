// Runs the C++ drop-in boundary ON THE GPU: class BASolver (GBA / KGBA / LBA) over a mock Map built
// from a scene file, and the matcher façade + the compiled FeatureMatching over a descriptor file.
// Results go to a file the Python test compares with the ctypes path and the oracle.
//   shim_gpu_check ba <scene.bin> <out.bin> gba|kgba|lba     shim_gpu_check match <desc.bin> <out.bin>
//   shim_gpu_check pose <batch.bin> <out.bin>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "xrsfm_mock.h"
#include "../../xrsfm_b200/shim/SiftMatchGPU_b200.h"
#include "../../xrsfm_b200/shim/ba_solver_b200.h"
#include "../../xrsfm_b200/shim/feature_matching_b200.h"
#include "../../xrsfm_b200/shim/pnp_b200.h"

template <class T>
static bool rd(FILE *f, T *p, size_t n = 1) { return fread(p, sizeof(T), n, f) == n; }
template <class T>
static void wr(FILE *f, const T *p, size_t n = 1) { fwrite(p, sizeof(T), n, f); }

static int run_ba(const char *scene, const char *out, const std::string &mode) {
    FILE *f = fopen(scene, "rb");
    if (!f) return 2;
    int32_t nf = 0, ntr = 0, gauge[2];
    double intr[4];
    rd(f, &nf), rd(f, &ntr), rd(f, gauge, 2), rd(f, intr, 4);
    mock::Map map;
    map.cameras_[0].model_id_ = 2;
    map.cameras_[0].params_ = {intr[0], intr[1], intr[2], intr[3]};
    map.init_id1 = gauge[0], map.init_id2 = gauge[1];
    map.frames_.resize(nf), map.tracks_.resize(ntr);
    for (int i = 0; i < nf; ++i) {
        auto &fr = map.frames_[i];
        fr.id = i, fr.registered = true;
        int32_t np = 0;
        rd(f, fr.Tcw.q.c.v, 4), rd(f, fr.Tcw.t.v, 3), rd(f, &np);
        fr.points.resize(np), fr.track_ids_.resize(np);
        for (int k = 0; k < np; ++k) {
            int32_t tid;
            rd(f, fr.points[k].v, 2), rd(f, &tid);
            fr.track_ids_[k] = tid;
            if (tid >= 0) map.tracks_[tid].observations_[i] = k;
        }
    }
    for (int t = 0; t < ntr; ++t) rd(f, map.tracks_[t].point3d_.v, 3);
    fclose(f);
    xrsfm_b200::BASolverT<mock::Map> solver(0);
    std::vector<int32_t> window;
    if (mode == "gba") {
        solver.GBA(map, true, false);
    } else if (mode == "kgba") {
        solver.KGBA(map, std::vector<int>{}, true);
    } else {
        const int frame_id = nf - 1;
        std::set<int> local;
        for (int id : CovisibilityNeibors(frame_id, map)) local.insert(id);
        for (int id : FindLocalBundle(frame_id, map)) local.insert(id);
        window.assign(local.begin(), local.end());
        solver.LBA(frame_id, map);
    }
    if (solver.last_status != XRB_OK) return 3;
    FILE *o = fopen(out, "wb");
    if (!o) return 2;
    for (auto &fr : map.frames_) wr(o, fr.Tcw.q.c.v, 4), wr(o, fr.Tcw.t.v, 3);
    for (auto &tr : map.tracks_) wr(o, tr.point3d_.v, 3);
    const int32_t nw = (int32_t)window.size();
    wr(o, &nw), wr(o, window.data(), window.size());
    fclose(o);
    return 0;
}

static int run_match(const char *desc, const char *out) {
    FILE *f = fopen(desc, "rb");
    if (!f) return 2;
    int32_t n_img = 0, n_pairs = 0;
    rd(f, &n_img), rd(f, &n_pairs);
    std::vector<mock::Frame> frames(n_img);
    for (int i = 0; i < n_img; ++i) {
        int32_t n = 0;
        rd(f, &n);
        frames[i].id = i;
        frames[i].uint_descs_.v.resize((size_t)n * 128);
        rd(f, frames[i].uint_descs_.v.data(), (size_t)n * 128);
        frames[i].points.resize(n);
        for (int k = 0; k < n; ++k) frames[i].points[k] = {{(double)k, (double)i}};
    }
    std::vector<std::pair<int, int>> pairs(n_pairs);
    for (auto &p : pairs) {
        int32_t ab[2];
        rd(f, ab, 2);
        p = {ab[0], ab[1]};
    }
    fclose(f);
    FILE *o = fopen(out, "wb");
    if (!o) return 2;
    // (a) the façade exactly as CreateSiftGPUMatcher + SiftMatch drive it (feature_processing.cc:53-154)
    SiftMatchGPU m;
    m = SiftMatchGPU(16384);
    m.SetLanguage(SiftMatchGPU::SIFTMATCH_CUDA_DEVICE0 + 0);
    if (!m.VerifyContextGL()) return 3;
    m.Allocate(16384, true);
    std::vector<uint32_t> buf(2 * 16384);
    for (auto &p : pairs) {
        m.SetDescriptors(0, (int)frames[p.first].uint_descs_.rows(), frames[p.first].uint_descs_.data());
        m.SetDescriptors(1, (int)frames[p.second].uint_descs_.rows(), frames[p.second].uint_descs_.data());
        const int32_t n = m.GetSiftMatch(16384, reinterpret_cast<uint32_t(*)[2]>(buf.data()), 0.7f, 0.8f, true);
        wr(o, &n), wr(o, buf.data(), 2 * (size_t)(n > 0 ? n : 0));
    }
    // (b) the compiled FeatureMatching; the verification stand-in keeps the even matches
    std::vector<mock::FramePair> frame_pairs;
    auto verify = [](const std::vector<mock::Vec2> &p1, const std::vector<mock::Vec2> &, mock::FramePair &fp) {
        fp.inlier_mask.assign(p1.size(), 0);
        fp.inlier_num = 0;
        for (size_t k = 0; k < p1.size(); k += 2) fp.inlier_mask[k] = 1, fp.inlier_num++;
    };
    const int rc = xrsfm_b200::FeatureMatching(frames, pairs, frame_pairs, true, verify);
    if (rc != XRB_OK) return 4;
    const int32_t nk = (int32_t)frame_pairs.size();
    wr(o, &nk);
    for (auto &fp : frame_pairs) {
        const int32_t hdr[4] = {fp.id1, fp.id2, (int32_t)fp.matches.size(), fp.inlier_num};
        wr(o, hdr, 4);
        for (auto &mt : fp.matches) {
            const int32_t ij[2] = {mt.id1, mt.id2};
            wr(o, ij, 2);
        }
    }
    // (c) the fully GPU variant: matching + batched LO-RANSAC verification (the mock's points carry no real two-view
    // geometry; the parity of the verification itself is tests/test_fm_gpu.py)
    std::vector<mock::FramePair> verified;
    const int rc2 = xrsfm_b200::FeatureMatching(frames, pairs, verified, true);
    const int32_t tail[2] = {rc2, (int32_t)verified.size()};
    wr(o, tail, 2);
    fclose(o);
    return 0;
}

// Pose refinement from C++: frames with their correspondences as RegisterImage holds them (pnp.cc:24-37).
static int run_pose(const char *in, const char *out) {
    FILE *f = fopen(in, "rb");
    if (!f) return 2;
    int32_t n = 0;
    rd(f, &n);
    std::vector<mock::Frame> frames(n);
    std::vector<mock::CameraT> cams(n);
    std::vector<std::vector<std::pair<int, int>>> id_pairs(n);
    std::vector<std::vector<mock::Vec3>> p3d(n);
    std::vector<std::vector<char>> masks(n);
    for (int i = 0; i < n; ++i) {
        int32_t m = 0, model = 0;
        double intr[8];
        rd(f, &m), rd(f, &model), rd(f, intr, 8);
        cams[i].model_id_ = model, cams[i].params_.assign(intr, intr + 8);
        rd(f, frames[i].Tcw.q.c.v, 4), rd(f, frames[i].Tcw.t.v, 3);
        frames[i].points.resize(m), p3d[i].resize(m), masks[i].resize(m);
        for (int k = 0; k < m; ++k) {
            uint8_t inl = 0;
            rd(f, frames[i].points[k].v, 2), rd(f, p3d[i][k].v, 3), rd(f, &inl);
            masks[i][k] = (char)inl;
            id_pairs[i].push_back({k, k});
        }
    }
    fclose(f);
    xrsfm_b200::PoseRefiner<mock::Frame> refiner(0);
    for (int i = 0; i < n; ++i) refiner.Add(frames[i], cams[i], id_pairs[i], p3d[i], masks[i]);
    if ((int)refiner.size() != n) return 4;
    if (refiner.Run(false) != XRB_OK) return 3;
    FILE *o = fopen(out, "wb");
    if (!o) return 2;
    for (int i = 0; i < n; ++i) {
        wr(o, frames[i].Tcw.q.c.v, 4), wr(o, frames[i].Tcw.t.v, 3);
        const double c[2] = {refiner.summaries[i].initial_cost, refiner.summaries[i].final_cost};
        wr(o, c, 2);
    }
    fclose(o);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 5 && !strcmp(argv[1], "ba")) return run_ba(argv[2], argv[3], argv[4]);
    if (argc >= 4 && !strcmp(argv[1], "match")) return run_match(argv[2], argv[3]);
    if (argc >= 4 && !strcmp(argv[1], "pose")) return run_pose(argv[2], argv[3]);
    std::fprintf(stderr, "usage: shim_gpu_check ba <scene> <out> gba|kgba|lba | match <desc> <out>\n");
    return 1;
}
