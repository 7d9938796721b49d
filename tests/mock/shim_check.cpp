#include <vector>
// Compile + link check of the drop-in shims against mock reference types.
#include <cstdio>
#include "xrsfm_mock.h"
#include "../../xrsfm_b200/shim/SiftMatchGPU_b200.h"
#include "../../xrsfm_b200/shim/ba_solver_b200.h"
#include "../../xrsfm_b200/shim/pnp_b200.h"

int main() {
    // the exact call sequence of CreateSiftGPUMatcher / SiftMatch (feature_processing.cc:53-154)
    SiftMatchGPU m;
    m = SiftMatchGPU(4096);
    m.SetLanguage(SiftMatchGPU::SIFTMATCH_CUDA_DEVICE0 + 0);
    const int ok = m.VerifyContextGL();
    std::printf("VerifyContextGL=%d (%s)\n", ok, ok ? "engine up" : xrb_last_error());
    mock::Map map;
    xrsfm_b200::FlatBA f = xrsfm_b200::Flatten(map, false, false);
    std::printf("flatten: %zu cams\n", f.frame_of_cam.size());
    {   // the LBA window as a flat problem (ba_solver.cc:358-391, 523-584), host logic only
        mock::Map m2;
        m2.cameras_[0].params_ = {700.0, 300.0, 200.0, 0.0};
        m2.frames_.resize(4);
        m2.tracks_.resize(3);
        m2.tracks_[1].angle_ = 7.0;                       // well-conditioned point: constant in LBA
        m2.tracks_[0].observations_[3] = 0, m2.tracks_[1].observations_[3] = 1;  // the new frame (3) sees tracks 0, 1
        for (int i = 0; i < 4; ++i) {
            auto &fr = m2.frames_[i];
            fr.id = i, fr.registered = true;
            fr.points = {{{1.0, 2.0}}, {{3.0, 4.0}}, {{5.0, 6.0}}};
            fr.track_ids_ = {0, 1, i == 0 ? -1 : 2};
        }
        m2.init_id1 = 0, m2.init_id2 = 9;                 // only one gauge frame inside the window
        xrsfm_b200::FlatBA lf = xrsfm_b200::FlattenLBA(m2, 3, std::vector<int>{1, 3}, std::vector<int>{0, 3, 2});
        std::printf("lba: cams=%zu obs=%zu pts=%zu fixed_t=%d%d%d%d fixed_pts=%d%d%d\n", lf.frame_of_cam.size(),
                    lf.obs_cam.size(), lf.track_of_pt.size(), lf.cam_t_fixed[0], lf.cam_t_fixed[1], lf.cam_t_fixed[2],
                    lf.cam_t_fixed[3], lf.pt_fixed[0], lf.pt_fixed[1], lf.pt_fixed[2]);
        m2.init_id1 = 8;                                  // no gauge frame inside: last two of ids2 (3, 2)
        lf = xrsfm_b200::FlattenLBA(m2, 3, std::vector<int>{1, 3}, std::vector<int>{0, 3, 2});
        std::printf("lba2: fixed_t=%d%d%d%d\n", lf.cam_t_fixed[0], lf.cam_t_fixed[1], lf.cam_t_fixed[2], lf.cam_t_fixed[3]);
    }
    if (false) {  // instantiate, never run without a GPU
        xrsfm_b200::LBA_Solve(map, 0, std::vector<int>{}, std::vector<int>{});
        xrsfm_b200::GBA(map, true, false);
        xrsfm_b200::KGBA_Solve(map);
        uint32_t buf[4][2];
        m.Allocate(16384, true);
        m.SetDescriptors(0, 0, nullptr);
        m.GetSiftMatch(4, buf, 0.7f, 0.8f, true);
    }
    // wire formats through the C ABI from C++ (host code only): ftr.bin write -> scan -> read
    {
        const char *path = "/tmp/xrb_shim_check_ftr.bin";
        const int64_t off[3] = {0, 2, 5};
        uint8_t desc[5 * 128];
        for (int i = 0; i < 5 * 128; ++i) desc[i] = (uint8_t)(i * 7);
        const char names[] = "a.jpg\0b.jpg";
        const int64_t noff[3] = {0, 6, 12};
        int rc = xrb_ftr_write(path, 2, off, desc, nullptr, names, noff);
        int32_t n = 0;
        int64_t total = 0, nb = 0;
        rc = rc ? rc : xrb_ftr_scan(path, &n, &total, &nb);
        int64_t off2[3] = {0, 0, 0};
        uint8_t desc2[5 * 128] = {0};
        rc = rc ? rc : xrb_ftr_read(path, n, off2, desc2, nullptr, nullptr, nullptr);
        bool same = rc == 0 && n == 2 && total == 5 && nb == 12 && off2[1] == 2 && off2[2] == 5;
        for (int i = 0; same && i < 5 * 128; ++i) same = desc[i] == desc2[i];
        std::printf("ftr roundtrip: %s\n", same ? "ok" : xrb_last_error());
        std::remove(path);
    }
    {
        // PoseRefiner (pnp.cc:38-71): queueing is host code; without a GPU Run() reports the failure and leaves the
        // frame's pose alone (with one, tests/test_shims.py runs the same object through shim_gpu_check)
        mock::Frame frame;
        frame.Tcw.q.c = {{0, 0, 0, 1}};
        frame.Tcw.t = {{0.5, -0.25, 2.0}};
        frame.points = {{{100, 50}}, {{200, 80}}, {{640, 190}}, {{900, 300}}};
        mock::CameraT cam;
        cam.model_id_ = 2, cam.params_ = {718.856, 607.19, 185.22, -0.02};
        std::vector<std::pair<int, int>> ids = {{0, 7}, {2, 9}, {3, 4}};
        std::vector<mock::Vec3> p3d = {{{1, 2, 10}}, {{0, 0, 12}}, {{3, 1, 9}}};
        std::vector<char> mask = {1, 0, 1};
        xrsfm_b200::PoseRefiner<mock::Frame> refiner(0);
        refiner.Add(frame, cam, ids, p3d, mask);
        const size_t queued = refiner.size();
        const int rc = refiner.Run(false);
        const bool untouched = frame.Tcw.t.v[0] == 0.5 && frame.Tcw.t.v[2] == 2.0 && frame.Tcw.q.c.v[3] == 1.0;
        std::printf("pose: queued=%zu max_it=%d status_ok=%d untouched=%d\n", queued, refiner.options.max_iterations,
                    rc == XRB_OK, (int)untouched);
    }
    return 0;
}
