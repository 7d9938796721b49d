// Compile + link check of the drop-in shims against mock reference types.
#include <cstdio>
#include "xrsfm_mock.h"
#include "../../xrsfm_b200/shim/SiftMatchGPU_b200.h"
#include "../../xrsfm_b200/shim/ba_solver_b200.h"

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
    if (false) {  // instantiate, never run without a GPU
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
    return 0;
}
