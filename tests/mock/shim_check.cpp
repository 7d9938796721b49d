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
    return 0;
}
