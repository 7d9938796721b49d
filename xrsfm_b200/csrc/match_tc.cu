// match_tc.cu — generation-2 score kernel (tcgen05 int8 MMA, accumulators in TMEM).
// Placeholder until the kernel lands: reports "not available" so the matcher stays on
// generation 1.
#include "match_kernels.cuh"

namespace xrb {

bool score_tc_available() { return false; }

int launch_score_tc(const PairDesc *, int, int, int, int, Top2State, Top2State, const int *,
                    cudaStream_t) {
    set_error("tcgen05 score kernel not built");
    return XRB_ERR_INVALID;
}

}  // namespace xrb
