// ba_api.cu — C ABI of path B (placeholder while the kernels land).
#include "common.cuh"

using namespace xrb;

extern "C" {
void xrb_ba_default_options(xrb_ba_options *o) {
    o->max_iterations = 50;
    o->function_tolerance = 1e-6;
    o->parameter_tolerance = 1e-8;
    o->gradient_tolerance = 1e-10;
    o->initial_radius = 1e4;
    o->huber_a = 5.99;
    o->min_depth = 1e-2;
    o->neg_depth_residual = 12.0;
    o->verbose = 0;
    o->fixed_iterations = 0;
}
xrb_ba_solver *xrb_ba_create(int device) {
    if (select_device(device) != XRB_OK) return nullptr;
    set_error("BA engine not built yet");
    return nullptr;
}
void xrb_ba_destroy(xrb_ba_solver *) {}
int xrb_ba_set_exchange(xrb_ba_solver *, int, int, xrb_allreduce_fn, void *) { return XRB_ERR_INVALID; }
int xrb_ba_solve(xrb_ba_solver *, const xrb_ba_problem *, const xrb_ba_options *, xrb_ba_summary *) { return XRB_ERR_INVALID; }
int xrb_ba_load(xrb_ba_solver *, const xrb_ba_problem *) { return XRB_ERR_INVALID; }
int xrb_ba_reset(xrb_ba_solver *) { return XRB_ERR_INVALID; }
int xrb_ba_run(xrb_ba_solver *, const xrb_ba_options *, xrb_ba_summary *, void *) { return XRB_ERR_INVALID; }
int xrb_ba_fetch(xrb_ba_solver *, xrb_ba_problem *) { return XRB_ERR_INVALID; }
int xrb_ba_residuals(xrb_ba_solver *, double *) { return XRB_ERR_INVALID; }
int xrb_ba_profile(const xrb_ba_solver *, double *, int64_t *) { return XRB_ERR_INVALID; }
}
