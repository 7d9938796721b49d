// ba_api.cu — C ABI of path B and the host-side Levenberg–Marquardt controller.
//
// The controller restates ceres::TrustRegionMinimizer + LevenbergMarquardtStrategy for the
// options BASolver selects (ba_solver.cc:70-77,624-634,665-670); all arithmetic over
// observations, points and the reduced camera system runs in the kernels of
// ba_kernels.cu / ba_chol.cu.  One device->host read of a few scalars per LM iteration is
// the only synchronisation.  Multi-GPU: points are sharded, cameras replicated, one SUM
// all-reduce of the packed reduced system per linear solve (xrb_ba_set_exchange).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "ba_structure.cuh"
#include "nccl_dyn.cuh"

using namespace xrb;

namespace {

struct Ev {  // owns its event: every exit path of a run releases it
    cudaEvent_t e = nullptr;
    Ev() = default;
    Ev(const Ev &) = delete;
    Ev &operator=(const Ev &) = delete;
    ~Ev() {
        if (e) cudaEventDestroy(e);
    }
    void rec(cudaStream_t s) {
        if (!e) cudaEventCreate(&e);
        cudaEventRecord(e, s);
    }
};

}  // namespace

struct xrb_ba_solver {
    int device = 0;
    int rank = 0, world = 1;
    xrb_allreduce_fn fn = nullptr;
    void *user = nullptr;
    NcclApi::Comm comm = nullptr;  // native exchange (xrb_ba_comm_init); takes precedence over fn
    std::vector<int32_t> shard_lo;  // [world + 1] point ranges of every rank (multi-GPU)
    cudaStream_t own_stream = nullptr, side_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    bool loaded = false;
    int C = 0, P_total = 0, O_total = 0, n_intr = 0;
    int P_local = 0, O_local = 0, p_lo = 0;
    int nc = 0, bw = 0;       // nc: padded dimension of the reduced camera system (a multiple of 64)
    int nc_true = 0, parts = 1;  // variable camera columns; independent interiors of the column order
    CholPlan plan;
    CholWorkspace chol_ws;
    // fused windowed Schur complement (sequence-like scenes), see ba_kernels.cu section 2b
    bool use_window = false;
    WindowPlanDev win;
    DevBuf d_anchor, d_win_pts, d_cta_ptr, d_cta_cam0, d_pS, d_pC, d_wstats;
    int cam_span = 0, max_track = 0;
    DevBuf d_holes, d_pat;
    int n_holes = 0;
    int n_var_q = 0, n_var_t = 0, n_var_pts = 0, n_res_blocks = 0;

    // device: problem
    DevBuf d_intr, d_intr_model, d_cam_intr, d_colq, d_colt, d_pt_ptr, d_obs_cam, d_obs_uv,
        d_pt_var, d_obs_orig;
    // Schur structure (ba_struct.cu) and per-observation records
    DevBuf d_obs_pt, d_cam_ptr, d_cam_obs, d_inc, d_blk_ptr, d_blk_cams, d_Tt, d_h;
    BAStructScratch W;  // ba_load.cu / ba_struct.cu working set, kept between loads
    int n_blocks = 0;
    int64_t n_inc = 0;
    // device: states
    DevBuf d_q[3], d_t[3], d_X[3];  // 0 = current, 1 = candidate, 2 = initial copy
    int cur = 0;
    // device: linear system.  E = [rhs | U | Ud | gc | scalE(8) | slots(world) | S tiles (original, then
    // fill) | n2c]: everything one linear solve exchanges is one contiguous prefix that ends with the
    // structurally non-zero tiles of S (the fill is zero on every rank; n2c travels once, at iteration 0)
    DevBuf d_E, d_Vinv, d_gp, d_sc, d_sp, d_linv, d_yc, d_step_p, d_scal, d_full;
    size_t off_rhs = 0, off_U = 0, off_Ud = 0, off_gc = 0, off_n2c = 0, off_scalE = 0, off_slots = 0, off_S = 0,
           off_exch_end = 0, E_count = 0;
    double *h_scal = nullptr;  // pinned mirror: [scalE(8) | slots(world) | scal2(8) | scalL(8)]

    double ms[6] = {0, 0, 0, 0, 0, 0};
    int64_t launches[6] = {0, 0, 0, 0, 0, 0};
    double ms_kernel[3] = {0, 0, 0};  // Schur split: k_lin, k_gather, k_cam_blocks
    int64_t n_steps = 0;              // compute_step calls of the last run

    BAProblemDev prob() const {
        BAProblemDev p;
        p.n_cams = C, p.n_pts_local = P_local, p.n_obs_local = O_local, p.nc = nc;
        p.intr = d_intr.as<double>(), p.intr_model = d_intr_model.as<int32_t>();
        p.cam_intr = d_cam_intr.as<int32_t>();
        p.colq = d_colq.as<int32_t>(), p.colt = d_colt.as<int32_t>();
        p.pt_ptr = d_pt_ptr.as<int32_t>(), p.obs_cam = d_obs_cam.as<int32_t>();
        p.obs_uv = d_obs_uv.as<double>(), p.pt_var = d_pt_var.as<uint8_t>();
        p.obs_pt = d_obs_pt.as<int32_t>(), p.cam_ptr = d_cam_ptr.as<int32_t>(), p.cam_obs = d_cam_obs.as<int32_t>();
        p.n_blocks = n_blocks, p.n_inc = n_inc;
        p.blk_ptr = d_blk_ptr.as<int32_t>(), p.blk_cams = d_blk_cams.as<int2>(), p.inc = d_inc.as<int2>();
        return p;
    }
    BAStateDev state(int i) const { return {d_q[i].as<double>(), d_t[i].as<double>(), d_X[i].as<double>()}; }
    BALinSys linsys() const {
        BALinSys L;
        double *E = d_E.as<double>();
        L.S = E + off_S, L.tm.nt = plan.d.nt, L.tm.tab = plan.d.tab, L.rhs = E + off_rhs, L.U = E + off_U, L.Ud = E + off_Ud, L.gc = E + off_gc, L.n2c = E + off_n2c;
        L.Vinv = d_Vinv.as<double>(), L.gp = d_gp.as<double>();
        L.sc = d_sc.as<double>(), L.sp = d_sp.as<double>();
        L.Tt = d_Tt.as<double>(), L.h = d_h.as<double>();
        return L;
    }
    double *scalE() const { return d_E.as<double>() + off_scalE; }
    double *slots() const { return d_E.as<double>() + off_slots; }
    double *scal2() const { return d_scal.as<double>(); }
    double *scalL() const { return d_scal.as<double>() + SC_COUNT; }

    // In-place SUM over ranks of `count` doubles, ordered on `st`: ncclAllReduce(ncclDouble,
    // ncclSum) on the solver's own stream — no host synchronisation — or the caller's hook.
    int exchange(double *buf, size_t count, cudaStream_t st) {
        if (world <= 1) return XRB_OK;
        if (comm) {
            const NcclApi *nc_ = nccl_api();
            const int r = nc_->AllReduce(buf, buf, count, NcclApi::kFloat64, NcclApi::kSum, comm, st);
            if (r != 0) {
                set_error("ncclAllReduce failed: %s", nc_->GetErrorString(r));
                return XRB_ERR_COMM;
            }
            launches[4]++;
            return XRB_OK;
        }
        if (!fn) {
            set_error("world > 1 but neither xrb_ba_comm_init nor an exchange hook was set");
            return XRB_ERR_COMM;
        }
        if (fn(buf, count, (void *)st, user) != 0) {
            set_error("exchange hook failed");
            return XRB_ERR_COMM;
        }
        launches[4]++;
        return XRB_OK;
    }
};

namespace {

template <class T>
int upload(DevBuf &b, const T *src, size_t n, cudaStream_t st) {
    int rc = b.reserve(std::max<size_t>(n, 1) * sizeof(T));
    if (rc) return rc;
    if (n) XRB_CUDA(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return XRB_OK;
}

struct LoadTimer {  // XRB_BA_DEBUG=1 prints where xrb_ba_load spends its time
    bool on = getenv("XRB_BA_DEBUG") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        cudaDeviceSynchronize();
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[xrb_ba_load] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

int do_load(xrb_ba_solver *s, const xrb_ba_problem *P) {
    XRB_CUDA(cudaSetDevice(s->device));
    LoadTimer lt;
    cudaStream_t st = s->own_stream;
    if (!P || P->n_cams < 0 || P->n_pts < 0 || P->n_obs < 0 || P->n_intr <= 0 || !P->cam_q || !P->cam_t ||
        !P->pts || !P->intr || !P->intr_model || !P->cam_intr || (P->n_obs && (!P->obs_cam || !P->obs_pt || !P->obs_uv))) {
        set_error("ba_load: null or negative field in xrb_ba_problem");
        return XRB_ERR_INVALID;
    }
    const int C = P->n_cams, NP = P->n_pts, NO = P->n_obs;
    for (int c = 0; c < C; ++c)
        if (P->cam_intr[c] < 0 || P->cam_intr[c] >= P->n_intr) {
            set_error("ba_load: camera %d references intrinsics %d out of range", c, P->cam_intr[c]);
            return XRB_ERR_INVALID;
        }
    for (int i = 0; i < P->n_intr; ++i)
        if (P->intr_model[i] < 0 || P->intr_model[i] > 4) {
            set_error("ba_load: unknown camera model id %d (camera_model.hpp defines 0..4)", P->intr_model[i]);
            return XRB_ERR_INVALID;
        }
    s->C = C, s->P_total = NP, s->O_total = NO, s->n_intr = P->n_intr;
    s->loaded = false;

    // ---- structure on the device (ba_load.cu): CSRs, variable flags, bandwidth, shard
    int rc;
    BAStructBufs out{&s->d_colq, &s->d_colt, &s->d_pt_ptr, &s->d_obs_cam, &s->d_obs_uv, &s->d_pt_var,
                     &s->d_obs_orig, &s->d_obs_pt, &s->d_cam_ptr, &s->d_cam_obs};
    BAStructInfo info{};
    if ((rc = ba_build_structure(P, s->rank, s->world, s->W, out, &info, st))) return rc;
    s->nc = info.nc, s->n_var_q = info.n_var_q, s->n_var_t = info.n_var_t, s->n_var_pts = info.n_var_pts;
    s->n_res_blocks = info.n_res_blocks, s->bw = info.bw;
    s->p_lo = info.p_lo, s->P_local = info.P_local, s->O_local = info.O_local;
    s->shard_lo = info.shard_lo;
    lt.lap("structure (device)");
    // ---- column order, tile pattern, symbolic factorisation (ba_plan.cu)
    {
        std::vector<int> widths, cam_of;
        for (int c = 0; c < C; ++c) {
            const int w = (info.h_colq[c] >= 0 ? 3 : 0) + (info.h_colt[c] >= 0 ? 3 : 0);
            if (w) widths.push_back(w), cam_of.push_back(c);
        }
        std::vector<int32_t> start;
        int n_pad = 64, parts = 1;
        const char *order_env = getenv("XRB_BA_ORDER");
        const bool allow_nd = !(order_env && strcmp(order_env, "natural") == 0);
        plan_column_order(widths, info.bw, allow_nd, start, n_pad, parts);
        std::vector<uint8_t> covered((size_t)n_pad, 0);
        for (size_t v = 0; v < cam_of.size(); ++v) {
            const int c = cam_of[v];
            int at = start[v];
            if (info.h_colq[c] >= 0) info.h_colq[c] = at, at += 3;
            if (info.h_colt[c] >= 0) info.h_colt[c] = at, at += 3;
            for (int j = start[v]; j < at; ++j) covered[j] = 1;
        }
        std::vector<int32_t> holes;
        for (int j = 0; j < n_pad; ++j)
            if (!covered[j]) holes.push_back(j);
        s->nc_true = info.nc, s->nc = n_pad, s->parts = parts, s->n_holes = (int)holes.size();
        const int nt = n_pad / 64;
        if (C) {
            XRB_CUDA(cudaMemcpyAsync(s->d_colq.p, info.h_colq.data(), (size_t)C * 4, cudaMemcpyHostToDevice, st));
            XRB_CUDA(cudaMemcpyAsync(s->d_colt.p, info.h_colt.data(), (size_t)C * 4, cudaMemcpyHostToDevice, st));
        }
        if ((rc = s->d_holes.reserve(std::max<size_t>(1, holes.size()) * 4))) return rc;
        if (!holes.empty())
            XRB_CUDA(cudaMemcpyAsync(s->d_holes.p, holes.data(), holes.size() * 4, cudaMemcpyHostToDevice, st));
        if ((rc = s->d_pat.reserve((size_t)nt * nt))) return rc;
        XRB_CUDA(cudaMemsetAsync(s->d_pat.p, 0, (size_t)nt * nt, st));
        if ((rc = launch_tile_pattern(NP, s->W.pt_ptr_g.as<int32_t>(), s->W.vals.as<int32_t>(), s->W.raw_cam.as<int32_t>(),
                                      s->W.pt_var_g.as<uint8_t>(), C, s->d_colq.as<int32_t>(), s->d_colt.as<int32_t>(), nt,
                                      s->d_pat.as<uint8_t>(), st)))
            return rc;
        std::vector<uint8_t> pat((size_t)nt * nt);
        XRB_CUDA(cudaMemcpyAsync(pat.data(), s->d_pat.p, pat.size(), cudaMemcpyDeviceToHost, st));
        XRB_CUDA(cudaStreamSynchronize(st));
        for (size_t v = 0; v < cam_of.size(); ++v) {  // a camera's own 6 x 6 block may straddle two tiles
            const int t0 = start[v] >> 6, t1 = (start[v] + widths[v] - 1) >> 6;
            pat[(size_t)t1 * nt + t0] = 1;
        }
        if ((rc = build_chol_plan(nt, pat.data(), s->plan.h))) {
            set_error("ba_load: symbolic factorisation failed");
            return rc;
        }
        if ((rc = s->plan.upload(st))) return rc;
        lt.lap("plan (order + symbolic)");
    }

    if ((rc = upload(s->d_intr, P->intr, 8 * (size_t)P->n_intr, st))) return rc;
    if ((rc = upload(s->d_intr_model, P->intr_model, (size_t)P->n_intr, st))) return rc;
    if ((rc = upload(s->d_cam_intr, P->cam_intr, (size_t)C, st))) return rc;
    if ((rc = upload(s->d_q[0], P->cam_q, 4 * (size_t)C, st))) return rc;
    if ((rc = upload(s->d_t[0], P->cam_t, 3 * (size_t)C, st))) return rc;
    if ((rc = upload(s->d_X[0], P->pts + 3 * (size_t)s->p_lo, 3 * (size_t)s->P_local, st))) return rc;
    for (int i = 1; i < 3; ++i) {  // candidate slot and the copy xrb_ba_reset restores
        if ((rc = s->d_q[i].reserve(std::max<size_t>(1, 4 * (size_t)C) * 8))) return rc;
        if ((rc = s->d_t[i].reserve(std::max<size_t>(1, 3 * (size_t)C) * 8))) return rc;
        if ((rc = s->d_X[i].reserve(std::max<size_t>(1, 3 * (size_t)s->P_local) * 8))) return rc;
        if (C) XRB_CUDA(cudaMemcpyAsync(s->d_q[i].p, s->d_q[0].p, 4 * (size_t)C * 8, cudaMemcpyDeviceToDevice, st));
        if (C) XRB_CUDA(cudaMemcpyAsync(s->d_t[i].p, s->d_t[0].p, 3 * (size_t)C * 8, cudaMemcpyDeviceToDevice, st));
        if (s->P_local)
            XRB_CUDA(cudaMemcpyAsync(s->d_X[i].p, s->d_X[0].p, 3 * (size_t)s->P_local * 8, cudaMemcpyDeviceToDevice, st));
    }
    s->cur = 0;
    lt.lap("state + intrinsics uploads");
    // ---- Schur structure: fused camera windows when every point sees a narrow range of cameras, else
    // per-observation records + block incidence lists
    {
        s->use_window = false;
        s->n_blocks = 0, s->n_inc = 0;
        BAProblemDev pd = s->prob();
        if ((rc = s->d_anchor.reserve(std::max<size_t>(1, (size_t)s->P_local) * 4))) return rc;
        if ((rc = s->d_wstats.reserve(16))) return rc;
        if ((rc = ba_launch_point_anchor(pd, s->d_anchor.as<int32_t>(), s->d_wstats.as<int32_t>(), st))) return rc;
        int32_t wst[4] = {0, 0, 0, 0};
        XRB_CUDA(cudaMemcpyAsync(wst, s->d_wstats.p, 16, cudaMemcpyDeviceToHost, st));
        XRB_CUDA(cudaStreamSynchronize(st));
        s->cam_span = wst[0], s->max_track = wst[1];
        // opt-in (XRB_BA_SCHUR=window): at C4 the fused kernel runs 9.7 ms against 6.2 ms for the three gather-path
        // kernels — its per-batch phases are latency-bound at 16 warps per SM
        const char *schur_env = getenv("XRB_BA_SCHUR");
        const bool want_window = schur_env && strcmp(schur_env, "window") == 0;
        if (want_window && s->P_local > 0 && C >= 2 * kWinCams && wst[0] <= kWinMaxSpan && wst[1] <= kWinMaxTrack && wst[2] == 0) {
            const int stride = kWinCams - wst[0], n_win = (C + stride - 1) / stride;
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
            const int parts = std::max(1, std::min(8, (4 * sms + n_win - 1) / n_win));
            const int n_ctas = n_win * parts;
            // points by window (counting sort on the host: O(points), once per load), each window split evenly
            std::vector<int32_t> anchor((size_t)s->P_local), cnt((size_t)n_win + 1, 0), pts((size_t)s->P_local);
            XRB_CUDA(cudaMemcpyAsync(anchor.data(), s->d_anchor.p, anchor.size() * 4, cudaMemcpyDeviceToHost, st));
            XRB_CUDA(cudaStreamSynchronize(st));
            // sorted by first camera (stable in the point id): the points of one batch then share their cameras,
            // so a batch touches ~ (span + 1)^2 / 2 camera pairs instead of the whole window's
            {
                std::vector<int32_t> per_cam((size_t)C + 1, 0);
                for (int p = 0; p < s->P_local; ++p) per_cam[(size_t)anchor[p] + 1]++;
                for (int c = 0; c < C; ++c) per_cam[c + 1] += per_cam[c];
                for (int w = 0; w <= n_win; ++w) cnt[w] = per_cam[std::min<int64_t>((int64_t)w * stride, C)];
                for (int p = 0; p < s->P_local; ++p) pts[per_cam[anchor[p]]++] = p;
            }
            std::vector<int32_t> cta_ptr((size_t)n_ctas + 1), cta_cam0((size_t)n_ctas);
            for (int w = 0; w < n_win; ++w) {
                const int64_t b = cnt[w], n = cnt[w + 1] - cnt[w];
                for (int part = 0; part < parts; ++part) {
                    cta_ptr[(size_t)w * parts + part] = (int32_t)(b + n * part / parts);
                    cta_cam0[(size_t)w * parts + part] = w * stride;
                }
            }
            cta_ptr[n_ctas] = s->P_local;
            if ((rc = upload(s->d_win_pts, pts.data(), pts.size(), st))) return rc;
            if ((rc = upload(s->d_cta_ptr, cta_ptr.data(), cta_ptr.size(), st))) return rc;
            if ((rc = upload(s->d_cta_cam0, cta_cam0.data(), cta_cam0.size(), st))) return rc;
            if ((rc = s->d_pS.reserve((size_t)n_ctas * kWinBlocks * 36 * 8))) return rc;
            if ((rc = s->d_pC.reserve((size_t)n_ctas * kWinCams * 54 * 8))) return rc;
            XRB_CUDA(cudaStreamSynchronize(st));  // the host vectors go out of scope
            s->win.win_pts = s->d_win_pts.as<int32_t>(), s->win.cta_ptr = s->d_cta_ptr.as<int32_t>();
            s->win.cta_cam0 = s->d_cta_cam0.as<int32_t>(), s->win.pS = s->d_pS.as<double>(), s->win.pC = s->d_pC.as<double>();
            s->win.n_ctas = n_ctas, s->win.n_win = n_win, s->win.parts = parts, s->win.stride = stride;
            s->use_window = true;
        }
        if (!s->use_window && (rc = s->d_Tt.reserve(std::max<size_t>(1, 18 * (size_t)s->O_local) * 8))) return rc;
        if (s->use_window && (rc = s->d_Tt.reserve(8))) return rc;
        if ((rc = s->d_h.reserve(std::max<size_t>(1, 3 * (size_t)s->P_local) * 8))) return rc;
        pd = s->prob();
        if ((rc = ba_build_block_lists(pd, s->W, s->d_inc, s->d_blk_ptr, s->d_blk_cams, &s->n_blocks, &s->n_inc, st)))
            return rc;
        lt.lap("block lists (device)");
    }
    // ---- linear-system storage
    const int nc = s->nc;
    const int nt = s->plan.h.nt;
    s->off_rhs = 0, s->off_U = s->off_rhs + (size_t)nt * 64;
    s->off_Ud = s->off_U + (size_t)nc * 6, s->off_gc = s->off_Ud + (size_t)nc * 6, s->off_scalE = s->off_gc + nc;
    s->off_slots = s->off_scalE + SC_COUNT;
    s->off_S = (s->off_slots + s->world + 1) / 2 * 2;  // tiles stay 16-byte aligned
    s->off_exch_end = s->off_S + (size_t)s->plan.h.n_tiles_orig * 4096;
    s->off_n2c = s->off_S + (size_t)s->plan.h.n_tiles * 4096;
    s->E_count = s->off_n2c + nc;
    if ((rc = s->d_E.reserve(s->E_count * 8))) return rc;
    if ((rc = s->d_Vinv.reserve(std::max<size_t>(1, 6 * (size_t)s->P_local) * 8))) return rc;
    if ((rc = s->d_gp.reserve(std::max<size_t>(1, 3 * (size_t)s->P_local) * 8))) return rc;
    if ((rc = s->d_sp.reserve(std::max<size_t>(1, 3 * (size_t)s->P_local) * 8))) return rc;
    if ((rc = s->d_step_p.reserve(std::max<size_t>(1, 3 * (size_t)s->P_local) * 8))) return rc;
    if ((rc = s->d_sc.reserve(std::max<size_t>(1, nc) * 8))) return rc;
    if ((rc = s->d_yc.reserve(std::max<size_t>(1, nc) * 8))) return rc;
    if ((rc = s->d_linv.reserve((size_t)(nt + 1) * 4 * 256 * 8))) return rc;
    if ((rc = s->d_scal.reserve(2 * SC_COUNT * 8))) return rc;
    if (!s->h_scal) XRB_CUDA(cudaMallocHost(&s->h_scal, (3 * SC_COUNT + 64) * sizeof(double)));
    if (s->world > 64) {
        set_error("world %d > 64 not supported", s->world);
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaStreamSynchronize(st));
    lt.lap("linear-system buffers");
    s->loaded = true;
    return XRB_OK;
}

struct StepOut {
    double model_cost_change, cand_cost, step_norm, cand_xnorm, grad_max;
    bool ok;
};

// One linear solve + candidate evaluation from the state in slot `cur`; candidate in slot 1-cur.
int compute_step(xrb_ba_solver *s, const BAConsts &k, double radius, cudaStream_t st, StepOut &out,
                 Ev ev[10]) {
    const BAProblemDev P = s->prob();
    const BALinSys L = s->linsys();
    const BAStateDev x = s->state(s->cur), cand = s->state(1 - s->cur);
    const double inv_radius = 1.0 / radius;
    int rc;
    ev[0].rec(st);
    // zero S, rhs, U, gc, scalE + slots (n2c untouched), scal2 + scalL
    XRB_CUDA(cudaMemsetAsync(s->d_E.p, 0, s->off_n2c * 8, st));
    XRB_CUDA(cudaMemsetAsync(s->d_scal.p, 0, 2 * SC_COUNT * 8, st));
    if (s->use_window) {
        // sequence-like scene: one fused kernel (+ two small reductions), no per-observation records
        if ((rc = ba_launch_schur_window(P, x, k, L, inv_radius, s->scalE(), s->win, st))) return rc;
        ev[7].rec(st);
        ev[8].rec(st);
        s->launches[0] += 3;
    } else {
        if ((rc = ba_launch_lin(P, x, k, L, inv_radius, s->scalE(), st))) return rc;
        ev[7].rec(st);
        // the camera-major reduction reads the records but never S: it runs beside the gather
        XRB_CUDA(cudaEventRecord(s->ev_fork, st));
        XRB_CUDA(cudaStreamWaitEvent(s->side_stream, s->ev_fork, 0));
        if ((rc = ba_launch_cam_blocks(P, x, k, L, s->side_stream))) return rc;
        XRB_CUDA(cudaEventRecord(s->ev_join, s->side_stream));
        if ((rc = ba_launch_gather(P, L, st))) return rc;
        ev[8].rec(st);
        XRB_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
        s->launches[0] += 3;
    }
    ev[1].rec(st);
    if (s->world > 1) {
        XRB_CUDA(cudaMemcpyAsync(s->slots() + s->rank, s->scalE() + SC_GRAD_MAX_PT, 8, cudaMemcpyDeviceToDevice, st));
        // S | U | gc | scalars | per-rank slots: ONE all-reduce per linear solve
        if ((rc = s->exchange(s->d_E.as<double>(), s->off_exch_end, st))) return rc;
    }
    ev[2].rec(st);
    if ((rc = ba_launch_set_holes(L, s->d_holes.as<int32_t>(), s->n_holes, st))) return rc;
    if ((rc = ba_launch_cam_diag(P, x, L, inv_radius, s->scalL(), st))) return rc;
    if ((rc = ba_launch_tile_cholesky_solve(s->plan.d, s->chol_ws, L.S, L.rhs, s->d_linv.as<double>(), s->d_yc.as<double>(),
                                            s->scalL() + SC_FAIL, st, &s->launches[1])))
        return rc;
    s->launches[1]++;
    ev[3].rec(st);
    if ((rc = ba_launch_backsub(P, x, cand, k, L, s->d_yc.as<double>(), s->d_step_p.as<double>(), s->scal2(), st))) return rc;
    if ((rc = ba_launch_cam_update(P, x, cand, L, s->d_yc.as<double>(), s->scal2(), s->rank == 0, st))) return rc;
    s->launches[2] += 2;
    ev[4].rec(st);
    if ((rc = ba_launch_cost(P, cand, k, 0, s->scal2() + SC_CAND_COST, st))) return rc;
    s->launches[3]++;
    ev[5].rec(st);
    if ((rc = s->exchange(s->scal2(), SC_COUNT, st))) return rc;
    ev[6].rec(st);
    double *h = s->h_scal;
    XRB_CUDA(cudaMemcpyAsync(h, s->scalE(), (SC_COUNT + s->world) * 8, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaMemcpyAsync(h + SC_COUNT + 64, s->d_scal.p, 2 * SC_COUNT * 8, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    const double *hE = h, *hslots = h + SC_COUNT, *h2 = h + SC_COUNT + 64, *hL = h2 + SC_COUNT;
    double gpt = 0.0;
    if (s->world > 1)
        for (int r = 0; r < s->world; ++r) gpt = std::max(gpt, hslots[r]);
    else
        gpt = hE[SC_GRAD_MAX_PT];
    out.grad_max = std::max(gpt, hL[SC_GRAD_MAX_CAM]);
    out.model_cost_change = h2[SC_MODEL_CHANGE];
    out.cand_cost = h2[SC_CAND_COST];
    out.step_norm = std::sqrt(h2[SC_STEP_NORM2]);
    out.cand_xnorm = std::sqrt(h2[SC_XNORM2]);
    out.ok = hE[SC_FAIL] == 0.0 && h2[SC_FAIL] == 0.0 && hL[SC_FAIL] == 0.0 && std::isfinite(out.model_cost_change) &&
             std::isfinite(out.cand_cost) && std::isfinite(out.step_norm);
    // phase timings
    s->n_steps++;
    {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev[0].e, ev[7].e) == cudaSuccess) s->ms_kernel[0] += t;  // memsets + k_lin
        if (cudaEventElapsedTime(&t, ev[7].e, ev[8].e) == cudaSuccess) s->ms_kernel[1] += t;
        if (cudaEventElapsedTime(&t, ev[8].e, ev[1].e) == cudaSuccess) s->ms_kernel[2] += t;
    }
    static const int phase_of[6] = {0, 4, 1, 2, 3, 4};
    for (int i = 0; i < 6; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev[i].e, ev[i + 1].e) == cudaSuccess) s->ms[phase_of[i]] += t;
    }
    return XRB_OK;
}

int reduce_scalar_cost(xrb_ba_solver *s, const BAConsts &k, int state_slot, int mode, cudaStream_t st, double &cost) {
    XRB_CUDA(cudaMemsetAsync(s->d_scal.p, 0, 2 * SC_COUNT * 8, st));
    int rc = ba_launch_cost(s->prob(), s->state(state_slot), k, mode, s->scal2() + SC_COST, st);
    if (rc) return rc;
    s->launches[3]++;
    if ((rc = s->exchange(s->scal2(), SC_COUNT, st))) return rc;
    XRB_CUDA(cudaMemcpyAsync(s->h_scal, s->scal2(), SC_COUNT * 8, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    cost = s->h_scal[SC_COST];
    return XRB_OK;
}

int do_run(xrb_ba_solver *s, const xrb_ba_options *O, xrb_ba_summary *sum, cudaStream_t st) {
    if (!s->loaded) {
        set_error("ba_run: no problem loaded");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(s->device));
    const auto t0 = std::chrono::steady_clock::now();
    memset(sum, 0, sizeof *sum);
    for (int i = 0; i < 6; ++i) s->ms[i] = 0.0, s->launches[i] = 0;
    s->ms_kernel[0] = s->ms_kernel[1] = s->ms_kernel[2] = 0.0, s->n_steps = 0;
    BAConsts k{O->huber_a, O->huber_a * O->huber_a, O->min_depth, O->neg_depth_residual};
    sum->num_residuals_reduced = 2 * s->n_res_blocks;
    sum->num_effective_parameters_reduced = 3 * (s->n_var_q + s->n_var_t + s->n_var_pts);
    sum->termination_type = XRB_BA_NO_CONVERGENCE;
    Ev ev[10], ev_run[2];
    ev_run[0].rec(st);
    int rc;
    double fixed_cost = 0.0, x_cost = 0.0;
    if ((rc = reduce_scalar_cost(s, k, s->cur, 1, st, fixed_cost))) return rc;
    sum->fixed_cost = fixed_cost;
    if (s->n_res_blocks == 0 || sum->num_effective_parameters_reduced == 0) {
        sum->initial_cost = sum->final_cost = fixed_cost;
        sum->termination_type = XRB_BA_CONVERGENCE;
        return XRB_OK;
    }
    if ((rc = reduce_scalar_cost(s, k, s->cur, 0, st, x_cost))) return rc;
    // ---- Jacobi scaling from the iteration-0 Jacobian, and ||x|| over the variable blocks of the start point
    // (both on the device: the scales start at 1, the norm is one reduction; nothing is copied back but a double)
    double xnorm;
    {
        const BALinSys L = s->linsys();
        XRB_CUDA(cudaMemsetAsync(s->d_scal.p, 0, 2 * SC_COUNT * 8, st));
        if ((rc = ba_launch_start_norm(s->prob(), s->state(s->cur), L, s->rank == 0, s->scal2() + SC_XNORM2, st))) return rc;
        if ((rc = ba_launch_colnorm(s->prob(), s->state(s->cur), k, L, st))) return rc;
        if ((rc = s->exchange(L.n2c, (size_t)s->nc, st))) return rc;
        if ((rc = ba_launch_finish_scaling(s->prob(), L, st))) return rc;
        s->launches[0] += 3;
        if ((rc = s->exchange(s->scal2(), SC_COUNT, st))) return rc;
        double n2 = 0.0;
        XRB_CUDA(cudaMemcpyAsync(&n2, s->scal2() + SC_XNORM2, 8, cudaMemcpyDeviceToHost, st));
        XRB_CUDA(cudaStreamSynchronize(st));
        xnorm = std::sqrt(n2);
    }

    double min_cost = 0.0;
    bool have_min = false;
    auto log_iter = [&](const xrb_ba_iteration &it) {
        if (sum->n_iterations_logged < 128) sum->iterations[sum->n_iterations_logged++] = it;
        if (!have_min || it.cost < min_cost) min_cost = it.cost, have_min = true;
        if (it.step_is_successful) sum->num_successful_steps++; else sum->num_unsuccessful_steps++;
        if (O->verbose && s->rank == 0)
            printf("%4d % .6e % .2e % .2e % .2e % .2e % .2e\n", it.iteration, it.cost, it.cost_change,
                   it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius);
    };

    double radius = O->initial_radius, decrease_factor = 2.0;
    xrb_ba_iteration it;
    memset(&it, 0, sizeof it);
    it.iteration = 0, it.step_is_valid = 1, it.step_is_successful = 1;
    it.cost = x_cost + fixed_cost, it.trust_region_radius = radius;
    sum->initial_cost = it.cost;
    StepOut so;
    if ((rc = compute_step(s, k, radius, st, so, ev))) return rc;
    bool have_step = true;
    it.gradient_max_norm = so.grad_max;
    log_iter(it);
    int iteration = 0, consecutive_invalid = 0;
    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue
        if (iteration >= O->max_iterations) { sum->termination_type = XRB_BA_NO_CONVERGENCE; break; }
        if (!O->fixed_iterations && it.gradient_max_norm <= O->gradient_tolerance) { sum->termination_type = XRB_BA_CONVERGENCE; break; }
        if (radius <= 1e-32) { sum->termination_type = XRB_BA_CONVERGENCE; break; }
        iteration++;
        xrb_ba_iteration cur;
        memset(&cur, 0, sizeof cur);
        cur.iteration = iteration;
        if (!have_step && (rc = compute_step(s, k, radius, st, so, ev))) return rc;
        have_step = false;
        sum->num_lm_iterations++;
        cur.model_cost_change = so.model_cost_change;
        cur.step_is_valid = so.ok && so.model_cost_change > 0.0;
        if (!cur.step_is_valid) {  // HandleInvalidStep
            if (++consecutive_invalid >= 5) { sum->termination_type = XRB_BA_FAILURE; break; }
            radius *= 0.5;
            cur.cost = x_cost + fixed_cost;
            cur.gradient_max_norm = it.gradient_max_norm;
            cur.trust_region_radius = radius;
            it = cur;
            log_iter(it);
            continue;
        }
        consecutive_invalid = 0;
        cur.step_norm = so.step_norm;
        // ParameterToleranceReached, then FunctionToleranceReached — before accept/reject
        if (!O->fixed_iterations && cur.step_norm <= O->parameter_tolerance * (xnorm + O->parameter_tolerance)) {
            sum->termination_type = XRB_BA_CONVERGENCE;
            break;
        }
        cur.cost_change = x_cost - so.cand_cost;
        if (!O->fixed_iterations && std::fabs(cur.cost_change) <= O->function_tolerance * x_cost) {
            sum->termination_type = XRB_BA_CONVERGENCE;
            break;
        }
        cur.relative_decrease = cur.cost_change / so.model_cost_change;
        const double cand_cost = so.cand_cost;
        if (cur.relative_decrease > 1e-3) {  // HandleSuccessfulStep
            s->cur = 1 - s->cur;
            xnorm = so.cand_xnorm;
            x_cost = cand_cost;
            cur.step_is_successful = 1;
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * cur.relative_decrease - 1.0, 3));
            radius = std::min(1e16, radius);
            decrease_factor = 2.0;
            if ((rc = compute_step(s, k, radius, st, so, ev))) return rc;  // also the gradient at the new point
            have_step = true;
            cur.gradient_max_norm = so.grad_max;
        } else {  // StepRejected
            radius = radius / decrease_factor;
            decrease_factor *= 2.0;
            cur.gradient_max_norm = it.gradient_max_norm;
        }
        cur.cost = cand_cost + fixed_cost;
        cur.trust_region_radius = radius;
        it = cur;
        log_iter(it);
    }
    sum->final_cost = have_min ? std::min(sum->initial_cost, min_cost) : sum->initial_cost;
    ev_run[1].rec(st);
    XRB_CUDA(cudaStreamSynchronize(st));
    float tr = 0.f;
    cudaEventElapsedTime(&tr, ev_run[0].e, ev_run[1].e);
    s->ms[5] = tr;
    sum->linear_solver_seconds = (s->ms[0] + s->ms[1] + s->ms[2]) * 1e-3;
    sum->residual_seconds = s->ms[3] * 1e-3;
    sum->total_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return XRB_OK;
}

int do_fetch(xrb_ba_solver *s, xrb_ba_problem *P) {
    if (!s->loaded || !P || P->n_cams != s->C || P->n_pts != s->P_total) {
        set_error("ba_fetch: problem does not match the loaded one");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->own_stream;
    XRB_CUDA(cudaMemcpyAsync(P->cam_q, s->d_q[s->cur].p, 4 * (size_t)s->C * 8, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaMemcpyAsync(P->cam_t, s->d_t[s->cur].p, 3 * (size_t)s->C * 8, cudaMemcpyDeviceToHost, st));
    if (s->world <= 1) {
        if (s->P_local)
            XRB_CUDA(cudaMemcpyAsync(P->pts, s->d_X[s->cur].p, 3 * (size_t)s->P_local * 8, cudaMemcpyDeviceToHost, st));
    } else {  // every rank returns every point
        int rc = s->d_full.reserve(std::max<size_t>(1, 3 * (size_t)s->P_total) * 8);
        if (rc) return rc;
        if (s->comm && (int)s->shard_lo.size() == s->world + 1) {
            // all-gather of unequal shards: one broadcast per owner, grouped into one NCCL operation
            const NcclApi *nc_ = nccl_api();
            if (s->P_local)
                XRB_CUDA(cudaMemcpyAsync(s->d_full.as<double>() + 3 * (size_t)s->p_lo, s->d_X[s->cur].p,
                                         3 * (size_t)s->P_local * 8, cudaMemcpyDeviceToDevice, st));
            int r = nc_->GroupStart();
            for (int o = 0; o < s->world && r == 0; ++o) {
                const size_t lo = s->shard_lo[o], n = (size_t)s->shard_lo[o + 1] - lo;
                if (!n) continue;
                double *seg = s->d_full.as<double>() + 3 * lo;
                r = nc_->Broadcast(seg, seg, 3 * n, NcclApi::kFloat64, o, s->comm, st);
            }
            const int r2 = nc_->GroupEnd();
            if (r != 0 || r2 != 0) {
                set_error("ncclBroadcast (shard gather) failed: %s", nc_->GetErrorString(r ? r : r2));
                return XRB_ERR_COMM;
            }
            s->launches[4]++;
        } else {  // hook: zero-padded shards summed
            XRB_CUDA(cudaMemsetAsync(s->d_full.p, 0, 3 * (size_t)s->P_total * 8, st));
            if (s->P_local)
                XRB_CUDA(cudaMemcpyAsync(s->d_full.as<double>() + 3 * (size_t)s->p_lo, s->d_X[s->cur].p,
                                         3 * (size_t)s->P_local * 8, cudaMemcpyDeviceToDevice, st));
            if ((rc = s->exchange(s->d_full.as<double>(), 3 * (size_t)s->P_total, st))) return rc;
        }
        XRB_CUDA(cudaMemcpyAsync(P->pts, s->d_full.p, 3 * (size_t)s->P_total * 8, cudaMemcpyDeviceToHost, st));
    }
    XRB_CUDA(cudaStreamSynchronize(st));
    return XRB_OK;
}

}  // namespace

extern "C" {

void xrb_ba_default_options(xrb_ba_options *o) {
    if (!o) return;
    o->max_iterations = 50;        // ceres::Solver::Options defaults ...
    o->function_tolerance = 1e-6;
    o->parameter_tolerance = 1e-8;
    o->gradient_tolerance = 1e-10;
    o->initial_radius = 1e4;
    o->huber_a = 5.99;             // ... and the reference's constants: ba_solver.cc:343,
    o->min_depth = 1e-2;           // cost_factor_ceres.h:29
    o->neg_depth_residual = 12.0;  // cost_factor_ceres.h:31
    o->verbose = 0;
    o->fixed_iterations = 0;
}

int xrb_ba_shard_range(int32_t n_pts, const int32_t *obs_per_point, int rank, int world, int32_t *lo,
                       int32_t *hi) {
    if (n_pts < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi || (n_pts && !obs_per_point)) {
        set_error("ba_shard_range: bad arguments");
        return XRB_ERR_INVALID;
    }
    std::vector<double> w((size_t)n_pts + 1, 0.0);
    for (int p = 0; p < n_pts; ++p) {
        const double k = obs_per_point[p];
        w[p + 1] = w[p] + k * k + 4.0 * k;
    }
    auto cut = [&](int r) {
        if (r <= 0) return 0;
        if (r >= world) return (int)n_pts;
        const double target = w[n_pts] * r / world;
        return (int)(std::lower_bound(w.begin(), w.end(), target) - w.begin());
    };
    *lo = std::min<int>(n_pts, cut(rank));
    *hi = std::max<int>(*lo, std::min<int>(n_pts, cut(rank + 1)));
    return XRB_OK;
}

xrb_ba_solver *xrb_ba_create(int device) {
    if (select_device(device) != XRB_OK) return nullptr;
    xrb_ba_solver *s = new xrb_ba_solver();
    s->device = device;
    if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        delete s;
        return nullptr;
    }
    return s;
}

void xrb_ba_destroy(xrb_ba_solver *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->own_stream);
    DevBuf *bufs[] = {&s->d_intr, &s->d_intr_model, &s->d_cam_intr, &s->d_colq, &s->d_colt, &s->d_pt_ptr,
                      &s->d_obs_cam, &s->d_obs_uv, &s->d_pt_var, &s->d_obs_orig, &s->d_E, &s->d_Vinv, &s->d_gp,
                      &s->d_sc, &s->d_sp, &s->d_linv, &s->d_yc, &s->d_step_p, &s->d_scal, &s->d_full,
                      &s->d_obs_pt, &s->d_cam_ptr, &s->d_cam_obs, &s->d_inc, &s->d_blk_ptr, &s->d_blk_cams,
                      &s->d_Tt, &s->d_h, &s->d_holes, &s->d_pat, &s->d_anchor, &s->d_win_pts, &s->d_cta_ptr,
                      &s->d_cta_cam0, &s->d_pS, &s->d_pC, &s->d_wstats};
    for (DevBuf *b : bufs) b->release();
    s->plan.release();
    s->chol_ws.release();
    s->W.release();
    for (int i = 0; i < 3; ++i) s->d_q[i].release(), s->d_t[i].release(), s->d_X[i].release();
    if (s->h_scal) cudaFreeHost(s->h_scal);
    if (s->comm)
        if (const NcclApi *nc_ = nccl_api()) nc_->CommDestroy(s->comm);
    cudaStreamDestroy(s->own_stream);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    if (s->ev_fork) cudaEventDestroy(s->ev_fork);
    if (s->ev_join) cudaEventDestroy(s->ev_join);
    delete s;
}

int xrb_ba_set_exchange(xrb_ba_solver *s, int rank, int world, xrb_allreduce_fn fn, void *user) {
    if (!s || world < 1 || rank < 0 || rank >= world || (world > 1 && !fn)) {
        set_error("ba_set_exchange: bad rank/world/hook");
        return XRB_ERR_INVALID;
    }
    s->rank = rank, s->world = world, s->fn = fn, s->user = user;
    s->loaded = false;  // sharding changes
    return XRB_OK;
}

int xrb_nccl_unique_id(uint8_t id[XRB_NCCL_ID_BYTES]) {
    static_assert(sizeof(NcclApi::UniqueId) == XRB_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
    const NcclApi *nc_ = nccl_api();
    if (!nc_ || !id) return nc_ ? XRB_ERR_INVALID : XRB_ERR_COMM;
    NcclApi::UniqueId u;
    const int r = nc_->GetUniqueId(&u);
    if (r != 0) {
        set_error("ncclGetUniqueId failed: %s", nc_->GetErrorString(r));
        return XRB_ERR_COMM;
    }
    memcpy(id, u.internal, XRB_NCCL_ID_BYTES);
    return XRB_OK;
}

int xrb_ba_comm_init(xrb_ba_solver *s, const uint8_t id[XRB_NCCL_ID_BYTES], int rank, int world) {
    if (!s || !id || world < 1 || rank < 0 || rank >= world || world > 64) {
        set_error("ba_comm_init: bad rank/world (world <= 64)");
        return XRB_ERR_INVALID;
    }
    XRB_CUDA(cudaSetDevice(s->device));
    if (s->comm) {
        if (const NcclApi *nc_ = nccl_api()) nc_->CommDestroy(s->comm);
        s->comm = nullptr;
    }
    s->rank = rank, s->world = world, s->loaded = false;
    if (world == 1) return XRB_OK;
    const NcclApi *nc_ = nccl_api();
    if (!nc_) return XRB_ERR_COMM;
    NcclApi::UniqueId u;
    memcpy(u.internal, id, XRB_NCCL_ID_BYTES);
    const int r = nc_->CommInitRank(&s->comm, world, u, rank);
    if (r != 0) {
        s->comm = nullptr;
        set_error("ncclCommInitRank failed: %s", nc_->GetErrorString(r));
        return XRB_ERR_COMM;
    }
    return XRB_OK;
}

int xrb_ba_load(xrb_ba_solver *s, const xrb_ba_problem *prob) {
    if (!s) return XRB_ERR_INVALID;
    return do_load(s, prob);
}

int xrb_ba_reset(xrb_ba_solver *s) {
    if (!s || !s->loaded) return XRB_ERR_INVALID;
    XRB_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->own_stream;
    s->cur = 0;
    XRB_CUDA(cudaMemcpyAsync(s->d_q[0].p, s->d_q[2].p, 4 * (size_t)s->C * 8, cudaMemcpyDeviceToDevice, st));
    XRB_CUDA(cudaMemcpyAsync(s->d_t[0].p, s->d_t[2].p, 3 * (size_t)s->C * 8, cudaMemcpyDeviceToDevice, st));
    if (s->P_local)
        XRB_CUDA(cudaMemcpyAsync(s->d_X[0].p, s->d_X[2].p, 3 * (size_t)s->P_local * 8, cudaMemcpyDeviceToDevice, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    return XRB_OK;
}

int xrb_ba_run(xrb_ba_solver *s, const xrb_ba_options *opt, xrb_ba_summary *summary, void *stream) {
    if (!s || !opt || !summary) return XRB_ERR_INVALID;
    return do_run(s, opt, summary, stream ? (cudaStream_t)stream : s->own_stream);
}

int xrb_ba_fetch(xrb_ba_solver *s, xrb_ba_problem *prob) {
    if (!s) return XRB_ERR_INVALID;
    return do_fetch(s, prob);
}

int xrb_ba_solve(xrb_ba_solver *s, const xrb_ba_problem *prob, const xrb_ba_options *opt,
                 xrb_ba_summary *summary) {
    if (!s || !prob || !opt || !summary) return XRB_ERR_INVALID;
    const auto t0 = std::chrono::steady_clock::now();
    int rc = do_load(s, prob);
    if (rc) return rc;
    if ((rc = do_run(s, opt, summary, s->own_stream))) return rc;
    if ((rc = do_fetch(s, const_cast<xrb_ba_problem *>(prob)))) return rc;
    summary->total_time_in_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return XRB_OK;
}

int xrb_ba_residuals(xrb_ba_solver *s, double *out) {
    if (!s || !s->loaded || !out) return XRB_ERR_INVALID;
    XRB_CUDA(cudaSetDevice(s->device));
    if (s->world > 1) {
        set_error("ba_residuals: single-GPU only");
        return XRB_ERR_INVALID;
    }
    cudaStream_t st = s->own_stream;
    DevBuf tmp;
    int rc = tmp.reserve(std::max<size_t>(1, 2 * (size_t)s->O_total) * 8);
    if (rc) return rc;
    BAConsts k{5.99, 5.99 * 5.99, 1e-2, 12.0};
    rc = ba_launch_residuals(s->prob(), s->state(s->cur), k, s->d_obs_orig.as<int32_t>(), tmp.as<double>(), st);
    if (rc == XRB_OK && cudaMemcpyAsync(out, tmp.p, 2 * (size_t)s->O_total * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess)
        rc = XRB_ERR_CUDA;
    cudaStreamSynchronize(st);
    tmp.release();
    return rc;
}

/* Many small problems (local-BA windows) on one device: see include/xrsfm_b200.h */
int xrb_ba_solve_batch(int device, int n_problems, const xrb_ba_problem *problems, const xrb_ba_options *opt,
                       xrb_ba_summary *summaries, int n_workers) {
    if (n_problems < 0 || !opt || (n_problems && (!problems || !summaries))) {
        set_error("ba_solve_batch: bad arguments");
        return XRB_ERR_INVALID;
    }
    if (n_problems == 0) return XRB_OK;
    n_workers = std::max(1, std::min(std::min(n_workers <= 0 ? 8 : n_workers, 32), n_problems));
    // one engine per worker and device, kept for the life of the process (their buffers are grow-only)
    static std::mutex pool_mutex;
    static std::map<int, std::vector<xrb_ba_solver *>> pool;
    std::vector<xrb_ba_solver *> engines;
    {
        std::lock_guard<std::mutex> lock(pool_mutex);
        std::vector<xrb_ba_solver *> &mine = pool[device];
        while ((int)mine.size() < n_workers) {
            xrb_ba_solver *h = xrb_ba_create(device);
            if (!h) return XRB_ERR_NO_DEVICE;
            mine.push_back(h);
        }
        engines.assign(mine.begin(), mine.begin() + n_workers);
    }
    static std::mutex run_mutex;  // one batch at a time per process: the engines are not shared between batches
    std::lock_guard<std::mutex> run_lock(run_mutex);
    std::atomic<int> next{0}, first_rc{XRB_OK};
    std::string first_error;
    std::mutex err_mutex;
    auto work = [&](xrb_ba_solver *h) {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n_problems) break;
            const int rc = xrb_ba_solve(h, &problems[i], opt, &summaries[i]);
            if (rc != XRB_OK) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (first_rc.load() == XRB_OK) first_rc = rc, first_error = xrb_last_error();
            }
        }
    };
    std::vector<std::thread> threads;
    for (int w = 1; w < n_workers; ++w) threads.emplace_back(work, engines[w]);
    work(engines[0]);
    for (auto &t : threads) t.join();
    if (first_rc.load() != XRB_OK) set_error("ba_solve_batch: %s", first_error.c_str());
    return first_rc.load();
}

/* Post-BA filter on the solver's current state (single GPU): see include/xrsfm_b200.h */
int xrb_ba_filter_points3d(xrb_ba_solver *s, double max_re, double deg, uint8_t *keep_obs, uint8_t *pt_outlier,
                           double *pt_error, double *pt_angle, int32_t counts[2]) {
    if (!s || !s->loaded || !keep_obs || !pt_outlier || !pt_error || !pt_angle || !counts) return XRB_ERR_INVALID;
    XRB_CUDA(cudaSetDevice(s->device));
    if (s->world > 1) {
        set_error("ba_filter_points3d: single-GPU only");
        return XRB_ERR_INVALID;
    }
    cudaStream_t st = s->own_stream;
    const size_t NO = (size_t)s->O_total, NP = (size_t)s->P_total;
    DevBuf tmp;
    const size_t o_ctr = 0, o_err = o_ctr + 24 * (size_t)std::max(1, s->C), o_ang = o_err + 8 * std::max<size_t>(1, NP);
    const size_t o_order = o_ang + 8 * std::max<size_t>(1, NP), o_cnt = o_order + 4 * std::max<size_t>(1, NO);
    const size_t o_flag = o_cnt + 16, o_keep = o_flag + std::max<size_t>(1, NO), o_out = o_keep + std::max<size_t>(1, NO);
    int rc = tmp.reserve(o_out + std::max<size_t>(1, NP));
    if (rc) return rc;
    char *b = tmp.as<char>();
    rc = ba_launch_filter(s->prob(), s->state(s->cur), s->d_obs_orig.as<int32_t>(), reinterpret_cast<double *>(b + o_ctr),
                          reinterpret_cast<int32_t *>(b + o_order), reinterpret_cast<uint8_t *>(b + o_flag), max_re, deg,
                          reinterpret_cast<uint8_t *>(b + o_keep), reinterpret_cast<uint8_t *>(b + o_out),
                          reinterpret_cast<double *>(b + o_err), reinterpret_cast<double *>(b + o_ang),
                          reinterpret_cast<int32_t *>(b + o_cnt), st);
    if (rc == XRB_OK) {
        bool ok = true;
        if (NO) ok &= cudaMemcpyAsync(keep_obs, b + o_keep, NO, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (NP) {
            ok &= cudaMemcpyAsync(pt_outlier, b + o_out, NP, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok &= cudaMemcpyAsync(pt_error, b + o_err, NP * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok &= cudaMemcpyAsync(pt_angle, b + o_ang, NP * 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        }
        ok &= cudaMemcpyAsync(counts, b + o_cnt, 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
        if (!ok) rc = XRB_ERR_CUDA;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = XRB_ERR_CUDA;
    tmp.release();
    return rc;
}

/* debug hook (not part of the reference surface): timeline of the Cholesky kernels */
int xrb_debug_chol_trace(int enable, int64_t *out, int cap_records) {
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    return ba_tile_cholesky_trace(enable, reinterpret_cast<long long *>(out), cap_records);
}

/* debug hook (not part of the reference surface): factor + solve one dense-stored SPD system with
 * the tile solver, on the current device.  The tile pattern is taken from the non-zeros of A. */
int xrb_debug_tile_solve(int n, int bw, const double *A, const double *rhs, double *x_out, int reps,
                         double *ms_out) {
    if (n <= 0 || !A || !rhs || !x_out || reps < 1) return XRB_ERR_INVALID;
    XRB_CUDA(cudaFree(0));
    const int nt = (n + 63) / 64, np = nt * 64;
    std::vector<uint8_t> pat((size_t)nt * nt, 0);
    for (int r = 0; r < n; ++r)
        for (int c = std::max(0, r - bw); c <= r; ++c)
            if (A[(size_t)r * n + c] != 0.0) pat[(size_t)(r >> 6) * nt + (c >> 6)] = 1;
    CholPlan plan;
    int rc = build_chol_plan(nt, pat.data(), plan.h);
    if (rc) return rc;
    cudaStream_t st;
    XRB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if ((rc = plan.upload(st))) return rc;
    const size_t nS = (size_t)plan.h.n_tiles * 4096, nR = (size_t)np;
    std::vector<double> packed(nS + nR, 0.0);
    auto at = [&](int r, int c) { return (size_t)plan.h.tab[(size_t)(r >> 6) * nt + (c >> 6)] * 4096 + (size_t)((r & 63) * 64 + (c & 63)); };
    for (int r = 0; r < n; ++r)
        for (int c = std::max(0, r - bw); c <= r; ++c)
            if (A[(size_t)r * n + c] != 0.0) packed[at(r, c)] = A[(size_t)r * n + c];
    for (int r = n; r < np; ++r) packed[at(r, r)] = 1.0;
    for (int r = 0; r < n; ++r) packed[nS + r] = rhs[r];
    DevBuf E, dinv, x, fail;
    CholWorkspace ws;
    if ((rc = E.reserve(packed.size() * 8)) || (rc = dinv.reserve((size_t)(nt + 1) * 1024 * 8)) ||
        (rc = x.reserve(nR * 8)) || (rc = fail.reserve(8)))
        return rc;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e30f;
    for (int it = 0; it < reps && rc == XRB_OK; ++it) {
        cudaMemcpyAsync(E.p, packed.data(), packed.size() * 8, cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(fail.p, 0, 8, st);
        cudaMemsetAsync(x.p, 0, nR * 8, st);
        cudaEventRecord(e0, st);
        rc = ba_launch_tile_cholesky_solve(plan.d, ws, E.as<double>(), E.as<double>() + nS, dinv.as<double>(), x.as<double>(),
                                           fail.as<double>(), st, nullptr);
        cudaEventRecord(e1, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = XRB_ERR_CUDA;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    double f = 0.0;
    if (rc == XRB_OK) {
        cudaMemcpy(x_out, x.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(&f, fail.p, 8, cudaMemcpyDeviceToHost);
        if (f != 0.0) {
            set_error("tile solve: %s", f == 2.0 ? "a dependency wait gave up (aborted)" : "non-positive pivot");
            rc = XRB_ERR_NUMERIC;
        }
    } else if (rc == XRB_ERR_CUDA) {
        set_error("tile solve: %s", cudaGetErrorString(cudaGetLastError()));
    }
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    E.release(), dinv.release(), x.release(), fail.release(), plan.release(), ws.release();
    return rc;
}

/* debug hooks, host only (no device needed): the symbolic plan of a tile pattern and the column order.
 * counts[16] = nt, n_tiles, n_tiles_orig, n_f, n_w, n_b, n_wb, n_far, n_chain_f, n_chain_b, depth_f, depth_b.
 * Output arrays may be null; each is filled up to its capacity (in int32 elements). */
int xrb_debug_chol_plan(int nt, const uint8_t *pat, int32_t *counts, int32_t *tab, int32_t *ftasks, int cap_f,
                        int32_t *wtasks, int cap_w, int32_t *btasks, int cap_b, int32_t *wbtasks, int cap_wb,
                        int32_t *far_rows, int32_t *far_slots, int cap_far) {
    if (nt <= 0 || !pat || !counts) return XRB_ERR_INVALID;
    CholPlanHost H;
    const int rc = build_chol_plan(nt, pat, H);
    if (rc) return rc;
    const int32_t c[16] = {H.nt, H.n_tiles, H.n_tiles_orig, H.n_f, H.n_w, H.n_b, H.n_wb, (int32_t)H.far_rows.size(),
                           H.n_chain_f, H.n_chain_b, H.depth_f, H.depth_b, 0, 0, 0, 0};
    memcpy(counts, c, sizeof c);
    auto copy = [](int32_t *dst, int cap, const std::vector<int32_t> &v) {
        if (dst && cap > 0 && !v.empty()) memcpy(dst, v.data(), std::min<size_t>(cap, v.size()) * 4);
    };
    copy(tab, nt * nt, H.tab), copy(ftasks, cap_f, H.ftasks), copy(wtasks, cap_w, H.wtasks), copy(btasks, cap_b, H.btasks);
    copy(wbtasks, cap_wb, H.wbtasks), copy(far_rows, cap_far, H.far_rows), copy(far_slots, cap_far, H.far_slots);
    return XRB_OK;
}

int xrb_debug_column_order(int n_cams, const int32_t *widths, int bw, int allow_nd, int32_t *start, int32_t *n_pad,
                           int32_t *parts) {
    if (n_cams < 0 || (n_cams && (!widths || !start)) || !n_pad || !parts) return XRB_ERR_INVALID;
    std::vector<int> w(widths, widths + n_cams);
    std::vector<int32_t> st;
    int np = 0, pa = 1;
    plan_column_order(w, bw, allow_nd != 0, st, np, pa);
    for (int v = 0; v < n_cams; ++v) start[v] = st[v];
    *n_pad = np, *parts = pa;
    return XRB_OK;
}

int xrb_ba_profile_detail(const xrb_ba_solver *s, double *out, int n) {
    if (!s || !out || n < 8) return XRB_ERR_INVALID;
    out[0] = s->ms_kernel[0], out[1] = s->ms_kernel[1], out[2] = s->ms_kernel[2];
    out[3] = (double)s->n_steps, out[4] = (double)s->n_blocks, out[5] = (double)s->n_inc;
    out[6] = (double)s->nc_true, out[7] = (double)s->bw;
    if (n >= 16) {
        const CholPlanHost &h = s->plan.h;
        out[8] = s->parts, out[9] = h.nt, out[10] = h.n_tiles, out[11] = h.n_tiles_orig, out[12] = h.flops;
        out[13] = h.depth_f, out[14] = h.depth_b, out[15] = h.n_chain_f;
    }
    if (n >= 20) out[16] = s->use_window ? s->win.n_ctas : 0, out[17] = s->win.stride, out[18] = s->cam_span, out[19] = s->max_track;
    return XRB_OK;
}

int xrb_ba_profile(const xrb_ba_solver *s, double ms[6], int64_t launches[6]) {
    if (!s) return XRB_ERR_INVALID;
    for (int i = 0; i < 6; ++i) {
        if (ms) ms[i] = s->ms[i];
        if (launches) launches[i] = s->launches[i];
    }
    return XRB_OK;
}

}  // extern "C"
