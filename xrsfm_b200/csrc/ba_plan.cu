// ba_plan.cu — column order, tile pattern, symbolic factorisation and task lists of the reduced
// camera system (see ba_plan.cuh).  Host code except k_tile_pattern; runs once per xrb_ba_load.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <numeric>

#include "ba_plan.cuh"

namespace xrb {

namespace {

inline int round_up64(int v) { return (v + 63) / 64 * 64; }

__global__ void k_tile_pattern(int n_pts, const int32_t *__restrict__ pt_ptr, const int32_t *__restrict__ pt_obs,
                               const int32_t *__restrict__ raw_cam, const uint8_t *__restrict__ pt_var,
                               const int32_t *__restrict__ colq, const int32_t *__restrict__ colt, int nt,
                               uint8_t *__restrict__ pat) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts || !pt_var[p]) return;
    const int k0 = pt_ptr[p], k1 = pt_ptr[p + 1];
    for (int a = k0; a < k1; ++a) {
        const int ca = raw_cam[pt_obs[a]];
        const int qa = colq[ca], ta = colt[ca];
        if (qa < 0 && ta < 0) continue;
        const int a_lo = (qa >= 0 ? (ta >= 0 ? min(qa, ta) : qa) : ta) >> 6;
        const int a_hi = (max(qa, ta) + 2) >> 6;
        for (int b = k0; b < a; ++b) {
            const int cb = raw_cam[pt_obs[b]];
            const int qb = colq[cb], tb = colt[cb];
            if (qb < 0 && tb < 0) continue;
            const int b_lo = (qb >= 0 ? (tb >= 0 ? min(qb, tb) : qb) : tb) >> 6;
            const int b_hi = (max(qb, tb) + 2) >> 6;
            for (int x = a_lo; x <= a_hi; ++x)
                for (int y = b_lo; y <= b_hi; ++y) {
                    const int i = max(x, y), j = min(x, y);
                    if (!pat[(size_t)i * nt + j]) pat[(size_t)i * nt + j] = 1;  // benign race: every writer stores 1
                }
        }
    }
}

}  // namespace

int launch_tile_pattern(int n_pts, const int32_t *pt_ptr, const int32_t *pt_obs, const int32_t *raw_cam,
                        const uint8_t *pt_var, int n_cams, const int32_t *colq, const int32_t *colt, int nt, uint8_t *pat,
                        cudaStream_t st) {
    (void)n_cams;
    if (n_pts > 0 && nt > 0) {
        k_tile_pattern<<<(n_pts + 127) / 128, 128, 0, st>>>(n_pts, pt_ptr, pt_obs, raw_cam, pt_var, colq, colt, nt, pat);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

void plan_column_order(const std::vector<int> &widths, int bw, bool allow_nd, std::vector<int32_t> &start, int &n_pad,
                       int &parts) {
    const int V = (int)widths.size();
    start.assign(V, 0);
    int W = 0;
    for (int v = 0; v < V; ++v) start[v] = W, W += widths[v];
    n_pad = std::max(64, round_up64(W));
    parts = 1;
    const int nt_nat = n_pad / 64;
    if (!allow_nd || bw <= 0 || nt_nat < 8) return;
    // one-level nested dissection of a band: P interiors, P - 1 separators of >= bw columns
    const int sepw = bw + 5;  // a separator is whole cameras: up to 5 columns beyond the minimum
    const int sep_tiles = (sepw + 63) / 64;
    int bestP = 0, best_depth = nt_nat;
    for (int P = 2; P <= 64; ++P) {
        const int interior = (W - (P - 1) * sepw) / P;
        if (interior < 2 * std::max(sepw, 64)) break;
        const int depth = (interior + 63) / 64 + 1 + (P - 1) * sep_tiles;
        if (depth < best_depth) best_depth = depth, bestP = P;
    }
    if (bestP < 2 || best_depth * 10 > nt_nat * 6) return;
    const int P = bestP;
    // walk the cameras: interior 0, separator 0, interior 1, ...
    std::vector<int> part(V, 0);  // 2p = interior p, 2p + 1 = separator p
    {
        const double target = (double)(W - (P - 1) * bw) / P;
        int v = 0;
        for (int p = 0; p < P; ++p) {
            int acc = 0;
            if (p == P - 1) {
                for (; v < V; ++v) part[v] = 2 * p;
                break;
            }
            for (; v < V && acc < target; ++v) part[v] = 2 * p, acc += widths[v];
            acc = 0;
            for (; v < V && acc < bw; ++v) part[v] = 2 * p + 1, acc += widths[v];
        }
    }
    // elimination order of the parts: interiors first (independent of each other), then the separators — which,
    // once the interiors are gone, form a path (s_p is coupled to s_p+1 through interior p+1) — in odd-even
    // (cyclic reduction) order: every other separator of the remaining path is independent of the others, so the
    // separator system costs log2(P) levels instead of a chain of P - 1
    std::vector<int> order;
    for (int p = 0; p < P; ++p) order.push_back(2 * p);
    {
        std::vector<int> path;
        for (int p = 0; p + 1 < P; ++p) path.push_back(p);
        while (!path.empty()) {
            std::vector<int> rest;
            for (size_t i = 0; i < path.size(); ++i) {
                if (i % 2 == 0)
                    order.push_back(2 * path[i] + 1);
                else
                    rest.push_back(path[i]);
            }
            path.swap(rest);
        }
    }
    int cur = 0;
    for (const int id : order) {
        bool any = false;
        for (int v = 0; v < V; ++v) {
            if (part[v] != id) continue;
            if (!any) cur = round_up64(cur), any = true;
            start[v] = cur, cur += widths[v];
        }
    }
    n_pad = std::max(64, round_up64(cur));
    parts = P;
}

int build_chol_plan(int nt, const uint8_t *pat, CholPlanHost &H) {
    H = CholPlanHost();
    H.nt = nt;
    if (nt <= 0) return XRB_OK;
    // ---- symbolic factorisation on tile columns
    std::vector<std::vector<int>> st(nt), children(nt);
    std::vector<uint8_t> is_orig((size_t)nt * nt, 0);
    for (int k = 0; k < nt; ++k) {
        std::vector<int> &S = st[k];
        for (int i = k + 1; i < nt; ++i)
            if (pat[(size_t)i * nt + k]) S.push_back(i), is_orig[(size_t)i * nt + k] = 1;
        is_orig[(size_t)k * nt + k] = 1;
        for (int c : children[k])
            for (int i : st[c])
                if (i > k) S.push_back(i);
        std::sort(S.begin(), S.end());
        S.erase(std::unique(S.begin(), S.end()), S.end());
        if (!S.empty()) children[S[0]].push_back(k);
    }
    // ---- slots: original tiles first (column-major), fill tiles after
    H.tab.assign((size_t)nt * nt, -1);
    int ns = 0;
    for (int k = 0; k < nt; ++k) {
        H.tab[(size_t)k * nt + k] = ns++;
        for (int i : st[k])
            if (is_orig[(size_t)i * nt + k]) H.tab[(size_t)i * nt + k] = ns++;
    }
    H.n_tiles_orig = ns;
    for (int k = 0; k < nt; ++k)
        for (int i : st[k])
            if (!is_orig[(size_t)i * nt + k]) H.tab[(size_t)i * nt + k] = ns++;
    H.n_tiles = ns;
    auto slot = [&](int i, int j) { return H.tab[(size_t)i * nt + j]; };
    H.colptr.assign(nt + 1, 0);
    for (int k = 0; k < nt; ++k) {
        H.colptr[k + 1] = H.colptr[k] + (int)st[k].size();
        for (int i : st[k]) H.rowidx.push_back(i), H.slot.push_back(slot(i, k));
    }
    // ---- kp[k]: the last column that updates the diagonal tile k
    std::vector<int> kp(nt, -1);
    for (int k = 0; k < nt; ++k)
        for (int i : st[k]) kp[i] = k;
    // ---- forward tasks.  Sweep 1 walks the columns in order and sequences the updates of a tile by k (the
    // natural order) to learn WHEN each update's operands are ready; the updates of a tile are then re-sequenced
    // by that readiness (any fixed order is legal: an update never feeds the operands of another update of the
    // same tile) — otherwise a separator tile that two interiors update would take the second interior's
    // contributions only after all of the first's, although both run at the same time.  Sweep 2 computes the
    // final longest-path levels for that sequencing.
    struct Rec {
        int32_t v[8];
    };
    struct UTask {
        int i, j, k, s_ij, ready, seq, lev;
    };
    std::vector<UTask> U;
    std::vector<std::vector<int>> upd_of(ns);  // tile -> its updates (ids into U), later in final order
    double flops = (double)nt * (64.0 * 64 * 64 / 3.0);
    const double t3 = 64.0 * 64 * 64;
    {
        std::vector<int> levF(nt, 0), lastlev(ns, 0), pl(nt, 0);
        for (int k = 0; k < nt; ++k) {
            if (kp[k] < 0) levF[k] = 1;
            for (int i : st[k]) {
                const int s_ik = slot(i, k);
                if (kp[i] == k) {
                    levF[i] = 1 + std::max(levF[k], std::max(lastlev[s_ik], lastlev[slot(i, i)]));
                    pl[i] = levF[i];
                    flops += 2.0 * t3;  // substitution + diagonal update
                } else {
                    pl[i] = 1 + std::max(levF[k], lastlev[s_ik]);
                    flops += t3;
                }
            }
            const std::vector<int> &S = st[k];
            for (size_t a2 = 0; a2 < S.size(); ++a2)
                for (size_t b2 = 0; b2 <= a2; ++b2) {
                    const int i = S[a2], j = S[b2];
                    if (i == j && kp[i] == k) continue;
                    const int s_ij = slot(i, j);
                    if (s_ij < 0) return XRB_ERR_INVALID;  // cannot happen: fill is closed under pair updates
                    const int ready = 1 + std::max(pl[i], pl[j]);
                    const int lev = std::max(ready, 1 + lastlev[s_ij]);
                    upd_of[s_ij].push_back((int)U.size());
                    U.push_back({i, j, k, s_ij, ready, 0, 0});
                    lastlev[s_ij] = lev;
                    flops += 2.0 * t3;
                }
        }
    }
    for (int sl = 0; sl < ns; ++sl) {
        std::vector<int> &v = upd_of[sl];
        std::stable_sort(v.begin(), v.end(), [&](int x, int y) { return U[x].ready != U[y].ready ? U[x].ready < U[y].ready : U[x].k < U[y].k; });
        for (size_t q = 0; q < v.size(); ++q) U[v[q]].seq = (int)q;
    }
    // Fused tasks PU(i, k) = P(i, k) followed, in the same CTA, by U(i, pk, k) with pk = parent(k) = the next column
    // that meets row i: the row's fill then advances one task per column, like the chain, instead of two
    // (substitution, then update) — in a dissected band every separator row is such a row.
    std::vector<int> fused_u(ns, -1);            // tile (i, k) -> id of its fused update, or -1
    std::vector<signed char> is_fused(U.size(), 0);
    for (size_t u = 0; u < U.size(); ++u) {
        const UTask &t = U[u];
        const int pk = st[t.k][0];
        if (t.j == pk && t.i > pk && kp[t.i] != t.k) fused_u[slot(t.i, t.k)] = (int)u, is_fused[u] = 1;
    }
    // sweep 2: memoised longest path.  Dependencies always have a smaller level, so the recursion depth is
    // bounded by the depth of the plan.
    std::vector<int> levF(nt, 0), levP(ns, 0);
    std::vector<signed char> stF(nt, 0), stP(ns, 0), stU(U.size(), 0);  // 0 new, 1 open, 2 done
    bool cyclic = false;
    struct Eval {
        std::function<int(int)> F, P, Uf;
    } ev;
    auto last_upd_level = [&](int sl) { return upd_of[sl].empty() ? 0 : ev.Uf(upd_of[sl].back()); };
    ev.F = [&](int k) -> int {
        if (stF[k] == 2) return levF[k];
        if (stF[k] == 1) { cyclic = true; return 0; }
        stF[k] = 1;
        int lev = last_upd_level(slot(k, k));
        if (kp[k] >= 0) lev = std::max(lev, std::max(ev.F(kp[k]), last_upd_level(slot(k, kp[k]))));
        levF[k] = lev + 1, stF[k] = 2;
        return levF[k];
    };
    std::function<int(int, int)> producer;
    std::vector<int> tile_i(ns, 0), tile_k(ns, 0);
    for (int k = 0; k < nt; ++k) {
        tile_i[slot(k, k)] = k, tile_k[slot(k, k)] = k;
        for (int i : st[k]) tile_i[slot(i, k)] = i, tile_k[slot(i, k)] = k;
    }
    ev.P = [&](int sl) -> int {  // substitution of tile sl = (i, k), not merged
        if (stP[sl] == 2) return levP[sl];
        if (stP[sl] == 1) { cyclic = true; return 0; }
        stP[sl] = 1;
        int lev = std::max(ev.F(tile_k[sl]), last_upd_level(sl));
        if (fused_u[sl] >= 0) {  // the fused update's own dependencies: the other panel tile, its predecessor in the tile
            const UTask &t = U[fused_u[sl]];
            lev = std::max(lev, producer(t.j, t.k));
            if (t.seq > 0) lev = std::max(lev, ev.Uf(upd_of[t.s_ij][t.seq - 1]));
        }
        levP[sl] = lev + 1, stP[sl] = 2;
        return levP[sl];
    };
    producer = [&](int i, int k) { return kp[i] == k ? ev.F(i) : ev.P(slot(i, k)); };
    ev.Uf = [&](int u) -> int {
        if (is_fused[u]) return U[u].lev = ev.P(slot(U[u].i, U[u].k));
        if (stU[u] == 2) return U[u].lev;
        if (stU[u] == 1) { cyclic = true; return 0; }
        stU[u] = 1;
        int lev = std::max(producer(U[u].i, U[u].k), producer(U[u].j, U[u].k));
        if (U[u].seq > 0) lev = std::max(lev, ev.Uf(upd_of[U[u].s_ij][U[u].seq - 1]));
        U[u].lev = lev + 1, stU[u] = 2;
        return U[u].lev;
    };
    std::vector<Rec> F, Wt;
    for (int k = 0; k < nt; ++k) {
        const int lf = ev.F(k);
        if (kp[k] < 0)
            F.push_back({{k, -1, slot(k, k), -1, -1, (int)upd_of[slot(k, k)].size(), 0, lf}});
        else
            F.push_back({{k, kp[k], slot(k, k), slot(k, kp[k]), slot(kp[k], kp[k]), (int)upd_of[slot(k, k)].size(),
                          (int)upd_of[slot(k, kp[k])].size(), lf}});
        for (int i : st[k])
            if (kp[i] != k) {
                const int sl = slot(i, k);
                if (fused_u[sl] >= 0) {
                    const UTask &t = U[fused_u[sl]];
                    if (upd_of[sl].size() > 0xFFFF || t.seq > 0x7FFF) return XRB_ERR_INVALID;
                    Wt.push_back({{TASK_PU, sl, slot(k, k), t.s_ij, (int)upd_of[sl].size() | (t.seq << 16), k, slot(t.j, k), ev.P(sl)}});
                } else {
                    Wt.push_back({{TASK_P, sl, slot(k, k), -1, (int)upd_of[sl].size(), k, i, ev.P(sl)}});
                }
            }
    }
    for (size_t u = 0; u < U.size(); ++u) {
        const UTask &t = U[u];
        const int lev = ev.Uf((int)u);
        if (!is_fused[u]) Wt.push_back({{TASK_U, slot(t.i, t.k), slot(t.j, t.k), t.s_ij, t.seq, t.k, t.i, lev}});
    }
    if (cyclic) return XRB_ERR_INVALID;
    std::stable_sort(F.begin(), F.end(), [](const Rec &a, const Rec &b) { return a.v[7] != b.v[7] ? a.v[7] < b.v[7] : a.v[0] < b.v[0]; });
    std::stable_sort(Wt.begin(), Wt.end(), [](const Rec &a, const Rec &b) {
        if (a.v[7] != b.v[7]) return a.v[7] < b.v[7];
        if (a.v[5] != b.v[5]) return a.v[5] < b.v[5];
        return a.v[0] < b.v[0];
    });
    H.n_f = (int)F.size(), H.n_w = (int)Wt.size();
    H.ftasks.resize((size_t)H.n_f * 8), H.wtasks.resize((size_t)H.n_w * 8);
    for (int t = 0; t < H.n_f; ++t) memcpy(&H.ftasks[(size_t)t * 8], F[t].v, 32);
    for (int t = 0; t < H.n_w; ++t) memcpy(&H.wtasks[(size_t)t * 8], Wt[t].v, 32);
    {
        int width = 0, run = 0, prev = -1;
        for (const Rec &r : F) {
            run = r.v[7] == prev ? run + 1 : 1, prev = r.v[7];
            width = std::max(width, run);
            H.depth_f = std::max(H.depth_f, r.v[7]);
        }
        H.n_chain_f = std::max(1, std::min(width, kMaxChainCtas));
    }
    // ---- backward substitution: x_k needs x_i for every i in struct(k)
    std::vector<int> blev(nt, 1);
    for (int k = nt - 1; k >= 0; --k)
        for (int i : st[k]) blev[k] = std::max(blev[k], blev[i] + 1);
    std::vector<int> order(nt);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return blev[a] != blev[b] ? blev[a] < blev[b] : a > b; });
    {
        int width = 0, run = 0, prev = -1;
        for (int k : order) {
            run = blev[k] == prev ? run + 1 : 1, prev = blev[k];
            width = std::max(width, run);
            H.depth_b = std::max(H.depth_b, blev[k]);
        }
        H.n_chain_b = std::max(1, std::min(width, kMaxChainCtas));
    }
    for (int k : order) {
        const std::vector<int> &S = st[k];
        const int n_near = std::min<int>(kNearTiles, (int)S.size());
        const bool has_far = (int)S.size() > n_near;
        int32_t rec[12] = {k, slot(k, k), n_near, has_far ? 1 : 0, -1, -1, -1, blev[k], -1, -1, -1, 0};
        for (int a = 0; a < n_near; ++a) rec[4 + a] = S[a], rec[8 + a] = slot(S[a], k);
        H.btasks.insert(H.btasks.end(), rec, rec + 12);
        H.n_b++;
        if (has_far) {
            std::vector<int> far(S.begin() + n_near, S.end());
            std::stable_sort(far.begin(), far.end(), [&](int a, int b) { return blev[a] != blev[b] ? blev[a] < blev[b] : a > b; });
            const int begin = (int)H.far_rows.size();
            for (int i : far) H.far_rows.push_back(i), H.far_slots.push_back(slot(i, k));
            const int32_t wrec[4] = {k, begin, (int)H.far_rows.size(), blev[k]};
            H.wbtasks.insert(H.wbtasks.end(), wrec, wrec + 4);
            H.n_wb++;
        }
        flops += 2.0 * 2.0 * 64 * 64 * (1.0 + S.size());
    }
    H.flops = flops;
    return XRB_OK;
}

int CholPlan::upload(cudaStream_t st) {
    auto pad4 = [](size_t n) { return (n + 3) / 4 * 4; };
    const size_t n_tab = pad4(h.tab.size()), n_f = pad4(h.ftasks.size()), n_w = pad4(h.wtasks.size()),
                 n_b = pad4(h.btasks.size()), n_wb = pad4(h.wbtasks.size()), n_fr = pad4(h.far_rows.size()),
                 n_fs = pad4(h.far_slots.size());
    const size_t total = n_tab + n_f + n_w + n_b + n_wb + n_fr + n_fs + 4;
    int rc = buf.reserve(total * 4);
    if (rc) return rc;
    std::vector<int32_t> flat(total, 0);
    size_t o = 0;
    auto put = [&](const std::vector<int32_t> &v, size_t padded) {
        const size_t at = o;
        if (!v.empty()) memcpy(&flat[o], v.data(), v.size() * 4);
        o += padded;
        return at;
    };
    const size_t o_tab = put(h.tab, n_tab), o_f = put(h.ftasks, n_f), o_w = put(h.wtasks, n_w), o_b = put(h.btasks, n_b),
                 o_wb = put(h.wbtasks, n_wb), o_fr = put(h.far_rows, n_fr), o_fs = put(h.far_slots, n_fs);
    XRB_CUDA(cudaMemcpyAsync(buf.p, flat.data(), total * 4, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaStreamSynchronize(st));  // `flat` is pageable
    const int32_t *base = buf.as<int32_t>();
    d.nt = h.nt, d.n_tiles = h.n_tiles;
    d.tab = base + o_tab;
    d.ftasks = reinterpret_cast<const int4 *>(base + o_f), d.wtasks = reinterpret_cast<const int4 *>(base + o_w);
    d.btasks = reinterpret_cast<const int4 *>(base + o_b), d.wbtasks = reinterpret_cast<const int4 *>(base + o_wb);
    d.far_rows = base + o_fr, d.far_slots = base + o_fs;
    d.n_f = h.n_f, d.n_w = h.n_w, d.n_b = h.n_b, d.n_wb = h.n_wb;
    d.n_chain_f = h.n_chain_f, d.n_chain_b = h.n_chain_b;
    return XRB_OK;
}

}  // namespace xrb
