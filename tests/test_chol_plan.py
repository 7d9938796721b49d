"""Host side of the reduced-camera-system solver (xrsfm_b200/csrc/ba_plan.cu), no GPU needed.

The plan (tile slots with fill, F / P / U task lists with the flag values they wait for, backward
B / W lists) is executed here by a small discrete simulator with numpy tiles: CTAs pull tasks from their
queues IN ORDER and may only run a task whose flags are satisfied — exactly the device protocol of
ba_tilechol.cu.  The simulator must never deadlock and must reproduce numpy's Cholesky solve.
"""
import ctypes as C

import numpy as np
import pytest

from xrsfm_b200 import _lib

B = 4  # simulator tile edge (the plan is tile-size agnostic)


def get_plan(pat):
    nt = pat.shape[0]
    lib = _lib.lib()
    counts = np.zeros(16, dtype=np.int32)
    pat = np.ascontiguousarray(pat, dtype=np.uint8)
    _lib.check(lib.xrb_debug_chol_plan(nt, pat.ctypes.data, counts.ctypes.data, None, None, 0, None, 0, None, 0, None, 0,
                                       None, None, 0), "plan")
    nt_, n_tiles, n_orig, n_f, n_w, n_b, n_wb, n_far, ncf, ncb, depth_f, depth_b = (int(v) for v in counts[:12])
    tab = np.zeros(nt * nt, dtype=np.int32)
    ft = np.zeros(max(1, n_f) * 8, dtype=np.int32)
    wt = np.zeros(max(1, n_w) * 8, dtype=np.int32)
    bt = np.zeros(max(1, n_b) * 12, dtype=np.int32)
    wbt = np.zeros(max(1, n_wb) * 4, dtype=np.int32)
    fr = np.zeros(max(1, n_far), dtype=np.int32)
    fs = np.zeros(max(1, n_far), dtype=np.int32)
    _lib.check(lib.xrb_debug_chol_plan(nt, pat.ctypes.data, counts.ctypes.data, tab.ctypes.data, ft.ctypes.data, ft.size,
                                       wt.ctypes.data, wt.size, bt.ctypes.data, bt.size, wbt.ctypes.data, wbt.size,
                                       fr.ctypes.data, fs.ctypes.data, fr.size), "plan")
    return dict(nt=nt, n_tiles=n_tiles, n_orig=n_orig, tab=tab.reshape(nt, nt), F=ft.reshape(-1, 8)[:n_f],
                W=wt.reshape(-1, 8)[:n_w], Bt=bt.reshape(-1, 12)[:n_b], WB=wbt.reshape(-1, 4)[:n_wb], far_rows=fr,
                far_slots=fs, n_chain_f=ncf, n_chain_b=ncb, depth_f=depth_f, depth_b=depth_b)


def spd_from_pattern(pat, rng):
    nt = pat.shape[0]
    n = nt * B
    A = np.zeros((n, n))
    for i in range(nt):
        for j in range(i + 1):
            if pat[i, j] or i == j:
                blk = rng.standard_normal((B, B))
                A[i * B:(i + 1) * B, j * B:(j + 1) * B] = blk
    A = np.tril(A)
    A = A + A.T
    A[np.arange(n), np.arange(n)] = np.abs(A).sum(axis=1) + 1.0
    return A


def run_queues(queues, n_ctas, ready, execute):
    """queues: list of task lists; n_ctas[q] CTAs serve queue q in order.  Returns executed count."""
    heads = [0] * len(queues)
    held = [[None] * n for n in n_ctas]
    done = 0
    total = sum(len(q) for q in queues)
    while done < total:
        progress = False
        for q, tasks in enumerate(queues):
            for c in range(n_ctas[q]):
                if held[q][c] is None and heads[q] < len(tasks):
                    held[q][c] = tasks[heads[q]]
                    heads[q] += 1
                    progress = True
        for q in range(len(queues)):
            for c in range(n_ctas[q]):
                t = held[q][c]
                if t is not None and ready(q, t):
                    execute(q, t)
                    held[q][c] = None
                    done += 1
                    progress = True
        assert progress, "deadlock: every CTA holds a task whose flags can never be satisfied"
    return done


def simulate(pat, rng, n_workers=5):
    P = get_plan(pat)
    nt, tab = P["nt"], P["tab"]
    A = spd_from_pattern(pat, rng)
    n = nt * B
    rhs0 = rng.standard_normal(n)
    tiles = np.zeros((P["n_tiles"], B, B))
    for i in range(nt):
        for j in range(i + 1):
            blk = A[i * B:(i + 1) * B, j * B:(j + 1) * B]
            if tab[i, j] >= 0:
                tiles[tab[i, j]] = blk
                if np.any(blk != 0) and not (pat[i, j] or i == j):
                    raise AssertionError("non-zero outside the pattern")
            else:
                assert not np.any(blk != 0)
    # original tiles come first
    for i in range(nt):
        for j in range(i + 1):
            if pat[i, j] or i == j:
                assert 0 <= tab[i, j] < P["n_orig"]
    rhs = rhs0.copy().reshape(nt, B)
    diag_done = np.zeros(nt, dtype=int)
    pdone = np.zeros(P["n_tiles"], dtype=int)
    upd = np.zeros(P["n_tiles"], dtype=int)

    def ready(q, t):
        if q == 0:
            k, kp, s_kk, s_kkp, s_kpkp, need_kk, need_kkp, _ = t
            if upd[s_kk] < need_kk:
                return False
            return kp < 0 or (upd[s_kkp] >= need_kkp and diag_done[kp])
        typ, s_ik, s_jk, s_ij, seq, k, i, _ = t
        if typ == 0:
            return upd[s_ik] >= seq and diag_done[k]
        if typ == 2:  # fused P(i,k) + U(i,pk,k): the simulator runs it in one piece, so it needs everything
            need, useq, s_pk_k = seq & 0xFFFF, seq >> 16, i
            return upd[s_ik] >= need and diag_done[k] and pdone[s_pk_k] and upd[s_ij] >= useq
        return pdone[s_ik] and pdone[s_jk] and upd[s_ij] >= seq

    def execute(q, t):
        if q == 0:
            k, kp, s_kk, s_kkp, s_kpkp, need_kk, need_kkp, _ = t
            if kp >= 0:
                assert upd[s_kkp] == need_kkp and upd[s_kk] == need_kk
                X = np.linalg.solve(tiles[s_kpkp], tiles[s_kkp].T).T
                tiles[s_kkp] = X
                rhs[k] -= X @ rhs[kp]
                tiles[s_kk] -= X @ X.T
                pdone[s_kkp] = 1
            tiles[s_kk] = np.linalg.cholesky(tiles[s_kk])
            rhs[k] = np.linalg.solve(tiles[s_kk], rhs[k])
            assert not diag_done[k]
            diag_done[k] = 1
            return
        typ, s_ik, s_jk, s_ij, seq, k, i, _ = t
        if typ == 0:
            assert upd[s_ik] == seq and not pdone[s_ik]
            tiles[s_ik] = np.linalg.solve(tiles[s_jk], tiles[s_ik].T).T
            pdone[s_ik] = 1
        elif typ == 2:
            need, useq, s_pk_k = seq & 0xFFFF, seq >> 16, i
            assert upd[s_ik] == need and not pdone[s_ik] and upd[s_ij] == useq
            tiles[s_ik] = np.linalg.solve(tiles[s_jk], tiles[s_ik].T).T   # s_jk holds the diagonal tile's slot
            pdone[s_ik] = 1
            tiles[s_ij] -= tiles[s_ik] @ tiles[s_pk_k].T
            upd[s_ij] = useq + 1
        else:
            assert upd[s_ij] == seq, "updates must arrive in sequence"
            tiles[s_ij] -= tiles[s_ik] @ tiles[s_jk].T
            if s_ik == s_jk:
                rhs[i] -= tiles[s_ik] @ rhs[k]
            upd[s_ij] = seq + 1

    run_queues([list(map(tuple, P["F"])), list(map(tuple, P["W"]))], [P["n_chain_f"], n_workers], ready, execute)
    assert diag_done.all()
    Lref = np.linalg.cholesky(A)
    for i in range(nt):
        for j in range(i + 1):
            ref = Lref[i * B:(i + 1) * B, j * B:(j + 1) * B]
            if tab[i, j] >= 0:
                got = tiles[tab[i, j]] if i != j else np.tril(tiles[tab[i, j]])
                np.testing.assert_allclose(got, ref, atol=1e-10)
            else:
                assert np.abs(ref).max() < 1e-12, "fill outside the symbolic structure"
    yref = np.linalg.solve(Lref, rhs0)
    np.testing.assert_allclose(rhs.reshape(-1), yref, atol=1e-9)

    # ---- backward substitution
    x = np.zeros((nt, B))
    xdone = np.zeros(nt, dtype=int)
    wdone = np.zeros(nt, dtype=int)
    wpart = np.zeros((nt, B))

    def ready_b(q, t):
        if q == 0:
            k, s_kk, n_near, has_far = t[:4]
            return all(xdone[t[4 + a]] for a in range(n_near)) and (not has_far or wdone[k])
        k, b, e, _ = t
        return all(xdone[P["far_rows"][u]] for u in range(b, e))

    def execute_b(q, t):
        if q == 0:
            k, s_kk, n_near, has_far = t[:4]
            s = rhs[k].copy()
            for a in range(n_near):
                s -= tiles[t[8 + a]].T @ x[t[4 + a]]
            if has_far:
                s -= wpart[k]
            x[k] = np.linalg.solve(np.tril(tiles[s_kk]).T, s)
            xdone[k] = 1
            return
        k, b, e, _ = t
        acc = np.zeros(B)
        for u in range(b, e):
            acc += tiles[P["far_slots"][u]].T @ x[P["far_rows"][u]]
        wpart[k] = acc
        wdone[k] = 1

    run_queues([list(map(tuple, P["Bt"])), list(map(tuple, P["WB"]))], [P["n_chain_b"], n_workers], ready_b, execute_b)
    np.testing.assert_allclose(x.reshape(-1), np.linalg.solve(A, rhs0), atol=1e-8)
    return P


def band_pattern(nt, h):
    pat = np.zeros((nt, nt), dtype=np.uint8)
    for i in range(nt):
        for j in range(max(0, i - h), i + 1):
            pat[i, j] = 1
    return pat


def test_dense_pattern_is_one_chain():
    P = simulate(band_pattern(9, 9), np.random.default_rng(0))
    assert P["n_chain_f"] == 1 and P["n_tiles"] == 45 and P["n_orig"] == 45
    # the chain walks the diagonal: every F task after the first is coupled to its predecessor
    assert [tuple(f[:2]) for f in P["F"]] == [(0, -1)] + [(k, k - 1) for k in range(1, 9)]


def test_band_has_no_fill():
    P = simulate(band_pattern(12, 2), np.random.default_rng(1))
    assert P["n_tiles"] == P["n_orig"] == 12 + 11 + 10


def test_dissected_band_runs_interiors_in_parallel():
    # 3 interiors of 4 tiles + 2 separators, block tridiagonal inside, separators coupled to both neighbours
    nt, pat = 14, np.zeros((14, 14), dtype=np.uint8)
    for p in range(3):
        for a in range(4):
            k = 4 * p + a
            pat[k, k] = 1
            if a:
                pat[k, k - 1] = 1
    for s in range(2):
        sep = 12 + s
        pat[sep, sep] = 1
        pat[sep, 4 * s + 3] = 1      # end of the interior on its left
        pat[sep, 4 * (s + 1)] = 1    # start of the interior on its right
    P = simulate(pat, np.random.default_rng(2))
    assert P["n_chain_f"] == 3
    assert P["depth_f"] < 14 * 2  # far shorter than a walk down the whole diagonal would be


@pytest.mark.parametrize("seed", range(6))
def test_random_sparse_patterns(seed):
    rng = np.random.default_rng(100 + seed)
    nt = int(rng.integers(2, 14))
    pat = np.tril((rng.random((nt, nt)) < rng.choice([0.1, 0.3, 0.7])).astype(np.uint8))
    simulate(pat, rng, n_workers=int(rng.integers(1, 7)))


def test_single_tile():
    P = simulate(np.ones((1, 1), dtype=np.uint8), np.random.default_rng(3))
    assert len(P["W"]) == 0 and len(P["F"]) == 1


def column_order(widths, bw, allow_nd=1):
    lib = _lib.lib()
    w = np.asarray(widths, dtype=np.int32)
    start = np.zeros(len(w), dtype=np.int32)
    n_pad, parts = C.c_int32(0), C.c_int32(0)
    _lib.check(lib.xrb_debug_column_order(len(w), w.ctypes.data, bw, allow_nd, start.ctypes.data, C.byref(n_pad),
                                          C.byref(parts)), "order")
    return start, n_pad.value, parts.value


def test_column_order_natural_for_dense_and_small():
    start, n_pad, parts = column_order([6] * 499, 2993)
    assert parts == 1 and n_pad == 3008 and (start == 6 * np.arange(499)).all()
    start, n_pad, parts = column_order([6] * 20, 30)
    assert parts == 1


def test_column_order_dissects_a_narrow_band():
    V, bw = 2700, 59
    start, n_pad, parts = column_order([6] * V, bw)
    assert parts >= 8 and n_pad % 64 == 0
    # columns are disjoint
    used = np.zeros(n_pad, dtype=int)
    for s in start:
        used[s:s + 6] += 1
    assert used.max() == 1
    # cameras within one bandwidth of each other in natural order (they may share a point) must not sit in
    # two different interiors: build the tile pattern of the band and check the interiors are decoupled
    nt = n_pad // 64
    tile_of = start // 64
    tile_hi = (start + 5) // 64
    pat = np.zeros((nt, nt), dtype=np.uint8)
    span_cams = bw // 6 + 1
    for a in range(V):
        for b in range(max(0, a - span_cams + 1), a + 1):
            if 6 * (a - b) + 5 > bw:
                continue
            for x in (tile_of[a], tile_hi[a]):
                for y in (tile_of[b], tile_hi[b]):
                    pat[max(x, y), min(x, y)] = 1
    P = get_plan(pat)
    nt_nat = (6 * V + 63) // 64
    assert P["depth_f"] * 3 < nt_nat, (P["depth_f"], nt_nat)
    assert P["n_chain_f"] >= 8
    assert column_order([6] * V, bw, allow_nd=0)[2] == 1
