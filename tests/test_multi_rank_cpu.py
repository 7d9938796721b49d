"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: the point sharding the BA
engine uses (xrb_ba_shard_range — pure host code in the product library) and the pair
sharding bench.py uses for matching.  The device-side exchange itself is covered by the
`gpu` tests run with 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xrsfm_b200 import _lib, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sc = synth.make_sequential_scene(40, 3000, 6, 5)
        k = np.bincount(sc.obs_pt, minlength=sc.n_pts).astype(np.int32)
        k[::7] += 9  # uneven work
        lo, hi = np.zeros(1, np.int32), np.zeros(1, np.int32)
        rc = _lib.lib().xrb_ba_shard_range(sc.n_pts, k.ctypes.data, rank, world, lo.ctypes.data, hi.ctypes.data)
        assert rc == 0
        mine = torch.tensor([int(lo[0]), int(hi[0])])
        allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allr, mine)
        # what the exchange hook does: SUM of zero-padded shards reproduces the whole vector
        full = torch.zeros(sc.n_pts, dtype=torch.float64)
        full[int(lo[0]): int(hi[0])] = torch.from_numpy(sc.pts[int(lo[0]): int(hi[0]), 0])
        dist.all_reduce(full)
        work = float((k[int(lo[0]): int(hi[0])].astype(np.float64) ** 2 + 4.0 * k[int(lo[0]): int(hi[0])]).sum())
        w = torch.tensor([work], dtype=torch.float64)
        ws = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(ws, w)
        # matching: pair list sharded round-robin, union is the whole list, no collective needed
        pairs = synth.sequential_pairs(30, window=5, n_retrieval=2, seed=1)
        mine_pairs = pairs[rank::world]
        cnt = torch.tensor([mine_pairs.shape[0]])
        dist.all_reduce(cnt)
        if rank == 0:
            q.put(dict(ranges=[t.tolist() for t in allr], gathered_ok=bool(np.array_equal(full.numpy(), sc.pts[:, 0])),
                       work=[float(x) for x in ws], n_pts=sc.n_pts, n_pairs=int(cnt.item()), n_pairs_all=int(pairs.shape[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_point_and_pair_sharding_across_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = res["ranges"]
    assert ranges[0][0] == 0 and ranges[-1][1] == res["n_pts"]
    for a, b in zip(ranges, ranges[1:]):
        assert a[1] == b[0]  # contiguous, disjoint, complete
    assert res["gathered_ok"]
    w = np.array(res["work"])
    assert w.max() / w.mean() < 1.05  # balanced by sum k^2 + 4k
    assert res["n_pairs"] == res["n_pairs_all"]


def test_shard_range_edge_cases():
    lib = _lib.lib()
    lo, hi = np.zeros(1, np.int32), np.zeros(1, np.int32)
    k = np.array([3, 3, 3, 3], dtype=np.int32)
    assert lib.xrb_ba_shard_range(4, k.ctypes.data, 0, 1, lo.ctypes.data, hi.ctypes.data) == 0
    assert (lo[0], hi[0]) == (0, 4)
    # more ranks than points: ranges stay valid (possibly empty) and cover everything
    cover = []
    for r in range(8):
        assert lib.xrb_ba_shard_range(4, k.ctypes.data, r, 8, lo.ctypes.data, hi.ctypes.data) == 0
        assert 0 <= lo[0] <= hi[0] <= 4
        cover.extend(range(lo[0], hi[0]))
    assert cover == [0, 1, 2, 3]
    assert lib.xrb_ba_shard_range(4, k.ctypes.data, 3, 2, lo.ctypes.data, hi.ctypes.data) != 0
    assert lib.xrb_ba_shard_range(0, None, 0, 2, lo.ctypes.data, hi.ctypes.data) == 0
