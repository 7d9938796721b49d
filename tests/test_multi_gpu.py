"""2-GPU test of the BA exchange path (points sharded, SUM all-reduce of the reduced camera
system over NCCL — the library's own communicator, and the caller-hook variant) and of
pair-sharded matching.  Needs >= 2 GPUs: run with
`gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _CudaPtr:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def _worker(rank, world, port, q, mode="native"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from tests import oracle_lib as ol
        from xrsfm_b200 import ba, matching, synth
        sc = synth.make_sphere_scene(30, 4000, 8, 77)
        work = sc.copy_state()
        solver = ba.BASolver(device=rank)

        if mode == "native":
            def bcast(raw):
                t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{rank}")
                if rank == 0:
                    t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
                dist.broadcast(t, 0)
                return bytes(t.cpu().numpy().tobytes())

            solver.comm_init(rank, world, bcast)
        else:
            def allreduce(ptr, count, stream):
                # the contract: the reduction is ordered on the solver's stream, no host sync
                t = torch.as_tensor(_CudaPtr(ptr, count), device=f"cuda:{rank}")
                with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                    dist.all_reduce(t)

            solver.set_exchange(rank, world, allreduce)
        s = solver.solve_scene(work, **ol.GBA_ACCURATE)
        # matching: each rank matches its share of the pair list
        imgs, _ = synth.make_images(6, 600, seed=9)
        pairs = synth.sequential_pairs(6, window=3, n_retrieval=1, seed=0)
        m = matching.SiftMatchGPU(600)
        m.SetLanguage(matching.SiftMatchGPU.SIFTMATCH_CUDA_DEVICE0 + rank)
        assert m.VerifyContextGL() == 1
        m.upload_images(imgs)
        mine = pairs[rank::world]
        off, mm = m.match_pairs(mine)
        ok = all(np.array_equal(mm[off[k]: off[k + 1]], ol.match_pair(imgs[a], imgs[b])) for k, (a, b) in enumerate(mine))
        q.put((rank, work.cam_q.copy(), work.cam_t.copy(), work.pts.copy(), s.final_cost, s.num_lm_iterations,
               s.termination_type, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["native", "hook"])
def test_two_gpu_ba_equals_single_gpu_and_oracle(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from tests import oracle_lib as ol
    from xrsfm_b200 import ba, synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    sc = synth.make_sphere_scene(30, 4000, 8, 77)
    single = sc.copy_state()
    s1 = ba.BASolver(device=0).solve_scene(single, **ol.GBA_ACCURATE)
    ref = sc.copy_state()
    s_ref = ol.ba_solve(ref, ol.ba_options(**ol.GBA_ACCURATE))
    for rank, q_, t_, X_, cost, iters, term, match_ok in res:
        assert match_ok
        assert term == s1.termination_type == s_ref.termination_type
        assert iters == s1.num_lm_iterations == s_ref.num_lm_iterations
        assert cost == pytest.approx(s1.final_cost, rel=1e-9)
        # every rank returns every point (shards gathered: grouped ncclBroadcast / hook)
        assert np.abs(X_ - single.pts).max() <= 1e-8 * max(1.0, np.abs(single.pts).max())
        assert np.abs(q_ - single.cam_q).max() <= 1e-9 and np.abs(t_ - single.cam_t).max() <= 1e-8
        assert np.abs(X_ - ref.pts).max() <= 1e-5 * max(1.0, np.abs(ref.pts).max())
    # both ranks hold bit-identical results (replicated Cholesky on identical all-reduced data)
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][3], res[1][3])
