"""World-size-2 gloo test (CPU) of the multi-GPU host logic: clip sharding covers every clip exactly once, the
max-over-ranks timing reduction and the result gather agree with a single-process run."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _clip_result(c):
    g = torch.Generator().manual_seed(100 + c)
    return torch.rand(4, 6, generator=g)


def _worker(rank, world, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dvmvs import sharding
    mine = sharding.clips_of_rank(n_clips, rank, world)
    local = {c: _clip_result(c) for c in mine}
    t = sharding.max_over_ranks(10.0 + rank)
    total = sharding.sum_over_ranks(len(mine))
    gathered = sharding.gather_clip_results(local, n_clips)
    ok = all(torch.equal(gathered[c], _clip_result(c)) for c in range(n_clips))
    q.put((rank, mine, t, total, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_clip_sharding_world_size_2():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "deep-video-mvs_b200"))
    world, n_clips = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    owned = sorted(c for _, mine, _, _, _ in results for c in mine)
    assert owned == list(range(n_clips))                       # every clip exactly once
    for rank, mine, t, total, ok in results:
        assert mine == list(range(rank, n_clips, world))
        assert t == 11.0 and total == n_clips and ok


def test_single_process_fallbacks():
    from dvmvs import sharding
    assert sharding.clips_of_rank(7, 0, 1) == list(range(7))
    assert sharding.max_over_ranks(3.5) == 3.5
    res = {c: _clip_result(c) for c in range(3)}
    assert all(torch.equal(a, b) for a, b in zip(sharding.gather_clip_results(res, 3), [res[0], res[1], res[2]]))
