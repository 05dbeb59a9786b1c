"""`-m "not gpu"`: the N>1 path on CPU with the gloo backend, world_size 2 -- pair sharding and the path's single
collective (final pose/match-count all-gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from far_b200.pipeline import shard_pairs, gather_poses


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_pairs(n_pairs, world, rank)
    # each rank "computes" poses for its shard: pose[b] is a function of the global pair id only
    ids = torch.arange(lo, hi, dtype=torch.float32)
    pose = ids[:, None, None] + torch.arange(12, dtype=torch.float32).reshape(1, 3, 4) / 100
    counts = (ids * 7 + 3).to(torch.int64)
    poses, cnts = gather_poses(pose, counts)
    if rank == 0:
        q.put((poses, cnts))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_pairs_covers_everything():
    for n in (1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_pairs(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_gather_poses_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_pairs = 7  # ragged: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    poses, cnts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = torch.arange(n_pairs, dtype=torch.float32)
    assert poses.shape == (n_pairs, 3, 4)
    assert torch.allclose(poses, ids[:, None, None] + torch.arange(12, dtype=torch.float32).reshape(1, 3, 4) / 100)
    assert torch.equal(cnts, (ids * 7 + 3).to(torch.int64))
