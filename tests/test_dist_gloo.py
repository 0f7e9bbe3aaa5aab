"""CPU, world_size 2 over gloo: the multi-GPU sharding algebra.  Each rank takes
its shard (robot_camera_calibration_b200.dist.shard_scene), forms the partial
reduced system the way the library does (oracle arithmetic), the partials are
all-reduced, and the sum must equal the reduced system of the full problem."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle_reduced, rel_fro, to_oracle
from robot_camera_calibration_b200.dist import owner_ranges, shard_scene
from robot_camera_calibration_b200.scenes import make_scene


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, elim, model, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    kw = dict(n_cam=2, model="rig") if model == "rig" else {}
    scene = make_scene(9, 14, 0.8, seed=31, **kw)
    local, (lo, hi) = shard_scene(scene, rank, world, eliminate=elim)
    S, b, *_ = oracle_reduced(to_oracle(local), elim == "views", 1e4)
    t = torch.from_numpy(np.concatenate([S.ravel(), b]))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    n_obs = torch.tensor([local.n_blocks])
    dist.all_reduce(n_obs)
    if rank == 0:
        Sf, bf, *_ = oracle_reduced(to_oracle(scene), elim == "views", 1e4)
        n = len(bf)
        q.put((rel_fro(t[:n * n].numpy().reshape(n, n), Sf), rel_fro(t[n * n:].numpy(), bf),
               int(n_obs.item()), scene.n_blocks))
    dist.destroy_process_group()


@pytest.mark.parametrize("elim,model", [("views", "single"), ("markers", "single"), ("views", "rig")])
def test_partial_reduced_systems_sum_to_the_full_one(elim, model):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, elim, model, q)) for r in range(2)]
    for p in procs:
        p.start()
    eS, eb, n_sum, n_all = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert n_sum == n_all                 # every observation block lives on exactly one rank
    assert eS < 1e-10 and eb < 1e-10


def test_owner_ranges_cover_and_balance():
    rng = np.random.default_rng(0)
    counts = rng.integers(0, 50, 1000)
    for world in (1, 2, 3, 8):
        r = owner_ranges(counts, world)
        assert r[0][0] == 0 and r[-1][1] == len(counts)
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        loads = [counts[lo:hi].sum() for lo, hi in r]
        assert max(loads) - min(loads) <= 2 * counts.max()
    assert owner_ranges(np.zeros(5, int), 2)[-1][1] == 5
    assert owner_ranges(np.array([7]), 4) [-1][1] == 1
