"""world_size-2 test of the walker-sharding logic on CPU (gloo).  The evaluator
is a stand-in NumPy function (the product's evaluator is the device plan); what
is tested is the slicing / padding / all-gather / replicated accept logic: the
chain must be bitwise identical to the unsharded run."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _evaluator(q):
    q = np.atleast_2d(q)
    lnp = -0.5 * np.sum((q - 0.3) ** 2 / np.array([1.0, 0.5, 2.0]), axis=1)
    flux = np.cos(q[:, :1]) * np.arange(1, 6)[None, :]
    return lnp, flux


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, W, nsteps, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from naima_b200.parallel import ShardedSampler

    p0 = np.random.default_rng(4).normal(size=(W, 3))
    s = ShardedSampler(W, 3, _evaluator, seed=11)
    assert (s.rank, s.world) == (rank, world)
    s.run_mcmc(p0, nsteps)
    assert s.collectives == 2 * nsteps + 1
    np.save(os.path.join(out, "chain%d.npy" % rank), s.get_chain())
    np.save(os.path.join(out, "lp%d.npy" % rank), s.get_log_prob())
    b = s.get_blobs()
    np.save(os.path.join(out, "blob%d.npy" % rank), np.array([b[-1, w][0] for w in range(W)]))
    dist.destroy_process_group()


@pytest.mark.parametrize("W", [12, 14])  # 14: half-ensemble of 7 does not divide by 2 -> padding
def test_sharded_sampler_matches_single_process(tmp_path, W):
    import torch.multiprocessing as mp

    from naima_b200.parallel import ShardedSampler, shard_bounds

    assert shard_bounds(7, 2) == (4, [(0, 4), (4, 7)])
    assert shard_bounds(8, 4) == (2, [(0, 2), (2, 4), (4, 6), (6, 8)])
    assert shard_bounds(3, 4) == (1, [(0, 1), (1, 2), (2, 3), (3, 3)])
    nsteps = 9
    mp.spawn(_worker, args=(2, _free_port(), W, nsteps, str(tmp_path)), nprocs=2, join=True)
    p0 = np.random.default_rng(4).normal(size=(W, 3))
    ref = ShardedSampler(W, 3, _evaluator, seed=11)  # world 1: no process group
    ref.run_mcmc(p0, nsteps)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / ("chain%d.npy" % r)), ref.get_chain())
        assert np.array_equal(np.load(tmp_path / ("lp%d.npy" % r)), ref.get_log_prob())
        b = ref.get_blobs()
        want = np.array([b[-1, w][0] for w in range(W)])
        assert np.array_equal(np.load(tmp_path / ("blob%d.npy" % r)), want)
