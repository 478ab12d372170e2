"""Data-parallel logic on CPU: world_size 2 over gloo.  Each rank runs the ORACLE on its shard of the global
minibatch; the sum-allreduced, 1/W-scaled gradients must equal the single-process full-batch gradients, and W
replicas stepping with them must stay bit-identical to each other (SURVEY.md §8e invariance test)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import taper_ref as R
    from taper_b200.dp import shard_permutation
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, steps, n = 16, 3, 200
        rng = np.random.default_rng(0)                       # same seed on every rank: identical data, weights, permutation
        X = rng.random((n, 784)).astype(np.float32)
        Y = rng.integers(0, 10, n).astype(np.float32)
        perm = rng.permutation(n)
        model = R.build_mlp([784, 32, 10], np.random.default_rng(1))
        full = R.build_mlp([784, 32, 10], np.random.default_rng(1))
        opt = R.Adam(model.parameters(), 1e-3)
        opt_full = R.Adam(full.parameters(), 1e-3)
        mine = shard_permutation(perm, rank, world, b)
        assert len(mine) == (n // (world * b)) * b
        for s in range(steps):
            idx = mine[s * b:(s + 1) * b]
            gidx = perm[s * world * b:(s + 1) * world * b]
            assert np.array_equal(idx, gidx[rank * b:(rank + 1) * b])
            # local backward on the shard
            R.Tape.reset()
            loss = R.cross_entropy_loss(model.forward(R.Tensor.new(X[idx], (b, 784))), R.Tensor.new(Y[idx], (b,)))
            loss.backward()
            flat = np.concatenate([p.grad() for p in model.parameters()])
            t = torch.from_numpy(flat)
            dist.all_reduce(t)                               # the one exchange step of the path
            avg = t.numpy() / np.float32(world)
            # single-process reference on the global minibatch
            R.Tape.reset()
            lf = R.cross_entropy_loss(full.forward(R.Tensor.new(X[gidx], (world * b, 784))), R.Tensor.new(Y[gidx], (world * b,)))
            lf.backward()
            ref = np.concatenate([p.grad() for p in full.parameters()])
            assert np.max(np.abs(avg - ref)) <= 1e-4 * max(np.max(np.abs(ref)), 1e-6)
            off = 0
            for p in model.parameters():
                p.set_grad(avg[off:off + p.data().size]); off += p.data().size
            opt.step(); opt.zero_grad()
            opt_full.step(); opt_full.zero_grad()
        # replicas stay identical across ranks
        mine_p = torch.from_numpy(np.concatenate([p.data() for p in model.parameters()]).copy())
        gathered = [torch.empty_like(mine_p) for _ in range(world)]
        dist.all_gather(gathered, mine_p)
        for g in gathered:
            assert torch.equal(g, gathered[0])
        q.put((rank, "ok"))
    except Exception as e:                                   # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_dp_invariance_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_permutation_partitions_every_global_batch():
    from taper_b200.dp import shard_permutation
    perm = np.random.default_rng(0).permutation(1000)
    for world in (1, 2, 4, 8):
        b = 16
        shards = [shard_permutation(perm, r, world, b) for r in range(world)]
        steps = 1000 // (world * b)
        assert all(len(s) == steps * b for s in shards)
        for s in range(steps):
            got = np.concatenate([sh[s * b:(s + 1) * b] for sh in shards])
            assert np.array_equal(got, perm[s * world * b:(s + 1) * world * b])
