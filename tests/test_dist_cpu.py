"""world_size-2 gloo test (CPU) of the multi-GPU host logic of bench.py: scan pairs are sharded by rank with no
overlap, the step time is the MAX over ranks, and the reported value is the whole-job aggregate."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert bench.dist_env() == (rank, world, rank)
    seeds = bench.shard_seeds(bench.WORKLOADS["pretrain"], n_batches=3, rank=rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds)
    flat = [s for r in gathered for b in r for s in b]
    assert len(flat) == len(set(flat)) == world * 3 * 4, "every rank must get disjoint scan pairs"
    t = torch.tensor([10.0 + 5 * rank], dtype=torch.float64)  # rank 1 is slower
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = bench.aggregate_value(batch=4, steps=10, world=world, total_ms=float(t))
    out[rank] = (float(t), value)
    dist.destroy_process_group()


def test_sharding_and_max_over_ranks():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, 29541, out), nprocs=world, join=True)
        assert out[0] == out[1]
        t, value = out[0]
        assert t == 15.0 and abs(value - 4 * 10 * 2 / 0.015) < 1e-6


def _allreduce_worker(rank, world, port, out):
    import importlib
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tdist = importlib.import_module("tmae_b200.dist")
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2, 2))]
    ps[0].grad = torch.full((3, 4), float(rank + 1))
    ps[1].grad = torch.arange(5.0) * (rank + 1)
    ps[2].grad = None if rank == 1 else torch.ones(2, 2)  # missing on one rank: counts as zeros
    tdist.allreduce_gradients(ps)
    out[rank] = [p.grad.clone() for p in ps]
    dist.destroy_process_group()


def test_allreduce_gradients_gloo_world2():
    """The pretraining gradient exchange (one flat all-reduce, mean over ranks) on 2 CPU ranks over gloo."""
    import torch
    import torch.multiprocessing as mp
    import tmae_b200  # noqa: F401  (package alias)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_allreduce_worker, args=(2, 29533, out), nprocs=2, join=True)
    for r in range(2):
        g0, g1, g2 = out[r]
        assert torch.equal(g0, torch.full((3, 4), 1.5))
        assert torch.equal(g1, torch.arange(5.0) * 1.5)
        assert torch.equal(g2, torch.full((2, 2), 0.5))


def _overlap_worker(rank, world, port, out):
    import importlib
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tdist = importlib.import_module("tmae_b200.dist")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
    unused = torch.nn.Parameter(torch.zeros(7))          # never receives a gradient on any rank: must keep grad None
    params = list(net.parameters()) + [unused]
    og = tdist.OverlappedGradients(params, world, n_buckets=3).attach()
    res = []
    for it in range(3):                                   # first step (presence agreed in finish) and steady state (hooks fire buckets)
        x = torch.full((2, 6), float(rank + 1 + it))
        net(x).sum().backward()
        og.finish()
        res.append([None if p.grad is None else p.grad.clone() for p in params])
        for p in params:
            p.grad = None
    out[rank] = res
    dist.destroy_process_group()


def test_overlapped_gradients_gloo_world2():
    """OverlappedGradients (bucketed all-reduce fired from post-accumulate hooks during backward) == the mean of the ranks' gradients;
    a parameter unused on every rank keeps grad None (as under DDP with find_unused_parameters=False)."""
    import torch
    import torch.multiprocessing as mp
    import tmae_b200  # noqa: F401
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_overlap_worker, args=(2, 29537, out), nprocs=2, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
    for it in range(3):
        ref = None
        for rank in range(2):
            net.zero_grad()
            net(torch.full((2, 6), float(rank + 1 + it))).sum().backward()
            g = [p.grad.clone() for p in net.parameters()]
            ref = g if ref is None else [a + b for a, b in zip(ref, g)]
        ref = [r / 2 for r in ref]
        for rank in range(2):
            got = out[rank][it]
            assert got[-1] is None
            for a, b in zip(got[:-1], ref):
                assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)


def _pipelined_worker(rank, world, port, out):
    import importlib
    import os
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tdist = importlib.import_module("tmae_b200.dist")
    res = {}
    for mode in ("single", "per_bucket"):
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 4), torch.nn.ReLU(), torch.nn.Linear(4, 3))
        params = list(net.parameters())
        fg = tdist.FlatGradients(params, world, n_buckets=3)
        if mode == "single":
            opts = [torch.optim.AdamW(params, lr=0.05, weight_decay=0.01)]
        else:
            opts = [torch.optim.AdamW(ps, lr=0.05, weight_decay=0.01) for ps in fg.bucket_params()]
            assert sum(len(ps) for ps in fg.bucket_params()) == len(params) and len(opts) == 3
        for it in range(3):
            net(torch.full((2, 6), float(rank + 1 + it))).sum().backward()
            if mode == "single":
                fg.reduce()
                opts[0].step()
            else:
                fg.reduce(step_fns=[o.step for o in opts])
            for o in opts:
                o.zero_grad(set_to_none=True)
        res[mode] = [p.detach().clone() for p in params]
    out[rank] = res
    dist.destroy_process_group()


def test_per_bucket_optimizer_steps_equal_one_step_gloo_world2():
    """FlatGradients.reduce(step_fns=...): one optimizer per bucket, stepped as its bucket's all-reduce completes, leaves the same
    parameters as one optimizer stepped behind the whole exchange, on both ranks."""
    import torch
    import torch.multiprocessing as mp
    import tmae_b200  # noqa: F401
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_pipelined_worker, args=(2, 29539, out), nprocs=2, join=True)
    for a, b in zip(out[0]["single"], out[0]["per_bucket"]):
        assert torch.equal(a, b)
    for a, b in zip(out[0]["per_bucket"], out[1]["per_bucket"]):
        assert torch.equal(a, b)


def test_balanced_batches_are_the_ranks_own_scans_sorted_by_cost():
    """bench.make_batches(balanced=True): every rank keeps exactly its own scan pairs (nothing changes owner), batched in ascending
    order of occupied pillars, so that lockstep steps line up across ranks."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from tmae_b200 import synth
    w = dict(bench.WORKLOADS["pretrain"], n_points=4000)
    shape = synth.SHAPES[w.get("shape", "once")]
    for rank in (0, 1):
        plain = bench.make_batches(w, 3, rank, balanced=False)
        bal = bench.make_batches(w, 3, rank, balanced=True)

        def scans(batches):
            out = []
            for pts, ptsp in batches:
                for b in range(w["batch"]):
                    out.append((pts[pts[:, 0] == b][:, 1:].numpy(), ptsp[ptsp[:, 0] == b][:, 1:].numpy()))
            return out
        a, b = scans(plain), scans(bal)
        key = lambda pr: (pr[0].shape[0], float(pr[0].sum()))
        assert sorted(map(key, a)) == sorted(map(key, b)), "the rank's scans changed"
        costs = [bench.scan_cost(pr, shape) for pr in b]
        assert costs == sorted(costs), "batches are not in ascending cost order"
