"""World-size-2 gloo test of the data-parallel training plumbing (CPU only, no kernels): the flat gradient bucket
averages per-rank gradients exactly like DistributedDataParallel would, skips frozen parameters, fills missing
gradients with zeros, and the construction-time broadcast makes rank 1 equal to rank 0."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from damvsnet_b200 import training


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model(seed):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    m[2].bias.requires_grad_(False)          # a frozen parameter stays out of the bucket
    return m


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _make_model(seed=rank)            # different init per rank ...
    training.broadcast_module_state([model])  # ... until rank 0's state is broadcast
    state = torch.cat([p.detach().flatten() for p in model.parameters()])
    bucket = training.GradientBucket(model.parameters())
    x = torch.full((4, 5), float(rank + 1))
    model(x).sum().backward()
    if rank == 1:
        model[0].bias.grad = None              # a parameter without gradient on one rank counts as zero
    local = [None if p.grad is None else p.grad.clone() for p in bucket.params]
    finish = bucket.allreduce(async_op=True)
    finish()
    ret[rank] = (state, local, [p.grad.clone() for p in bucket.params], len(bucket.params), bucket.numel)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_matches_mean_of_local_gradients():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        s0, l0, g0, n0, numel0 = ret[0]
        s1, l1, g1, n1, _ = ret[1]
    assert torch.equal(s0, s1), "broadcast_module_state did not synchronise the ranks"
    assert n0 == n1 == 5 and numel0 == 5 * 7 + 7 + 7 + 7 + 7 * 3          # frozen bias excluded
    for a, b, ga, gb in zip(l0, l1, g0, g1):
        za = a if a is not None else torch.zeros_like(ga)
        zb = b if b is not None else torch.zeros_like(ga)
        want = (za + zb) / 2
        torch.testing.assert_close(ga, want)
        torch.testing.assert_close(gb, want)


def test_bucket_is_noop_without_process_group():
    m = _make_model(0)
    m(torch.ones(4, 5)).sum().backward()
    b = training.GradientBucket(m.parameters())
    before = [p.grad.clone() for p in b.params]
    assert b.allreduce() is None
    for p, g in zip(b.params, before):
        assert torch.equal(p.grad, g)


def test_depth_loss_matches_reference_formula():
    g = torch.Generator().manual_seed(0)
    outs = [{"depth": torch.rand(2, 4, 6, generator=g) * 100} for _ in range(3)]
    gts = [torch.rand(2, 4, 6, generator=g) * 100 for _ in range(3)]
    masks = [(torch.rand(2, 4, 6, generator=g) > 0.3).float() for _ in range(3)]
    got = training.depth_loss(outs, gts, masks, (0.5, 1.0, 2.0))
    want = sum(w * torch.nn.functional.smooth_l1_loss(o["depth"][m > 0.5], t[m > 0.5]) for o, t, m, w in zip(outs, gts, masks, (0.5, 1.0, 2.0)))
    torch.testing.assert_close(got, want)


def _overlap_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    stages = [torch.nn.Linear(4, 4) for _ in range(3)]
    ob = training.OverlappedBuckets([list(m.parameters()) for m in stages])
    outs = []
    for step in range(2):                                  # hooks and counters survive several steps
        for m in stages:
            m.zero_grad(set_to_none=False)
        x = torch.full((2, 4), float(rank + 1 + step))
        (stages[0](x).sum() + 2 * stages[1](x).sum() + 3 * stages[2](x).sum()).backward()
        local = [p.grad.clone() for m in stages for p in m.parameters()]
        ob.finish()
        outs.append((local, [p.grad.clone() for m in stages for p in m.parameters()]))
    ret[rank] = outs
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_overlapped_stage_buckets_average_gradients():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_overlap_worker, args=(world, port, ret), nprocs=world, join=True)
        r0, r1 = ret[0], ret[1]
    for step in range(2):
        (l0, g0), (l1, g1) = r0[step], r1[step]
        for a, b, ga, gb in zip(l0, l1, g0, g1):
            torch.testing.assert_close(ga, (a + b) / 2)
            torch.testing.assert_close(gb, (a + b) / 2)


def _skew_worker(rank, world, port, ret):
    """One rank's middle stage never gets a gradient (its hooks never complete that bucket): the collectives must still
    pair up -- every rank issues them in the same order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    stages = [torch.nn.Linear(4, 4) for _ in range(3)]
    ob = training.OverlappedBuckets([list(m.parameters()) for m in stages])
    for m in stages:
        m.zero_grad(set_to_none=False)
    x = torch.full((2, 4), float(rank + 1))
    loss = stages[0](x).sum() + 3 * stages[2](x).sum()
    if rank == 0:
        loss = loss + 2 * stages[1](x).sum()              # rank 1: stage 1 takes no part in the loss
    loss.backward()
    local = [torch.zeros_like(p) if p.grad is None else p.grad.clone() for m in stages for p in m.parameters()]
    ob.finish()
    got = [p.grad.clone() for m in stages for p in m.parameters()]
    # a second backward before finish() is refused instead of silently corrupting the averages
    raised = False
    (stages[0](x).sum() + stages[1](x).sum() + stages[2](x).sum()).backward()
    try:
        stages[2](x).sum().backward()
    except RuntimeError as exc:
        raised = "second backward" in str(exc)
    ob.reset()
    ob.close()
    n_hooks_left = len(ob._handles)
    stages[2](x).sum().backward()                            # hooks are gone: nothing fires, nothing raises
    ret[rank] = (local, got, raised, n_hooks_left)
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_buckets_fixed_order_guard_and_close():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_skew_worker, args=(world, port, ret), nprocs=world, join=True)
        (l0, g0, raised0, h0), (l1, g1, raised1, h1) = ret[0], ret[1]
    for a, b, ga, gb in zip(l0, l1, g0, g1):
        torch.testing.assert_close(ga, (a + b) / 2)
        torch.testing.assert_close(gb, (a + b) / 2)
    assert raised0 and raised1
    assert h0 == 0 and h1 == 0
