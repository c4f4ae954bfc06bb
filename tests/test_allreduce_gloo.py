"""CPU, world_size=2 (gloo): the training path's only collective -- one flat all-reduce (sum) of the gradients."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avvad.train import allreduce_gradients


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(4096, 13), (7,), (3, 5, 2)]
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
    for p in params:
        p.grad = torch.randn(p.shape, generator=g)
    mine = [p.grad.clone() for p in params]
    bucket = allreduce_gradients(params)
    assert bucket is not None and bucket.numel() == sum(p.numel() for p in params)
    torch.save({"mine": mine, "summed": [p.grad.clone() for p in params]}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_allreduce_sums_gradients(tmp_path):
    port = 29900 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "r0.pt"), torch.load(tmp_path / "r1.pt")
    for a, b, s0, s1 in zip(r0["mine"], r1["mine"], r0["summed"], r1["summed"]):
        assert torch.allclose(s0, a + b) and torch.equal(s0, s1)


def test_allreduce_is_a_noop_without_process_group():
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    assert allreduce_gradients([p]) is None and torch.equal(p.grad, torch.ones(3))


def _arena_worker(rank, world, port, out_dir):
    from avvad.train import GradientArena
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    frozen = torch.nn.Linear(3, 2)
    for p in frozen.parameters():
        p.requires_grad = False
    arena = GradientArena(list(lin.parameters()) + list(frozen.parameters()))
    assert arena.flat.numel() == 18 and arena.attached()
    x = torch.randn(4, 5, generator=torch.Generator().manual_seed(10 + rank))
    frozen(lin(x)).sum().backward()                     # autograd accumulates INTO the arena views
    assert arena.attached()
    local = arena.flat.clone()
    assert torch.equal(local[:15].view(3, 5), lin.weight.grad)
    arena.all_reduce()
    summed = arena.flat.clone()
    # a caller that replaces .grad (zero_grad(set_to_none=True) + backward) is re-gathered, not ignored
    lin.zero_grad(set_to_none=True)
    frozen(lin(x)).sum().backward()
    assert not arena.attached()
    arena.all_reduce()
    assert arena.attached() and torch.allclose(arena.flat, summed)
    arena.zero()
    assert float(lin.weight.grad.abs().sum()) == 0.0
    torch.save({"local": local, "summed": summed}, os.path.join(out_dir, f"a{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_arena_allreduce_in_place(tmp_path):
    """The Trainer's path: .grad tensors are views into one flat buffer, the all-reduce runs in place on it."""
    port = 29400 + os.getpid() % 500
    mp.spawn(_arena_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "a0.pt"), torch.load(tmp_path / "a1.pt")
    assert torch.allclose(r0["summed"], r0["local"] + r1["local"]) and torch.equal(r0["summed"], r1["summed"])
