"""CPU tier, world_size 2 over gloo: the host logic of the data-parallel path -- rank-sharded bucketed batches
(disjoint ids, same number of batches on every rank) and the flat-arena gradient mean (DDP semantics)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from torch import nn

        from scal_sdt_b200 import GradExchange, ParamArena
        from scal_sdt_b200.bucket import DEFAULT_BUCKET_CONFIG, AspectSampler
        ex = GradExchange.from_torch_distributed(torch.device("cpu"))
        assert (ex.rank, ex.world) == (rank, world)
        # identical parameters on every rank, rank-dependent gradients -> arena mean
        torch.manual_seed(0)
        net = nn.Sequential(nn.Linear(6, 10), nn.Linear(10, 3))
        arena = ParamArena([{"params": list(net[0].parameters()), "lr": 1e-3}, {"params": list(net[1].parameters())}])
        for i, p in enumerate(net.parameters()):
            p.grad.fill_(float(rank + 1) * (i + 1))
        ex.all_reduce_mean_(arena.grads)
        expect = (1 + world) / 2.0
        ok_grad = all(torch.allclose(p.grad, torch.full_like(p.grad, expect * (i + 1))) for i, p in enumerate(net.parameters()))
        # chunked exchange driven by autograd hooks (the overlapped all-reduce of a native fine-tune): same means as one
        # all-reduce of the arena, every chunk sent exactly once, chunks cover the arena without gaps
        from scal_sdt_b200.comm import OverlappedExchange
        torch.manual_seed(1)
        net2 = nn.Sequential(nn.Linear(5, 7), nn.Tanh(), nn.Linear(7, 7), nn.Tanh(), nn.Linear(7, 2), nn.Linear(2, 2))
        for p in net2[5].parameters():
            p.requires_grad_(False)                           # a frozen tail: never in the arena
        arena2 = ParamArena([{"params": [p for p in net2.parameters() if p.requires_grad]}])
        ov = OverlappedExchange(arena2, ex, chunk_bytes=160)
        cover = [(c["begin"], c["end"]) for c in ov.chunks]
        ok_cover = cover[0][0] == 0 and cover[-1][1] == arena2.numel and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        x = torch.full((3, 5), float(rank + 1))
        arena2.zero_grad()
        ov.begin()
        net2(x).sum().backward()
        sent_in_backward = ov.launched
        local = None
        ov.finish()
        got = arena2.grads.clone()
        # reference: the same local gradients reduced in one piece
        arena2.zero_grad()
        net2(x).sum().backward()
        ex.all_reduce_mean_(arena2.grads)
        ok_overlap = (ok_cover and len(ov.chunks) >= 3 and sent_in_backward == len(ov.chunks) and ov.launched == len(ov.chunks)
                      and torch.allclose(got, arena2.grads, rtol=0, atol=0))
        ok_grad = ok_grad and ok_overlap
        # sharded sampler
        sizes = [(512, 512), (768, 512), (512, 768), (640, 448)]
        rs = np.random.RandomState(0)
        idmap = {i: sizes[int(k)] for i, k in enumerate(rs.randint(0, 4, size=203))}
        s = AspectSampler(idmap, 512, DEFAULT_BUCKET_CONFIG, 4, 114514, world, rank)
        batches = list(s.batches())
        ids = sorted(i for b, _ in batches for i in b)
        gathered = [None] * world
        dist.all_gather_object(gathered, (len(batches), ids))
        q.put((rank, ok_grad, gathered))
    finally:
        dist.destroy_process_group()


def test_world2_gradient_mean_and_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_grad, gathered in results:
        assert ok_grad
        (n0, ids0), (n1, ids1) = gathered
        assert n0 == n1 == 25                      # 203 ids -> 200 usable -> 100 per rank -> 25 batches of 4
        assert not set(ids0) & set(ids1) and len(ids0) == len(ids1) == 100
