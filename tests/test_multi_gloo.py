"""CPU: the N>1 path (replicate scene / shard views / one gradient all-reduce) on world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gs_b200 import multi


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert multi.world() == (rank, world)
        views = multi.shard_views(7)
        assert views == list(range(rank, 7, world))
        # scene replication from rank 0
        t = torch.arange(12, dtype=torch.float32).view(4, 3) if rank == 0 else torch.zeros(4, 3)
        ids = torch.arange(8, dtype=torch.uint8) if rank == 0 else torch.zeros(8, dtype=torch.uint8)
        multi.broadcast_scene([t, ids])
        assert torch.equal(t, torch.arange(12, dtype=torch.float32).view(4, 3)) and torch.equal(ids, torch.arange(8, dtype=torch.uint8))
        # per-view accumulation then ONE all-reduce; statistics are per view (non-linear), folded before the reduction
        P, M = 5, 4
        acc = multi.GradAccumulator(P, M, "cpu")
        bufs = acc.buffers()
        assert [tuple(b.shape) for b in bufs] == [(P, 3), (P, 3), (P, 1), (P, 3), (P, 6), (P, M, 3), (P, 3), (P, 4)]
        assert acc.floats_per_gaussian == 14 + 3 * M and acc.payload_floats == P * (14 + 3 * M + 2)    # 62 floats at M = 16 + 2 statistics
        for v in views:
            radii = torch.tensor([v % 2, 1, 0, 2, 3 * (v + 1)])
            for b in bufs:
                b += (v + 1)
            acc.view_means2D.fill_(float(v + 1))            # what the backward writes for this view (GsbGrads.dL_dmeans2D_view)
            acc.observe_view(radii)
        if rank == 0:
            acc.all_reduce()
        else:                                               # the asynchronous form: launch, then join
            acc.all_reduce(async_op=True)
            acc.wait()
        tot = sum(v + 1 for v in range(7))
        own = sum(v + 1 for v in views)
        for name, b in zip(multi.GRAD_NAMES, acc.buffers()):
            # dL_dcolors / dL_dcov3D are not part of the communicated payload: they keep this rank's own sum
            assert torch.all(b == (own if name in ("dL_dcolors", "dL_dcov3D") else tot)), name
        assert float(acc.denom[1]) == 7 and float(acc.denom[2]) == 0
        exp_norm = sum(math_sqrt2(v + 1) for v in range(7))
        assert abs(float(acc.xyz_gradient_accum[1]) - exp_norm) < 1e-4
        assert float(acc.max_radii2D[4]) == 21.0
        # degree-banded payload: Gaussians ordered by degree (2 of degree 0, 1 of degree 1, 2 of degree 1's successor ...): only the active
        # coefficients travel, the result equals the dense reduction on the active part and the inactive part stays zero
        counts = [2, 2, 0, 1] if M == 16 else [3, 2, 0, 0]
        accb = multi.GradAccumulator(P, M, "cpu", band_counts=counts)
        dense = multi.GradAccumulator(P, M, "cpu")
        g = torch.Generator().manual_seed(100 + rank)
        sh_local = torch.randn(P, M, 3, generator=g)
        start = 0
        for d, c in enumerate(counts):                           # what the kernels guarantee: zero outside each Gaussian's active bands
            sh_local[start:start + c, (d + 1) ** 2:, :] = 0.0
            start += c
        for a_ in (accb, dense):
            a_.buffers()[5].copy_(sh_local)
            a_.buffers()[3].fill_(float(rank + 1))
        assert accb.payload_floats < dense.payload_floats
        accb.all_reduce()
        dense.all_reduce()
        assert torch.equal(accb.buffers()[5], dense.buffers()[5]) and torch.equal(accb.buffers()[3], dense.buffers()[3])
        start = 0
        for d, c in enumerate(counts):
            assert float(accb.buffers()[5][start:start + c, (d + 1) ** 2:, :].abs().sum()) == 0.0
            start += c
        imgs = multi.gather_images(torch.full((3, 2, 2), float(rank)))
        if rank == 0:
            assert [float(i.mean()) for i in imgs] == [0.0, 1.0]
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def math_sqrt2(x):
    return (2 * x * x) ** 0.5


def test_view_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_defaults():
    assert multi.world() == (0, 1)
    assert multi.shard_views(5) == [0, 1, 2, 3, 4]
    acc = multi.GradAccumulator(3, 1, "cpu")
    acc.all_reduce()
    assert acc.small.numel() == 3 * (3 + 1 + 3 + 3 + 4 + 2) and acc.sh.numel() == 3 * 3 and acc.local.numel() == 3 * 9
