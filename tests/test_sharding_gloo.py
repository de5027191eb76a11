"""CPU, world_size 2 over gloo: host logic of the multi-GPU sampling driver (shard bounds, per-rank seeds,
ragged final gather).  The per-shard sampler is stubbed; the CUDA path itself is covered by the gpu tests."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from v_diffusion_b200.generate import shard_bounds, sample_sharded


def test_shard_bounds_partition_everything():
    for total in (1, 7, 8, 50000, 4096 * 8 + 3):
        for world in (1, 2, 3, 8):
            if total < world:
                continue
            edges = [shard_bounds(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, batch, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, _ = shard_bounds(total, rank, world)
    calls = []

    def sample_fn(n, gen):
        # each "image" carries (rank, running index in shard, a draw from the per-rank generator)
        i0 = sum(calls)
        calls.append(n)
        x = torch.zeros(n, 3, 2, 2)
        x[:, 0] = rank
        x[:, 1] = (torch.arange(n) + i0 + start).view(n, 1, 1).float()
        x[:, 2] = torch.rand(n, generator=gen).view(n, 1, 1)
        return x
    full = sample_sharded(sample_fn, total, batch, seed=100)
    torch.save((full, calls), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_sharded_world2(tmp_path):
    total, batch, world = 11, 4, 2
    mp.spawn(_worker, args=(world, _free_port(), total, batch, str(tmp_path)), nprocs=world, join=True)
    (f0, c0), (f1, c1) = (torch.load(tmp_path / f"r{r}.pt") for r in range(world))
    assert torch.equal(f0, f1) and f0.shape == (11, 3, 2, 2)              # every rank holds the full gather
    assert c0 == [4, 2] and c1 == [4, 1]                                   # rank 0 owns 6 samples, rank 1 owns 5
    assert f0[:, 1, 0, 0].tolist() == list(range(11))                      # global order preserved
    assert f0[:6, 0, 0, 0].eq(0).all() and f0[6:, 0, 0, 0].eq(1).all()
    g0 = torch.rand(4, generator=torch.Generator().manual_seed(100))      # per-rank seeds: seed + rank
    g1 = torch.rand(4, generator=torch.Generator().manual_seed(101))
    assert torch.allclose(f0[:4, 2, 0, 0], g0) and torch.allclose(f0[6:10, 2, 0, 0], g1)


def _queue_worker(rank, world, port, total, batch, out_dir):
    import time
    from v_diffusion_b200.generate import sample_balanced
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ran = []

    def sample_fn(n, gen):
        time.sleep(0.05 if rank == 0 else 0.2)           # rank 1 is the slow GPU
        ran.append(n)
        x = torch.zeros(n, 2, 2, 2)
        x[:, 0] = torch.rand(n, generator=gen).view(n, 1, 1)
        x[:, 1] = rank
        return x
    full = sample_balanced(sample_fn, total, batch, seed=100)
    again = sample_balanced(sample_fn, total, batch, seed=100)            # a second queue on the same store
    torch.save((full, again, ran), os.path.join(out_dir, f"q{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sample_balanced_world2_matches_single_process(tmp_path):
    """The dynamic batch queue: every batch runs exactly once, the faster rank runs more of them, and the images do
    not depend on who ran what (batch k is seeded seed + k) -- equal to a single-process run."""
    from v_diffusion_b200.generate import sample_balanced
    total, batch, world = 38, 4, 2
    mp.spawn(_queue_worker, args=(world, _free_port(), total, batch, str(tmp_path)), nprocs=world, join=True)
    (f0, a0, r0), (f1, a1, r1) = (torch.load(tmp_path / f"q{r}.pt") for r in range(world))
    assert torch.equal(f0, f1) and f0.shape == (38, 2, 2, 2)
    assert torch.equal(f0[:, 0], a0[:, 0])                                 # the second call drew the same numbers
    assert sum(r0) + sum(r1) == 2 * total and len(r0) > len(r1)            # nothing twice, the fast rank did more

    def sample_fn(n, gen):
        x = torch.zeros(n, 2, 2, 2)
        x[:, 0] = torch.rand(n, generator=gen).view(n, 1, 1)
        return x
    single = sample_balanced(sample_fn, total, batch, seed=100)
    assert torch.equal(single[:, 0], f0[:, 0])
    want = torch.cat([torch.rand(min(4, total - 4 * k), generator=torch.Generator().manual_seed(100 + k)) for k in range(10)])
    assert torch.equal(single[:, 0, 0, 0], want)
