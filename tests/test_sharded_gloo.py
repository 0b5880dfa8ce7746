"""world_size-2 gloo test of the patch-sharded driver (host logic of SURVEY.md section 8 e)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import multimodal_isic_b200 as pkg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = np.arange(1, n + 1, dtype=float)  # increasing cost -> unequal shard lengths

    def fake_extract(lo, hi):  # stands in for the per-rank GPU engine
        idx = torch.arange(lo, hi, dtype=torch.float64)
        return idx[:, None] * 10 + torch.arange(F, dtype=torch.float64)[None, :], (idx % 3 == 0).to(torch.int32)

    full, status, bounds = pkg.sharded_extract(fake_extract, n, costs)
    ok = torch.equal(full, torch.arange(n, dtype=torch.float64)[:, None] * 10 + torch.arange(F, dtype=torch.float64))
    ok = ok and torch.equal(status, (torch.arange(n) % 3 == 0).to(torch.int32))
    q.put((rank, bool(ok), bounds))
    dist.destroy_process_group()


def test_sharded_extract_world2():
    world, n, F = 2, 37, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, F, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res)
    b = res[0][2]
    assert b == res[1][2] and b[0] == 0 and b[-1] == n and b[1] > n // 2  # cost-balanced, not count-balanced


def _worker_overlapped(rank, world, port, rows, F, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    ok = True
    for pieces, layout in ((1, None), (2, "contiguous"), (3, None), (rows + 5, None), (1, "block_cyclic"), (11, "block_cyclic"),
                           ((0.5, 0.3, 0.15, 0.05), None), ((0.6, 0.4), "contiguous")):
        og = pkg.OverlappedGather(rows, F, world, "cpu", pieces=pieces, layout=layout)

        def fake_extract(lo, hi, out, status, og=og):  # stands in for Engine.extract_device on this rank's slice
            idx = torch.as_tensor(og.global_index(rank, np.arange(lo, hi)), dtype=torch.float64)
            out.copy_(idx[:, None] * 10 + torch.arange(F, dtype=torch.float64)[None, :])
            status.copy_((idx % 3 == 0).to(torch.int32))

        out, status = torch.zeros((rows, F), dtype=torch.float64), torch.zeros(rows, dtype=torch.int32)
        gathered = torch.zeros((world * rows, F), dtype=torch.float64)
        og.run(fake_extract, out, status, gathered)
        want = torch.arange(world * rows, dtype=torch.float64)[:, None] * 10 + torch.arange(F, dtype=torch.float64)
        ok = ok and torch.equal(gathered, want) and og.bounds[0] == 0 and og.bounds[-1] == rows
        ok = ok and sorted(int(og.global_index(r, l)) for r in range(world) for l in range(rows)) == list(range(world * rows))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_overlapped_gather_world2():
    """The sliced extract + all-gather pipeline bench.py uses for N > 1 (side stream on GPUs; same slicing on gloo)."""
    world, rows, F = 2, 11, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlapped, args=(r, world, port, rows, F, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert all(ok for _, ok in res)
