"""torchrun worker of tests/test_gpu_multi.py: the patch-sharded multi-GPU data plane on REAL engines over NCCL.

The reference fans records over a process pool and gets the results back IN ORDER
(/root/reference/RadiomicExtractor.py:60-65: ``pool.imap``).  Here every rank extracts its shard on its own
GPU and the rows are all-gathered; this worker checks, on every rank, that the gathered matrix equals the
rows one GPU computes for the whole list (bit for bit), for both drivers:

* ``OverlappedGather`` (equal shards, sliced, the all-gather of slice k under the extraction of slice k+1:
  what ``bench.py --gpus N`` runs), and
* ``OverlappedGather.run_chunked`` (one extraction call, per-chunk completion events: radb_set_chunk_events), and
* ``sharded_extract`` (cost-balanced contiguous shards of unequal length, padded all-gather).

plus 16 rows spot-checked against the oracle on rank 0.  Prints ``NCCL_WORKER_OK`` per rank.
Launch:  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/nccl_worker.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import multimodal_isic_b200 as pkg  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    settings = {"label": 255, "binWidth": 25, "force2D": False}
    ex = pkg.RadiomicsExtractor({"setting": settings}, device=local)
    F = ex.engine.F

    # ---- the whole list (every rank builds the same one from the same seed, on the host)
    rows = 1536  # per rank
    n = world * rows
    imgs, masks = pkg.synth.make_patches(256, 64, seed=11)
    reps = (n + 255) // 256
    rng = np.random.default_rng(5)
    order = rng.permutation(256 * reps)[:n]
    imgs = np.tile(imgs, (reps, 1, 1))[order]
    masks = np.tile(masks, (reps, 1, 1))[order]
    # unequal costs: shrink the ROI of the second half of the list (cheaper patches at the end)
    masks[n // 2:, :24] = 0
    d_img, d_msk = torch.as_tensor(imgs).to(dev), torch.as_tensor(masks).to(dev)
    want, want_st = ex.engine.extract_device(d_img, d_msk)  # one GPU, the whole list
    torch.cuda.synchronize()

    # ---- 1. OverlappedGather: equal shards, sliced, overlapped all-gather
    for pieces, layout in ((1, "contiguous"), (2, "contiguous"), (4, "block_cyclic"), ((0.5, 0.3, 0.15, 0.05), "block_cyclic")):
        out = torch.zeros((rows, F), dtype=torch.float64, device=dev)
        status = torch.zeros((rows,), dtype=torch.int32, device=dev)
        gathered = torch.zeros((n, F), dtype=torch.float64, device=dev)
        og = pkg.OverlappedGather(rows, F, world, dev, pieces=pieces, layout=layout)

        def extract_slice(a, b, o, s, og=og):
            g0 = int(og.global_index(rank, a))  # a slice is a contiguous run of global rows in both layouts
            ex.engine.extract_device(d_img[g0:g0 + (b - a)], d_msk[g0:g0 + (b - a)], o, s)

        for _ in range(2):  # twice: buffer re-use across steps
            og.run(extract_slice, out, status, gathered)
        torch.cuda.synchronize()
        assert torch.equal(gathered.view(torch.int64), want.view(torch.int64)), "OverlappedGather(%d, %s) rows differ" % (pieces, layout)

    # ---- 1b. one extraction call per step, the all-gather of chunk k on the engine's completion event of chunk k
    # (what bench.py --gpus N runs by default): 4 chunks per shard
    ex.engine.set_chunk(rows // 4)
    bounds = pkg.OverlappedGather.chunk_bounds(ex.engine, rows, 64, 64)
    assert len(bounds) == 5, bounds
    og = pkg.OverlappedGather(rows, F, world, dev, bounds=bounds)
    out = torch.zeros((rows, F), dtype=torch.float64, device=dev)
    status = torch.zeros((rows,), dtype=torch.int32, device=dev)
    gathered = torch.zeros((n, F), dtype=torch.float64, device=dev)
    gidx = torch.as_tensor(og.global_index(rank, np.arange(rows)), device=dev)  # my local patches' global rows
    my_img, my_msk = d_img.index_select(0, gidx), d_msk.index_select(0, gidx)
    for _ in range(3):
        gathered.zero_()
        og.run_chunked(ex.engine, my_img, my_msk, out, status, gathered)
    torch.cuda.synchronize()
    assert torch.equal(gathered.view(torch.int64), want.view(torch.int64)), "run_chunked rows differ"
    ex.engine.set_chunk(0)

    # ---- 2. sharded_extract: cost-balanced shards of unequal length
    costs = (masks == 255).reshape(n, -1).sum(1).astype(np.float64)

    def extract_fn(a, b):
        return ex.engine.extract_device(d_img[a:b], d_msk[a:b])

    full, st, bounds = pkg.sharded_extract(extract_fn, n, costs)
    torch.cuda.synchronize()
    lens = [bounds[r + 1] - bounds[r] for r in range(world)]
    assert len(set(lens)) > 1, "the cost model should give unequal shards: %s" % lens
    assert torch.equal(full.view(torch.int64), want.view(torch.int64)), "sharded_extract rows differ"
    assert torch.equal(st, want_st)

    # ---- 3. oracle spot checks (rank 0): 16 rows of the gathered matrix
    if rank == 0:
        from oracle import cmatrices, radiomics_oracle as orc

        cmatrices.build()
        got = full.cpu().numpy()
        for b in np.linspace(0, n - 1, 16).astype(int):
            ref = np.array(list(orc.execute(imgs[b], masks[b], settings, matrix_backend=cmatrices).values()))
            np.testing.assert_allclose(got[b], ref, rtol=1e-6, atol=1e-9)
    dist.barrier()
    print("NCCL_WORKER_OK rank %d/%d shards %s" % (rank, world, lens), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
