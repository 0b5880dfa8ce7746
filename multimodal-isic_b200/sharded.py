"""Patch-sharded multi-GPU driver (SURVEY.md section 8 e).  One process per GPU
(``torch.distributed``); the patch list is split into contiguous cost-balanced shards, every
rank runs the single-GPU engine on its shard, and the only collective is an all-gather of the
``[rows, F]`` float64 feature block (NCCL over NVLink on GPUs; gloo in the CPU tests, where
the per-rank compute is injected)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(costs, world):
    """Contiguous shards with near-equal total cost.  ``costs``: per-patch cost estimate
    (ROI pixel count or H*W).  Returns ``world + 1`` boundaries; deterministic on every rank."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if n == 0:
        return [0] * (world + 1)
    cum = np.concatenate([[0.0], np.cumsum(np.maximum(costs, 1e-9))])
    targets = cum[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cum, targets, side="left")
    # pick the boundary (inner or inner-1) closest to the target
    inner = np.where((inner > 0) & (np.abs(cum[np.maximum(inner - 1, 0)] - targets) < np.abs(cum[np.minimum(inner, n)] - targets)),
                     inner - 1, inner)
    b = [0] + [int(x) for x in np.clip(inner, 0, n)] + [n]
    for i in range(1, len(b)):
        b[i] = max(b[i], b[i - 1])
    return b


def all_gather_rows(local, rows_per_rank, group=None):
    """All-gather of variable-length row blocks: pads to the longest shard, one
    ``all_gather_into_tensor``, then trims.  ``local``: ``[rows_per_rank[rank], F]``."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    F = local.shape[1]
    mx = int(max(rows_per_rank)) if len(rows_per_rank) else 0
    pad = torch.zeros((mx, F), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty((world * mx, F), dtype=local.dtype, device=local.device)
    if world == 1:
        gathered.copy_(pad)
    elif local.is_cuda:
        dist.all_gather_into_tensor(gathered, pad, group=group)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        gathered = torch.cat(parts)
    assert rows_per_rank[rank] == local.shape[0]
    return torch.cat([gathered[r * mx: r * mx + int(rows_per_rank[r])] for r in range(world)])


def sharded_extract(extract_fn, n_patches, costs=None, group=None, F=None):
    """Run ``extract_fn(lo, hi) -> (features [hi-lo, F], status [hi-lo])`` on this rank's shard and
    all-gather to the full ``[n_patches, F]`` matrix (original order) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if costs is None:
        costs = np.ones(n_patches)
    b = shard_bounds(costs, world)
    lo, hi = b[rank], b[rank + 1]
    feats, status = extract_fn(lo, hi)
    if world == 1:
        return feats, status, b
    rows = [b[r + 1] - b[r] for r in range(world)]
    full = all_gather_rows(feats, rows, group)
    st = all_gather_rows(status.view(-1, 1).to(torch.int32), rows, group).view(-1)
    return full, st, b


class OverlappedGather:
    """Equal shards (weak scaling: every rank owns ``rows`` patches): the local shard is extracted in slices, and
    the all-gather of slice k (NCCL over NVLink, on a side stream) overlaps the extraction of slice k+1, so only the
    last slice's collective is exposed.  ``pieces`` = number of equal slices, or a sequence of fractions
    (e.g. ``(0.5, 0.3, 0.15, 0.05)``: a small last slice keeps the exposed collective small, large early slices keep
    the thread-level reduction kernels at several waves).  Slices alternate between two extraction streams, so the
    latency-bound reduction kernels of slice k run under the build kernel of slice k+1.

    Global order (``layout="block_cyclic"``, the default): the global patch list is dealt to the ranks slice by slice --
    slice k of rank r holds global rows ``world * bounds[k] + r * n_k + i`` (``n_k`` = slice length).
    ``all_gather_into_tensor`` of slice k then writes the contiguous global rows
    ``[world * bounds[k], world * bounds[k + 1])`` in place: no staging buffer and no re-ordering copy.
    ``layout="contiguous"`` keeps rank r's shard at global rows ``[r * rows, (r + 1) * rows)`` (gathered through a
    staging buffer and copied).  ``global_index(r, local)`` maps either way.

    ``extract_fn(lo, hi, out, status)`` writes the rows of local patches ``[lo, hi)`` into the given
    slices, asynchronously on the current stream (``Engine.extract_device`` does).  On CPU tensors
    (gloo tests) the same slicing runs without streams."""

    def __init__(self, rows, F, world, device, pieces=2, group=None, layout=None, bounds=None):
        self.rows, self.F, self.world, self.group = int(rows), int(F), int(world), group
        if bounds is not None:  # explicit slice boundaries (run_chunked: the engine's own chunk boundaries)
            self.bounds = [int(b) for b in bounds]
            assert self.bounds[0] == 0 and self.bounds[-1] == self.rows and all(a < b for a, b in zip(self.bounds, self.bounds[1:]))
        elif isinstance(pieces, (list, tuple)):
            cum, b = 0.0, [0]
            for f in pieces[:-1]:
                cum += float(f)
                b.append(min(self.rows, max(b[-1], int(round(self.rows * cum)))))
            b.append(self.rows)
            self.bounds = [b[0]] + [x for i, x in enumerate(b[1:]) if x > b[i]]  # drop empty slices
        else:
            k = max(1, min(int(pieces), self.rows or 1))
            self.bounds = [self.rows * i // k for i in range(k + 1)]
        self.pieces = len(self.bounds) - 1
        self.layout = layout or "block_cyclic"
        if self.layout not in ("block_cyclic", "contiguous"):
            raise ValueError("layout must be block_cyclic or contiguous")
        self.cuda = torch.device(device).type == "cuda"
        mx = max((self.bounds[k + 1] - self.bounds[k] for k in range(self.pieces)), default=0)
        # contiguous layout: one staging buffer per slice (slice k's collective may still run when slice k+1 is extracted)
        self.tmp = ([torch.empty((self.world, mx, self.F), dtype=torch.float64, device=device) for _ in range(self.pieces)]
                    if self.layout == "contiguous" else None)
        self.skip_collective = False  # measurement aid (bench.py attribution): run the slices, leave the collective out
        # high priority: the collective's few CTAs must get SM slots while the extraction's grids keep every SM full
        self.comm = torch.cuda.Stream(device=device, priority=-1) if self.cuda and self.world > 1 else None
        self.work = [torch.cuda.Stream(device=device) for _ in range(2)] if self.cuda and self.pieces > 1 else None

    def global_index(self, rank, local):
        """Global row of local patch ``local`` of rank ``rank`` (NumPy arrays or ints)."""
        if self.layout == "contiguous":
            return rank * self.rows + local
        b = np.asarray(self.bounds)
        k = np.searchsorted(b, local, side="right") - 1
        return self.world * b[k] + rank * (b[k + 1] - b[k]) + (local - b[k])

    def _slice_view(self, gathered, k):
        lo, hi = self.bounds[k], self.bounds[k + 1]
        return gathered[self.world * lo: self.world * hi]

    def run_chunked(self, engine, images, masks, out, status, gathered):
        """ONE extraction call for the whole shard: the engine pipelines its chunks (build kernel of chunk k + 1 under
        the reduction kernels of chunk k, exactly as on one GPU) and records a completion event per chunk
        (radb_set_chunk_events); the all-gather of chunk k's rows waits for event k on the communication stream.
        ``bounds`` must be the engine's chunk boundaries (``chunk_bounds``); block-cyclic layout only."""
        assert self.layout == "block_cyclic" and self.cuda
        cur = torch.cuda.current_stream(out.device)
        if getattr(self, "_events", None) is None:
            self._events = [torch.cuda.Event() for _ in range(self.pieces)]
            for e in self._events:
                e.record(cur)  # materialise the CUDA events
        if self.comm is not None:
            self.comm.wait_stream(cur)  # the previous step's consumers of `gathered` are done
        engine.set_chunk_events(self._events)
        engine.extract_device(images, masks, out, status)
        for k in range(self.pieces):
            lo, hi = self.bounds[k], self.bounds[k + 1]
            if self.world == 1:
                cur.wait_event(self._events[k])
                self._slice_view(gathered, k).copy_(out[lo:hi])
                continue
            if self.skip_collective:
                continue
            self.comm.wait_event(self._events[k])
            with torch.cuda.stream(self.comm):
                dist.all_gather_into_tensor(self._slice_view(gathered, k), out[lo:hi], group=self.group)  # in place
        if self.comm is not None:
            cur.wait_stream(self.comm)
        return gathered

    @staticmethod
    def chunk_bounds(engine, rows, H, W, dtype=torch.uint8):
        """Boundaries of the chunks a dense call with ``rows`` patches will use (for ``bounds=``)."""
        n = engine.chunk_rows(rows, H, W, dtype)
        return list(range(0, rows, n)) + [rows]

    def run(self, extract_fn, out, status, gathered):
        cur = torch.cuda.current_stream(out.device) if self.cuda else None
        if self.comm is not None:
            self.comm.wait_stream(cur)  # the previous step's consumers of `gathered` / tmp are done
        if self.work is not None:
            for w in self.work:
                w.wait_stream(cur)
        g = gathered.view(self.world, self.rows, self.F) if self.layout == "contiguous" else None
        for k in range(self.pieces):
            lo, hi = self.bounds[k], self.bounds[k + 1]
            ws = self.work[k % 2] if self.work is not None else cur
            if self.work is not None:
                with torch.cuda.stream(ws):
                    extract_fn(lo, hi, out[lo:hi], status[lo:hi])
            else:
                extract_fn(lo, hi, out[lo:hi], status[lo:hi])
            if self.world == 1:
                dst = g[0, lo:hi] if g is not None else self._slice_view(gathered, k)
                if self.work is not None:
                    with torch.cuda.stream(ws):
                        dst.copy_(out[lo:hi])
                else:
                    dst.copy_(out[lo:hi])
                continue
            if self.skip_collective:
                continue
            if self.comm is not None:
                ev = torch.cuda.Event()
                ev.record(ws)
                self.comm.wait_event(ev)
                with torch.cuda.stream(self.comm):
                    if g is None:
                        dist.all_gather_into_tensor(self._slice_view(gathered, k), out[lo:hi], group=self.group)  # in place
                    else:
                        flat = self.tmp[k].view(-1, self.F)[: self.world * (hi - lo)].view(self.world, hi - lo, self.F)
                        dist.all_gather_into_tensor(flat, out[lo:hi], group=self.group)
                        g[:, lo:hi].copy_(flat)
            else:
                parts = [torch.empty((hi - lo, self.F), dtype=out.dtype) for _ in range(self.world)]
                dist.all_gather(parts, out[lo:hi].contiguous(), group=self.group)
                n = hi - lo
                for r in range(self.world):
                    (g[r, lo:hi] if g is not None else self._slice_view(gathered, k)[r * n:(r + 1) * n]).copy_(parts[r])
        if self.cuda:
            if self.work is not None:
                for w in self.work:
                    cur.wait_stream(w)
            if self.comm is not None:
                cur.wait_stream(self.comm)
        return gathered
