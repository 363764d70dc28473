"""Multi-GPU plumbing for the batched decoder.

Meshes are independent units, so the decode path itself has NO collective: each rank decodes its own bin.  The only
exchange is moving the COMPRESSED blobs (~5 B/vertex) from the rank that ingested them to the ranks that decode them:

  ingest rank   walk tapes (the few hundred header / directory bytes per blob, corto_b200.walk_tape) -> LPT plan
                (crt_shard_lpt over 4·nface + nvert·nattr) -> blobs laid out bin after bin in ONE device arena
  exchange      ONE grouped send/recv (torch.distributed.batch_isend_irecv = ncclGroupStart .. ncclSend/ncclRecv .. ncclGroupEnd
                over NVLink on GPUs; gloo in the CPU tests): rank r receives its bin's slice of the arena, contiguous, in device
                memory.  The tapes travel as (small) metadata.
  decode rank   corto_b200.BatchDecoder.from_device(tapes, lens, arena): the directory is rebuilt from the tapes, the payload
                never returns to a host.

Outputs stay sharded on the GPU that produced them; `gather_rows` moves an output arena to one rank when a consumer needs
that (bounded by that GPU's NVLink ingress — timed separately by bench.py).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import shard_lpt, walk_tape


def _round16(n):
    return (int(n) + 15) // 16 * 16


def plan(blobs, world):
    """LPT assignment of HOST blobs to ranks from their headers only (no decode).
    Returns (rank_of[n], tapes[n], cost[n]) — cost is the model the LPT balances (4·nface + nvert·nattr)."""
    tapes, nv, nf, na = [], [], [], []
    for b in blobs:
        t, v, f, a = walk_tape(b)
        tapes.append(t); nv.append(v); nf.append(f); na.append(a)
    rank_of = shard_lpt(nv, nf, na, world)
    cost = 4 * np.asarray(nf, dtype=np.int64) + np.asarray(nv, dtype=np.int64) * np.asarray(na, dtype=np.int64)
    return rank_of, tapes, cost


class Ingest:
    """What the ingest rank prepares once per batch: the plan, the tapes, and ONE arena (host, optionally device) in which
    every rank's bin is a contiguous 16-byte-aligned slice."""

    def __init__(self, blobs, world, device=None, pin=False):
        blobs = [np.ascontiguousarray(b, dtype=np.uint8) for b in blobs]
        self.world = world
        self.rank_of, self.tapes, self.cost = plan(blobs, world)
        self.ids = [[i for i in range(len(blobs)) if self.rank_of[i] == r] for r in range(world)]
        self.lens = [int(len(b)) for b in blobs]
        self.slice_off, self.slice_len = [], []
        tot = 0
        for r in range(world):
            n = sum(_round16(self.lens[i]) for i in self.ids[r])
            self.slice_off.append(tot); self.slice_len.append(n)
            tot += n
        host = torch.empty(max(tot, 16), dtype=torch.uint8)
        if pin:
            host = host.pin_memory()
        hv = host.numpy()
        for r in range(world):
            o = self.slice_off[r]
            for i in self.ids[r]:
                hv[o:o + self.lens[i]] = blobs[i]
                o += _round16(self.lens[i])
        self.host = host
        self.arena = host.to(device, non_blocking=False) if device is not None and str(device) != "cpu" else host
        load = np.array([self.cost[self.ids[r]].sum() if self.ids[r] else 0 for r in range(world)], dtype=np.float64)
        self.load_max_over_mean = float(load.max() / load.mean()) if load.mean() > 0 else 1.0

    def meta(self):
        return dict(ids=self.ids, lens=self.lens, tapes=[t.tobytes() for t in self.tapes], slice_len=self.slice_len,
                    load_max_over_mean=self.load_max_over_mean)


def scatter_blobs(ingest, src=0, device=None, group=None, meta=None, out=None):
    """Rank `src` passes its `Ingest`; every rank returns dict(arena, tapes, lens, ids, meta) for ITS bin, the arena in
    `device` memory.  ONE grouped send/recv moves the payload.  `meta` (from an earlier call) skips the metadata broadcast;
    `out` (the arena of an earlier call) receives in place, so a BatchDecoder built over it stays valid."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = torch.device(device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu"))
    if meta is None:
        box = [ingest.meta() if rank == src else None]
        dist.broadcast_object_list(box, src=src, group=group)
        meta = box[0]
    ids = meta["ids"][rank]
    lens = [meta["lens"][i] for i in ids]
    tapes = [np.frombuffer(meta["tapes"][i], dtype=np.uint8) for i in ids]
    n = meta["slice_len"][rank]
    if rank == src:
        arena_all = ingest.arena
        ops = [dist.P2POp(dist.isend, arena_all[ingest.slice_off[r]: ingest.slice_off[r] + ingest.slice_len[r]], r, group)
               for r in range(world) if r != src and ingest.slice_len[r]]
        arena = arena_all[ingest.slice_off[src]: ingest.slice_off[src] + max(n, 16)]      # its own bin: a view, nothing moves
    else:
        arena = out if out is not None else torch.empty(max(n, 16), dtype=torch.uint8, device=device)
        ops = [dist.P2POp(dist.irecv, arena[:n], src, group)] if n else []
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return dict(arena=arena, tapes=tapes, lens=lens, ids=ids, meta=meta)


def gather_rows(x, rows_per_rank, dst=0, group=None):
    """Concatenate the per-rank output arenas `x` (rows_per_rank[r] rows on rank r) on rank `dst` with one grouped exchange.
    Returns the gathered tensor on `dst`, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if rank == dst:
        out = torch.empty((int(sum(rows_per_rank)),) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        offs = np.concatenate([[0], np.cumsum(rows_per_rank)]).astype(np.int64)
        ops = [dist.P2POp(dist.irecv, out[offs[r]:offs[r + 1]], r, group) for r in range(world) if r != dst and rows_per_rank[r]]
        out[offs[dst]:offs[dst + 1]].copy_(x[:rows_per_rank[dst]])
    else:
        out = None
        ops = [dist.P2POp(dist.isend, x[:rows_per_rank[rank]].contiguous(), dst, group)] if rows_per_rank[rank] else []
    if ops:
        for q in dist.batch_isend_irecv(ops):
            q.wait()
    return out
