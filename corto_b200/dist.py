"""Multi-GPU plumbing for the batched decoder.

Meshes are independent units, so the decode path itself has NO collective: each rank decodes its own bin.  The only
exchange is moving the COMPRESSED blobs (~5 B/vertex) from the rank that ingested them to the ranks that decode them —
a ragged scatter done with point-to-point send/recv over the process group (NCCL over NVLink on GPUs, gloo in the CPU
tests).  Outputs stay sharded on the GPU that produced them (a gather would be bounded by one GPU's NVLink ingress and
is left to the consumer).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import shard_lpt, lib
import ctypes as C


def plan(blobs, world):
    """LPT assignment of blobs to ranks from their headers only (no decode).  Returns int array rank_of[n]."""
    L = lib()
    nv, nf, na = [], [], []
    for b in blobs:
        a = np.ascontiguousarray(b)
        ptrs = (C.c_void_p * 1)(a.ctypes.data)
        lens = (C.c_int * 1)(len(a))
        h = L.crt_batch_create(1, ptrs, lens)
        if not h:
            raise RuntimeError(L.crt_last_error().decode())
        v, f, m = C.c_uint32(), C.c_uint32(), C.c_uint32()
        L.crt_batch_mesh_info(h, 0, C.byref(v), C.byref(f), C.byref(m))
        L.crt_batch_destroy(h)
        nv.append(v.value); nf.append(f.value); na.append(bin(m.value & 0x1f).count("1"))
    return shard_lpt(nv, nf, na, world)


def scatter_blobs(blobs, src=0, device=None, group=None):
    """Rank `src` holds `blobs` (list of uint8 arrays); every rank returns (its blobs, their global indices)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    device = device or ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    meta = [None]
    if rank == src:
        blobs = [_aligned(b) for b in blobs]          # the header parse wants 4-byte aligned memory
        rank_of = plan(blobs, world)
        meta = [(rank_of.tolist(), [len(b) for b in blobs])]
    dist.broadcast_object_list(meta, src=src, group=group)
    rank_of, lens = meta[0]
    ids = [i for i, r in enumerate(rank_of) if r == rank]
    mine = []
    if rank == src:
        reqs = []
        for i, r in enumerate(rank_of):
            if r == src:
                continue
            t = torch.from_numpy(np.ascontiguousarray(blobs[i]).copy()).to(device)
            reqs.append(dist.isend(t, dst=r, group=group))
        for q in reqs:
            q.wait()
        mine = [np.ascontiguousarray(blobs[i]) for i in ids]
    else:
        for i in ids:
            t = torch.empty(lens[i], dtype=torch.uint8, device=device)
            dist.recv(t, src=src, group=group)
            mine.append(t.cpu().numpy())
    return mine, ids


def _aligned(b):
    b = np.ascontiguousarray(b, dtype=np.uint8)
    if b.ctypes.data % 16 == 0:
        return b
    raw = np.empty(b.size + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off:off + b.size]
    out[:] = b
    return out
