"""BASELINE.json workloads as lists of .crt blobs (test infrastructure: encoded by the reference Encoder).

SURVEY §8d stand-ins (bunny / Proserpina / Nile are not in the reference repo):
  c1  185^2 grid (34,225 v), pos14                                      x1
  c2  358^2 grid (128,164 v / 254,898 f), pos14 + uv12 + normal10 ESTIMATED   x batch, seeds differ per mesh
  c3  409^2 cloud (167,281 v), pos14 + colour 6/6/6/6 + normal10 DIFF   x batch
  c4  grids with nvert in [8K, 256K], all attributes, 1-4 groups, 10% holes, 10% two components
  c5  3163^2 grid (10.0 M v), pos14 + normal10 BORDER                   x1
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import meshgen as mg
from . import refshim


def _c2(seed):
    return refshim.encode(mg.grid(358, seed), pos_bits=14, uv_bits=12, normal_bits=10, normal_pred=1, with_colors=False)[0]


def _c3(seed):
    return refshim.encode(mg.cloud(409, seed), pos_bits=14, normal_bits=10, normal_pred=0, color_bits=(6, 6, 6, 6), with_uv=False)[0]


def _c4(seed):
    rs = np.random.RandomState(seed)
    W = int(rs.randint(90, 507))
    kind = rs.uniform()
    if kind < 0.1:
        m = mg.punch_hole(mg.grid(W, seed), W)
    elif kind < 0.2:
        m = mg.two_components(int(W * 0.85), seed)
    else:
        m = mg.grid(W, seed)
    ng = int(rs.randint(1, 5))
    return refshim.encode(m, pos_bits=14, uv_bits=12, normal_bits=10, normal_pred=seed % 3, color_bits=(6, 6, 6, 6),
                          groups=mg.random_groups(m.nface, ng, seed))[0]


def _c1(seed):
    return refshim.encode(mg.grid(185, seed), pos_bits=14, with_uv=False, with_normals=False, with_colors=False)[0]


def _c5(seed):
    return refshim.encode(mg.grid(3163, seed, with_attrs=True), pos_bits=14, normal_bits=10, normal_pred=2, with_uv=False,
                          with_colors=False)[0]


def _tarta(seed):
    return refshim.aligned_blob(open(refshim.TARTA, 'rb').read())


_BUILDERS = dict(c1=_c1, c2=_c2, c3=_c3, c4=_c4, c5=_c5, tarta=_tarta)

DESCRIPTION = dict(
    c1="185^2 grid 34K verts pos14 (bunny stand-in)",
    c2="358^2 grid 128K verts pos14/uv12/normal10-ESTIMATED (Proserpina stand-in)",
    c3="409^2 cloud 167K verts pos14/color6/normal10-DIFF (Nile stand-in)",
    c4="mixed grids 8K-256K verts, all attributes + groups",
    c5="3163^2 grid 10M verts pos14/normal10-BORDER",
    tarta="html/models/tarta.crt (real scan, 1.71M verts / 3.29M faces, pos+uv), replicated",
)


def build(workload, batch, seed0=1, distinct=None, threads=None):
    """Return `batch` blobs of `workload`.  `distinct` (<= batch) limits how many different seeds are encoded; the rest
    are repeats in round-robin order (encoding is the slow part of set-up, not something the benchmark measures)."""
    fn = _BUILDERS[workload]
    if not refshim.available():
        # no reference encoder on this machine: fall back to the committed pre-encoded blob of this workload (seed 1), replicated
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "bench", "%s_seed1.crt" % workload)
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref is not built and there is no pre-encoded fallback for workload " + workload)
        blob = refshim.aligned_blob(open(path, "rb").read())
        return [blob] * batch
    distinct = batch if distinct is None else max(1, min(distinct, batch))
    if workload == 'tarta':
        distinct = 1
    threads = threads or min(32, os.cpu_count() or 1)
    seeds = [seed0 + i for i in range(distinct)]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        uniq = list(ex.map(fn, seeds))
    return [uniq[i % distinct] for i in range(batch)]
