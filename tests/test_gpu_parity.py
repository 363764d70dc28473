"""GPU parity: CUDA path (through the C ABI) vs the oracle, bit for bit, on every fixture category.

`-m gpu`.  Ground truth = oracle/liboracle.so (pinned to the reference by tests/test_oracle.py) and, when the prebuilt
reference shim travelled with the snapshot, the unmodified reference itself.
"""
import glob
import os

import numpy as np
import pytest

import corto_b200
from tests import cases
from oracle import pyoracle, refshim

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same(name, a, b):
    assert a.shape == b.shape and a.dtype == b.dtype, (name, a.shape, b.shape, a.dtype, b.dtype)
    av, bv = a.reshape(a.shape[0], -1).view(np.uint8), b.reshape(b.shape[0], -1).view(np.uint8)
    if not np.array_equal(av, bv):
        bad = np.nonzero((av != bv).any(1))[0]
        raise AssertionError("%s: %d of %d rows differ, first %d: got %s want %s" % (name, len(bad), a.shape[0], bad[0], a[bad[0]], b[bad[0]]))


def _check_blob(blob, tag):
    info = pyoracle.info(blob)
    for var in cases.VARIANTS:
        if not cases.applicable(var, info["attrs"], info["nvert"], info["nface"]):
            continue
        want = pyoracle.decode(blob, **var)
        kw = dict(index16=var.get("index16", False), normals16=var.get("normals16", False), color_components=var.get("color_out"))
        got = corto_b200.Decoder(blob).decode(**kw)
        for k, w in want.items():
            if isinstance(w, np.ndarray):
                _same("%s%s/%s" % (tag, var, k), got[k], w)
        if refshim.available():
            r = refshim.decode(blob, **var)
            for k, w in r.items():
                if isinstance(w, np.ndarray):
                    _same("%s%s/%s(ref)" % (tag, var, k), got[k], w)


@pytest.mark.parametrize("name,builder", cases.SMALL, ids=[c[0] for c in cases.SMALL])
def test_small(name, builder):
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    _check_blob(builder(), name)


@pytest.mark.parametrize("name,builder", cases.MEDIUM, ids=[c[0] for c in cases.MEDIUM])
def test_medium(name, builder):
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    _check_blob(builder(), name)


@pytest.mark.parametrize("n_lat,n_lon", [(10, 12), (8, 60), (6, 300), (4, 3000)])
@pytest.mark.parametrize("pred", [1, 2])
def test_high_valence_normals(n_lat, n_lon, pred):
    """ESTIMATED / BORDER normals on UV spheres whose poles have valence n_lon: 12 and 60 go through the per-vertex overflow chains
    (valence > 8 slots), 300 and 3000 through the in-order face scan of a fan pole — all bit-exact, and fast (the first version of
    this path rescanned the whole overflow list per incident face: quadratic)."""
    import time
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    from oracle import meshgen as mg
    blob = refshim.encode(mg.sphere(n_lat, n_lon, 3), pos_bits=14, normal_bits=10, normal_pred=pred, with_uv=False, with_colors=False)[0]
    want = refshim.decode(blob)
    corto_b200.Decoder(blob).decode()                          # warm-up (context, allocations)
    t0 = time.perf_counter()
    got = corto_b200.Decoder(blob).decode()
    dt = time.perf_counter() - t0
    for k, w in want.items():
        if isinstance(w, np.ndarray):
            _same("sphere %dx%d pred %d/%s" % (n_lat, n_lon, pred, k), got[k], w)
    assert dt < 0.5, dt
    want16 = refshim.decode(blob, normals16=True)
    got16 = corto_b200.Decoder(blob).decode(normals16=True)
    _same("sphere int16 normals", got16["normal"], want16["normal"])


def test_batch_from_device_arena():
    """crt_batch_create_device: the blobs exist in DEVICE memory only (one arena, as they arrive from the ingest rank over NVLink),
    the directory comes from their walk tapes; decode, re-walk + decode again — every array equals the oracle's."""
    import torch
    blobs = [np.frombuffer(open(p, "rb").read(), dtype=np.uint8) for p in sorted(glob.glob(os.path.join(GOLDEN, "*.crt")))]
    blobs = [b for b in blobs if not [a for a in pyoracle.info(corto_b200._aligned_copy(b))["attrs"] if a["name"] == "radius"]]
    assert len(blobs) > 20
    tapes = [corto_b200.walk_tape(b)[0] for b in blobs]
    lens = [len(b) for b in blobs]
    host = np.zeros(sum((n + 15) // 16 * 16 for n in lens) + 16, dtype=np.uint8)
    o = 0
    for b in blobs:
        host[o:o + len(b)] = b
        o += (len(b) + 15) // 16 * 16
    arena = torch.from_numpy(host).cuda()
    del host
    bd = corto_b200.BatchDecoder.from_device(tapes, lens, arena, color_components=4)
    bd.allocate(fill=0xA5)
    bd.upload()
    for _ in range(2):
        bd.decode()
        torch.cuda.synchronize()
        rc, st = bd.status()
        assert rc == 0, st
        for i, blob in enumerate(blobs):
            want = pyoracle.decode(corto_b200._aligned_copy(blob), color_out=4)
            got = bd.mesh_outputs(i)
            for k, w in want.items():
                if isinstance(w, np.ndarray):
                    _same("device-arena[%d]/%s" % (i, k), got[k], w)
        bd.rewalk()


def test_batch_mixed():
    """All small + medium blobs in ONE device-resident batch; every mesh's slices must match the oracle."""
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    import torch
    blobs = [b() for _, b in cases.SMALL + cases.MEDIUM]
    blobs = [b for b in blobs if len([a for a in pyoracle.info(b)["attrs"] if a["name"] == "radius"]) == 0]
    bd = corto_b200.BatchDecoder(blobs, color_components=4)
    bd.allocate(fill=0xA5)
    bd.upload()
    bd.decode()
    torch.cuda.synchronize()
    rc, st = bd.status()
    assert rc == 0, st
    for i, blob in enumerate(blobs):
        want = pyoracle.decode(blob, color_out=4)
        got = bd.mesh_outputs(i)
        for k, w in want.items():
            if isinstance(w, np.ndarray):
                _same("batch[%d]/%s" % (i, k), got[k], w)
    # second decode of the same batch object must give the same answer (scratch / tickets reset correctly)
    bd.decode()
    torch.cuda.synchronize()
    for i in (0, len(blobs) - 1):
        want = pyoracle.decode(blobs[i], color_out=4)
        got = bd.mesh_outputs(i)
        for k, w in want.items():
            if isinstance(w, np.ndarray):
                _same("batch2[%d]/%s" % (i, k), got[k], w)


def test_batch_baseline_sizes():
    """BASELINE configs[1] / configs[2] at their real mesh sizes in one batch large enough for the big-batch code paths bench.py
    times (one CTA per unpack chain at 6 CTAs per SM, one warp per (mesh, attribute) in the delta inverse): 2 distinct 128 K-vertex
    meshes x 128 (= the 256 meshes of configs[1]) and 2 distinct 167 K-point clouds x 16, every copy compared with the oracle's
    decode of its blob."""
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    import torch
    from oracle import workloads
    distinct = [workloads._c2(1), workloads._c2(2), workloads._c3(1), workloads._c3(2)]
    want = [pyoracle.decode(b, color_out=4) for b in distinct]
    order = ([0, 1] * 128) + ([2, 3] * 16)
    bd = corto_b200.BatchDecoder([distinct[k] for k in order], color_components=4)
    bd.allocate(fill=0xA5)
    bd.upload()
    bd.decode()
    torch.cuda.synchronize()
    rc, st = bd.status()
    assert rc == 0, st
    for i, k in enumerate(order):
        got = bd.mesh_outputs(i)
        for name, w in want[k].items():
            if isinstance(w, np.ndarray):
                _same("baseline[%d]/%s" % (i, name), got[name], w)


def test_tarta():
    import os
    if not os.path.exists(refshim.TARTA):
        pytest.skip("tarta.crt fixture not present")
    blob = refshim.aligned_blob(open(refshim.TARTA, "rb").read())
    want = pyoracle.decode(blob)
    got = corto_b200.Decoder(blob).decode()
    for k, w in want.items():
        if isinstance(w, np.ndarray):
            _same("tarta/" + k, got[k], w)


@pytest.mark.parametrize("mode", ["1", "2", "3", "4", "split0", "split1", "deltacta", "deltaseq", "deltawarp", "tunseq", "unpackchain", "overlap1", "overlap2", "overlap3", "adjagg1"])
def test_alternative_clers_machines(mode, tmp_path):
    """CORTO_CLERS=1 (single-warp lazy-front machine), =2 / =3 (leader/follower without / with window steps for every mesh) and =4
    (the CTA machine for every mesh, irregular ones included; the default hands those to the leader/follower kernel) stay
    bit-exact: they are the A/B baselines DESIGN.md section 5 quotes, selected once per process by the environment."""
    import os
    import subprocess
    import sys
    script = tmp_path / "alt.py"
    script.write_text('''
import sys, os, glob
sys.path.insert(0, %r)
import numpy as np, corto_b200
from oracle import pyoracle, refshim
for path in sorted(glob.glob(os.path.join(%r, "*.crt"))):
    blob = refshim.aligned_blob(open(path, "rb").read())
    want = pyoracle.decode(blob); got = corto_b200.Decoder(blob).decode()
    for k, w in want.items():
        if isinstance(w, np.ndarray):
            assert np.array_equal(got[k].view(np.uint8).reshape(-1), w.view(np.uint8).reshape(-1)), (path, k)
print("ok")
''' % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")))
    # split0 / split1: the delta inverse with one warp per (mesh, attribute) / per component (the host picks by batch size);
    # overlap1: the side-stream stage overlap (CORTO_OVERLAP, off by default)
    extra = {"CORTO_CLERS": mode}
    if mode.startswith("split"):                  # the warp-per-chain delta kernel, per (mesh, attribute) / per component
        extra = {"CORTO_DELTA_SPLIT": mode[-1]}
    elif mode == "deltacta":                      # the block-wide delta kernel (sub-blocks + pointer doubling through shared memory)
        extra = {"CORTO_DELTA": "cta"}
    elif mode == "deltawarp":                     # the warp-per-chain delta kernel for every mesh (default: segmented-scan rounds for regular meshes)
        extra = {"CORTO_DELTA": "warp"}
    elif mode == "tunseq":                        # the one-thread Tunstall dictionary build (default: warp-cooperative)
        extra = {"CORTO_TUN": "seq"}
    elif mode == "deltaseq":                      # ... with every round on its sequential-warp path
        extra = {"CORTO_DELTA": "seq"}
    elif mode == "unpackchain":                   # one CTA per unpack chain (the default only for batches with >= 2 x SMs chains)
        extra = {"CORTO_UNPACK": "chain"}
    elif mode.startswith("overlap"):
        extra = {"CORTO_OVERLAP": mode[-1]}
    elif mode.startswith("adjagg"):
        extra = {"CORTO_ADJ_AGG": mode[-1]}       # warp-aggregated adjacency atomics (off by default)
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, **extra), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
