"""CPU tests of the product's host side (no GPU): the C ABI library loads and exports every declared symbol, header
parsing / directory walk / error behaviour mirror the reference, the device cores pass on the CPU (host emulation),
and the multi-rank sharding works over gloo with world_size 2."""
import ctypes as C
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import corto_b200
from tests import cases
from oracle import pyoracle, refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _golden(name):
    return refshim.aligned_blob(open(os.path.join(GOLDEN, name + ".crt"), "rb").read())


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "corto_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", hdr))
    names -= {"defined"}
    assert len(names) > 60
    lib = C.CDLL(corto_b200.LIB_PATH)
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing


def _header_attr_spans(blob):
    """(start, end) byte spans of the attribute entries in a .crt header."""
    b = blob.tobytes()
    p = 9
    nexif = int.from_bytes(b[p:p + 4], "little"); p += 4
    for _ in range(2 * nexif):
        p += 2 + int.from_bytes(b[p:p + 2], "little")
    nattr = int.from_bytes(b[p:p + 4], "little"); p += 4
    spans = []
    for _ in range(nattr):
        s0 = p
        p += 2 + int.from_bytes(b[p:p + 2], "little") + 4 + 4 + 3
        spans.append((s0, p))
    return spans


def swapped_header(blob, i=0, j=1):
    """The same blob with header entries i and j exchanged (the payload stays in the original, sorted, order)."""
    sp = _header_attr_spans(blob)
    b = blob.tobytes()
    (a0, a1), (b0, b1) = sp[i], sp[j]
    out = b[:a0] + b[b0:b1] + b[a1:b0] + b[a0:a1] + b[b1:]
    assert len(out) == len(b)
    return refshim.aligned_blob(out)


def test_header_attributes_follow_std_map_order():
    """The reference visits attributes in std::map (sorted-name) order whatever the header order (decoder.cpp:72-86,168)."""
    blob = _golden("grid_est")
    names = list(corto_b200.Decoder(blob).attributes)
    assert names == sorted(names) and len(names) >= 4
    sw = swapped_header(blob, 0, 2)
    d = corto_b200.Decoder(sw)
    assert list(d.attributes) == names
    for k in names:
        assert d.attributes[k] == corto_b200.Decoder(blob).attributes[k]


def test_no_cpu_fallback():
    """Without a CUDA device decode must fail loudly (CRT_E_CUDA), never fall back to a CPU path."""
    if corto_b200.device_available():
        pytest.skip("a GPU is visible here")
    d = corto_b200.Decoder(_golden("grid_est"))
    with pytest.raises(corto_b200.CortoError) as e:
        d.decode()
    assert e.value.code == -9


def test_product_does_not_touch_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "corto_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
            src = open(path, errors="replace").read()
            assert "liboracle" not in src and "pyoracle" not in src and "refshim" not in src and "libcorto_ref" not in src, path


@pytest.mark.parametrize("name", [c[0] for c in cases.SMALL])
def test_header_matches_reference_ctor(name):
    """crt::Decoder ctor fields (decoder.cpp:41-89): nvert, nface, attribute table in wire order."""
    blob = _golden(name)
    d = corto_b200.Decoder(blob)
    i = pyoracle.info(blob)
    assert (d.nvert, d.nface) == (i["nvert"], i["nface"])
    assert list(d.attributes) == [a["name"] for a in i["attrs"]]
    for a in i["attrs"]:
        mine = d.attributes[a["name"]]
        assert (mine["codec"], mine["N"], mine["strategy"]) == (a["codec"], a["N"], a["strategy"])
        assert np.float32(mine["q"]) == np.float32(a["q"])
    g = d.groups
    assert g and g[-1]["end"] == d.nface or d.nface == 0


def test_groups_and_properties():
    d = corto_b200.Decoder(_golden("groups3"))
    g = d.groups
    assert len(g) == 3 and [x["end"] for x in g] == sorted(x["end"] for x in g) and g[-1]["end"] == d.nface
    assert g[1]["properties"] == {"material": "m1"}      # oracle/ref_shim.cpp gives odd groups a material property


def test_error_behaviour():
    blob = _golden("grid_pos")
    L = corto_b200.lib()
    bad = blob.copy(); bad[0] ^= 0xFF
    assert not L.crt_new_decoder(len(bad), bad.ctypes.data_as(C.c_void_p))
    assert b"Not a crt file" in L.crt_last_error()                           # decoder.cpp:50-51
    raw = np.empty(len(blob) + 8, dtype=np.uint8)
    off = (-raw.ctypes.data) % 4 + 1                                         # misaligned by one
    mis = raw[off:off + len(blob)]; mis[:] = blob
    assert not L.crt_new_decoder(len(mis), mis.ctypes.data_as(C.c_void_p))
    assert b"alignegned" in L.crt_last_error()                               # decoder.cpp:43-44 (sic)


@pytest.mark.parametrize("name", ["grid_est", "cloud_all", "none_entropy", "torus"])
def test_truncated_blobs_are_rejected_not_read_out_of_bounds(name):
    """The reference has no bounds checks; the directory walk must reject every truncation."""
    blob = _golden(name)
    L = corto_b200.lib()
    for cut in list(range(0, 64)) + list(range(64, len(blob) - 1, max(1, len(blob) // 97))):
        part = refshim.aligned_blob(blob[:cut].tobytes())
        ptrs = (C.c_void_p * 1)(part.ctypes.data)
        lens = (C.c_int * 1)(cut)
        h = L.crt_batch_create(1, ptrs, lens)
        assert not h, "truncated at %d of %d accepted" % (cut, len(blob))
    ptrs = (C.c_void_p * 1)(blob.ctypes.data)
    lens = (C.c_int * 1)(len(blob))
    h = L.crt_batch_create(1, ptrs, lens)
    assert h
    L.crt_batch_destroy(h)


@pytest.mark.parametrize("name", ["grid_est", "cloud_all", "groups4_border", "torus"])
def test_mutated_directories_never_escape_the_blob(name):
    """Random byte edits in the header / block headers (sizes, counts, string lengths, nsym ...) through the C ABI: the host walk
    either rejects the blob or builds a batch that answers its queries; the product library must survive all of it (the
    sanitizer run below checks the same walk for out-of-bounds reads byte by byte)."""
    blob = _golden(name)
    L = corto_b200.lib()
    rs = np.random.RandomState(1234)
    n = len(blob)
    accepted = 0
    for trial in range(1500):
        bad = refshim.aligned_blob(blob.tobytes())
        # most edits land in the first 256 bytes (header, group table, first block headers), the rest anywhere
        for _ in range(int(rs.randint(1, 4))):
            pos = int(rs.randint(0, min(n, 256))) if rs.uniform() < 0.7 else int(rs.randint(0, n))
            bad[pos] = rs.randint(0, 256) if rs.uniform() < 0.5 else (0xFF if rs.uniform() < 0.5 else 0x00)
        ptrs = (C.c_void_p * 1)(bad.ctypes.data)
        lens = (C.c_int * 1)(n)
        h = L.crt_batch_create(1, ptrs, lens)
        if h:
            accepted += 1
            nv, nf = C.c_uint32(), C.c_uint32()
            mask = C.c_uint32()
            assert L.crt_batch_mesh_info(h, 0, C.byref(nv), C.byref(nf), C.byref(mask)) == 0
            assert L.crt_batch_total_bytes(h) == n
            L.crt_batch_destroy(h)
    assert accepted > 0            # payload-only edits must still parse (the walk reads no payload byte)


def test_walk_under_sanitizers(tmp_path):
    """crt_walk.cpp under AddressSanitizer + UBSan: every truncation and 3000 random edits per fixture, each in a heap buffer of
    exactly the blob's size (tests/host_emul/walk_fuzz.cpp).  Any out-of-bounds read aborts the run.  The same for the walk REPLAYED
    from a tape (no blob: crt_batch_create_device): the replay equals the direct walk, every truncated tape is refused, 3000 edited
    tapes stay in bounds."""
    exe = str(tmp_path / "walk_fuzz")
    src = [os.path.join(ROOT, "tests", "host_emul", "walk_fuzz.cpp"), os.path.join(ROOT, "corto_b200", "csrc", "crt_walk.cpp")]
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", os.path.join(ROOT, "include"),
                        "-o", exe] + src, capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("sanitizer build not available here: " + r.stderr[-200:])
    for name in ("grid_est", "cloud_all", "groups4_border", "torus", "none_entropy", "const_color", "radius_both"):
        p = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", name + ".crt"), "3000"], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and "accepted" in p.stdout, (name, p.stdout[-300:], p.stderr[-2000:])


def test_batch_layout_without_gpu():
    blobs = [_golden(n) for n in ("grid_est", "cloud_all", "torus", "triangle")]
    bd_ptrs = (C.c_void_p * len(blobs))(*[b.ctypes.data for b in blobs])
    lens = (C.c_int * len(blobs))(*[len(b) for b in blobs])
    L = corto_b200.lib()
    h = L.crt_batch_create(len(blobs), bd_ptrs, lens)
    assert h
    vb, fb = L.crt_batch_vert_base(h), L.crt_batch_face_base(h)
    nv = [pyoracle.info(b)["nvert"] for b in blobs]
    nf = [pyoracle.info(b)["nface"] for b in blobs]
    assert [vb[i] for i in range(5)] == list(np.concatenate([[0], np.cumsum(nv)]))
    assert [fb[i] for i in range(5)] == list(np.concatenate([[0], np.cumsum(nf)]))
    assert L.crt_batch_total_bytes(h) == sum(len(b) for b in blobs)
    L.crt_batch_destroy(h)


# ---- the kernels' sequential cores on the CPU ------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul():
    so = os.path.join(ROOT, "tests", "host_emul", "libhost_emul.so")
    src = [os.path.join(ROOT, "tests", "host_emul", "host_emul.cpp"), os.path.join(ROOT, "corto_b200", "csrc", "crt_walk.cpp")]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so] + src)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name", [c[0] for c in cases.SMALL])
def test_device_cores_on_cpu(emul, name):
    """tun_build_seq vs the oracle's dictionaries for every entropy block; clers_decode_seq and clers_run (the kernel's
    lazy-front machine, with tiny rings so that every reach-back / drain path runs) vs the oracle's faces + predictions."""
    blob = _golden(name)
    out = np.zeros(64 * 5, dtype=np.uint32)
    n = emul.emul_walk_blocks(_p(blob), len(blob), _p(out), 64)
    assert n > 0
    for po, nsym, size, csize, do in out[:n * 5].reshape(n, 5).tolist():
        if nsym <= 1:
            continue
        probs = np.ascontiguousarray(blob[po:po + 2 * nsym])
        i1, l1, t1, u1 = pyoracle.tunstall_tables(probs)
        i2 = np.zeros(256, np.int32); l2 = np.zeros(256, np.int32); t2 = np.zeros(8192, np.uint8)
        emul.emul_tunstall_tables(_p(probs), nsym, _p(i2), _p(l2), _p(t2))
        assert np.array_equal(i1, i2) and np.array_equal(l1, l2) and np.array_equal(t1[:u1], t2[:u1])
    o = pyoracle.decode(blob, debug=True)
    if o["nface"]:
        # (R, Q): 0,0 = clers_decode_seq; R<0 = single-warp machine clers_run; R>0,Q<0 = leader/follower (clers_lead + clers_follow),
        # with EMUL_VEC also the host transcriptions of the kernel's warp-wide window steps (lead_vector / follow_vector / lead_pop_vector)
        for vec in (False, True):
            if vec:
                os.environ["EMUL_VEC"] = "1"
            else:
                os.environ.pop("EMUL_VEC", None)
            for ring in ((0, 0), (-64, 64), (-128, 64), (-4096, 2048), (64, -64), (256, -256), (4096, -2048)):
                faces = np.zeros((o["nface"], 3), np.uint32); pred = np.zeros((o["nvert"], 3), np.uint32)
                rc = emul.emul_clers(_p(blob), len(blob), _p(o["clers"]), len(o["clers"]), _p(faces), _p(pred), ring[0], ring[1])
                assert rc == 0 and np.array_equal(faces, o["index"]) and np.array_equal(pred[1:], o["prediction"][1:]), (name, ring, vec)
        os.environ.pop("EMUL_VEC", None)


@pytest.mark.parametrize("name", [c[0] for c in cases.SMALL])
def test_merged_clers_machine_on_cpu(emul, name):
    """clers_merged (the scalar machine of k_clers_cta) + host transcriptions of its CTA-wide window / pop steps vs the oracle's
    faces + predictions: scalar only and with windows, several window widths and ring sizes (tiny rings force the write-back /
    reach-back paths), run thresholds 2 and 4."""
    blob = _golden(name)
    o = pyoracle.decode(blob, debug=True)
    if not o["nface"]:
        return
    try:
        for vec, W, R, runmin in ((False, 16, 64, 4), (True, 16, 64, 2), (True, 32, 128, 4), (True, 256, 1024, 4), (True, 256, 2048, 2), (True, 8, 32, 3)):
            os.environ.pop("EMUL_VEC", None)
            if vec:
                os.environ["EMUL_VEC"] = "1"
            os.environ["EMUL_W"] = str(W); os.environ["EMUL_RUNMIN"] = str(runmin)
            faces = np.zeros((o["nface"], 3), np.uint32); pred = np.zeros((o["nvert"], 3), np.uint32)
            rc = emul.emul_clers(_p(blob), len(blob), _p(o["clers"]), len(o["clers"]), _p(faces), _p(pred), R, 0)
            assert rc == 0 and np.array_equal(faces, o["index"]) and np.array_equal(pred[1:], o["prediction"][1:]), (name, vec, W, R, runmin)
    finally:
        for k in ("EMUL_VEC", "EMUL_W", "EMUL_RUNMIN"):
            os.environ.pop(k, None)


def _emul_proto2(emul, blob, o, combos):
    try:
        os.environ["EMUL_PROTO"] = "2"
        for W, R in combos:
            os.environ["EMUL_W"] = str(W)
            faces = np.zeros((o["nface"], 3), np.uint32); pred = np.zeros((o["nvert"], 3), np.uint32)
            rc = emul.emul_clers(_p(blob), len(blob), _p(o["clers"]), len(o["clers"]), _p(faces), _p(pred), R, 0)
            assert rc == 0 and np.array_equal(faces, o["index"]) and np.array_equal(pred[1:], o["prediction"][1:]), (W, R, rc)
    finally:
        for k in ("EMUL_PROTO", "EMUL_W"):
            os.environ.pop(k, None)


@pytest.mark.parametrize("name", [c[0] for c in cases.SMALL])
def test_cta_step_protocol_on_cpu(emul, name):
    """The CURRENT step protocol of k_clers_cta (crt_clers_cta.cu, v7.3) transcribed for the host — windows of any width over runs of
    >= 1 symbols that retire the gate on BOUNDARY / DELAY, consecutive prev chains verified in the ring or in the reach-back store,
    pops that consume BOUNDARY / DELAY, one scalar symbol (clers_merged) for the rest, write-back with the kernel's KEEP / ROOM rule —
    against the oracle's faces + predictions, for window widths 8 .. 1024 and rings down to 32 entries."""
    blob = _golden(name)
    o = pyoracle.decode(blob, debug=True)
    if not o["nface"]:
        return
    _emul_proto2(emul, blob, o, ((16, 64), (8, 32), (32, 128), (256, 1024), (256, 2048), (512, 2048), (1024, 4096), (1024, 8192)))


def test_cta_step_protocol_on_cpu_large(emul):
    """The same on a configs[1]-sized grid (256 K symbols, strips of ~700) and on the reference's real scan (3.7 M symbols, runs of ~6,
    DELAY / RIGHT / SPLIT everywhere), with the kernel's own ring sizes."""
    from oracle import refshim
    done = 0
    if refshim.available():
        from oracle import workloads
        blob = workloads._c2(3)
        _emul_proto2(emul, blob, pyoracle.decode(blob, debug=True), ((1024, 4096), (512, 2048), (256, 1024)))
        blob = workloads._c4(10)                     # a grid with a punched hole: DELAY / SPLIT / RIGHT between the strips
        _emul_proto2(emul, blob, pyoracle.decode(blob, debug=True), ((512, 2048), (1024, 8192)))
        done += 1
    if os.path.exists(refshim.TARTA):
        blob = refshim.aligned_blob(open(refshim.TARTA, "rb").read())
        _emul_proto2(emul, blob, pyoracle.decode(blob, debug=True), ((1024, 8192), (256, 2048)))
        done += 1
    if not done:
        pytest.skip("needs oracle/_ref (reference encoder, tarta.crt)")


# ---- multi-rank sharding over gloo, world_size 2 ---------------------------------------------------------------------
def test_shard_lpt_balances():
    rs = np.random.RandomState(1)
    nv = rs.randint(8000, 256000, 512).astype(np.uint32)
    nf = (nv * 2).astype(np.uint32)
    na = np.full(512, 4, dtype=np.uint32)
    for world in (1, 2, 4, 8):
        r = corto_b200.shard_lpt(nv, nf, na, world)
        assert r.min() == 0 and r.max() == world - 1
        load = np.array([(4.0 * nf[r == k] + nv[r == k] * 4.0).sum() for k in range(world)])
        assert load.max() / load.mean() < 1.02


def test_walk_tapes_rebuild_the_directory():
    """crt_walk_tape + crt_batch_create_device: the directory rebuilt from the tapes (no blob byte available) is the one the host
    walk finds, for every fixture; a truncated tape is rejected."""
    import glob
    L = corto_b200.lib()
    paths = sorted(glob.glob(os.path.join(GOLDEN, "*.crt")))
    blobs = [np.frombuffer(open(p, "rb").read(), dtype=np.uint8) for p in paths]
    al = [corto_b200._aligned_copy(b) for b in blobs]
    n = len(al)
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in al])
    lens = (C.c_int * n)(*[len(b) for b in al])
    h0 = L.crt_batch_create(n, ptrs, lens)
    assert h0
    tapes = [corto_b200.walk_tape(b)[0] for b in blobs]
    assert max(len(t) for t in tapes) < 2048          # header + group table + block headers, never payload
    tp = (C.c_void_p * n)(*[t.ctypes.data for t in tapes])
    tl = (C.c_int * n)(*[len(t) for t in tapes])
    fake_arena = C.c_void_p(1 << 20)                 # never dereferenced by create (16-byte aligned device address stand-in)
    h1 = L.crt_batch_create_device(n, tp, tl, lens, fake_arena)
    assert h1, L.crt_last_error()
    assert L.crt_batch_directory_signature(h0) == L.crt_batch_directory_signature(h1)
    assert L.crt_batch_total_verts(h0) == L.crt_batch_total_verts(h1) and L.crt_batch_total_faces(h0) == L.crt_batch_total_faces(h1)
    L.crt_batch_destroy(h1)
    # a tape cut short: the replay must fail cleanly
    short = tapes[3][:len(tapes[3]) // 2].copy()
    tp2 = (C.c_void_p * 1)(short.ctypes.data)
    assert not L.crt_batch_create_device(1, tp2, (C.c_int * 1)(len(short)), (C.c_int * 1)(len(al[3])), fake_arena)
    L.crt_batch_destroy(h0)


def test_scatter_two_ranks_gloo(tmp_path):
    """corto_b200.dist: rank 0 holds the blobs, LPT-shards them from their walk tapes, ONE grouped send/recv delivers every rank's
    bin as a contiguous arena; each rank rebuilds its directory from the tapes (crt_batch_create_device)."""
    script = tmp_path / "w.py"
    script.write_text('''
import os, sys, glob, hashlib, ctypes as C
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import corto_b200
from corto_b200 import dist as cd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
paths = sorted(glob.glob(os.path.join(%r, "*.crt")))
ing = None
if rank == 0:
    blobs = [np.frombuffer(open(p, "rb").read(), dtype=np.uint8) for p in paths]
    ing = cd.Ingest(blobs, world, device="cpu")
    assert 1.0 <= ing.load_max_over_mean < 1.5
got = cd.scatter_blobs(ing, src=0, device="cpu")
ids, lens, tapes, arena = got["ids"], got["lens"], got["tapes"], got["arena"].numpy()
all_ids = [None] * world
dist.all_gather_object(all_ids, ids)
flat = sorted(i for x in all_ids for i in x)
assert flat == list(range(len(paths))), flat
o = 0
for i, n in zip(ids, lens):
    assert hashlib.sha1(arena[o:o + n].tobytes()).hexdigest() == hashlib.sha1(open(paths[i], "rb").read()).hexdigest()
    assert o %% 16 == 0
    o += (n + 15) // 16 * 16
assert len(ids) > 0
# the directory from the tapes == the directory from the bytes that arrived
L = corto_b200.lib()
n = len(ids)
tp = (C.c_void_p * n)(*[t.ctypes.data for t in tapes]); tl = (C.c_int * n)(*[len(t) for t in tapes]); bl = (C.c_int * n)(*lens)
h1 = L.crt_batch_create_device(n, tp, tl, bl, C.c_void_p(1 << 20))
assert h1, L.crt_last_error()
local = [corto_b200._aligned_copy(np.frombuffer(open(paths[i], "rb").read(), dtype=np.uint8)) for i in ids]
h0 = L.crt_batch_create(n, (C.c_void_p * n)(*[b.ctypes.data for b in local]), bl)
assert L.crt_batch_directory_signature(h0) == L.crt_batch_directory_signature(h1)
# gather_rows: every rank contributes rank+1 rows
rows = [r + 1 for r in range(world)]
x = torch.full((rows[rank], 3), float(rank))
g = cd.gather_rows(x, rows, dst=0)
if rank == 0:
    assert g.shape == (sum(rows), 3) and float(g[-1, 0]) == world - 1 and float(g[0, 0]) == 0.0
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok", n)
''' % (ROOT, GOLDEN))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
