"""GPU parity at the BASELINE configs' own sizes and launch configurations (`-m gpu`), against the unmodified reference:

  * configs[3]: a 640-mesh batch of mixed meshes (8K..256K vertices, all attributes, 1-4 groups, holes, two components) — more
    than 4 x SMs meshes, i.e. the small-ring launch configuration of both CLERS kernels (k_clers_cta R = 2048, k_clers_lf RB = 1024);
  * configs[4]: one 10 M-vertex mesh with BORDER normals, every array compared by FNV digest;
  * tarta.crt x 4 in one batch;
  * generic attributes dequantised to INT32 / UINT32 (vertex_attribute.h:200-203, 220-223);
  * the reference shims' own entry points (newDecoder ... decode, CreateDecoder / DecodeMesh) called through ctypes.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import corto_b200
from oracle import pyoracle, refshim

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

C4_SCRIPT = r'''
import sys, os
sys.path.insert(0, %(root)r)
import numpy as np, torch
import corto_b200
from oracle import refshim, workloads, meshgen as mg
# distinct blobs: the configs[3] generator + the extremes it names (256 K vertices, 4 groups, a hole, two components, every prediction)
distinct = [workloads._c4(s) for s in (3, 4, 5, 7, 11, 12, 13, 20)]
big = mg.punch_hole(mg.grid(506, 99), 506)
distinct.append(refshim.encode(big, pos_bits=14, uv_bits=12, normal_bits=10, normal_pred=2, color_bits=(6, 6, 6, 6), groups=mg.random_groups(big.nface, 4, 99))[0])
full = mg.grid(506, 97)
distinct.append(refshim.encode(full, pos_bits=14, uv_bits=12, normal_bits=10, normal_pred=0, color_bits=(6, 6, 6, 6), groups=mg.random_groups(full.nface, 4, 97))[0])
two = mg.two_components(430, 98)
distinct.append(refshim.encode(two, pos_bits=14, uv_bits=12, normal_bits=10, normal_pred=1, color_bits=(6, 6, 6, 6), groups=mg.random_groups(two.nface, 3, 98))[0])
want = [refshim.decode(b, color_out=4) for b in distinct]
assert max(w["nvert"] for w in want) > 250000
n = 640
order = [i %% len(distinct) for i in range(n)]
bd = corto_b200.BatchDecoder([distinct[k] for k in order], color_components=4)
bd.allocate(fill=0xA5)
bd.upload(); bd.decode()
torch.cuda.synchronize()
rc, st = bd.status()
assert rc == 0, st
dev = [{k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).cuda() for k, v in w.items() if isinstance(v, np.ndarray)} for w in want]
for i, k in enumerate(order):
    v0, v1, f0, f1 = bd.vert_base[i], bd.vert_base[i + 1], bd.face_base[i], bd.face_base[i + 1]
    for name, w in dev[k].items():
        got = bd.out[name][f0:f1] if name == "index" else bd.out[name][v0:v1]
        assert torch.equal(got.view(torch.uint8).reshape(-1), w.view(torch.uint8).reshape(-1)), (i, k, name)
print("ok", n, int(bd.total_verts))
'''


@pytest.mark.timeout(900)
@pytest.mark.parametrize("clers", ["default", "3", "4"])
def test_c4_batch_640(clers, tmp_path):
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    script = tmp_path / "c4.py"
    script.write_text(C4_SCRIPT % dict(root=ROOT))
    env = dict(os.environ)
    if clers != "default":
        env["CORTO_CLERS"] = clers
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=850)
    assert r.returncode == 0 and "ok 640" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.timeout(900)
def test_c5_one_10m_vertex_mesh():
    """configs[4]: 3163^2 grid, pos14 + normal10 BORDER; digests of every array vs the reference decode of the same blob."""
    if not refshim.available():
        pytest.skip("needs the reference shim to encode")
    from oracle import workloads
    blob = workloads._c5(1)
    want = refshim.decode(blob)
    assert want["nvert"] > 10_000_000
    got = corto_b200.Decoder(blob).decode()
    for k, w in want.items():
        if isinstance(w, np.ndarray):
            assert got[k].shape == w.shape
            assert pyoracle.fnv1a64(got[k]) == pyoracle.fnv1a64(w), k
    # (the index has to stay bound: the reference's decodeFaces writes through faces32 / faces16 unconditionally)
    want16 = refshim.decode(blob, normals16=True, bind=["position", "normal", "index"])
    got16 = corto_b200.Decoder(blob).decode(normals16=True, bind=["position", "normal", "index"])
    assert pyoracle.fnv1a64(got16["normal"]) == pyoracle.fnv1a64(want16["normal"])


@pytest.mark.timeout(600)
def test_tarta_batch_of_4():
    import torch
    if not os.path.exists(refshim.TARTA):
        pytest.skip("tarta.crt fixture not present")
    blob = refshim.aligned_blob(open(refshim.TARTA, "rb").read())
    want = pyoracle.decode(blob)
    bd = corto_b200.BatchDecoder([blob] * 4)
    bd.allocate(fill=0xA5)
    bd.upload(); bd.decode()
    torch.cuda.synchronize()
    rc, st = bd.status()
    assert rc == 0, st
    for i in range(4):
        got = bd.mesh_outputs(i)
        for k, w in want.items():
            if isinstance(w, np.ndarray):
                assert np.array_equal(got[k].view(np.uint8).reshape(-1), w.view(np.uint8).reshape(-1)), (i, k)


@pytest.mark.timeout(600)
def test_regular_and_irregular_meshes_in_one_batch():
    """The default CLERS launch splits a batch by a sample of each stream: regular meshes stay on k_clers_cta, irregular ones (the
    real scan) go to k_clers_lf behind it.  Both kinds in one batch, twice (the deferral flags live in the zeroed control region)."""
    import torch
    from oracle import workloads
    if not os.path.exists(refshim.TARTA) or not refshim.available():
        pytest.skip("needs tarta.crt and the reference encoder")
    tarta = refshim.aligned_blob(open(refshim.TARTA, "rb").read())
    grid = workloads._c2(7)
    blobs = [tarta, grid, tarta, workloads._c2(8)]
    want = [pyoracle.decode(b) for b in (tarta, grid, tarta, blobs[3])]
    bd = corto_b200.BatchDecoder(blobs)
    bd.allocate(fill=0xA5)
    bd.upload()
    for _ in range(2):
        bd.decode()
        torch.cuda.synchronize()
        rc, st = bd.status()
        assert rc == 0, st
        for i, w in enumerate(want):
            got = bd.mesh_outputs(i)
            for k, x in w.items():
                if isinstance(x, np.ndarray):
                    assert np.array_equal(got[k].view(np.uint8).reshape(-1), x.view(np.uint8).reshape(-1)), (i, k)
        bd.rewalk()


@pytest.mark.parametrize("name,formats", [
    ("grid_diff", {"position": corto_b200.INT32, "uv": corto_b200.UINT32, "radius": corto_b200.INT32}),
    ("cloud_all", {"position": corto_b200.UINT32, "uv": corto_b200.INT32, "radius": corto_b200.UINT32}),
    ("radius_both", {"position": corto_b200.FLOAT, "radius": corto_b200.UINT32}),
])
def test_generic_integer_formats(name, formats):
    """GenericAttr::dequantize INT32 / UINT32: `u32 *= q` (vertex_attribute.h:200-203, 220-223) — mesh path (k_dequant) and the
    fused point-cloud path."""
    blob = refshim.aligned_blob(open(os.path.join(GOLDEN, name + ".crt"), "rb").read())
    want = pyoracle.decode(blob, formats=formats, bind=list(formats))
    got = corto_b200.Decoder(blob).decode(formats=formats, bind=list(formats) + ["index"])
    for k in formats:
        assert np.array_equal(got[k].view(np.uint32).reshape(-1), want[k].view(np.uint32).reshape(-1)), k
    if refshim.available():
        ref = refshim.decode_formats(blob, formats)
        for k in formats:
            assert np.array_equal(got[k].view(np.uint32).reshape(-1), ref[k].view(np.uint32).reshape(-1)), k + "(ref)"


def _raw_lib():
    L = C.CDLL(corto_b200.LIB_PATH)
    vp, ci = C.c_void_p, C.c_int
    L.newDecoder.restype = vp; L.newDecoder.argtypes = [ci, vp]
    for f in ("deleteDecoder", "decode"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = None
    for f in ("nvert", "nface", "ngroups", "hasNormal", "hasColor", "hasUv"):
        getattr(L, f).argtypes = [vp]; getattr(L, f).restype = ci
    L.hasAttr.argtypes = [vp, C.c_char_p]; L.hasAttr.restype = ci
    L.groups.argtypes = [vp, vp]
    for f in ("setPositions", "setNormals32", "setNormals16", "setUvs", "setIndex16", "setIndex32"):
        getattr(L, f).argtypes = [vp, vp]; getattr(L, f).restype = None
    L.setColors.argtypes = [vp, vp, ci]; L.setColors.restype = None
    L.CreateDecoder.restype = vp; L.CreateDecoder.argtypes = [ci, vp, vp]
    L.DestroyDecoder.argtypes = [vp]; L.DestroyDecoder.restype = None
    L.DecodeMesh.restype = ci; L.DecodeMesh.argtypes = [vp, vp, vp, vp, vp, vp]
    return L


@pytest.mark.parametrize("name", ["grid_est", "groups3", "torus", "cloud_all"])
def test_reference_named_wasm_shim(name):
    """newDecoder / set* / decode / deleteDecoder exactly as html/js/emscripten/post.js drives them (emcorto.cpp:14-89)."""
    L = _raw_lib()
    blob = refshim.aligned_blob(open(os.path.join(GOLDEN, name + ".crt"), "rb").read())
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = L.newDecoder(len(blob), blob.ctypes.data)
    assert d
    nv, nf = L.nvert(d), L.nface(d)
    pos = np.zeros((nv, 3), np.float32); L.setPositions(d, pos.ctypes.data)
    nrm = np.zeros((nv, 3), np.float32); col = np.zeros((nv, 4), np.uint8); uv = np.zeros((nv, 2), np.float32)
    idx = np.zeros((nf, 3), np.uint32)
    if L.hasNormal(d): L.setNormals32(d, nrm.ctypes.data)
    if L.hasColor(d): L.setColors(d, col.ctypes.data, 4)
    if L.hasUv(d): L.setUvs(d, uv.ctypes.data)
    if nf: L.setIndex32(d, idx.ctypes.data)
    L.decode(d)
    ng = L.ngroups(d)
    ends = np.zeros(max(ng, 1), np.int32); L.groups(d, ends.ctypes.data)
    assert np.array_equal(pos.view(np.uint32), gold["default/position"].view(np.uint32))
    if L.hasNormal(d): assert np.array_equal(nrm.view(np.uint32), gold["default/normal"].view(np.uint32))
    if L.hasColor(d): assert np.array_equal(col, gold["color_out4/color"])
    if L.hasUv(d): assert np.array_equal(uv.view(np.uint32), gold["default/uv"].view(np.uint32))
    if nf:
        assert np.array_equal(idx, gold["default/index"])
        assert ng >= 1 and ends[ng - 1] == nf
    assert L.hasAttr(d, b"position") == 1 and L.hasAttr(d, b"nosuch") == 0
    L.deleteDecoder(d)
    # int16 normals + u16 index through the same shim
    if "index161_normals161/normal" in gold.files and nv < 65536:
        d = L.newDecoder(len(blob), blob.ctypes.data)
        pos = np.zeros((nv, 3), np.float32); L.setPositions(d, pos.ctypes.data)
        n16 = np.full((nv, 3), 0xA5A5, np.uint16).view(np.int16); L.setNormals16(d, n16.ctypes.data)
        i16 = np.zeros((nf, 3), np.uint16)
        if nf: L.setIndex16(d, i16.ctypes.data)
        L.decode(d)
        assert np.array_equal(n16, gold["index161_normals161/normal"])
        if nf: assert np.array_equal(i16, gold["index161_normals161/index"])
        L.deleteDecoder(d)


@pytest.mark.parametrize("name", ["grid_est", "groups3", "cloud_all"])
def test_reference_named_unity_shim(name):
    """CreateDecoder / DecodeMesh / DestroyDecoder as unity/CortoMeshLoader.cs P/Invokes them (corto_codec.cpp:6-57): info =
    (nface, nvert); point clouds return -1; colours come out as float r,g,b,a = u8 / 255."""
    L = _raw_lib()
    blob = refshim.aligned_blob(open(os.path.join(GOLDEN, name + ".crt"), "rb").read())
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    info = np.zeros(2, np.float32)
    d = L.CreateDecoder(len(blob), blob.ctypes.data, info.ctypes.data)
    assert d
    nf, nv = int(info[0]), int(info[1])
    pos = np.zeros((nv, 3), np.float32); nrm = np.zeros((nv, 3), np.float32); col = np.zeros((nv, 4), np.float32); uv = np.zeros((nv, 2), np.float32)
    idx = np.zeros((max(nf, 1), 3), np.int32)
    rc = L.DecodeMesh(d, pos.ctypes.data, idx.ctypes.data, nrm.ctypes.data, col.ctypes.data, uv.ctypes.data)
    if nf == 0:
        assert rc == -1
    else:
        assert rc == nf
        assert np.array_equal(pos.view(np.uint32), gold["default/position"].view(np.uint32))
        assert np.array_equal(idx.view(np.uint32)[:nf], gold["default/index"])
        if "default/normal" in gold.files: assert np.array_equal(nrm.view(np.uint32), gold["default/normal"].view(np.uint32))
        if "default/uv" in gold.files: assert np.array_equal(uv.view(np.uint32), gold["default/uv"].view(np.uint32))
        if "color_out4/color" in gold.files: assert np.array_equal(col, gold["color_out4/color"].astype(np.float32) / np.float32(255.0))
    L.DestroyDecoder(d)


def test_unsorted_header_decodes_in_std_map_order():
    """Header entries out of order (hand-built file): the reference walks its std::map in name order (decoder.cpp:168)."""
    from tests.test_host import swapped_header
    blob = refshim.aligned_blob(open(os.path.join(GOLDEN, "grid_est.crt"), "rb").read())
    gold = np.load(os.path.join(GOLDEN, "grid_est.npz"))
    for i, j in ((0, 2), (1, 3)):
        got = corto_b200.Decoder(swapped_header(blob, i, j)).decode()
        for k in ("position", "normal", "color", "uv", "radius", "index"):
            assert np.array_equal(got[k].view(np.uint8).reshape(-1), gold["default/" + k].view(np.uint8).reshape(-1)), k
