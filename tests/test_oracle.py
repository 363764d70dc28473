"""CPU tests (no GPU): the oracle is pinned to the reference's own outputs.

 * golden vectors: tests/golden/*.crt + *.npz were produced by the UNMODIFIED reference Encoder/Decoder
   (tests/golden/make_golden.py); the C restatement must reproduce every array bit for bit;
 * where the reference shim was built (oracle/_ref), the same on freshly encoded blobs incl. multi-tile sizes and the
   reference's only shipped .crt (html/models/tarta.crt) against committed FNV digests.
"""
import glob
import json
import os

import numpy as np
import pytest

from tests import cases
from oracle import pyoracle, refshim

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _aligned(raw):
    return refshim.aligned_blob(raw)


def _vkey(var):
    return "_".join("%s%s" % (k, int(v)) for k, v in sorted(var.items())) or "default"


def _eq(a, b):
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.reshape(-1).view(np.uint8), b.reshape(-1).view(np.uint8))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.crt"))), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_matches_golden(path):
    blob = _aligned(open(path, "rb").read())
    gold = np.load(path[:-4] + ".npz")
    info = pyoracle.info(blob)
    checked = 0
    for var in cases.VARIANTS:
        if not cases.applicable(var, info["attrs"], info["nvert"], info["nface"]):
            continue
        out = pyoracle.decode(blob, **var)
        for k, v in out.items():
            if isinstance(v, np.ndarray):
                key = _vkey(var) + "/" + k
                assert key in gold.files, key
                assert _eq(v, gold[key]), (os.path.basename(path), key)
                checked += 1
    assert checked > 0


def test_golden_set_is_complete():
    names = {os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.crt"))}
    assert names == {c[0] for c in cases.SMALL}


@pytest.mark.parametrize("name,builder", cases.SMALL + cases.MEDIUM, ids=[c[0] for c in cases.SMALL + cases.MEDIUM])
def test_oracle_matches_reference(ref, name, builder):
    blob = builder()
    info = pyoracle.info(blob)
    for var in cases.VARIANTS:
        if not cases.applicable(var, info["attrs"], info["nvert"], info["nface"]):
            continue
        r = ref.decode(blob, debug=True, **var)
        o = pyoracle.decode(blob, debug=True, **var)
        for k, v in r.items():
            if isinstance(v, np.ndarray):
                if k == "prediction":
                    assert _eq(v[1:], o[k][1:]), (name, var, k)     # prediction[0] is never read
                else:
                    assert _eq(v, o[k]), (name, var, k)


def test_colour_4_to_3_self_corruption(ref):
    """N=4 colours decoded with out_components=3 overwrite unread inputs in the reference's in-place loop.
    The oracle restates that loop, so it reproduces the corruption exactly."""
    blob = cases.SMALL[1][1]()
    r = ref.decode(blob, color_out=3)
    o = pyoracle.decode(blob, color_out=3)
    assert _eq(r["color"], o["color"])


def test_tarta_digests(ref):
    if not os.path.exists(ref.TARTA):
        pytest.skip("tarta.crt not copied to oracle/_ref")
    want = json.load(open(os.path.join(GOLDEN, "tarta.json")))
    blob = _aligned(open(ref.TARTA, "rb").read())
    out = pyoracle.decode(blob)
    assert out["nvert"] == want["nvert"] and out["nface"] == want["nface"]
    for k in ("position", "uv", "index"):
        assert "%016x" % pyoracle.fnv1a64(out[k]) == want[k], k


def test_tunstall_table_properties():
    """Dictionary invariants on random probability tables: 256 words, offsets inside the text, low-entropy branch."""
    rs = np.random.RandomState(0)
    for trial in range(200):
        n = int(rs.randint(2, 40))
        p = np.sort(rs.randint(0, 256, n))[::-1].astype(np.uint8)
        if trial % 5 == 0:
            p[0] = 250; p[1:] = rs.randint(0, 3, n - 1)     # low entropy -> run >= 16 branch (tunstall.cpp:149)
        syms = rs.permutation(256)[:n].astype(np.uint8)
        probs = np.stack([syms, p], 1).reshape(-1)
        idx, ln, tab, used = pyoracle.tunstall_tables(probs)
        assert used <= 8192 and (ln >= 1).all() and ((idx + ln) <= used).all()
        words = {bytes(tab[i:i + l]) for i, l in zip(idx, ln)}
        assert len(words) >= 2


def test_unsorted_header_follows_std_map_order():
    """A hand-built header with attribute entries out of order: the reference decodes in std::map (sorted-name) order
    (decoder.cpp:72-86,168), so the arrays equal those of the sorted original; the restatement does the same."""
    if not refshim.available():
        pytest.skip("needs the reference shim")
    from tests.test_host import swapped_header
    blob = refshim.aligned_blob(open(os.path.join(GOLDEN, "grid_est.crt"), "rb").read())
    want = refshim.decode(blob)
    for i, j in ((0, 2), (1, 3), (0, 4)):
        sw = swapped_header(blob, i, j)
        ref = refshim.decode(sw)
        got = pyoracle.decode(sw)
        for k, w in want.items():
            if isinstance(w, np.ndarray):
                assert np.array_equal(ref[k].view(np.uint8), w.view(np.uint8)), k
                assert np.array_equal(got[k].view(np.uint8), w.view(np.uint8)), k
