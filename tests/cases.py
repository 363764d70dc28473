"""Fixture categories for the parity tests (SURVEY §8c list), built with the reference Encoder through oracle.refshim.

Each case is (name, builder) where builder() -> blob (uint8 ndarray, 16-byte aligned).  `VARIANTS` are the decode-side
bindings every blob is decoded with.  tests/golden/make_golden.py freezes the SMALL cases into committed fixtures so
they can be checked where the reference shim is absent.
"""
import numpy as np
from oracle import meshgen as mg


def _enc(mesh, **kw):
    from oracle import refshim
    return refshim.encode(mesh, **kw)[0]


def _const_color(m, rgba=(200, 100, 50, 255)):
    m.colors = np.tile(np.array(rgba, dtype=np.uint8), (m.nvert, 1))
    return m


def _flat(m):
    m.pos[:, 2] = 0.0
    return m


def _far(m, off=1.0e6):
    m.pos += np.float32(off)          # large quantised coordinates -> big logs, fp32 rounding in normal estimation
    return m


def _triangle():
    return mg.Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float32), np.array([[0, 1, 2]], dtype=np.uint32),
                   np.array([[0, 0, 1]] * 3, dtype=np.float32), np.array([[0, 0], [1, 0], [0, 1]], dtype=np.float32),
                   np.array([[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255]], dtype=np.uint8), np.ones(3, dtype=np.float32))


def _tiny_cloud(n):
    rs = np.random.RandomState(n)
    p = rs.uniform(0, 10, (n, 3))
    nr = rs.normal(size=(n, 3)); nr /= np.linalg.norm(nr, axis=1, keepdims=True)
    return mg.Mesh(p, None, nr, rs.uniform(0, 1, (n, 2)), rs.randint(0, 255, (n, 4)), np.ones(n))


SMALL = [
    ("grid_pos", lambda: _enc(mg.grid(17, 1), with_uv=False, with_normals=False, with_colors=False)),
    ("grid_est", lambda: _enc(mg.grid(21, 2), normal_pred=1, with_radius=True)),
    ("grid_diff", lambda: _enc(mg.grid(19, 3), normal_pred=0, with_radius=True)),
    ("grid_border", lambda: _enc(mg.grid(23, 4), normal_pred=2)),
    ("groups3", lambda: _enc(mg.grid(18, 5), groups=mg.random_groups(2 * 17 * 17, 3, 5))),
    ("groups4_border", lambda: _enc(mg.grid(22, 6), normal_pred=2, groups=mg.random_groups(2 * 21 * 21, 4, 9))),
    ("hole_border", lambda: _enc(mg.punch_hole(mg.grid(24, 7), 24), normal_pred=2)),
    ("twocomp", lambda: _enc(mg.two_components(14, 8))),
    ("torus", lambda: _enc(mg.torus(24, 12, 9), normal_pred=2)),
    ("sphere", lambda: _enc(mg.sphere(10, 14, 10))),
    ("bowtie", lambda: _enc(mg.bowtie(9, 11), normal_pred=0)),
    ("flat_fan_est", lambda: _enc(mg.flat_fan(20, 12), normal_pred=1)),
    ("flat_fan_border", lambda: _enc(mg.flat_fan(20, 13), normal_pred=2)),
    ("flat_plane_border", lambda: _enc(_flat(mg.grid(16, 14, jitter=0.0)), normal_pred=2)),
    ("far_est", lambda: _enc(_far(mg.grid(16, 15)), normal_pred=1, pos_bits=20)),
    ("color3", lambda: _enc(mg.grid(15, 16), color_comps=3, color_bits=(5, 6, 5))),
    ("const_color", lambda: _enc(_const_color(mg.grid(40, 17)), normal_pred=0)),
    ("none_entropy", lambda: _enc(mg.grid(16, 18), entropy=0, with_radius=True)),
    ("none_entropy_cloud", lambda: _enc(mg.cloud(16, 19), entropy=0, normal_pred=0)),
    ("cloud_all", lambda: _enc(mg.cloud(25, 20), normal_pred=0, with_radius=True)),
    ("cloud_pos", lambda: _enc(mg.cloud(18, 21), with_uv=False, with_normals=False, with_colors=False)),
    ("radius_parallel", lambda: _enc(mg.grid(14, 22), with_radius=True, radius_strategy=1)),
    ("radius_correlated", lambda: _enc(mg.grid(14, 23), with_radius=True, radius_strategy=2)),
    ("radius_both", lambda: _enc(mg.grid(14, 24), with_radius=True, radius_strategy=3)),
    ("normal_bits16", lambda: _enc(mg.grid(15, 25), normal_bits=16, normal_pred=1)),
    ("grid2x2", lambda: _enc(mg.grid(2, 26))),
    ("grid3x3", lambda: _enc(mg.grid(3, 27), normal_pred=2)),
    ("triangle", lambda: _enc(_triangle())),
    ("cloud1", lambda: _enc(_tiny_cloud(1), normal_pred=0)),
    ("cloud2", lambda: _enc(_tiny_cloud(2), normal_pred=0)),
    ("cloud3", lambda: _enc(_tiny_cloud(3), normal_pred=0)),
]

# bigger ones: exercise multi-tile scans / look-back chains (built on the fly where the reference shim exists)
MEDIUM = [
    ("grid185_pos", lambda: _enc(mg.grid(185, 1), with_uv=False, with_normals=False, with_colors=False)),   # BASELINE configs[0] stand-in
    ("grid120_est", lambda: _enc(mg.grid(120, 31), normal_pred=1, with_radius=True)),
    ("grid120_border_groups", lambda: _enc(mg.punch_hole(mg.grid(120, 32), 120), normal_pred=2,
                                            groups=mg.random_groups(2 * 119 * 119 - 2 * 30 * 30, 4, 3))),
    ("torus_big", lambda: _enc(mg.torus(200, 80, 33), normal_pred=1)),
    ("cloud150", lambda: _enc(mg.cloud(150, 34), normal_pred=0, with_radius=True)),
    ("const_color_big", lambda: _enc(_const_color(mg.grid(150, 35)), normal_pred=0)),
    ("none_entropy_big", lambda: _enc(mg.grid(100, 36), entropy=0)),
    ("far_border_big", lambda: _enc(_far(mg.grid(90, 37), 3.0e5), normal_pred=2, pos_bits=22)),
]

# decode-side bindings every blob goes through
VARIANTS = [
    dict(),
    dict(index16=True, normals16=True),
    dict(color_out=4),
    dict(color_out=3),
]


def applicable(variant, info_attrs, nvert, nface):
    """Skip variants that make no sense for a blob (u16 index needs nvert < 65536; colour variants need a colour)."""
    if variant.get("index16") and nvert > 65535:
        return False
    if "color_out" in variant:
        col = [a for a in info_attrs if a["codec"] == 3]
        if not col:
            return False
        if variant["color_out"] < col[0]["N"]:
            return False     # N=4 -> 3 self-corrupts in the reference; covered by its own oracle-only test
    return True
