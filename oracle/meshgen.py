"""Deterministic synthetic meshes and point clouds (test infrastructure).

Follows the generator SURVEY.md §8d specifies for the BASELINE configs: a W x W height-field grid,
x = col + U(-.2,.2), y = row + U(-.2,.2), z = 10 sin(.05x) cos(.04y) + U(-.2,.2), two triangles per
cell, analytic unit normals, uv = (col/W, row/W), colour = f(col,row,z), alpha 255 — plus a few
extra topologies (torus, sphere, bow-tie, holes, multi-component, degenerate fans) that drive the
CLERS automaton through DELAY / END / SPLIT / BOUNDARY and the x86 float corner cases.
numpy.random.RandomState is bit-stable across numpy versions, so a (kind, W, seed) triple always
yields the same arrays and therefore the same .crt blob out of the reference Encoder.
"""
import numpy as np


class Mesh:
    def __init__(self, pos, faces=None, normals=None, uv=None, colors=None, radius=None, groups=None):
        self.pos = np.ascontiguousarray(pos, dtype=np.float32)
        self.faces = None if faces is None else np.ascontiguousarray(faces, dtype=np.uint32)
        self.normals = None if normals is None else np.ascontiguousarray(normals, dtype=np.float32)
        self.uv = None if uv is None else np.ascontiguousarray(uv, dtype=np.float32)
        self.colors = None if colors is None else np.ascontiguousarray(colors, dtype=np.uint8)
        self.radius = None if radius is None else np.ascontiguousarray(radius, dtype=np.float32)
        self.groups = groups  # list of end-face indices or None

    @property
    def nvert(self):
        return self.pos.shape[0]

    @property
    def nface(self):
        return 0 if self.faces is None else self.faces.shape[0]


def _grid_faces(W, H=None):
    H = W if H is None else H
    r, c = np.meshgrid(np.arange(H - 1), np.arange(W - 1), indexing="ij")
    v00 = (r * W + c).ravel()
    v01 = v00 + 1
    v10 = v00 + W
    v11 = v10 + 1
    f = np.empty((v00.size * 2, 3), dtype=np.uint32)
    f[0::2] = np.stack([v00, v01, v11], 1)
    f[1::2] = np.stack([v00, v11, v10], 1)
    return f


def grid(W, seed=1, H=None, jitter=0.2, with_attrs=True):
    """SURVEY §8d height-field grid."""
    H = W if H is None else H
    rs = np.random.RandomState(seed)
    row, col = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    x = col + rs.uniform(-jitter, jitter, col.shape)
    y = row + rs.uniform(-jitter, jitter, col.shape)
    zs = 10.0 * np.sin(0.05 * x) * np.cos(0.04 * y)
    z = zs + rs.uniform(-jitter, jitter, col.shape)
    pos = np.stack([x, y, z], -1).reshape(-1, 3).astype(np.float32)
    faces = _grid_faces(W, H)
    if not with_attrs:
        return Mesh(pos, faces)
    # analytic normal of the smooth surface z = 10 sin(.05x) cos(.04y)
    dzdx = 10.0 * 0.05 * np.cos(0.05 * x) * np.cos(0.04 * y)
    dzdy = -10.0 * 0.04 * np.sin(0.05 * x) * np.sin(0.04 * y)
    n = np.stack([-dzdx, -dzdy, np.ones_like(dzdx)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    uv = np.stack([col / W, row / H], -1)
    colr = np.stack([(col * 255 / max(W - 1, 1)), (row * 255 / max(H - 1, 1)), (z - z.min()) * 255 / max(np.ptp(z), 1e-6),
                     np.full_like(col, 255.0)], -1)
    radius = (1.0 + 0.5 * np.sin(0.1 * x + 0.07 * y))
    return Mesh(pos, faces, n.reshape(-1, 3), uv.reshape(-1, 2), colr.reshape(-1, 4).astype(np.uint8),
                radius.reshape(-1))


def cloud(W, seed=1):
    """Point cloud with the grid's vertices (nface = 0); the encoder Morton-sorts it."""
    m = grid(W, seed)
    m.faces = None
    return m


def punch_hole(m, W, frac=0.25):
    """Remove a block of cells from a W x W grid mesh -> interior boundary loop."""
    lo, hi = int(W * (0.5 - frac / 2)), int(W * (0.5 + frac / 2))
    cell = np.arange(m.nface) // 2
    r, c = cell // (W - 1), cell % (W - 1)
    keep = ~((r >= lo) & (r < hi) & (c >= lo) & (c < hi))
    m.faces = np.ascontiguousarray(m.faces[keep])
    return m


def two_components(W, seed=1):
    """Two disjoint grids in one mesh."""
    a, b = grid(W, seed), grid(max(W // 2, 2), seed + 1000)
    off = a.nvert
    b.pos[:, 0] += W + 10
    cat = lambda x, y: np.concatenate([x, y], 0)
    return Mesh(cat(a.pos, b.pos), cat(a.faces, b.faces + off), cat(a.normals, b.normals), cat(a.uv, b.uv),
                cat(a.colors, b.colors), cat(a.radius, b.radius))


def torus(W, H, seed=1, R=10.0, r=3.0):
    """Closed genus-1 surface: drives END and non-initial SPLIT symbols."""
    rs = np.random.RandomState(seed)
    i, j = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    u = 2 * np.pi * j / W
    v = 2 * np.pi * i / H
    x = (R + r * np.cos(v)) * np.cos(u)
    y = (R + r * np.cos(v)) * np.sin(u)
    z = r * np.sin(v)
    pos = np.stack([x, y, z], -1).reshape(-1, 3) + rs.uniform(-0.01, 0.01, (W * H, 3))
    n = np.stack([np.cos(v) * np.cos(u), np.cos(v) * np.sin(u), np.sin(v)], -1).reshape(-1, 3)
    v00 = (i * W + j).ravel()
    v01 = (i * W + (j + 1) % W).ravel()
    v10 = (((i + 1) % H) * W + j).ravel()
    v11 = (((i + 1) % H) * W + (j + 1) % W).ravel()
    f = np.empty((v00.size * 2, 3), dtype=np.uint32)
    f[0::2] = np.stack([v00, v01, v11], 1)
    f[1::2] = np.stack([v00, v11, v10], 1)
    uv = np.stack([j / W, i / H], -1).reshape(-1, 2)
    col = np.stack([j * 255 // W, i * 255 // H, (j + i) % 256, np.full_like(i, 255)], -1).reshape(-1, 4)
    return Mesh(pos, f, n, uv, col.astype(np.uint8), np.ones(W * H))


def sphere(n_lat, n_lon, seed=1, R=5.0):
    """Closed genus-0 UV sphere with pole fans."""
    rs = np.random.RandomState(seed)
    pts = [[0, 0, R]]
    for a in range(1, n_lat):
        th = np.pi * a / n_lat
        for b in range(n_lon):
            ph = 2 * np.pi * b / n_lon
            pts.append([R * np.sin(th) * np.cos(ph), R * np.sin(th) * np.sin(ph), R * np.cos(th)])
    pts.append([0, 0, -R])
    pos = np.array(pts) + rs.uniform(-0.01, 0.01, (len(pts), 3))
    f = []
    ring = lambda a, b: 1 + (a - 1) * n_lon + (b % n_lon)
    for b in range(n_lon):
        f.append([0, ring(1, b), ring(1, b + 1)])
    for a in range(1, n_lat - 1):
        for b in range(n_lon):
            f.append([ring(a, b), ring(a + 1, b), ring(a + 1, b + 1)])
            f.append([ring(a, b), ring(a + 1, b + 1), ring(a, b + 1)])
    last = len(pts) - 1
    for b in range(n_lon):
        f.append([last, ring(n_lat - 1, b + 1), ring(n_lat - 1, b)])
    nrm = pos / np.linalg.norm(pos, axis=1, keepdims=True)
    uv = np.stack([np.arctan2(nrm[:, 1], nrm[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(nrm[:, 2], -1, 1)) / np.pi], -1)
    col = np.clip((nrm * 0.5 + 0.5) * 255, 0, 255)
    col = np.concatenate([col, np.full((len(pts), 1), 255.0)], 1)
    return Mesh(pos, np.array(f, dtype=np.uint32), nrm, uv, col.astype(np.uint8), np.ones(len(pts)))


def bowtie(W, seed=1):
    """Two grids sharing ONE vertex (non-manifold vertex): second component starts with a SPLIT triangle."""
    a, b = grid(W, seed), grid(W, seed + 7)
    b.pos[:, 0] += (W - 1)
    b.pos[:, 1] += (W - 1)
    off = a.nvert
    fb = b.faces + off
    shared_a = a.nvert - 1           # last vertex of a == first vertex of b
    fb[fb == off] = shared_a
    cat = lambda x, y: np.concatenate([x, y], 0)
    m = Mesh(cat(a.pos, b.pos), cat(a.faces, fb), cat(a.normals, b.normals), cat(a.uv, b.uv), cat(a.colors, b.colors),
             cat(a.radius, b.radius))
    return m


def flat_fan(W, seed=1):
    """Grid whose z is constant and x,y exactly on the lattice in a band: after quantisation whole
    rows are collinear-free but zero-area fans appear when a row is squashed (y equal) -> zero normals (H6/H7)."""
    m = grid(W, seed)
    p = m.pos.reshape(W, W, 3)
    band = slice(W // 3, W // 3 + 3)
    p[band, :, 1] = p[W // 3, :, 1].mean()      # three rows share the same y
    p[band, :, 2] = 0.0                        # and the same z -> all triangles between them are degenerate (collinear)
    p[band, :, 0] = np.arange(W)[None, :]
    m.pos = np.ascontiguousarray(p.reshape(-1, 3))
    return m


def random_groups(nface, ngroups, seed):
    rs = np.random.RandomState(seed)
    if ngroups <= 1:
        return [nface]
    cuts = sorted(set(int(c) for c in rs.randint(1, max(nface - 1, 2), ngroups - 1)))
    return cuts + [nface]
