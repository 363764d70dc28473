"""ctypes binding for oracle/_ref/libcorto_ref.so — the UNMODIFIED reference compiled in place.

TEST INFRASTRUCTURE: ground truth for parity, fixture generator, and the `reference` CPU baseline.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("CORTO_REFSHIM_SO") or os.path.join(_HERE, "_ref", "libcorto_ref.so")   # override: sanitizer builds
TARTA = os.path.join(_HERE, "_ref", "tarta.crt")
_lib = None


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libcorto_ref.so missing: run `make -C oracle` where /root/reference exists")
        L = C.CDLL(_SO)
        L.ref_encode.restype = C.c_void_p
        L.ref_encode.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_float,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_info.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.ref_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p]
        L.ref_decode_debug.argtypes = L.ref_decode.argtypes + [C.c_void_p] * 4
        L.ref_decode_fmt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_decode_bench.restype = C.c_double
        L.ref_decode_bench.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def aligned_blob(data):
    """bytes -> uint8 ndarray whose base address is 16-byte aligned (reference needs 4, decoder.cpp:43)."""
    n = len(data)
    raw = np.empty(n + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off:off + n]
    out[:] = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return out


def encode(mesh, pos_bits=14, pos_q=0.0, uv_bits=12, normal_bits=10, normal_pred=1, color_bits=(6, 6, 6, 6), color_comps=4,
           with_uv=True, with_normals=True, with_colors=True, with_radius=False, radius_q=1.0 / 64, radius_strategy=0,
           groups=None, entropy=1):
    """Encode with the reference Encoder.  Returns (blob uint8 ndarray 16B-aligned, nvert, nface)."""
    L = lib()
    nface = mesh.nface
    uv = mesh.uv if (with_uv and mesh.uv is not None) else None
    nrm = mesh.normals if (with_normals and mesh.normals is not None) else None
    col = None
    if with_colors and mesh.colors is not None:
        col = np.ascontiguousarray(mesh.colors[:, :color_comps])
    rad = mesh.radius if (with_radius and mesh.radius is not None) else None
    cb = np.array(list(color_bits) + [8] * (4 - len(color_bits)), dtype=np.int32)
    groups = groups if groups is not None else (mesh.groups or [])
    g = np.array(groups, dtype=np.int32)
    ol, ov, of = C.c_uint32(), C.c_uint32(), C.c_uint32()
    err = C.create_string_buffer(256)
    ptr = L.ref_encode(mesh.nvert, nface, _p(mesh.pos), _p(mesh.faces), pos_bits, pos_q, _p(uv), 2.0 ** -uv_bits,
                       _p(nrm), normal_bits, normal_pred, _p(col), color_comps, _p(cb), _p(rad), radius_q, radius_strategy,
                       _p(g), len(groups), entropy, C.byref(ol), C.byref(ov), C.byref(of), err, 256)
    if not ptr:
        raise RuntimeError("reference encoder: " + err.value.decode())
    blob = aligned_blob(C.string_at(ptr, ol.value))
    L.ref_free(ptr)
    return blob, ov.value, of.value


def info(blob):
    nv, nf, m, cc = C.c_uint32(), C.c_uint32(), C.c_int(), C.c_int(4)
    if lib().ref_info(_p(blob), len(blob), C.byref(nv), C.byref(nf), C.byref(m), C.byref(cc)) != 0:
        raise RuntimeError("reference decoder rejected blob")
    return dict(nvert=nv.value, nface=nf.value, mask=m.value, color_comps=cc.value)


def decode(blob, index16=False, normals16=False, color_out=None, bind=None, debug=False, sentinel=0xA5):
    """Decode with the reference Decoder.  Returns dict of numpy arrays (bit patterns are what matter).
    bind: iterable of names to bind (default: everything present).  Outputs are pre-filled with `sentinel`
    bytes so that untouched elements are visible."""
    i = info(blob)
    nv, nf, mask = i["nvert"], i["nface"], i["mask"]
    names = {"position": 1, "normal": 2, "color": 4, "uv": 8, "radius": 16}
    want = set(n for n, b in names.items() if mask & b) if bind is None else set(bind)
    if nf and (bind is None or "index" in bind):
        want.add("index")
    cc = color_out or i["color_comps"]

    def buf(shape, dt):
        a = np.empty(shape, dtype=dt)
        a.view(np.uint8)[...] = sentinel
        return a
    out = {}
    if "position" in want: out["position"] = buf((nv, 3), np.float32)
    if "normal" in want: out["normal"] = buf((nv, 3), np.int16 if normals16 else np.float32)
    if "color" in want: out["color"] = buf((nv, max(cc, i["color_comps"])), np.uint8)
    if "uv" in want: out["uv"] = buf((nv, 2), np.float32)
    if "radius" in want: out["radius"] = buf((nv,), np.float32)
    if "index" in want: out["index"] = buf((nf, 3), np.uint16 if index16 else np.uint32)
    n32 = out.get("normal") if not normals16 else None
    n16 = out.get("normal") if normals16 else None
    args = [_p(blob), len(blob), _p(out.get("position")), _p(out.get("index")), int(index16), _p(n32), _p(n16),
            _p(out.get("color")), cc, _p(out.get("uv")), _p(out.get("radius"))]
    if debug:
        clers = np.zeros(nf * 4 + 64, dtype=np.uint8)
        ncl = C.c_uint32()
        pred = np.zeros((nv, 3), dtype=np.uint32)
        grp = np.zeros(4096, dtype=np.int32)
        ng = lib().ref_decode_debug(*args, _p(clers), C.byref(ncl), _p(pred), _p(grp))
        out["clers"] = clers[:ncl.value].copy()
        out["prediction"] = pred
    else:
        ng = lib().ref_decode(*args)
    if ng < 0:
        raise RuntimeError("reference decoder threw")
    if "color" in out:
        # the reference writes nvert*out_components bytes packed at the front of the buffer
        out["color"] = out["color"].reshape(-1)[:nv * cc].reshape(nv, cc).copy()
    out["nvert"], out["nface"], out["ngroups"] = nv, nf, ng
    return out


def decode_formats(blob, formats, sentinel=0xA5):
    """Generic attributes (position / uv / radius) decoded by the reference with an explicit output Format per attribute
    (formats: name -> Format enum value; names absent from it are left unbound).  Returns dict of uint32/float32 arrays."""
    i = info(blob)
    nv, mask = i["nvert"], i["mask"]
    comps = {"position": 3, "uv": 2, "radius": 1}
    bits = {"position": 1, "uv": 8, "radius": 16}
    out = {}
    for name, fmt in formats.items():
        if not (mask & bits[name]):
            continue
        a = np.empty((nv, comps[name]) if comps[name] > 1 else (nv,), dtype=np.float32 if fmt == 6 else np.uint32)
        a.view(np.uint8)[...] = sentinel
        out[name] = a
    rc = lib().ref_decode_fmt(_p(blob), len(blob), _p(out.get("position")), formats.get("position", 6), _p(out.get("uv")), formats.get("uv", 6),
                              _p(out.get("radius")), formats.get("radius", 6))
    if rc:
        raise RuntimeError("reference decoder threw")
    return out


def decode_bench(blobs, nthreads=1, repeats=3):
    """Best wall seconds for one pass of reference ctor+set*+decode() over all blobs."""
    n = len(blobs)
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in blobs])
    lens = (C.c_int * n)(*[len(b) for b in blobs])
    return lib().ref_decode_bench(n, ptrs, lens, nthreads, repeats)
