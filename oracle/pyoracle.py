"""ctypes binding for oracle/liboracle.so — the plain-C restatement of the reference decode path.

TEST INFRASTRUCTURE (checker / `port` CPU baseline).  Same calling convention as oracle.refshim.decode so the
two are interchangeable in tests.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

F_UINT32, F_INT32, F_UINT16, F_INT16, F_UINT8, F_INT8, F_FLOAT, F_DOUBLE = range(8)
MAX_ATTR = 16


class OAttrInfo(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("codec", C.c_int), ("q", C.c_float), ("N", C.c_int), ("format", C.c_int),
                ("strategy", C.c_int)]


class OInfo(C.Structure):
    _fields_ = [("nvert", C.c_uint32), ("nface", C.c_uint32), ("entropy", C.c_int), ("nattr", C.c_int),
                ("attr", OAttrInfo * MAX_ATTR), ("body", C.c_uint32)]


class OBind(C.Structure):
    _fields_ = [("buffer", C.c_void_p), ("format", C.c_int), ("out_components", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.crt_oracle_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(OInfo)]
        L.crt_oracle_decode.argtypes = [C.c_void_p, C.c_int, C.POINTER(OBind), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.crt_oracle_fnv1a64.restype = C.c_uint64
        L.crt_oracle_fnv1a64.argtypes = [C.c_void_p, C.c_uint64]
        L.crt_oracle_tunstall_tables.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.crt_oracle_decode_all.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def info(blob):
    i = OInfo()
    rc = lib().crt_oracle_info(_p(blob), len(blob), C.byref(i))
    if rc:
        raise RuntimeError("oracle: bad blob (%d)" % rc)
    attrs = [dict(name=i.attr[k].name.decode(), codec=i.attr[k].codec, q=i.attr[k].q, N=i.attr[k].N, format=i.attr[k].format,
                  strategy=i.attr[k].strategy) for k in range(i.nattr)]
    return dict(nvert=i.nvert, nface=i.nface, entropy=i.entropy, attrs=attrs)


def fnv1a64(arr):
    a = np.ascontiguousarray(arr)
    return lib().crt_oracle_fnv1a64(_p(a), a.nbytes)


def decode(blob, index16=False, normals16=False, color_out=None, bind=None, debug=False, sentinel=0xA5, formats=None):
    """Decode with the C restatement.  Mirrors oracle.refshim.decode (same dict of arrays)."""
    i = info(blob)
    nv, nf = i["nvert"], i["nface"]
    names = [a["name"] for a in i["attrs"]]
    want = set(names) if bind is None else set(bind)
    if nf and (bind is None or "index" in bind):
        want.add("index")
    formats = formats or {}

    def buf(shape, dt):
        a = np.empty(shape, dtype=dt)
        a.view(np.uint8)[...] = sentinel
        return a
    out = {}
    binds = (OBind * MAX_ATTR)()
    cc = None
    for k, a in enumerate(i["attrs"]):
        nm = a["name"]
        if nm not in want:
            continue
        if a["codec"] == 2:
            arr = buf((nv, 3), np.int16 if normals16 else np.float32)
            binds[k].format = F_INT16 if normals16 else F_FLOAT
        elif a["codec"] == 3:
            cc = color_out or a["N"]
            arr = buf((nv * max(cc, a["N"]),), np.uint8)
            binds[k].format = F_UINT8
            binds[k].out_components = cc
        else:
            fmt = formats.get(nm, F_FLOAT)
            arr = buf((nv, a["N"]) if a["N"] > 1 else (nv,), np.float32 if fmt == F_FLOAT else np.uint32)
            binds[k].format = fmt
        binds[k].buffer = arr.ctypes.data
        out[nm] = arr
    if "index" in want:
        out["index"] = buf((nf, 3), np.uint16 if index16 else np.uint32)
    clers = np.zeros(nf * 4 + 64, dtype=np.uint8) if debug else None
    ncl = C.c_uint32()
    pred = np.zeros((nv, 3), dtype=np.uint32) if debug else None
    rc = lib().crt_oracle_decode(_p(blob), len(blob), binds, _p(out.get("index")), int(index16), _p(clers), C.byref(ncl), _p(pred))
    if rc:
        raise RuntimeError("oracle decode failed (%d)" % rc)
    if "color" in out:
        out["color"] = out["color"][:nv * cc].reshape(nv, cc).copy()
    if debug:
        out["clers"] = clers[:ncl.value].copy()
        out["prediction"] = pred
    out["nvert"], out["nface"] = nv, nf
    return out


def tunstall_tables(probs):
    """probs: uint8 array of (symbol, prob) pairs.  Returns (index[256], lengths[256], table[8192], used)."""
    probs = np.ascontiguousarray(probs, dtype=np.uint8)
    idx = np.zeros(256, dtype=np.int32)
    ln = np.zeros(256, dtype=np.int32)
    tab = np.zeros(8192, dtype=np.uint8)
    used = lib().crt_oracle_tunstall_tables(_p(probs), probs.size // 2, _p(idx), _p(ln), _p(tab))
    return idx, ln, tab, used


def decode_all(blobs):
    n = len(blobs)
    ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in blobs])
    lens = (C.c_int * n)(*[len(b) for b in blobs])
    rc = lib().crt_oracle_decode_all(n, ptrs, lens)
    if rc:
        raise RuntimeError("oracle decode_all failed")
