"""corto_b200 — B200-native (sm_100a) implementation of corto's .crt DECODE path.

This package is only the thin host-side mirror of the reference's decoder interface
(``crt::Decoder``, include/corto/decoder.h:38-73) over the C ABI of ``lib/libcorto_b200.so``
(include/corto_b200.h).  All decode work runs in hand-written CUDA kernels (corto_b200/csrc).  There is
no CPU fallback: if the library is missing or no CUDA device is visible, decoding raises.

    dec = corto_b200.Decoder(blob)            # header parse (decoder.cpp:41-89)
    out = dec.decode()                         # dict of numpy arrays; H2D, kernels, D2H

    bd = corto_b200.BatchDecoder(blobs)        # many blobs, outputs stay in HBM (torch tensors)
    bd.upload(); bd.decode(); torch.cuda.synchronize()
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcorto_b200.so")

UINT32, INT32, UINT16, INT16, UINT8, INT8, FLOAT, DOUBLE = range(8)
NORMAL_DIFF, NORMAL_ESTIMATED, NORMAL_BORDER = 0, 1, 2
HAS_POSITION, HAS_NORMAL, HAS_COLOR, HAS_UV, HAS_OTHER, HAS_INDEX = 1, 2, 4, 8, 16, 32

_lib = None


class CortoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (code %d)" % (msg, code))
        self.code = code


def lib():
    """Load the CUDA library; fail loudly if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("corto_b200: %s not built (run __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, ci, cu32, cu64 = C.c_void_p, C.c_int, C.c_uint32, C.c_uint64
        L.crt_last_error.restype = C.c_char_p
        L.crt_new_decoder.restype = vp; L.crt_new_decoder.argtypes = [ci, vp]
        L.crt_delete_decoder.argtypes = [vp]
        L.crt_nvert.restype = cu32; L.crt_nvert.argtypes = [vp]
        L.crt_nface.restype = cu32; L.crt_nface.argtypes = [vp]
        L.crt_ngroups.argtypes = [vp]; L.crt_groups.argtypes = [vp, vp]
        L.crt_group_nprops.argtypes = [vp, ci]
        L.crt_group_prop.restype = C.c_char_p; L.crt_group_prop.argtypes = [vp, ci, ci, C.POINTER(C.c_char_p)]
        L.crt_nexif.argtypes = [vp]
        L.crt_exif.restype = C.c_char_p; L.crt_exif.argtypes = [vp, ci, C.POINTER(C.c_char_p)]
        L.crt_has_attr.argtypes = [vp, C.c_char_p]
        L.crt_nattr.argtypes = [vp]
        L.crt_attr_info.restype = C.c_char_p
        L.crt_attr_info.argtypes = [vp, ci, C.POINTER(ci), C.POINTER(C.c_float), C.POINTER(ci), C.POINTER(ci), C.POINTER(ci)]
        for f in ("crt_set_positions", "crt_set_normals32", "crt_set_normals16", "crt_set_uvs", "crt_set_index32", "crt_set_index16"):
            getattr(L, f).argtypes = [vp, vp]
        L.crt_set_colors.argtypes = [vp, vp, ci]
        L.crt_set_attribute.argtypes = [vp, C.c_char_p, vp, ci]
        L.crt_decode.argtypes = [vp]
        L.crt_normal_prediction.argtypes = [vp]
        L.crt_color_q.argtypes = [vp, vp]
        L.crt_batch_create.restype = vp; L.crt_batch_create.argtypes = [ci, vp, vp]
        L.crt_batch_destroy.argtypes = [vp]
        L.crt_batch_count.argtypes = [vp]
        L.crt_batch_mesh_info.argtypes = [vp, ci, C.POINTER(cu32), C.POINTER(cu32), C.POINTER(cu32)]
        for f in ("crt_batch_total_verts", "crt_batch_total_faces", "crt_batch_total_bytes"):
            getattr(L, f).restype = cu64; getattr(L, f).argtypes = [vp]
        L.crt_batch_vert_base.restype = C.POINTER(cu64); L.crt_batch_vert_base.argtypes = [vp]
        L.crt_batch_face_base.restype = C.POINTER(cu64); L.crt_batch_face_base.argtypes = [vp]
        L.crt_batch_bind.argtypes = [vp, C.c_char_p, vp, ci, ci]
        L.crt_batch_attr_components.argtypes = [vp, C.c_char_p]
        L.crt_batch_upload.argtypes = [vp, vp]
        L.crt_batch_rewalk.argtypes = [vp, vp]
        L.crt_batch_decode.argtypes = [vp, vp]
        L.crt_batch_status.argtypes = [vp, vp]
        L.crt_batch_launches.argtypes = [vp]
        L.crt_batch_set_profiling.argtypes = [vp, ci]
        L.crt_batch_stage_times.argtypes = [vp, vp, vp, ci]
        L.crt_batch_debug_clers.argtypes = [vp, ci, vp, cu32, C.POINTER(cu32)]
        L.crt_batch_debug_prediction.argtypes = [vp, ci, vp]
        L.crt_shard_lpt.argtypes = [ci, vp, vp, vp, ci, vp]
        L.crt_walk_tape.argtypes = [vp, ci, vp, ci, C.POINTER(cu32), C.POINTER(cu32), C.POINTER(cu32)]
        L.crt_batch_create_device.restype = vp; L.crt_batch_create_device.argtypes = [ci, vp, vp, vp, vp]
        L.crt_batch_directory_signature.restype = cu64; L.crt_batch_directory_signature.argtypes = [vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise CortoError(rc, lib().crt_last_error().decode(errors="replace"))


def device_available():
    return bool(lib().crt_device_available())


def _aligned_copy(data):
    """bytes-like -> uint8 ndarray whose base is 16-byte aligned (decoder.cpp:43 demands 4)."""
    src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    if isinstance(data, np.ndarray) and data.ctypes.data % 4 == 0 and data.flags.c_contiguous:
        return src
    raw = np.empty(src.size + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    out = raw[off:off + src.size]
    out[:] = src
    return out


class Decoder:
    """Mirror of crt::Decoder (include/corto/decoder.h:38-73) with host (numpy) buffers.

    The reference's set* calls bind caller-allocated arrays; ``setPositions`` etc. do the same here with numpy
    arrays, and ``decode()`` with nothing bound allocates every array present in the stream and returns them.
    """

    def __init__(self, blob):
        self._blob = _aligned_copy(blob)                       # borrowed by the C side: keep alive
        self._h = lib().crt_new_decoder(len(self._blob), self._blob.ctypes.data_as(C.c_void_p))
        if not self._h:
            raise CortoError(-1, lib().crt_last_error().decode())
        self.nvert = lib().crt_nvert(self._h)
        self.nface = lib().crt_nface(self._h)
        self._keep = {}
        self.attributes = {}
        for i in range(lib().crt_nattr(self._h)):
            codec, q, n, fmt, strat = C.c_int(), C.c_float(), C.c_int(), C.c_int(), C.c_int()
            name = lib().crt_attr_info(self._h, i, C.byref(codec), C.byref(q), C.byref(n), C.byref(fmt), C.byref(strat)).decode()
            self.attributes[name] = dict(codec=codec.value, q=q.value, N=n.value, format=fmt.value, strategy=strat.value)
        self.exif = {}
        for i in range(lib().crt_nexif(self._h)):
            v = C.c_char_p()
            k = lib().crt_exif(self._h, i, C.byref(v))
            self.exif[k.decode()] = v.value.decode()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().crt_delete_decoder(self._h)
            self._h = None

    def hasAttr(self, name):
        return bool(lib().crt_has_attr(self._h, name.encode()))

    @property
    def groups(self):
        n = lib().crt_ngroups(self._h)
        ends = (C.c_int * max(n, 1))()
        lib().crt_groups(self._h, ends)
        out = []
        for g in range(n):
            props = {}
            for i in range(lib().crt_group_nprops(self._h, g)):
                v = C.c_char_p()
                k = lib().crt_group_prop(self._h, g, i, C.byref(v))
                props[k.decode()] = v.value.decode()
            out.append(dict(end=ends[g], properties=props))
        return out

    def _bind(self, key, arr, fn, *extra):
        self._keep[key] = arr
        return bool(fn(self._h, arr.ctypes.data_as(C.c_void_p), *extra))

    def setPositions(self, arr): return self._bind("position", arr, lib().crt_set_positions)
    def setNormals(self, arr):
        return self._bind("normal", arr, lib().crt_set_normals16 if arr.dtype == np.int16 else lib().crt_set_normals32)
    def setUvs(self, arr): return self._bind("uv", arr, lib().crt_set_uvs)
    def setColors(self, arr, components=4): return self._bind("color", arr, lib().crt_set_colors, components)

    def setAttribute(self, name, arr, fmt):
        self._keep[name] = arr
        return bool(lib().crt_set_attribute(self._h, name.encode(), arr.ctypes.data_as(C.c_void_p), fmt))

    def setIndex(self, arr):
        self._keep["index"] = arr
        (lib().crt_set_index16 if arr.dtype == np.uint16 else lib().crt_set_index32)(self._h, arr.ctypes.data_as(C.c_void_p))

    def decode(self, index16=False, normals16=False, color_components=None, bind=None, sentinel=0xA5, formats=None):
        """Decode.  If nothing was bound with set*, allocate outputs for everything in ``bind`` (default: all attributes
        present + index), pre-filled with ``sentinel`` bytes, and return them as a dict."""
        auto = not self._keep
        if auto:
            want = set(self.attributes) | ({"index"} if self.nface else set()) if bind is None else set(bind)
            formats = formats or {}

            def buf(shape, dt):
                a = np.empty(shape, dtype=dt)
                a.view(np.uint8)[...] = sentinel
                return a
            for name, a in self.attributes.items():
                if name not in want:
                    continue
                if a["codec"] == 2:
                    self.setNormals(buf((self.nvert, 3), np.int16 if normals16 else np.float32))
                elif a["codec"] == 3:
                    cc = color_components or a["N"]
                    self.setColors(buf((self.nvert, cc), np.uint8), cc)
                else:
                    fmt = formats.get(name, FLOAT)
                    shape = (self.nvert, a["N"]) if a["N"] > 1 else (self.nvert,)
                    self.setAttribute(name, buf(shape, np.float32 if fmt == FLOAT else np.uint32), fmt)
            if "index" in want and self.nface:
                self.setIndex(buf((self.nface, 3), np.uint16 if index16 else np.uint32))
        _check(lib().crt_decode(self._h))
        out = dict(self._keep)
        out["nvert"], out["nface"] = self.nvert, self.nface
        return out

    @property
    def normal_prediction(self):
        return lib().crt_normal_prediction(self._h)

    @property
    def color_q(self):
        q = (C.c_int * 4)()
        lib().crt_color_q(self._h, q)
        return list(q)


class BatchDecoder:
    """Batched, device-resident decode (include/corto_b200.h group 3).  Outputs are torch CUDA tensors laid out as
    flat arenas, meshes concatenated in batch order; ``vert_base`` / ``face_base`` locate mesh i."""

    def __init__(self, blobs, normals16=False, index16=False, color_components=4, bind=None, device=None, _device_arena=None):
        import torch
        self.torch = torch
        if _device_arena is None:
            self._blobs = [_aligned_copy(b) for b in blobs]
            n = len(self._blobs)
            ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in self._blobs])
            lens = (C.c_int * max(n, 1))(*[len(b) for b in self._blobs])
            self._h = lib().crt_batch_create(n, ptrs, lens)
        else:
            # blobs = (tapes, blob lengths): the payload is in `_device_arena` (a CUDA uint8 tensor) already
            tapes, blob_lens = blobs
            self._arena = _device_arena
            self._tapes = [np.ascontiguousarray(t, dtype=np.uint8) for t in tapes]
            n = len(self._tapes)
            ptrs = (C.c_void_p * max(n, 1))(*[t.ctypes.data for t in self._tapes])
            tl = (C.c_int * max(n, 1))(*[len(t) for t in self._tapes])
            bl = (C.c_int * max(n, 1))(*[int(x) for x in blob_lens])
            self._h = lib().crt_batch_create_device(n, ptrs, tl, bl, C.c_void_p(_device_arena.data_ptr()))
        if not self._h:
            raise CortoError(-1, lib().crt_last_error().decode())
        self.n = n
        self.total_verts = lib().crt_batch_total_verts(self._h)
        self.total_faces = lib().crt_batch_total_faces(self._h)
        self.total_bytes = lib().crt_batch_total_bytes(self._h)
        vb, fb = lib().crt_batch_vert_base(self._h), lib().crt_batch_face_base(self._h)
        self.vert_base = np.array([vb[i] for i in range(n + 1)], dtype=np.int64)
        self.face_base = np.array([fb[i] for i in range(n + 1)], dtype=np.int64)
        self.mask = 0
        self.info = []
        for i in range(n):
            nv, nf, m = C.c_uint32(), C.c_uint32(), C.c_uint32()
            lib().crt_batch_mesh_info(self._h, i, C.byref(nv), C.byref(nf), C.byref(m))
            self.info.append((nv.value, nf.value, m.value))
            self.mask |= m.value
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.normals16, self.index16, self.color_components = normals16, index16, color_components
        self.out = {}
        self._want = bind
        self.launches = 0

    @classmethod
    def from_device(cls, tapes, blob_lens, arena, **kw):
        """Batch over blobs that are ALREADY in device memory: `arena` is a CUDA uint8 tensor holding them back to back (each at
        the sum of the 16-byte-rounded lengths before it), `tapes` their walk tapes (`walk_tape` on the rank that had the host
        copy).  The arena is borrowed; no payload byte returns to the host (include/corto_b200.h: crt_batch_create_device)."""
        return cls((tapes, blob_lens), _device_arena=arena, **kw)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().crt_batch_destroy(self._h)
            self._h = None

    def allocate(self, fill=None):
        """Allocate the output arenas for every attribute present in the batch (or the ``bind`` subset) and bind them."""
        t, V, F = self.torch, int(self.total_verts), int(self.total_faces)
        want = self._want

        def mk(name, shape, dtype, fmt, comps=0):
            if want is not None and name not in want:
                return
            x = t.empty(shape, dtype=dtype, device=self.device)
            if fill is not None:
                x.view(t.uint8).fill_(fill)
            self.out[name] = x
            _check(lib().crt_batch_bind(self._h, name.encode(), C.c_void_p(x.data_ptr()), fmt, comps))
        for name, n in (("position", 3), ("uv", 2)):      # the arenas below are sized for these component counts
            have = lib().crt_batch_attr_components(self._h, name.encode())
            if have not in (0, n):
                raise CortoError(-8, "attribute '%s' has %d components, BatchDecoder.allocate lays out %d" % (name, have, n))
        if self.mask & HAS_POSITION: mk("position", (V, 3), t.float32, FLOAT)
        if self.mask & HAS_UV: mk("uv", (V, 2), t.float32, FLOAT)
        if self.mask & HAS_NORMAL:
            mk("normal", (V, 3), t.int16 if self.normals16 else t.float32, INT16 if self.normals16 else FLOAT)
        if self.mask & HAS_COLOR: mk("color", (V, self.color_components), t.uint8, UINT8, self.color_components)
        if self.mask & HAS_INDEX:
            mk("index", (F, 3), t.int16 if self.index16 else t.int32, UINT16 if self.index16 else UINT32)
        return self.out

    def bind(self, name, tensor, fmt, components=0):
        self.out[name] = tensor
        _check(lib().crt_batch_bind(self._h, name.encode(), C.c_void_p(tensor.data_ptr()), fmt, components))

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def upload(self):
        if not self.out:
            self.allocate()
        _check(lib().crt_batch_upload(self._h, self._stream()))

    def rewalk(self):
        _check(lib().crt_batch_rewalk(self._h, self._stream()))

    def decode(self):
        _check(lib().crt_batch_decode(self._h, self._stream()))
        self.launches = lib().crt_batch_launches(self._h)

    def status(self):
        st = (C.c_int * max(self.n, 1))()
        rc = lib().crt_batch_status(self._h, st)
        return rc, list(st)[:self.n]

    def set_profiling(self, on=True):
        lib().crt_batch_set_profiling(self._h, int(on))

    def stage_times(self):
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        k = lib().crt_batch_stage_times(self._h, names, ms, 32)
        return [(names[i].decode(), ms[i]) for i in range(k)]

    def debug_clers(self, i):
        cap = self.info[i][1] * 4 + 64
        buf = np.zeros(cap, dtype=np.uint8)
        n = C.c_uint32()
        _check(lib().crt_batch_debug_clers(self._h, i, buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
        return buf[:n.value].copy()

    def debug_prediction(self, i):
        out = np.zeros((self.info[i][0], 3), dtype=np.uint32)
        _check(lib().crt_batch_debug_prediction(self._h, i, out.ctypes.data_as(C.c_void_p)))
        return out

    def mesh_outputs(self, i):
        """Slices of the arenas belonging to mesh i, as numpy arrays (D2H)."""
        v0, v1, f0, f1 = self.vert_base[i], self.vert_base[i + 1], self.face_base[i], self.face_base[i + 1]
        res = {}
        for k, x in self.out.items():
            sl = x[f0:f1] if k == "index" else x[v0:v1]
            a = sl.cpu().numpy()
            if k == "index":
                a = a.view(np.uint16 if self.index16 else np.uint32)
            res[k] = a
        return res


def walk_tape(blob):
    """(tape, nvert, nface, nattr) of one HOST blob: the bytes the header parse + directory walk read (no payload)."""
    b = _aligned_copy(blob)
    nv, nf, na = C.c_uint32(), C.c_uint32(), C.c_uint32()
    buf = np.empty(4096, dtype=np.uint8)
    n = lib().crt_walk_tape(b.ctypes.data_as(C.c_void_p), len(b), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(nv), C.byref(nf), C.byref(na))
    if n < 0:
        raise CortoError(n, lib().crt_last_error().decode(errors="replace"))
    if n > buf.size:
        buf = np.empty(n, dtype=np.uint8)
        n = lib().crt_walk_tape(b.ctypes.data_as(C.c_void_p), len(b), buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(nv), C.byref(nf), C.byref(na))
    return buf[:n].copy(), nv.value, nf.value, na.value


def shard_lpt(nvert, nface, nattr, world):
    """Longest-processing-time assignment of blobs to ranks.  Returns an int array of ranks."""
    n = len(nvert)
    a = np.ascontiguousarray(nvert, dtype=np.uint32)
    f = np.ascontiguousarray(nface, dtype=np.uint32)
    t = np.ascontiguousarray(nattr, dtype=np.uint32)
    out = np.zeros(n, dtype=np.int32)
    _check(lib().crt_shard_lpt(n, a.ctypes.data_as(C.c_void_p), f.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), world,
                               out.ctypes.data_as(C.c_void_p)))
    return out
