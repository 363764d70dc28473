"""Freeze the SMALL fixture categories (tests/cases.py) into committed golden vectors.

Run where /root/reference exists (oracle/_ref built):   python tests/golden/make_golden.py
For every case writes  <name>.crt  (the blob produced by the reference Encoder) and  <name>.npz  (the arrays the
UNMODIFIED reference Decoder produced for it, for each decode variant: default, u16 index + i16 normals, colour 4,
colour 3).  Output buffers were pre-filled with 0xA5 bytes, so elements the reference leaves untouched are
part of the pin.  Also writes tarta.json: FNV-1a/64 digests of the reference decode of html/models/tarta.crt, the only
.crt the reference ships (the file itself is 5.7 MB and is copied to oracle/_ref/ by the oracle Makefile instead).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests import cases          # noqa: E402
from oracle import refshim, pyoracle   # noqa: E402


def vkey(var):
    return "_".join("%s%s" % (k, int(v)) for k, v in sorted(var.items())) or "default"


def main():
    for name, build in cases.SMALL:
        blob = build()
        with open(os.path.join(HERE, name + ".crt"), "wb") as f:
            f.write(blob.tobytes())
        info = pyoracle.info(blob)
        arrays = {}
        for var in cases.VARIANTS:
            if not cases.applicable(var, info["attrs"], info["nvert"], info["nface"]):
                continue
            out = refshim.decode(blob, **var)
            for k, v in out.items():
                if isinstance(v, np.ndarray):
                    arrays[vkey(var) + "/" + k] = v
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    if os.path.exists(refshim.TARTA):
        blob = refshim.aligned_blob(open(refshim.TARTA, "rb").read())
        out = refshim.decode(blob)
        dig = {k: "%016x" % pyoracle.fnv1a64(v) for k, v in out.items() if isinstance(v, np.ndarray)}
        dig.update(nvert=out["nvert"], nface=out["nface"], bytes=len(blob))
        json.dump(dig, open(os.path.join(HERE, "tarta.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(cases.SMALL), "golden cases to", HERE)


if __name__ == "__main__":
    main()
