"""GPU robustness: corrupted payloads must end in a clean per-mesh status (or a clean decode), never in a hang, a fault or an
out-of-bounds write.  The reference has no bounds checks at all; here the host walk rejects structural damage
(tests/test_host.py) and the kernels clamp every index that comes out of the payload."""
import os

import numpy as np
import pytest

import corto_b200
from oracle import refshim

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _blob(name):
    return refshim.aligned_blob(open(os.path.join(GOLDEN, name + ".crt"), "rb").read())


@pytest.mark.timeout(240)
@pytest.mark.parametrize("name", ["grid_est", "torus", "groups4_border", "cloud_all", "const_color"])
def test_corrupted_payload_terminates(name):
    import torch
    clean = _blob(name)
    rs = np.random.RandomState(1234)
    blobs = []
    for k in range(24):
        b = clean.copy()
        lo = len(b) // 3                                  # past the header / group table: payload bytes only
        for _ in range(1 + k % 5):
            b[rs.randint(lo, len(b))] ^= np.uint8(rs.randint(1, 256))
        # keep only damage the host walk accepts (otherwise there is nothing for the GPU to do)
        try:
            corto_b200.Decoder(b)
            blobs.append(refshim.aligned_blob(b.tobytes()))
        except corto_b200.CortoError:
            pass
    ok = 0
    for b in blobs:
        try:
            bd = corto_b200.BatchDecoder([b, clean])
        except corto_b200.CortoError:
            continue                                      # rejected by the bounds-checked directory walk
        bd.allocate(fill=0x5A)
        bd.upload(); bd.decode()
        torch.cuda.synchronize()
        rc, st = bd.status()
        assert st[1] == 0                                 # the clean neighbour in the same batch is unaffected
        assert st[0] in (0, -5)
        ok += 1
    assert ok > 0
    # and the device is still healthy
    good = corto_b200.Decoder(clean).decode()
    assert good["nvert"] > 0
