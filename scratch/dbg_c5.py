import numpy as np, sys, os, faulthandler
faulthandler.enable()
sys.path.insert(0, os.getcwd())
from oracle import refshim, workloads
if len(sys.argv) > 1: refshim._SO = sys.argv[1]
blob = workloads._c5(1)
print("blob", len(blob), hash(blob.tobytes()) & 0xffffffff, flush=True)
import hashlib; print(hashlib.md5(blob.tobytes()).hexdigest(), flush=True)
w = refshim.decode(blob); print("decoded", w['nvert'], flush=True)
