import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import corto_b200
from oracle import refshim, workloads, meshgen as mg
distinct = [workloads._c4(s) for s in (3, 4, 5, 7, 11, 12, 13, 20)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 320
order = [i % len(distinct) for i in range(n)]
bd = corto_b200.BatchDecoder([distinct[k] for k in order], color_components=4)
bd.allocate(fill=0xA5)
bd.upload(); bd.decode()
torch.cuda.synchronize()
print(bd.status())
