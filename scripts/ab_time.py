"""A/B timing on the device clock: python scripts/ab_time.py [workload] [warm] [iters] [chunks]  (DOPF_LIB selects the library)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1] if len(sys.argv) > 1 else "target"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 100
chunks = int(sys.argv[4]) if len(sys.argv) > 4 else 5
prob, cfg = bench.make_case(pkg, wl, 0)
dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
dev.step(warm)
ms = []
for c in range(chunks):
    dev.step(iters)
    ms.append(dev.status.last_step_ms / iters)
print("%s %s: ms/iter per chunk %s  min %.4f median %.4f" % (os.environ.get("DOPF_LIB", "tree").split("/")[-1], wl, " ".join("%.4f" % m for m in ms), min(ms), float(np.median(ms))))
