"""Fraction of idle storages (D = C = 0 over the whole horizon): python scripts/count_idle.py workload"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1] if len(sys.argv) > 1 else "target"
prob, cfg = bench.make_case(pkg, wl, 0)
dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
for it in (5, 30, 60, 100, 200):
    dev.step(it - dev.status.iterations_done)
    r = dev.get_iterate(want=("D", "C", "E"))
    idle = (np.abs(r["D"]).sum(axis=1) == 0) & (np.abs(r["C"]).sum(axis=1) == 0)
    act = (np.abs(r["D"]) + np.abs(r["C"]) > 0).sum(axis=1)
    print(f"iteration {it}: idle storages {int(idle.sum())} of {idle.size}; active timesteps per non-idle storage: median {np.median(act[~idle]) if (~idle).any() else 0}; E>0 anywhere {(r['E'].max(axis=1) > 0).sum()}")
