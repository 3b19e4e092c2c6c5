"""A/B of library variants on the device clock, one line per library (DOPF_LIB selects it):
   mean ms of iterations 1-5 / 6-25 from the cold start (the driver's window), steady chunks, per-kernel split afterwards.
   python scripts/ab2.py [workload] [kernel-substring ...]"""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1] if len(sys.argv) > 1 else "target"
show = sys.argv[2:] or ["k_sto_warp"]
prob, cfg = bench.make_case(pkg, wl, 0)
dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
ms = []
for k in range(25):
    dev.step(1); ms.append(dev.status.last_step_ms)
dev.step(5); w = dev.status.last_step_ms / 5
dev.step(20); w2 = dev.status.last_step_ms / 20        # the driver's window: iterations 6..25 replayed as ONE graph batch is not possible here; see bench
dev.step(50)
ch = []
for c in range(3):
    dev.step(50); ch.append(dev.status.last_step_ms / 50)
prof = dev.profile_iteration()
kern = {}
for name, t in prof:
    kern[name] = kern.get(name, 0.0) + t
pick = {k: round(v, 4) for k, v in kern.items() if any(s in k for s in show)}
print("%-22s %s: it1-5 %.3f it6-25 %.3f | it101-250 %s | sum_kernels %.3f %s" % (
    os.environ.get("DOPF_LIB", "tree").split("/")[-1], wl, sum(ms[:5]) / 5, sum(ms[5:25]) / 20,
    " ".join("%.4f" % m for m in ch), sum(kern.values()), json.dumps(pick)))
dev.close()
for at in [int(x) for x in os.environ.get("AB_PROF_AT", "").split(",") if x]:
    dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
    dev.step(at - 1)
    kern = {}
    for name, t in dev.profile_iteration():
        kern[name] = kern.get(name, 0.0) + t
    print("   iteration %d: sum %.3f %s" % (at, sum(kern.values()), json.dumps({k: round(v, 4) for k, v in kern.items() if any(s in k for s in show)})))
    dev.close()
