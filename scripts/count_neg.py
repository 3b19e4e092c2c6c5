"""How many (line, t) rows have an always-active slack hinge (b < 0 => W != 0)?  python scripts/count_neg.py workload iters"""
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
for it in (5, 20, 60, 100):
    dev.step(it - dev.status.iterations_done)
    r = dev.get_iterate()
    F, U, K = r["flow"], r["avgU"], r["avgK"]
    g2w = cfg["gamma"] / (2.0 * cfg["flow_weight"])
    bp = (prob.fmax[:, None] - F) + g2w * U
    bm = (prob.fmax[:, None] + F) + g2w * K
    neg = (bp < 0) | (bm < 0)
    print(f"iteration {it}: rows with W != 0: {int(neg.sum())} of {neg.size}; max per t {int(neg.sum(axis=0).max())}; lines {int(neg.any(axis=1).sum())}")
