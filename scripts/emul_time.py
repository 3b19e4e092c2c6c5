"""Sequential CPU run of the library's per-element device code (tests/host_emul) on a whole bench workload: seconds per iteration and the
correction counters (status = iteration, converged, gen_corrected, sto_corrected, tight rows, wide rows, error, sto_cold) - to be read beside
profiles/r2_transient_target.log.   python scripts/emul_time.py target"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
import bench
pkg = g.load_package()
from tests.host_emul import emul
from dopf_b200 import multi
wl = sys.argv[1]
prob, cfg = bench.make_case(pkg, wl, 0)
sub, gi, si = multi.shard_problem(prob, 0, 1)       # node-sorted
t0 = time.time()
e = emul.EmulADMM(sub, gamma=cfg["gamma"], flow_weight=cfg["flow_weight"], hcap=64)
print("create", time.time() - t0, flush=True)
for k in range(3):
    t0 = time.time(); e.iterate(); dt = time.time() - t0
    print(k, "iteration s", round(dt, 3), "status", e.status.tolist(), flush=True)
