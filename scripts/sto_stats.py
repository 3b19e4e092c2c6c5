"""Statistics of the warp storage solver (needs a -DDOPF_STATS build): python scripts/sto_stats.py workload gs ws it,it,..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
import numpy as np
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1]; gs = float(sys.argv[2]); ws = float(sys.argv[3]); ats = [int(x) for x in sys.argv[4].split(",")]
prob, cfg = bench.make_case(pkg, wl, 0)
A = prob.G + prob.S
dev = DeviceADMM(prob, device=0, hinge_capacity=64, gamma=gs / A, flow_weight=ws / A)
names = ["storages", "rounds", "passes", "anchors", "free_steps", "newton_cap", "gave_up", "all_anchored", "all_clipped", "one_round", "one_round_le2_passes", "gt8_rounds", "max_rounds", "max_passes"]
done = 0
for at in ats:
    if at - 1 > done:
        dev.step(at - 1 - done); done = at - 1
    dev.debug_counters(True)
    dev.step(1); done += 1
    c = dev.debug_counters(True)
    for base, lab in ((0, "predict"), (16, "fix")):
        n = max(c[base], 1)
        print("it %d %s: " % (at, lab) + ", ".join("%s %d" % (names[i], c[base + i]) for i in range(len(names))) + " | per storage: rounds %.2f passes %.2f anchors %.1f free %.1f" % (c[base + 1] / n, c[base + 2] / n, c[base + 3] / n, c[base + 4] / n))
    it = dev.get_iterate(("D", "C", "E", "flow", "avgU", "avgK"))
    g2w = gs / (2 * ws)
    bp = prob.fmax[:, None] - it["flow"] + g2w * it["avgU"]; bm = prob.fmax[:, None] + it["flow"] + g2w * it["avgK"]
    print("   anchored hinge rows (W != 0): %.2f%% of L*T; rows with any: per t %.1f of %d lines" % (100 * ((bp < 0) | (bm < 0)).mean(), ((bp < 0) | (bm < 0)).sum(0).mean(), prob.L))
    idle = (np.abs(it["D"]).max(1) + np.abs(it["C"]).max(1)) == 0
    act = (it["D"] > 0) | (it["C"] > 0)
    print("   idle storages %d of %d; active steps per non-idle storage %.1f of %d; steps at E=0: %.1f%%, at emax: %.1f%%" % (idle.sum(), prob.S, act[~idle].sum(1).mean() if (~idle).any() else 0, prob.T,
          100 * (it["E"] <= 1e-9).mean(), 100 * (it["E"] >= prob.sto_emax[:, None] - 1e-9).mean()), flush=True)
