"""Cold-start transient of a workload: device time of every iteration 1..n and the per-kernel split at chosen iterations.
    python scripts/transient.py workload wscale n_iters prof_at,prof_at,...
(wscale: flow_weight = wscale/A; 10 = the reference's w/gamma ratio)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1]; wscale = float(sys.argv[2]); n = int(sys.argv[3])
prof_at = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 and sys.argv[4] else []
prob, cfg = bench.make_case(pkg, wl, 0)
A = prob.G + prob.S
cfg["flow_weight"] = wscale / A
if len(sys.argv) > 5: cfg["gamma"] = float(sys.argv[5]) / A
dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
g0 = s0 = q0 = 0
rows = []
for k in range(1, n + 1):
    dev.step(1); st = dev.status
    rows.append(dict(it=k, ms=round(st.last_step_ms, 4), gen_fix=st.gen_corrected - g0, sto_fix=st.sto_corrected - s0, seq=st.fix_sequential - q0,
                     cold=st.sto_cold, tight=st.tight_rows, wide=st.wide_rows, res=[float("%.3g" % x) for x in (st.res_lambda, st.res_mue, st.res_rho)]))
    g0, s0, q0 = st.gen_corrected, st.sto_corrected, st.fix_sequential
for r in rows:
    print(json.dumps(r))
ms = [r["ms"] for r in rows]
print("mean ms it 1-5 %.3f | 6-25 %.3f | 26-%d %.3f" % (sum(ms[:5]) / 5, sum(ms[5:25]) / max(1, len(ms[5:25])), n, sum(ms[25:]) / max(1, len(ms[25:]))))
dev.close()
for at in prof_at:
    dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
    if at > 1:
        dev.step(at - 1)
    prof = dev.profile_iteration()
    kern = {}
    for name, t in prof:
        kern[name] = kern.get(name, 0.0) + t
    print("profile of iteration %d (sum %.3f ms):" % (at, sum(kern.values())), json.dumps({k: round(v, 4) for k, v in sorted(kern.items(), key=lambda x: -x[1])}))
    dev.close()
