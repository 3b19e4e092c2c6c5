"""Scenario batch timing: python scripts/batch_probe.py C iters  (118-node / 24-period scenarios, BASELINE configs[3])"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
C = int(sys.argv[1]); iters = int(sys.argv[2])
N, L, G, S, T = 118, 186, 1000, 200, 24
A = G + S
d = pkg.cases.synthetic_scenarios(N=N, L=L, G=G, S=S, T=T, n_scen=C, seed=0)
prob = pkg.Problem.from_arrays(d)
dev = DeviceADMM(prob, device=0, gamma=0.03 / A, flow_weight=1.0 / A, hinge_capacity=64)
for k in range(4):
    st = dev.step(iters)
    print("C=%d it %d: %.4f ms/iter -> %.3e agent*t/s (launches %d) fixes gen %d sto %d cold %d" % (C, st.iterations_done, st.last_step_ms / iters, C * A * T * iters / (st.last_step_ms * 1e-3), st.launches_per_iteration, st.gen_corrected, st.sto_corrected, st.sto_cold), flush=True)
prof = dev.profile_iteration()
kern = {}
for name, t in prof:
    kern[name] = kern.get(name, 0.0) + t
print("profile (sum %.3f ms):" % sum(kern.values()), json.dumps({k: round(v, 4) for k, v in sorted(kern.items(), key=lambda x: -x[1])}))
one = DeviceADMM(prob.scenario(0), device=0, gamma=0.03 / A, flow_weight=1.0 / A, hinge_capacity=64)
one.step(iters); st = one.step(iters)
print("single scenario: %.4f ms/iter -> serial %d scenarios = %.3f ms" % (st.last_step_ms / iters, C, C * st.last_step_ms / iters))
