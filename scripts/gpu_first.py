"""First GPU bring-up: parity vs the oracle on several cases + rough timings (scratch tool)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
from oracle import oracle

def cmp(dev, ora):
    it = dev.get_iterate(); lam, mu, rho = dev.get_duals(0)
    return {k: float(np.abs(a - b).max()) if a.size else 0.0 for k, a, b in [
        ("P", it["P"], ora.P), ("D", it["D"], ora.D), ("C", it["C"], ora.C), ("E", it["E"], ora.E), ("inj", it["injection"], ora.inj),
        ("F", it["flow"], ora.flow), ("U", it["avgU"], ora.avgU), ("K", it["avgK"], ora.avgK), ("lam", lam, ora.lam), ("mu", mu, ora.mu), ("rho", rho, ora.rho)]}

g.smoke()
prob = pkg.Problem.from_structs(*pkg.cases.three_node())
for name in ["TNS", "big_gamma", "wrong_weight"]:
    gd = np.load(f"tests/golden/{name}.npz")
    dev = DeviceADMM(prob, gamma=float(gd["gamma"]), flow_weight=float(gd["flow_weight"]), device=0)
    K = gd["P"].shape[0]; worst = 0.0; stop = None
    for k in range(K):
        st = dev.step(1)
        if st.converged: stop = k + 1; break
        it = dev.get_iterate(("P", "D", "C"))
        worst = max(worst, np.abs(it["P"] - gd["P"][k]).max(), np.abs(it["D"] - gd["D"][k]).max(), np.abs(it["C"] - gd["C"][k]).max())
    print(name, "GPU vs golden worst", worst, "stop", stop, "gen_fix", st.gen_corrected, "sto_fix", st.sto_corrected, flush=True)

for (N, L, G, S, T, gam, w, iters) in [(12, 18, 30, 8, 6, None, None, 40), (12, 18, 30, 8, 6, 0.02, 10.0, 25), (118, 186, 1000, 200, 24, None, None, 25)]:
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=3)
    p = pkg.Problem.from_arrays(d); A = G + S
    gam = gam or 0.3 / A; w = w or 1.0 / A
    dev = DeviceADMM(p, gamma=gam, flow_weight=w, device=0, hinge_capacity=64); ora = oracle.OracleADMM(p, gam, flow_weight=w)
    worst = {}
    for k in range(iters):
        dev.step(1); ora.iterate(0)
        for kk, vv in cmp(dev, ora).items(): worst[kk] = max(worst.get(kk, 0), vv)
    st = dev.status
    print((N, L, G, S, T), "worst", {k: float("%.1e" % v) for k, v in worst.items()}, "fix", st.gen_corrected, st.sto_corrected, "rows", st.tight_rows, st.wide_rows, flush=True)

# rough timing
for (N, L, G, S, T) in [(118, 186, 1000, 200, 24), (2000, 3000, 20000, 5000, 96), (2000, 3000, 80000, 20000, 96)]:
    t0 = time.time(); d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=0); p = pkg.Problem.from_arrays(d); A = G + S
    t1 = time.time(); dev = DeviceADMM(p, gamma=0.3 / A, flow_weight=1.0 / A, device=0, hinge_capacity=64); t2 = time.time()
    dev.step(5)
    t3 = time.time(); dev.step(50); t4 = time.time()
    st = dev.status
    print((N, L, G, S, T), "gen %.1fs create %.1fs; %.3f ms/iter; units/s %.3e; fix %d %d rows %d/%d launches %d res %s" % (
        t1 - t0, t2 - t1, (t4 - t3) / 50 * 1e3, A * T * 50 / (t4 - t3), st.gen_corrected, st.sto_corrected, st.tight_rows, st.wide_rows, st.launches_per_iteration, (st.res_lambda, st.res_mue, st.res_rho)), flush=True)
    dev.close()
