"""Does the ADMM reach the reference's stop rule on synthetic cases, and does the converged point match the central LP?
    python scripts/converge_probe.py N L G S T max_iters gs:ws [gs:ws ...]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
from tests import central_lp
N, L, G, S, T, maxit = [int(x) for x in sys.argv[1:7]]
d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=0, congest_frac=0.0)
prob = pkg.Problem.from_arrays(d); A = G + S
t0 = time.time(); lp = central_lp.solve(prob); print("central LP status %s objective %.6f (%.1fs)" % (lp["status"], lp.get("objective", float("nan")), time.time() - t0), flush=True)
for pair in sys.argv[7:]:
    gs, ws = [float(x) for x in pair.split(":")]
    dev = DeviceADMM(prob, device=0, hinge_capacity=64, gamma=gs / A, flow_weight=ws / A)
    done = 0; ms = 0.0
    while done < maxit:
        st = dev.step(min(20000, maxit - done)); done = st.iterations_done; ms += st.last_step_ms
        print("  gamma=%g/A w=%g/A it %d res %.2e %.2e %.2e%s" % (gs, ws, done, st.res_lambda, st.res_mue, st.res_rho, " CONVERGED at iteration %d" % st.iteration if st.converged else ""), flush=True)
        if st.converged:
            break
    it = dev.get_iterate()
    cost = dev.total_costs()
    if lp["status"] == 0:
        print("  -> device time %.1f ms; total cost %.4f vs LP %.4f (rel %.2e); max|P-P_lp|/max P %.2e; max|flow-flow_lp| %.3g; imbalance %.2e; max overload %.3g" % (
            ms, cost, lp["objective"], abs(cost - lp["objective"]) / lp["objective"], np.abs(it["P"] - lp["P"]).max() / np.abs(lp["P"]).max(),
            np.abs(it["flow"] - lp["flow"]).max(), np.abs(it["injection"].sum(0)).max(), (np.abs(it["flow"]) - prob.fmax[:, None]).max()), flush=True)
    dev.close()
