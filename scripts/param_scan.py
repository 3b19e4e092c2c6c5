"""Stability / cost of the ADMM parameters on a workload: gamma = gs/A, flow_weight = ws/A.
    python scripts/param_scan.py workload iters gs:ws gs:ws ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1]; iters = int(sys.argv[2])
prob, cfg = bench.make_case(pkg, wl, 0)
A = prob.G + prob.S
for pair in sys.argv[3:]:
    gs, ws = [float(x) for x in pair.split(":")]
    dev = DeviceADMM(prob, device=0, hinge_capacity=64, gamma=gs / A, flow_weight=ws / A)
    out = []
    g0 = s0 = 0
    chunk = max(1, iters // 8)
    for c in range(8):
        try:
            dev.step(chunk)
        except Exception as e:
            out.append("ERR " + str(e)[:80]); break
        st = dev.status
        out.append("it %d: %.3f ms res %.2e %.2e %.2e genfix/it %.0f stofix/it %.0f tight %d%s" % (st.iterations_done, st.last_step_ms / chunk, st.res_lambda, st.res_mue, st.res_rho,
                   (st.gen_corrected - g0) / chunk, (st.sto_corrected - s0) / chunk, st.tight_rows, " CONVERGED" if st.converged else ""))
        g0, s0 = st.gen_corrected, st.sto_corrected
        if st.converged:
            break
    print("gamma=%g/A w=%g/A (w/gamma=%.1f):" % (gs, ws, ws / gs)); print("   " + "\n   ".join(out), flush=True)
    dev.close()
