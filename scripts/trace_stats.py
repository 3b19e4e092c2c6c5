"""Per-chunk statistics of a long run: python scripts/trace_stats.py workload chunks iters"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1]; chunks = int(sys.argv[2]); iters = int(sys.argv[3])
prob, cfg = bench.make_case(pkg, wl, 0)
dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
g0 = s0 = q0 = 0
for c in range(chunks):
    dev.step(iters); st = dev.status
    print("it %4d: %.4f ms/iter  gen fixes/iter %7.1f  sto fixes/iter %6.1f  seq fallbacks/iter %5.2f  cold(last) %d  rows %d/%d  res %.4g %.4g %.4g conv %d" % (
        st.iterations_done, st.last_step_ms / iters, (st.gen_corrected - g0) / iters, (st.sto_corrected - s0) / iters, (st.fix_sequential - q0) / iters,
        st.sto_cold, st.tight_rows, st.wide_rows, st.res_lambda, st.res_mue, st.res_rho, st.converged), flush=True)
    g0, s0, q0 = st.gen_corrected, st.sto_corrected, st.fix_sequential
