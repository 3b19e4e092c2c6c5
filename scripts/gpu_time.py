"""Timing of synthetic cases after a warm-up (scratch tool): python scripts/gpu_time.py [warm] [iters]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
warm = int(sys.argv[1]) if len(sys.argv) > 1 else 40
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for (N, L, G, S, T) in [(118, 186, 1000, 200, 24), (2000, 3000, 20000, 5000, 96), (2000, 3000, 80000, 20000, 96)]:
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=0); p = pkg.Problem.from_arrays(d); A = G + S
    dev = DeviceADMM(p, gamma=0.3 / A, flow_weight=1.0 / A, device=0, hinge_capacity=64)
    for chunk in range(3):
        g0, s0 = dev.status.gen_corrected, dev.status.sto_corrected
        t3 = time.time(); dev.step(warm if chunk == 0 else iters); t4 = time.time()
        n = warm if chunk == 0 else iters
        st = dev.status
        print((N, L, G, S, T), "chunk", chunk, "%.3f ms/iter; units/s %.3e; fix/iter gen %.0f sto %.1f rows %d/%d res %s" % (
            (t4 - t3) / n * 1e3, A * T * n / (t4 - t3), (st.gen_corrected - g0) / n, (st.sto_corrected - s0) / n, st.tight_rows, st.wide_rows,
            tuple(round(x, 4) for x in (st.res_lambda, st.res_mue, st.res_rho))), flush=True)
    dev.close()
