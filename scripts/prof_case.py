"""Run a few iterations of a synthetic case without the CUDA graph (for ncu / timing):
    python scripts/prof_case.py N L G S T iters [warm] [gamma_scale] [w_scale]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
N, L, G, S, T, iters = [int(x) for x in sys.argv[1:7]]
warm = int(sys.argv[7]) if len(sys.argv) > 7 else 0
gs = float(sys.argv[8]) if len(sys.argv) > 8 else 0.03
ws = float(sys.argv[9]) if len(sys.argv) > 9 else 1.0
d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=0); p = pkg.Problem.from_arrays(d); A = G + S
dev = DeviceADMM(p, gamma=gs / A, flow_weight=ws / A, device=0, hinge_capacity=64, use_graph=False)
if warm: dev.step(warm)
t0 = time.time(); dev.step(iters); t1 = time.time()
st = dev.status
print("ms/iter %.3f fix %d %d rows %d/%d" % ((t1 - t0) / iters * 1e3, st.gen_corrected, st.sto_corrected, st.tight_rows, st.wide_rows))
