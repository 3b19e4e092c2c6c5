"""Per-kernel CUDA-event times of one steady-state iteration: python scripts/prof_kernels.py workload [warm]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
wl = sys.argv[1] if len(sys.argv) > 1 else "target"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 60
prob, cfg = bench.make_case(pkg, wl, 0)
dev = DeviceADMM(prob, device=0, hinge_capacity=64, **cfg)
dev.step(warm)
for rep in range(3):
    prof = dev.profile_iteration()
    tot = sum(ms for _, ms in prof)
    print(f"--- {wl} iteration {dev.iteration}: total {tot*1e3:.1f} us; cold {dev.status.sto_cold} fixes(cum) {dev.status.gen_corrected}/{dev.status.sto_corrected} (sequential {dev.status.fix_sequential}) rows {dev.status.tight_rows}/{dev.status.wide_rows}")
    if rep == 2 or (len(sys.argv) > 3 and sys.argv[3] == 'all'):
        for name, ms in prof:
            print(f"   {name:24s} {ms*1e3:9.1f} us  {100*ms/tot:5.1f}%")
    dev.step(10)
