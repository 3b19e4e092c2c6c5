"""Stats build (-DDOPF_STATS), debug flag 128: start/end device timestamps of the straggler storages of k_sto_warp beside those
of the last storages drawn (~ the end of the kernel).  python scripts/strag_time.py it,it,..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import bench
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
ats = [int(x) for x in sys.argv[1].split(",")]
prob, cfg = bench.make_case(pkg, "target", 0)
dev = DeviceADMM(prob, device=0, hinge_capacity=64, debug_flags=128, **cfg)
dev.step(max(ats))
dev.close()
