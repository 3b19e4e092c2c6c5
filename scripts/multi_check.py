"""torchrun --nproc-per-node N scripts/multi_check.py : agent-partitioned run (graph-captured) == single-GPU run."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200.device import DeviceADMM
from dopf_b200 import multi
local = int(os.environ.get("LOCAL_RANK", "0")); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
for (N, L, G, S, T, iters, graph) in [(40, 60, 200, 40, 24, 30, True), (118, 186, 1000, 200, 24, 25, True), (118, 186, 1000, 200, 24, 25, False), (2000, 3000, 20000, 5000, 96, 12, True)]:
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=1); prob = pkg.Problem.from_arrays(d); A = G + S
    cfg = dict(gamma=0.03 / A, flow_weight=1.0 / A)
    part = multi.PartitionedADMM(prob, rank, world, local, hinge_capacity=64, graph=graph, **cfg)
    dev, gi, si = part.dev, part.gen_index, part.sto_index
    torch.cuda.synchronize(); t0 = time.time(); st = part.step(iters); torch.cuda.synchronize(); dt = time.time() - t0
    assert st.iterations_done == iters, (st.iterations_done, iters)
    assert (part.graph is not None) == graph, part.graph_error
    it = dev.get_iterate(); lam, mu, rho = dev.get_duals(0)
    if rank == 0:
        ref = DeviceADMM(prob, device=local, hinge_capacity=64, **cfg); ref.step(iters)
        rit = ref.get_iterate(); rl, rm, rr = ref.get_duals(0)
        err = dict(P=np.abs(it["P"] - rit["P"][gi]).max(), D=np.abs(it["D"] - rit["D"][si]).max() if len(si) else 0.0, inj=np.abs(it["injection"] - rit["injection"]).max(),
                   flow=np.abs(it["flow"] - rit["flow"]).max(), avgU=np.abs(it["avgU"] - rit["avgU"]).max(), lam=np.abs(lam - rl).max(), mu=np.abs(mu - rm).max(), rho=np.abs(rho - rr).max())
        print((N, L, G, S, T), f"world {world} graph {graph}: {dt / iters * 1e3:.3f} ms/iter (1 GPU: {ref.status.last_step_ms / iters:.3f}) max abs diff vs single GPU:", {k: float('%.2e' % v) for k, v in err.items()}, flush=True)
        assert max(err.values()) < 1e-6 * max(1.0, np.abs(rit["flow"]).max())
    part.close()
    torch.cuda.synchronize()
    dist.barrier()
# the same partition with the collectives inside libdopf (dopf_comm_init + the ordinary dopf_step; the host only broadcasts the NCCL id)
for (N, L, G, S, T, iters) in [(118, 186, 1000, 200, 24, 25), (2000, 3000, 20000, 5000, 96, 12)]:
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=1); prob = pkg.Problem.from_arrays(d); A = G + S
    cfg = dict(gamma=0.03 / A, flow_weight=1.0 / A)
    com = multi.LibraryCommADMM(prob, rank, world, local, hinge_capacity=64, **cfg)
    com.step(1); st = com.step(iters - 1)
    assert st.iterations_done == iters, (st.iterations_done, iters)
    ms = st.last_step_ms / (iters - 1)
    it = com.dev.get_iterate(); lam, mu, rho = com.dev.get_duals(0); gi, si = com.gen_index, com.sto_index
    if rank == 0:
        ref = DeviceADMM(prob, device=local, hinge_capacity=64, **cfg); ref.step(iters)
        rit = ref.get_iterate(); rl, rm, rr = ref.get_duals(0)
        err = dict(P=np.abs(it["P"] - rit["P"][gi]).max(), D=np.abs(it["D"] - rit["D"][si]).max() if len(si) else 0.0, inj=np.abs(it["injection"] - rit["injection"]).max(),
                   flow=np.abs(it["flow"] - rit["flow"]).max(), avgU=np.abs(it["avgU"] - rit["avgU"]).max(), lam=np.abs(lam - rl).max(), mu=np.abs(mu - rm).max(), rho=np.abs(rho - rr).max())
        print((N, L, G, S, T), f"world {world} library-owned NCCL (dopf_comm_init): {ms:.3f} ms/iter device time (1 GPU: {ref.status.last_step_ms / iters:.3f}) max abs diff vs single GPU:", {k: float('%.2e' % v) for k, v in err.items()}, flush=True)
        assert max(err.values()) < 1e-6 * max(1.0, np.abs(rit["flow"]).max())
        ref.close()
    com.close()
    torch.cuda.synchronize()
    dist.barrier()
sys.stdout.flush()
os._exit(0)       # a process group that carried a captured NCCL graph blocks in its destructor
