"""The agent-partitioned (multi-GPU) code path of libdopf - dopf_set_partition / dopf_step_phase / dopf_exchange_buffer -
against a single handle.  On a one-GPU box the ranks are `world` handles in one process and the all-reduces are plain
torch reductions over the exchange buffers (multi.LocalPartitionGroup): every line of the library's phase / exchange
code runs, only NCCL itself is replaced.  With >= 2 GPUs the real torch.distributed/NCCL path is launched with torchrun.
pytest -m gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.abs(a - b).max() / max(1.0, np.abs(b).max()) if a.size else 0.0


@pytest.mark.parametrize("dims,world,wscale,iters", [((40, 60, 200, 40, 24), 2, 1.0, 30), ((40, 60, 200, 40, 24), 3, 10.0, 20),
                                                     ((118, 186, 1000, 200, 24), 4, 1.0, 25), ((300, 450, 3000, 600, 96), 2, 1.0, 12)])
def test_partitioned_handles_equal_single_handle(pkg, dims, world, wscale, iters):
    from dopf_b200 import multi
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T = dims
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=1)
    prob = pkg.Problem.from_arrays(d); A = G + S
    cfg = dict(gamma=0.3 / A, flow_weight=wscale / A, hinge_capacity=64)
    grp = multi.LocalPartitionGroup(prob, world, device=0, **cfg)
    ref = DeviceADMM(prob, device=0, **cfg)
    st = grp.step(iters); ref.step(iters)
    assert st.iterations_done == iters == ref.status.iterations_done and st.iteration == ref.status.iteration
    rit = ref.get_iterate(); rl, rm, rr = ref.get_duals(0)
    tol = 1e-9
    for r, (dev, gi, si) in enumerate(grp.members):
        it = dev.get_iterate(); lam, mu, rho = dev.get_duals(0)
        assert _rel(it["P"], rit["P"][gi]) < tol and _rel(it["D"], rit["D"][si]) < tol and _rel(it["C"], rit["C"][si]) < tol
        assert _rel(it["E"], rit["E"][si]) < tol
        for k in ("injection", "flow", "avgU", "avgK"):        # the network / dual part is replicated on every rank
            assert _rel(it[k], rit[k]) < tol, (r, k)
        assert _rel(lam, rl) < tol and _rel(mu, rm) < tol and _rel(rho, rr) < tol
        assert ((mu == 0) == (rm == 0)).all() and ((rho == 0) == (rr == 0)).all()
    assert ref.status.gen_corrected > 0
    assert sum(dev.status.gen_corrected for dev, _, _ in grp.members) == ref.status.gen_corrected
    grp.close(); ref.close()


def test_capacity_error_of_a_partition_surfaces(pkg):
    """a device-side capacity error of a partitioned handle must reach the host (dopf_get_status)"""
    from dopf_b200 import multi
    from dopf_b200.device import DopfError
    raised = 0
    for seed in range(6):
        d = pkg.cases.synthetic_arrays(N=12, L=18, G=30, S=8, T=6, seed=seed, congest_frac=0.5)
        prob = pkg.Problem.from_arrays(d)
        grp = multi.LocalPartitionGroup(prob, 2, device=0, gamma=0.02, flow_weight=10.0, hinge_capacity=1)
        try:
            grp.step(40)
        except DopfError as e:
            assert "capacity" in str(e)
            raised += 1
        grp.close()
    assert raised > 0


def test_library_owned_communicator_single_rank(pkg):
    """dopf_comm_init / dopf_step with the NCCL collectives issued by libdopf itself (one rank: every collective still runs,
    the iteration - kernels and ncclAllReduce - is captured in the library's CUDA graph) == a plain single handle"""
    from dopf_b200 import multi
    from dopf_b200.device import DeviceADMM
    N, L, G, S, T = 118, 186, 1000, 200, 24
    d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=2)
    prob = pkg.Problem.from_arrays(d); A = G + S
    cfg = dict(gamma=0.3 / A, flow_weight=1.0 / A, hinge_capacity=64)
    com = multi.LibraryCommADMM(prob, 0, 1, 0, **cfg)
    ref = DeviceADMM(prob, device=0, **cfg)
    st = com.step(1); st = com.step(29); ref.step(30)
    assert st.iterations_done == 30 == ref.status.iterations_done
    it = com.dev.get_iterate(); rit = ref.get_iterate()
    gi, si = com.gen_index, com.sto_index
    assert _rel(it["P"], rit["P"][gi]) < 1e-9 and _rel(it["D"], rit["D"][si]) < 1e-9 and _rel(it["E"], rit["E"][si]) < 1e-9
    for k in ("injection", "flow", "avgU", "avgK"):
        assert _rel(it[k], rit[k]) < 1e-9, k
    for a, b in zip(com.dev.get_duals(0), ref.get_duals(0)):
        assert _rel(a, b) < 1e-9
    com.close(); ref.close()


def test_two_gpu_nccl_run_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the same library path is covered on one GPU by LocalPartitionGroup above)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(ROOT, "scripts", "multi_check.py")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
