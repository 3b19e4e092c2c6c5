"""The multi-GPU path on the CPU: agent sharding; a world_size-2 gloo run of the partitioned mode's four phases and three
exchanges with the product's own per-element device code compiled for the host (tests/host_emul) against the
single-process run of the same code; and a gloo run in which the oracle stands in for the kernels (exchange semantics)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.conftest import ROOT


def test_shard_bounds_and_problem(pkg):
    from dopf_b200 import multi
    for count, world in ((10, 3), (7, 8), (100000, 8), (0, 2)):
        spans = [multi.shard_bounds(count, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == count
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    d = pkg.cases.synthetic_arrays(N=12, L=18, G=31, S=7, T=6, seed=3)
    rng = np.random.default_rng(1)
    perm = rng.permutation(31)
    for k in ("gen_mc", "gen_pmax", "gen_node"):
        d[k] = d[k][perm]                                   # unsorted input
    prob = pkg.Problem.from_arrays(d)
    seen_g, seen_s = [], []
    for r in range(3):
        sub, gi, si = multi.shard_problem(prob, r, 3)
        assert np.all(np.diff(sub.gen_node) >= 0) and np.all(np.diff(sub.sto_node) >= 0)     # node-sorted blocks
        assert np.array_equal(sub.gen_pmax, prob.gen_pmax[gi]) and sub.ptdf is prob.ptdf
        seen_g += list(gi); seen_s += list(si)
    assert sorted(seen_g) == list(range(31)) and sorted(seen_s) == list(range(7))


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["DOPF_ROOT"])
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200 import multi
from oracle import oracle
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = pkg.cases.synthetic_arrays(N=8, L=11, G=14, S=5, T=5, seed=2); prob = pkg.Problem.from_arrays(d); A = prob.G + prob.S
gamma, w = 0.3 / A, 1.0 / A
full = oracle.OracleADMM(prob, gamma, flow_weight=w)
sub, gi, si = multi.shard_problem(prob, rank, world)
inj = -prob.demand.copy()
for k in range(6):
    prev = {a: getattr(full, a).copy() for a in ("P", "D", "C")}
    full.iterate(0)                                   # every rank knows the replicated network/dual state
    # this rank's share of the exchange: the injection of ITS agents (rank 0 carries the demand)
    loc = -prob.demand.copy() if rank == 0 else np.zeros_like(prob.demand)
    np.add.at(loc, sub.gen_node, full.P[gi]); np.add.at(loc, sub.sto_node, full.D[si] - full.C[si])
    t = torch.from_numpy(loc.copy()); dist.all_reduce(t)                       # DOPF_XBUF_INJ
    assert np.abs(t.numpy() - full.inj).max() < 1e-9, "summed injection differs"
    mv = np.zeros(prob.T)
    if len(gi): mv = np.maximum(mv, np.abs(full.P[gi] - prev["P"][gi]).max(0))
    if len(si): mv = np.maximum(mv, np.abs((full.D[si] - prev["D"][si]) - (full.C[si] - prev["C"][si])).max(0))
    t = torch.from_numpy(mv.copy()); dist.all_reduce(t, op=dist.ReduceOp.MAX)  # DOPF_XBUF_DMAX
    allmv = np.maximum(np.abs(full.P - prev["P"]).max(0), np.abs((full.D - prev["D"]) - (full.C - prev["C"])).max(0))
    assert np.abs(t.numpy() - allmv).max() < 1e-12, "move maxima differ"
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_two_rank_gloo_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, DOPF_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


_WORKER_EMUL = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["DOPF_ROOT"])
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
from dopf_b200 import multi
from tests.host_emul import emul
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, L, G, S, T = 10, 14, 26, 7, 6
d = pkg.cases.synthetic_arrays(N=N, L=L, G=G, S=S, T=T, seed=4, congest_frac=0.4); prob0 = pkg.Problem.from_arrays(d); A = G + S
go = np.argsort(prob0.gen_node, kind="stable"); so = np.argsort(prob0.sto_node, kind="stable")
for wscale in (1.0, 10.0):
    cfg = dict(gamma=0.3 / A, flow_weight=wscale / A, hcap=64)
    # this rank's block of the node-sorted agents; the exchange buffers are all-reduced exactly like multi.PartitionedADMM does
    sub, gi, si = multi.shard_problem(prob0, rank, world)
    me = emul.EmulADMM(sub, **cfg)
    me.partition_init(A)
    def allreduce(which, op):
        t = torch.from_numpy(me.exchange_buffer(which))       # shares memory with the emulation's buffer
        dist.all_reduce(t, op=op)
    allreduce(3, dist.ReduceOp.MAX)
    me.partition_finish_setup()
    # single-process run of the same device math on the whole (node-sorted) case
    full = pkg.Problem(prob0.N, prob0.L, prob0.T, G, S, prob0.ptdf, prob0.fmax, prob0.demand,
                       np.ascontiguousarray(prob0.gen_mc[go]), np.ascontiguousarray(prob0.gen_pmax[go]), np.ascontiguousarray(prob0.gen_node[go]),
                       np.ascontiguousarray(prob0.sto_mc[so]), np.ascontiguousarray(prob0.sto_pmax[so]), np.ascontiguousarray(prob0.sto_emax[so]),
                       np.ascontiguousarray(prob0.sto_node[so]))
    ref = emul.EmulADMM(full, **cfg)
    g0, g1 = multi.shard_bounds(G, rank, world); s0, s1 = multi.shard_bounds(S, rank, world)
    fixes = 0
    for k in range(40):
        me.phase(0); allreduce(0, dist.ReduceOp.MAX)
        me.phase(1); allreduce(1, dist.ReduceOp.SUM)
        me.phase(2); allreduce(2, dist.ReduceOp.SUM)
        me.phase(3)
        me.fetch(); ref.iterate()
        rel = lambda a, b: np.abs(a - b).max() / max(1.0, np.abs(b).max()) if a.size else 0.0
        worst = max(rel(me.P, ref.P[g0:g1]), rel(me.D, ref.D[s0:s1]), rel(me.C, ref.C[s0:s1]), rel(me.E, ref.E[s0:s1]), rel(me.inj, ref.inj), rel(me.flow, ref.flow),
                    rel(me.avgU, ref.avgU), rel(me.avgK, ref.avgK), rel(me.lam, ref.lam), rel(me.mu, ref.mu), rel(me.rho, ref.rho))
        assert worst < 1e-9, (wscale, k, worst)
        assert ((me.mu == 0) == (ref.mu == 0)).all() and ((me.rho == 0) == (ref.rho == 0)).all()
        assert me.status[6] == 0 and me.iteration == ref.iteration
    t = torch.tensor([int(me.status[2]) + int(me.status[3])]); dist.all_reduce(t)
    assert int(t.item()) == int(ref.status[2]) + int(ref.status[3]) and int(t.item()) > 0, "the correction pass must have run"
    assert int(ref.status[4]) > 0, "tight rows (exact slack corrections) must have occurred"
print("rank", rank, "ok")
dist.destroy_process_group()
'''


def test_two_rank_gloo_partition_of_the_device_math(tmp_path):
    """The partitioned mode's algebra on the CPU with the product's own per-element device code (dopf_bodies.h compiled for the
    host, tests/host_emul): two gloo ranks each hold a block of the agents, run the four phases of dopf_step_phase and all-reduce
    the three exchange buffers (move maxima, agents' injection without the demand, slack-row corrections + partial flows);
    every rank must reproduce the single-process iterate, duals and slack masks."""
    from tests.host_emul import emul
    emul.build()
    script = tmp_path / "worker_emul.py"
    script.write_text(_WORKER_EMUL)
    env = dict(os.environ, DOPF_ROOT=ROOT, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29534", str(script)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
