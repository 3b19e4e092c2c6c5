"""Multi-GPU helpers: agent-block partition of a case over ranks (one process per GPU).

Every rank keeps the full network data (PTDF, limits, demand) and a contiguous, node-sorted block of
the generators and of the storages (so a storage's whole horizon stays on one GPU).  Per iteration the
ranks exchange (NCCL all-reduce inside libdopf): the per-timestep maximum move, the nodal injection
and the exact slack row sums - SURVEY.md section 8(e) "agent block".
"""
import ctypes as C

import numpy as np

from .problem import Problem


def shard_bounds(count, rank, world):
    """contiguous block [lo, hi) of `count` items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(count, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_problem(prob: Problem, rank: int, world: int) -> Problem:
    """the rank's share of the agents (sorted by node first), full network data"""
    go = np.argsort(prob.gen_node, kind="stable"); so = np.argsort(prob.sto_node, kind="stable")
    g0, g1 = shard_bounds(prob.G, rank, world); s0, s1 = shard_bounds(prob.S, rank, world)
    gi, si = go[g0:g1], so[s0:s1]
    return Problem(prob.N, prob.L, prob.T, len(gi), len(si), prob.ptdf, prob.fmax, prob.demand,
                   np.ascontiguousarray(prob.gen_mc[gi]), np.ascontiguousarray(prob.gen_pmax[gi]), np.ascontiguousarray(prob.gen_node[gi]),
                   np.ascontiguousarray(prob.sto_mc[si]), np.ascontiguousarray(prob.sto_pmax[si]), np.ascontiguousarray(prob.sto_emax[si]),
                   np.ascontiguousarray(prob.sto_node[si])), gi, si


def connect(dev, total_agents, dist=None):
    """join the ranks of a torch.distributed process group into one libdopf communicator"""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        rc = dev.lib.dopf_comm_unique_id(uid.ctypes.data_as(C.c_void_p))
        if rc != 0:
            raise RuntimeError("dopf_comm_unique_id failed")
    backend = dist.get_backend()
    t = torch.from_numpy(uid)
    if backend == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    uid = t.cpu().numpy()
    dev._check(dev.lib.dopf_comm_init(dev.h, rank, world, uid.ctypes.data_as(C.c_void_p), int(total_agents)), "dopf_comm_init")
    return rank, world
