"""Multi-GPU helpers: agent-block partition of a case over ranks (one process per GPU).

Every rank keeps the full network data (PTDF, limits, demand) and a contiguous, node-sorted block of
the generators and of the storages (so a storage's whole horizon stays on one GPU).  Per iteration the
ranks all-reduce three device buffers of libdopf in place (the caller owns the collective:
torch.distributed / NCCL on the stream the library runs on): the per-timestep maximum move, the nodal
injection and the exact slack row sums - SURVEY.md section 8(e) "agent block".
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


class _DevBuf:
    """wraps a raw device pointer so that torch can view it (__cuda_array_interface__)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8", data=(int(ptr), False), version=2)


class PartitionedADMM:
    """One rank of an agent-partitioned run.  libdopf runs the four phases of an iteration on a torch CUDA stream;
    torch.distributed (NCCL) all-reduces the exchange buffers in place between them.  All exchange buffers are fixed
    device addresses, so one whole iteration - the library's kernels AND the three collectives - is captured once in a
    CUDA graph and replayed: no interpreter / launch overhead per iteration (`graph=False` keeps the eager path)."""

    def __init__(self, prob: Problem, rank, world, device, dist=None, graph=True, **cfg):
        import torch
        import torch.distributed as tdist
        from .device import DeviceADMM
        self.torch, self.dist = torch, dist or tdist
        self.rank, self.world = rank, world
        self.full = prob
        self.sub, self.gen_index, self.sto_index = shard_problem(prob, rank, world)
        self.dev = DeviceADMM(self.sub, device=device, use_graph=False, **cfg)
        self.stream = torch.cuda.Stream(device=device)
        self.want_graph, self.graph, self.graph_error = bool(graph), None, None
        d = self.dev
        d._check(d.lib.dopf_set_stream(d.h, C.c_void_p(self.stream.cuda_stream)), "dopf_set_stream")
        with torch.cuda.stream(self.stream):
            d._check(d.lib.dopf_set_partition(d.h, rank, world, prob.G + prob.S), "dopf_set_partition")
            self.bufs = {w: self._buffer(w) for w in (0, 1, 2, 3)}     # fixed addresses for the lifetime of the handle
            self._allreduce(1, self.dist.ReduceOp.SUM)     # initial injection of the agents (the library subtracts the demand)
            self._allreduce(3, self.dist.ReduceOp.MAX)     # per-node box ranges -> identical candidate rows on all ranks
            d._check(d.lib.dopf_step_phase(d.h, -1), "dopf_step_phase")
        self.stream.synchronize()

    def _buffer(self, which):
        ptr, n = C.c_void_p(), C.c_int64()
        self.dev._check(self.dev.lib.dopf_exchange_buffer(self.dev.h, which, C.byref(ptr), C.byref(n)), "dopf_exchange_buffer")
        return self.torch.as_tensor(_DevBuf(ptr.value, n.value), device=f"cuda:{self.torch.cuda.current_device()}")

    def _allreduce(self, which, op):
        if self.world > 1:
            self.dist.all_reduce(self.bufs[which], op=op)

    def _enqueue_iteration(self):
        d, lib, R = self.dev, self.dev.lib, self.dist.ReduceOp
        d._check(lib.dopf_step_phase(d.h, 0), "phase 0"); self._allreduce(0, R.MAX)
        d._check(lib.dopf_step_phase(d.h, 1), "phase 1"); self._allreduce(1, R.SUM)
        d._check(lib.dopf_step_phase(d.h, 2), "phase 2"); self._allreduce(2, R.SUM)
        d._check(lib.dopf_step_phase(d.h, 3), "phase 3")

    def _capture(self):
        """one iteration (kernels + collectives) -> CUDA graph; falls back to eager stepping if the capture fails"""
        torch = self.torch
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode="thread_local"):
                self._enqueue_iteration()
            self.graph = g
        except Exception as e:      # keep running eagerly, remember why
            self.graph, self.graph_error, self.want_graph = None, repr(e), False
            torch.cuda.synchronize()

    def step(self, iters=1, check_every=16):
        d, lib, R = self.dev, self.dev.lib, self.dist.ReduceOp
        done = 0
        with self.torch.cuda.stream(self.stream):
            if self.want_graph and self.graph is None and iters > 1:
                self._enqueue_iteration(); done += 1          # one eager iteration first (communicator warm-up)
                self.stream.synchronize()
                self._capture()
            while done < iters:
                n = min(check_every, iters - done)
                for _ in range(n):
                    if self.graph is not None:
                        self.graph.replay()
                    else:
                        self._enqueue_iteration()
                done += n
                rc = lib.dopf_get_status(d.h, C.byref(d.status))      # synchronises; < 0: device-side capacity error
                if self.world > 1:       # a failing rank must stop ALL ranks (the others would all-reduce its stale buffers)
                    flag = self.torch.tensor([rc], dtype=self.torch.int32, device=f"cuda:{self.torch.cuda.current_device()}")
                    self.dist.all_reduce(flag, op=R.MIN)
                    worst = int(flag.item())
                    if worst != 0 and rc == 0:
                        raise RuntimeError(f"partitioned run stopped: another rank reported rc={worst}")
                d._check(rc, "dopf_get_status")
                if d.status.converged:
                    break
        return d.status

    def close(self):
        self.graph = None
        self.dev.close()


class LibraryCommADMM:
    """One rank of an agent-partitioned run with the collectives INSIDE libdopf (dopf_comm_init, SURVEY.md 8(b) "n_gpus"): the
    host only hands the 128 opaque bytes of the NCCL id from rank 0 to the other ranks, then calls the ordinary dopf_step -
    the call surface a Julia driver uses.  `exchange_id(bytes_or_None) -> bytes` is any broadcast from rank 0
    (default: torch.distributed.broadcast_object_list on the default process group)."""

    def __init__(self, prob: Problem, rank, world, device, exchange_id=None, **cfg):
        from .device import DeviceADMM, DopfError
        self.rank, self.world, self.full = rank, world, prob
        self.sub, self.gen_index, self.sto_index = shard_problem(prob, rank, world)
        self.dev = DeviceADMM(self.sub, device=device, **cfg)
        d = self.dev
        idb = C.create_string_buffer(128)
        if rank == 0:
            rc = d.lib.dopf_comm_get_unique_id(idb)
            if rc != 0:
                raise DopfError(f"dopf_comm_get_unique_id rc={rc}: {d.lib.dopf_last_error(None).decode()}")
        if world > 1:
            if exchange_id is None:
                import torch.distributed as tdist

                def exchange_id(b):
                    box = [b]
                    tdist.broadcast_object_list(box, src=0)
                    return box[0]
            idb.raw = exchange_id(idb.raw if rank == 0 else None)
        d._check(d.lib.dopf_comm_init(d.h, idb, rank, world, prob.G + prob.S), "dopf_comm_init")

    def step(self, iters=1):
        return self.dev.step(iters)

    def close(self):
        self.dev.close()


class LocalPartitionGroup:
    """`world` partition handles inside ONE process on ONE device, stepped in lockstep with the all-reduces done by
    plain torch ops on the exchange buffers.  It drives exactly the library code of the multi-GPU mode
    (dopf_set_partition / dopf_step_phase / dopf_exchange_buffer) without NCCL, so the partitioned path can be
    checked against a single handle on a one-GPU box (tests/test_gpu_partition.py)."""

    def __init__(self, prob: Problem, world, device=0, **cfg):
        import torch
        from .device import DeviceADMM
        self.torch, self.world, self.full = torch, world, prob
        self.stream = torch.cuda.Stream(device=device)
        self.members = []
        for r in range(world):
            sub, gi, si = shard_problem(prob, r, world)
            dev = DeviceADMM(sub, device=device, use_graph=False, **cfg)
            dev._check(dev.lib.dopf_set_stream(dev.h, C.c_void_p(self.stream.cuda_stream)), "dopf_set_stream")
            self.members.append((dev, gi, si))
        with torch.cuda.stream(self.stream):
            for r, (dev, _, _) in enumerate(self.members):
                dev._check(dev.lib.dopf_set_partition(dev.h, r, world, prob.G + prob.S), "dopf_set_partition")
            self._allreduce(1, "sum"); self._allreduce(3, "max")
            for dev, _, _ in self.members:
                dev._check(dev.lib.dopf_step_phase(dev.h, -1), "dopf_step_phase")
        self.stream.synchronize()

    def _buffers(self, which):
        out = []
        for dev, _, _ in self.members:
            ptr, n = C.c_void_p(), C.c_int64()
            dev._check(dev.lib.dopf_exchange_buffer(dev.h, which, C.byref(ptr), C.byref(n)), "dopf_exchange_buffer")
            out.append(self.torch.as_tensor(_DevBuf(ptr.value, n.value), device=f"cuda:{self.torch.cuda.current_device()}"))
        return out

    def _allreduce(self, which, op):
        bufs = self._buffers(which)
        st = self.torch.stack(bufs)
        red = st.sum(0) if op == "sum" else st.max(0).values
        for b in bufs:
            b.copy_(red)

    def step(self, iters=1):
        with self.torch.cuda.stream(self.stream):
            for _ in range(iters):
                for phase, (which, op) in enumerate(((0, "max"), (1, "sum"), (2, "sum"), (None, None))):
                    for dev, _, _ in self.members:
                        dev._check(dev.lib.dopf_step_phase(dev.h, phase), f"phase {phase}")
                    if which is not None:
                        self._allreduce(which, op)
        for dev, _, _ in self.members:
            dev._check(dev.lib.dopf_get_status(dev.h, C.byref(dev.status)), "dopf_get_status")
        return self.members[0][0].status

    def close(self):
        for dev, _, _ in self.members:
            dev.close()


