"""Structure-of-arrays packing of a case: what crosses the C ABI in `dopf_problem`."""
from dataclasses import dataclass

import numpy as np

from .ptdf import calculate_ptdf


@dataclass
class Problem:
    N: int
    L: int
    T: int
    G: int
    S: int
    ptdf: np.ndarray      # [L,N] row-major
    fmax: np.ndarray      # [L]
    demand: np.ndarray    # [N,T]
    gen_mc: np.ndarray
    gen_pmax: np.ndarray
    gen_node: np.ndarray  # int32, 0-based
    sto_mc: np.ndarray
    sto_pmax: np.ndarray
    sto_emax: np.ndarray
    sto_node: np.ndarray

    @staticmethod
    def from_arrays(d):
        f = lambda k, shape: np.ascontiguousarray(np.asarray(d[k], dtype=np.float64).reshape(shape))
        i = lambda k, shape: np.ascontiguousarray(np.asarray(d[k], dtype=np.int32).reshape(shape))
        N, L, T, G, S = (int(d[k]) for k in "NLTGS")
        return Problem(N, L, T, G, S, f("ptdf", (L, N)), f("fmax", (L,)), f("demand", (N, T)),
                       f("gen_mc", (G,)), f("gen_pmax", (G,)), i("gen_node", (G,)),
                       f("sto_mc", (S,)), f("sto_pmax", (S,)), f("sto_emax", (S,)), i("sto_node", (S,)))

    @staticmethod
    def from_structs(nodes, generators, storages, lines):
        """What ADMM(...) derives from the structs (structures/admm.jl:29-60)."""
        idx = {id(n): k for k, n in enumerate(nodes)}
        T = len(nodes[0].demand)
        return Problem.from_arrays(dict(
            N=len(nodes), L=len(lines), T=T, G=len(generators), S=len(storages),
            ptdf=calculate_ptdf(nodes, lines), fmax=[l.max_capacity for l in lines],
            demand=[n.demand for n in nodes],
            gen_mc=[g.marginal_costs for g in generators], gen_pmax=[g.max_generation for g in generators],
            gen_node=[idx[id(g.node)] for g in generators],
            sto_mc=[s.marginal_costs for s in storages], sto_pmax=[s.max_power for s in storages],
            sto_emax=[s.max_level for s in storages], sto_node=[idx[id(s.node)] for s in storages]))
