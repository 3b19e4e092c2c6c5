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
    n_scen: int = 1       # > 1: batch of scenarios on one grid; demand [C,N,T], gen_mc/gen_pmax [C,G], sto_mc/sto_pmax/sto_emax [C,S]

    @staticmethod
    def from_arrays(d):
        f = lambda k, shape: np.ascontiguousarray(np.asarray(d[k], dtype=np.float64).reshape(shape))
        i = lambda k, shape: np.ascontiguousarray(np.asarray(d[k], dtype=np.int32).reshape(shape))
        N, L, T, G, S = (int(d[k]) for k in "NLTGS")
        C = int(d.get("n_scen", 1))
        lead = (C,) if C > 1 else ()
        return Problem(N, L, T, G, S, f("ptdf", (L, N)), f("fmax", (L,)), f("demand", lead + (N, T)),
                       f("gen_mc", lead + (G,)), f("gen_pmax", lead + (G,)), i("gen_node", (G,)),
                       f("sto_mc", lead + (S,)), f("sto_pmax", lead + (S,)), f("sto_emax", lead + (S,)), i("sto_node", (S,)), C)

    def scenario(self, c):
        """the single problem of scenario c of a batch"""
        if self.n_scen == 1:
            return self
        return Problem(self.N, self.L, self.T, self.G, self.S, self.ptdf, self.fmax, np.ascontiguousarray(self.demand[c]),
                       np.ascontiguousarray(self.gen_mc[c]), np.ascontiguousarray(self.gen_pmax[c]), self.gen_node,
                       np.ascontiguousarray(self.sto_mc[c]), np.ascontiguousarray(self.sto_pmax[c]), np.ascontiguousarray(self.sto_emax[c]), self.sto_node, 1)

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
