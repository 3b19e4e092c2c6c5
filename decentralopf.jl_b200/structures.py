"""Input structures, kept as the reference defines them.

Mirror of /root/reference/src/structures/network_elements.jl:1-30 (Node, Generator, Storage,
Line - all numeric inputs are integers there; floats are accepted here as well) and the
result/convergence containers of src/structures/{results,convergence,penalty_terms}.jl.
Objects are compared by identity, like Julia's mutable structs used as Dict keys
(structures/admm.jl:18-20).
"""
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass(eq=False)
class Node:  # network_elements.jl:1-5
    name: str
    demand: List[float]
    slack: bool


@dataclass(eq=False)
class Generator:  # network_elements.jl:7-13
    name: str
    marginal_costs: float
    max_generation: float
    plot_color: str
    node: Node


@dataclass(eq=False)
class Storage:  # network_elements.jl:15-22
    name: str
    marginal_costs: float
    max_power: float
    max_level: float
    plot_color: str
    node: Node


@dataclass(eq=False)
class Line:  # network_elements.jl:24-30
    name: str
    from_: Node  # `from` is a Python keyword
    to: Node
    max_capacity: float
    susceptance: float


@dataclass(eq=False)
class Convergence:  # structures/convergence.jl:1-20
    lambda_: bool = False
    lambda_res: list = field(default_factory=list)
    mue: bool = False
    mue_res: list = field(default_factory=list)
    rho: bool = False
    rho_res: list = field(default_factory=list)
    all: bool = False


@dataclass(eq=False)
class PenaltyTerm:  # structures/penalty_terms.jl:1-5 - values of the three penalty expressions per timestep
    energy_balance: np.ndarray
    upper_flow: np.ndarray
    lower_flow: np.ndarray


class _UnitResult:
    """penalty_term, U, K of a unit result (subproblems.jl:89-102) are evaluated on the device on first access
    (dopf_get_unit_penalty) instead of being copied for every agent in every iteration: U, K are L x T per agent."""
    _source = None
    _cache = None

    def _detail(self):
        if self._cache is None:
            if self._source is None:
                raise RuntimeError("no device source attached to this unit result")
            self._cache = self._source()
        return self._cache

    @property
    def penalty_term(self):
        d = self._detail()
        return PenaltyTerm(d["energy_balance"], d["upper_flow"], d["lower_flow"])

    @property
    def U(self):
        return self._detail()["U"]

    @property
    def K(self):
        return self._detail()["K"]


class ResultGenerator(_UnitResult):  # structures/results.jl:11-17
    def __init__(self, generator, generation, source=None):
        self.generator, self.generation, self._source = generator, generation, source


class ResultStorage(_UnitResult):  # structures/results.jl:1-9
    def __init__(self, storage, discharge, charge, level, source=None):
        self.storage, self.discharge, self.charge, self.level, self._source = storage, discharge, charge, level, source


@dataclass(eq=False)
class ResultNode:  # structures/results.jl:19-34 (the node's penalty terms are the sums over its units; not materialised)
    node: Node
    generation: np.ndarray
    discharge: np.ndarray
    charge: np.ndarray


class Result:  # structures/results.jl:36-48
    def __init__(self, unit_to_result, node_to_result, generation, discharge, charge, avg_U, avg_K, total_costs,
                 injection, line_utilization, totals_source=None):
        self.unit_to_result = unit_to_result        # unit object -> ResultGenerator | ResultStorage
        self.node_to_result = node_to_result        # Node -> ResultNode
        self.generation, self.discharge, self.charge = generation, discharge, charge      # [T]
        self.avg_U, self.avg_K = avg_U, avg_K                                              # [L,T]
        self.total_costs = total_costs
        self.injection = injection                  # [N,T]
        self.line_utilization = line_utilization    # [L,T]
        self._totals_source, self._totals = totals_source, None

    @property
    def penalty_term(self):
        """sum of the units' penalty terms (results.jl:73-76), evaluated on the device on first access"""
        if self._totals is None:
            if self._totals_source is None:
                raise RuntimeError("no device source attached to this result")
            self._totals = self._totals_source()
        t = self._totals
        return PenaltyTerm(t["energy_balance"], t["upper_flow"], t["lower_flow"])
