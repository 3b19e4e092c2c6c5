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
class ResultGenerator:  # structures/results.jl:11-17 (U,K per agent are never materialised)
    generator: Generator
    generation: np.ndarray


@dataclass(eq=False)
class ResultStorage:  # structures/results.jl:1-9
    storage: Storage
    discharge: np.ndarray
    charge: np.ndarray
    level: np.ndarray


@dataclass(eq=False)
class Result:  # structures/results.jl:36-48 (fields that the iteration or its readers use)
    unit_to_result: dict
    generation: np.ndarray      # [T]
    discharge: np.ndarray       # [T]
    charge: np.ndarray          # [T]
    avg_U: np.ndarray           # [L,T]
    avg_K: np.ndarray           # [L,T]
    total_costs: float
    injection: np.ndarray       # [N,T]
    line_utilization: np.ndarray  # [L,T]
