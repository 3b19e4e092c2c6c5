"""decentralopf.jl_b200 - B200-native ADMM iteration for DecentralOPF.jl's hot path.

The directory name carries a dot, so it is loaded by path: `from __graft_entry__ import
load_package; dopf = load_package()` (registers the package as `dopf_b200`).
"""
from .structures import (Convergence, Generator, Line, Node, PenaltyTerm, Result, ResultGenerator,  # noqa: F401
                         ResultNode, ResultStorage, Storage)
from .problem import Problem  # noqa: F401
from .ptdf import calculate_ptdf, ptdf_from_arrays  # noqa: F401
from . import cases  # noqa: F401
from .admm import (ADMM, calculate_iteration, check_convergence, export_results, get_average_slack_results,  # noqa: F401,E402
                   get_nodal_price, get_node_results, get_results, get_unit_results, optimize_all_subproblems, run, update_duals)
