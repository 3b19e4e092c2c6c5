"""Host-side mirror of the reference's ADMM object and driver functions (the drop-in boundary).

The reference driver, /root/reference/src/opf_admm_decentral.jl:5-9, reads

    admm = ADMM(0.3, nodes, generators, storages, lines)     # structures/admm.jl:23-62
    run!(admm)                                                # optimization/run.jl:1-5
    np = get_nodal_price(admm.iteration)                      # helpers/network_elements.jl:16-25

and here (Python cannot use `!`):

    admm = ADMM(0.3, nodes, generators, storages, lines)
    run(admm)
    np_ = get_nodal_price(admm.iteration)

Same names, argument meaning and semantics as /root/reference/src: `calculate_iteration`, `optimize_all_subproblems`,
`update_duals`, `check_convergence` (run.jl:7-16), the readers `admm.iteration`, `admm.lambdas[k]`, `admm.mues[k]`,
`admm.rhos[k]`, `admm.results[k].unit_to_result[unit].{generation|discharge|charge|level}`, `.injection`,
`.line_utilization`, `.avg_U/.avg_K`, `admm.convergence.{lambda_,mue,rho,all,*_res}`, the accessors of
helpers/results.jl and `export_results` (helpers/output.jl).  Like the reference, the helper functions read the
module-global `admm` (SURVEY.md section 1 "global-state quirk"): the constructor binds it.

Index convention: the reference is 1-based - `iteration` arguments are 1-based here as well and address the Python lists
at `iteration - 1` (`admm.lambdas[admm.iteration - 1]` are the duals used by iteration `admm.iteration`).

All numerics - the subproblem solves, aggregation, dual update and convergence check - run on the GPU behind the C ABI
(include/dopf.h); this module packs the structs, keeps the reference's history vectors when `trace=True` and builds
`Result` objects from the device iterate.  There is no CPU path.
"""
import os

import numpy as np

from .device import DeviceADMM
from .problem import Problem
from .structures import Convergence, PenaltyTerm, Result, ResultGenerator, ResultNode, ResultStorage

admm = None      # the reference's global `admm` (bound by ADMM.__init__, like `admm = ADMM(...)` in the driver script)


class ADMM:
    """structures/admm.jl:1-63.  `trace=True` keeps the reference's growing histories (every iterate is copied to the
    host, like `push!(admm.results, result)`); `trace=False` keeps only the newest iterate and the last two dual sets
    and lets `run` stay on the device until the stop rule fires."""

    def __init__(self, gamma, nodes, generators, storages, lines, *, flow_weight=10.0, prox_weight=1.0,
                 slack_mask_tol=1e-2, eps=1e-3, trace=True, device=-1, hinge_capacity=0):
        global admm
        self.iteration = 1                                       # admm.jl:29
        self.gamma = float(gamma)
        self.nodes, self.generators, self.storages, self.lines = nodes, generators, storages, lines
        self.problem = Problem.from_structs(nodes, generators, storages, lines)
        p = self.problem
        self.T = list(range(1, p.T + 1)); self.N = list(range(1, p.N + 1)); self.L = list(range(1, p.L + 1))
        self.lambdas = [np.zeros(p.T)]                           # admm.jl:34-36
        self.mues = [np.zeros((p.L, p.T))]
        self.rhos = [np.zeros((p.L, p.T))]
        self.f_max = p.fmax
        self.results = []
        self.convergence = Convergence()
        self.ptdf = p.ptdf
        self.total_demand = p.demand.sum(axis=0)
        self.node_id_to_demand = {i + 1: n.demand for i, n in enumerate(nodes)}
        self.node_to_id = {n: i + 1 for i, n in enumerate(nodes)}
        self.node_to_units = {}
        for u in list(generators) + list(storages):
            self.node_to_units.setdefault(u.node, []).append(u)
        self.trace = bool(trace)
        self.dev = DeviceADMM(p, gamma=gamma, flow_weight=flow_weight, prox_weight=prox_weight,
                              slack_mask_tol=slack_mask_tol, eps=eps, device=device, hinge_capacity=hinge_capacity)
        self._pending = False     # optimize_all_subproblems ran, check_convergence not yet
        self._status = self.dev.status
        admm = self

    # ---- Result of the newest device iterate (structures/results.jl:50-117) ----------------------------------------
    def _pull_result(self):
        it = self.dev.get_iterate()
        units, nodes = {}, {n: ResultNode(n, np.zeros(len(self.T)), np.zeros(len(self.T)), np.zeros(len(self.T))) for n in self.nodes}
        for i, g in enumerate(self.generators):
            units[g] = ResultGenerator(g, it["P"][i], _unit_source(self, 0, i))
            nodes[g.node].generation = nodes[g.node].generation + it["P"][i]
        for i, s in enumerate(self.storages):
            units[s] = ResultStorage(s, it["D"][i], it["C"][i], it["E"][i], _unit_source(self, 1, i))
            nodes[s.node].discharge = nodes[s.node].discharge + it["D"][i]
            nodes[s.node].charge = nodes[s.node].charge + it["C"][i]
        k = self.dev.status.iterations_done

        def totals():
            if self.dev.status.iterations_done != k:
                raise RuntimeError("result.penalty_term is evaluated on the device from the newest iterate only")
            return self.dev.penalty_totals()
        return Result(units, nodes, it["P"].sum(0), it["D"].sum(0), it["C"].sum(0), it["avgU"], it["avgK"],
                      self.dev.total_costs(), it["injection"], it["flow"], totals)


class _unit_source:
    """lazy per-unit penalty terms and private slacks U, K (subproblems.jl:89-102): computed on the device on first
    access, which must happen while the unit's iterate is still the newest one on the device"""

    def __init__(self, adm, kind, index):
        self.adm, self.kind, self.index, self.k = adm, kind, index, adm.dev.status.iterations_done

    def __call__(self):
        if self.adm.dev.status.iterations_done != self.k:
            raise RuntimeError("per-unit penalty terms / U / K are computed on the device from the newest iterate only; "
                               "read them before the next iteration runs")
        return self.adm.dev.unit_penalty(self.kind, self.index)


def _set_flags(a: ADMM, st):
    c = a.convergence
    c.lambda_, c.mue, c.rho, c.all = bool(st.conv_lambda), bool(st.conv_mue), bool(st.conv_rho), bool(st.converged)


def optimize_all_subproblems(a: ADMM):
    """optimize_all_subproblems!(admm) (subproblems.jl:1-17): all agent subproblems + Result(...) of this iteration.
    On the device the dual update and the stop rule are fused behind the aggregation (one CUDA graph per iteration), so
    the whole iteration runs here; update_duals / check_convergence then publish what the reference computes there."""
    if a._pending:
        raise RuntimeError("optimize_all_subproblems: update_duals / check_convergence of the previous call are outstanding")
    if a.convergence.all:
        return
    a._status = a.dev.step(1)
    a._pending = True
    res = a._pull_result()
    if a.trace:
        a.results.append(res)
    else:
        a.results = [res]


def update_duals(a: ADMM):
    """update_duals!(admm) (update_duals.jl:1-39): appends lambda/mue/rho of the next iteration"""
    if not a._pending:
        raise RuntimeError("update_duals: optimize_all_subproblems has not run for this iteration")
    lam, mu, rho = a.dev.get_duals(0)
    if a.trace:
        a.lambdas.append(lam); a.mues.append(mu); a.rhos.append(rho)
    else:
        a.lambdas = [a.lambdas[-1], lam]; a.mues = [a.mues[-1], mu]; a.rhos = [a.rhos[-1], rho]


def check_convergence(a: ADMM):
    """check_convergence!(admm) (convergence.jl:1-31): residual histories, flags, iteration counter (the comparison
    itself ran on the device right after the dual update)"""
    if not a._pending:
        raise RuntimeError("check_convergence: no iteration outstanding")
    st = a._status
    if a.iteration != 1:
        c = a.convergence
        c.lambda_res.append(np.abs(a.lambdas[-1] - a.lambdas[-2]))
        c.mue_res.append(np.abs(a.mues[-1] - a.mues[-2]))
        c.rho_res.append(np.abs(a.rhos[-1] - a.rhos[-2]))
        _set_flags(a, st)
    if not a.convergence.all:
        a.iteration += 1
    a._pending = False
    assert a.iteration == st.iteration, "host and device iteration counters diverged"


def calculate_iteration(a: ADMM):
    """calculate_iteration!(admm) (run.jl:7-16)"""
    if a.convergence.all:
        return
    optimize_all_subproblems(a)
    update_duals(a)
    check_convergence(a)


def run(a: ADMM, max_iterations=1_000_000):
    """run!(admm) (run.jl:1-5).  With `trace=False` the whole loop stays on the device (the stop rule is evaluated by
    the kernels themselves, one host synchronisation per 64 iterations)."""
    if a.trace:
        n = 0
        while not a.convergence.all and n < max_iterations:
            calculate_iteration(a)
            n += 1
        return
    st = a.dev.step(max_iterations)
    _set_flags(a, st)
    a.iteration = st.iteration
    a.results = [a._pull_result()]
    (l1, m1, r1), (l0, m0, r0) = a.dev.get_duals(1), a.dev.get_duals(0)
    a.lambdas, a.mues, a.rhos = [l1, l0], [m1, m0], [r1, r0]


def get_nodal_price(iteration=None, adm=None):
    """get_nodal_price(iteration) (network_elements.jl:16-25): lambda_t + sum_l (mue+rho)[l,t]*ptdf[l,:] with the duals of
    `iteration` (1-based; the driver passes admm.iteration = the duals used by the last executed iteration).  Reads
    the global `admm` like the reference; evaluated on the device."""
    if isinstance(iteration, ADMM):          # get_nodal_price(admm, k) convenience form
        iteration, adm = adm, iteration
    a = adm or admm
    if iteration is None:
        iteration = a.iteration
    if a.trace:
        k = iteration - 1
        if not 0 <= k < len(a.lambdas):
            raise IndexError(f"get_nodal_price: iteration {iteration} outside the history 1..{len(a.lambdas)}")
        return a.dev.nodal_price_of(a.lambdas[k], a.mues[k], a.rhos[k])
    done = a.dev.status.iterations_done
    if iteration == done + 1:
        return a.dev.nodal_price(0)
    if iteration == done and done >= 1:
        return a.dev.nodal_price(1)
    raise IndexError(f"get_nodal_price: with trace=False only the dual sets of iterations {done} and {done + 1} exist; "
                     f"construct ADMM(..., trace=True) for the full history")


# ---- accessors of helpers/results.jl (1-based `iteration`; zeros before the first result, results.jl:15-21) -----------
def get_results(iteration):
    if len(admm.results) == 0:
        z = np.zeros(len(admm.T)); return z, z, z
    r = admm.results[iteration - 1]
    return r.generation, r.discharge, r.charge


def get_unit_results(unit, iteration):
    from .structures import Storage
    if len(admm.results) == 0:
        z = np.zeros(len(admm.T))
        return (z, z) if isinstance(unit, Storage) else z
    r = admm.results[iteration - 1].unit_to_result[unit]
    return (r.discharge, r.charge) if isinstance(unit, Storage) else r.generation


def get_node_results(iteration, node):
    if len(admm.results) == 0:
        z = np.zeros(len(admm.T)); return z, z, z
    r = admm.results[iteration - 1].node_to_result[node]
    return r.generation, r.discharge, r.charge


def get_average_slack_results(iteration):
    if len(admm.results) == 0:
        z = np.zeros((len(admm.L), len(admm.T))); return z, z
    r = admm.results[iteration - 1]
    return r.avg_U, r.avg_K


# ---- export_results (helpers/output.jl:1-85): the reference's long-format CSVs -----------------------------------------
def _jl(x):
    """Julia's shortest round-trip printing of a Float64 (CSV.jl) for the common cases: integers keep a trailing .0"""
    r = repr(float(x))
    if r.endswith(".0") or "e" in r or "." in r or r in ("inf", "-inf", "nan"):
        return r
    return r + ".0"


def export_results(a: ADMM, filename, parent_dir="results/"):
    """export_results(admm, filename): <parent_dir><filename>_{duals,generators,storages}.csv with the reference's columns
    and row order (output.jl:14-85; E is not exported there either).  Needs the histories, i.e. trace=True."""
    if not a.trace:
        raise RuntimeError("export_results needs the iteration histories: construct ADMM(..., trace=True)")
    os.makedirs(parent_dir, exist_ok=True)
    n_it = a.iteration
    with open(os.path.join(parent_dir, filename + "_duals.csv"), "w") as f:
        f.write("iteration,dual,timestep,line,value\n")
        for name, vals in (("lambda", a.lambdas), ("rho", a.rhos), ("mue", a.mues)):
            for i in range(1, n_it + 1):
                for t in a.T:
                    if name == "lambda":
                        f.write(f"{i},{name},{t},,{_jl(vals[i - 1][t - 1])}\n")
                    else:
                        for l in a.L:
                            f.write(f"{i},{name},{t},{l},{_jl(vals[i - 1][l - 1, t - 1])}\n")
    with open(os.path.join(parent_dir, filename + "_generators.csv"), "w") as f:
        f.write("iteration,generator,timestep,generation\n")
        for g in a.generators:
            for i in range(1, n_it + 1):
                r = a.results[i - 1].unit_to_result[g]
                for t in a.T:
                    f.write(f"{i},{g.name},{t},{_jl(r.generation[t - 1])}\n")
    with open(os.path.join(parent_dir, filename + "_storages.csv"), "w") as f:
        f.write("iteration,storage,timestep,charge,discharge\n")
        for s in a.storages:
            for i in range(1, n_it + 1):
                r = a.results[i - 1].unit_to_result[s]
                for t in a.T:
                    f.write(f"{i},{s.name},{t},{_jl(r.charge[t - 1])},{_jl(r.discharge[t - 1])}\n")
