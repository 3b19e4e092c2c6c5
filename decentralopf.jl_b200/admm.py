"""Host-side mirror of the reference's ADMM object and driver functions.

    admm = ADMM(0.3, nodes, generators, storages, lines)     # structures/admm.jl:23-62
    run(admm)                                                 # run!            optimization/run.jl:1-5
    calculate_iteration(admm)                                 # calculate_iteration!  run.jl:7-16
    np_ = get_nodal_price(admm, admm.iteration)               # helpers/network_elements.jl:16-25

Same names, argument meaning and semantics as /root/reference/src (Python cannot use `!`).  The
subproblem solves, aggregation, dual update and convergence check all run on the GPU behind the C ABI
(include/dopf.h); this module only packs the structs, keeps the reference's history vectors
(`admm.lambdas`, `admm.mues`, `admm.rhos`, `admm.results`, `admm.convergence.*_res`) when
`trace=True`, and rebuilds `Result` objects lazily.
"""
import numpy as np

from .device import DeviceADMM
from .problem import Problem
from .structures import Convergence, Result, ResultGenerator, ResultStorage


class ADMM:
    def __init__(self, gamma, nodes, generators, storages, lines, *, flow_weight=10.0, prox_weight=1.0,
                 slack_mask_tol=1e-2, eps=1e-3, trace=True, device=-1, hinge_capacity=0):
        self.gamma = float(gamma)
        self.nodes, self.generators, self.storages, self.lines = nodes, generators, storages, lines
        self.problem = Problem.from_structs(nodes, generators, storages, lines)
        p = self.problem
        self.T = list(range(1, p.T + 1)); self.N = list(range(1, p.N + 1)); self.L = list(range(1, p.L + 1))
        self.ptdf = p.ptdf
        self.f_max = p.fmax
        self.total_demand = p.demand.sum(axis=0)
        self.node_to_id = {id(n): i + 1 for i, n in enumerate(nodes)}
        self.trace = trace
        self.convergence = Convergence()
        self.lambdas = [np.zeros(p.T)]                 # admm.jl:34-36
        self.mues = [np.zeros((p.L, p.T))]
        self.rhos = [np.zeros((p.L, p.T))]
        self.results = []
        self.dev = DeviceADMM(p, gamma=gamma, flow_weight=flow_weight, prox_weight=prox_weight,
                              slack_mask_tol=slack_mask_tol, eps=eps, device=device, hinge_capacity=hinge_capacity)

    @property
    def iteration(self):
        return self.dev.iteration

    def _pull_result(self):
        it = self.dev.get_iterate()
        units = {}
        for i, g in enumerate(self.generators):
            units[id(g)] = ResultGenerator(g, it["P"][i])
        for i, s in enumerate(self.storages):
            units[id(s)] = ResultStorage(s, it["D"][i], it["C"][i], it["E"][i])
        return Result(units, it["P"].sum(0), it["D"].sum(0), it["C"].sum(0), it["avgU"], it["avgK"],
                      self.dev.total_costs(), it["injection"], it["flow"])

    def result_of(self, unit, k=-1):
        """admm.results[k].unit_to_result[unit]"""
        return self.results[k].unit_to_result[id(unit)]


def calculate_iteration(admm: ADMM):
    """calculate_iteration!(admm): optimize_all_subproblems! + update_duals! + check_convergence!"""
    if admm.convergence.all:
        return
    st = admm.dev.step(1)
    if admm.trace:
        admm.results.append(admm._pull_result())
        lam, mu, rho = admm.dev.get_duals(0)
        admm.lambdas.append(lam); admm.mues.append(mu); admm.rhos.append(rho)
        if st.iterations_done > 1:
            admm.convergence.lambda_res.append(np.abs(admm.lambdas[-1] - admm.lambdas[-2]))
            admm.convergence.mue_res.append(np.abs(admm.mues[-1] - admm.mues[-2]))
            admm.convergence.rho_res.append(np.abs(admm.rhos[-1] - admm.rhos[-2]))
    c = admm.convergence
    c.lambda_, c.mue, c.rho, c.all = bool(st.conv_lambda), bool(st.conv_mue), bool(st.conv_rho), bool(st.converged)


def run(admm: ADMM, max_iterations=1_000_000):
    """run!(admm): iterate until admm.convergence.all.  Without tracing the whole loop stays on the
    device (convergence flag checked by the kernels themselves)."""
    if admm.trace:
        n = 0
        while not admm.convergence.all and n < max_iterations:
            calculate_iteration(admm)
            n += 1
    else:
        st = admm.dev.step(max_iterations)
        c = admm.convergence
        c.lambda_, c.mue, c.rho, c.all = bool(st.conv_lambda), bool(st.conv_mue), bool(st.conv_rho), bool(st.converged)
        admm.results = [admm._pull_result()]
        admm.lambdas = [admm.dev.get_duals(1)[0], admm.dev.get_duals(0)[0]]
        admm.mues = [admm.dev.get_duals(1)[1], admm.dev.get_duals(0)[1]]
        admm.rhos = [admm.dev.get_duals(1)[2], admm.dev.get_duals(0)[2]]


def get_nodal_price(admm: ADMM, iteration=None):
    """get_nodal_price(iteration): lambda_t + sum_l (mue+rho)[l,t]*ptdf[l,:] for the duals of `iteration`
    (the driver calls it with admm.iteration, i.e. the duals used by the last executed iteration)."""
    if iteration is None or iteration == admm.iteration:
        return admm.dev.nodal_price(1 if admm.convergence.all else 0)
    lam, mu, rho = admm.lambdas[iteration - 1], admm.mues[iteration - 1], admm.rhos[iteration - 1]
    return lam[None, :] + admm.ptdf.T @ (mu + rho)
