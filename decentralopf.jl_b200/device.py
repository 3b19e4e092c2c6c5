"""DeviceADMM: thin array-level host object over one libdopf handle (one GPU)."""
import ctypes as C
import os

import numpy as np

from . import _lib
from .problem import Problem


class DopfError(RuntimeError):
    pass


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceADMM:
    """Owns a `dopf_handle`.  Arrays are numpy float64, row-major, timestep contiguous."""

    def __init__(self, prob: Problem, gamma=0.3, flow_weight=10.0, prox_weight=1.0, slack_mask_tol=1e-2, eps=1e-3,
                 device=-1, hinge_capacity=0, use_graph=True, debug_flags=0, gemm_ksplit=0):
        self.lib = _lib.load()
        self.prob = prob
        p = prob
        self._keep = p  # inputs are copied by the library; keep them only for the duration of the call
        cp = _lib.DopfProblem(p.N, p.L, p.T, p.G, p.S,
                              p.ptdf.ctypes.data_as(_lib._dp), p.fmax.ctypes.data_as(_lib._dp), p.demand.ctypes.data_as(_lib._dp),
                              p.gen_mc.ctypes.data_as(_lib._dp), p.gen_pmax.ctypes.data_as(_lib._dp), p.gen_node.ctypes.data_as(_lib._ip),
                              p.sto_mc.ctypes.data_as(_lib._dp), p.sto_pmax.ctypes.data_as(_lib._dp), p.sto_emax.ctypes.data_as(_lib._dp),
                              p.sto_node.ctypes.data_as(_lib._ip))
        cfg = _lib.DopfConfig()
        self.lib.dopf_default_config(C.byref(cfg))
        cfg.gamma, cfg.flow_weight, cfg.prox_weight = float(gamma), float(flow_weight), float(prox_weight)
        cfg.slack_mask_tol, cfg.eps = float(slack_mask_tol), float(eps)
        cfg.device, cfg.hinge_capacity, cfg.use_graph = int(device), int(hinge_capacity), int(bool(use_graph))
        cfg.debug_flags = int(debug_flags) | int(os.environ.get("DOPF_DEBUG_FLAGS", "0"))
        cfg.n_scenarios, cfg.gemm_ksplit = int(getattr(prob, "n_scen", 1)), int(gemm_ksplit)
        self.C = cfg.n_scenarios
        self.h = C.c_void_p()
        rc = self.lib.dopf_create(C.byref(cp), C.byref(cfg), C.byref(self.h))
        if rc != 0:
            self.h = None
            raise DopfError(f"dopf_create rc={rc}: {self.lib.dopf_last_error(None).decode()}")
        self.status = _lib.DopfStatus()
        self.lib.dopf_get_status(self.h, C.byref(self.status))

    def _check(self, rc, what):
        if rc != 0:
            raise DopfError(f"{what} rc={rc}: {self.lib.dopf_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.dopf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stepping -------------------------------------------------------------------------
    def step(self, iters=1):
        self._check(self.lib.dopf_step(self.h, int(iters), C.byref(self.status)), "dopf_step")
        return self.status

    @property
    def iteration(self):
        return self.status.iteration

    @property
    def converged(self):
        return bool(self.status.converged)

    @property
    def residuals(self):
        return (self.status.res_lambda, self.status.res_mue, self.status.res_rho)

    # ---- state transfer --------------------------------------------------------------------
    def get_iterate(self, want=("P", "D", "C", "E", "injection", "flow", "avgU", "avgK"), out=None):
        p = self.prob
        lead = (self.C,) if self.C > 1 else ()
        shapes = dict(P=(p.G, p.T), D=(p.S, p.T), C=(p.S, p.T), E=(p.S, p.T), injection=(p.N, p.T),
                      flow=(p.L, p.T), avgU=(p.L, p.T), avgK=(p.L, p.T))
        shapes = {k: lead + v for k, v in shapes.items()}
        res = out if out is not None else {}
        for k in want:
            if k not in res:
                res[k] = np.empty(shapes[k])
        args = [_ptr(res.get(k)) if k in want else None for k in ("P", "D", "C", "E", "injection", "flow", "avgU", "avgK")]
        self._check(self.lib.dopf_get_iterate(self.h, *args), "dopf_get_iterate")
        return res

    def get_duals(self, which=0):
        p = self.prob
        lead = (self.C,) if self.C > 1 else ()
        lam = np.empty(lead + (p.T,)); mu = np.empty(lead + (p.L, p.T)); rho = np.empty(lead + (p.L, p.T))
        self._check(self.lib.dopf_get_duals(self.h, int(which), _ptr(lam), _ptr(mu), _ptr(rho)), "dopf_get_duals")
        return lam, mu, rho

    def set_state(self, iteration, P=None, D=None, C_=None, avgU=None, avgK=None, lam=None, mu=None, rho=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (P, D, C_, avgU, avgK, lam, mu, rho)]
        self._check(self.lib.dopf_set_state(self.h, int(iteration), *[_ptr(a) for a in arrs]), "dopf_set_state")
        self.lib.dopf_get_status(self.h, C.byref(self.status))

    def profile_iteration(self, cap=64):
        """One iteration with a CUDA-event pair around every kernel -> [(kernel name, ms)]."""
        import re
        ms = (C.c_float * cap)(); names = (C.c_char_p * cap)(); n = C.c_int32()
        self._check(self.lib.dopf_profile_iteration(self.h, cap, ms, names, C.byref(n)), "dopf_profile_iteration")
        self.lib.dopf_get_status(self.h, C.byref(self.status))
        out = []
        for i in range(n.value):
            m = re.match(r"\s*(k_\w+)(<[^<>]*>)?\s*<<<", names[i].decode())
            out.append(((m.group(1) + (m.group(2) or "")) if m else names[i].decode()[:40], float(ms[i])))
        return out

    def scenario_status(self):
        """per scenario: iteration [C], converged [C], residuals of the last check [C,3]"""
        it = np.zeros(self.C, dtype=np.int32); cv = np.zeros(self.C, dtype=np.int32); res = np.zeros((self.C, 3))
        self._check(self.lib.dopf_get_scenario_status(self.h, _ptr(it), _ptr(cv), _ptr(res)), "dopf_get_scenario_status")
        return it, cv, res

    def nodal_price(self, which=1):
        out = np.empty(((self.C,) if self.C > 1 else ()) + (self.prob.N, self.prob.T))
        self._check(self.lib.dopf_get_nodal_price(self.h, int(which), _ptr(out)), "dopf_get_nodal_price")
        return out

    def nodal_price_of(self, lam, mu, rho):
        """get_nodal_price for an arbitrary dual set of the caller's history"""
        out = np.empty(((self.C,) if self.C > 1 else ()) + (self.prob.N, self.prob.T))
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (lam, mu, rho)]
        self._check(self.lib.dopf_nodal_price_from(self.h, _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(out)), "dopf_nodal_price_from")
        return out

    def unit_penalty(self, kind, index, with_slacks=True):
        """penalty_term and private slacks U, K of one unit in the newest iterate (kind 0 generator, 1 storage)"""
        p = self.prob
        eb, up, lo = np.empty(p.T), np.empty(p.T), np.empty(p.T)
        U = np.empty((p.L, p.T)) if with_slacks else None
        K = np.empty((p.L, p.T)) if with_slacks else None
        self._check(self.lib.dopf_get_unit_penalty(self.h, int(kind), int(index), _ptr(eb), _ptr(up), _ptr(lo), _ptr(U), _ptr(K)), "dopf_get_unit_penalty")
        return dict(energy_balance=eb, upper_flow=up, lower_flow=lo, U=U, K=K)

    def penalty_totals(self):
        p = self.prob
        shp = ((self.C,) if self.C > 1 else ()) + (p.T,)
        eb, up, lo = np.empty(shp), np.empty(shp), np.empty(shp)
        self._check(self.lib.dopf_get_penalty_totals(self.h, _ptr(eb), _ptr(up), _ptr(lo)), "dopf_get_penalty_totals")
        return dict(energy_balance=eb, upper_flow=up, lower_flow=lo)

    def debug_counters(self, reset=True):
        out = (C.c_uint64 * 32)()
        self._check(self.lib.dopf_debug_counters(self.h, out, int(reset)), "dopf_debug_counters")
        return [int(x) for x in out]

    def total_costs(self):
        """result.total_costs (one value per scenario for a batch)"""
        v = (C.c_double * self.C)()
        self._check(self.lib.dopf_get_total_costs(self.h, v), "dopf_get_total_costs")
        return v[0] if self.C == 1 else np.array(list(v))
