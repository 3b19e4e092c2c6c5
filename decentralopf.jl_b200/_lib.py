"""ctypes binding of libdopf.so (the C ABI declared in include/dopf.h).

The product path has no CPU fallback: if the shared library is missing, or no CUDA device is
present, creation of an instance raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("DOPF_LIB", os.path.join(_HERE, "libdopf.so"))
_SRCS = [os.path.join(_HERE, "csrc", f) for f in ("dopf_kernels.cu", "dopf_api.cu", "dopf_ptdf.cu")]
_HDRS = [os.path.join(_HERE, "csrc", f) for f in ("dopf_math.h", "dopf_bodies.h", "dopf_kernels.h", "dopf_sto_warp.cuh")] + \
        [os.path.join(ROOT, "include", "dopf.h")]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class DopfProblem(C.Structure):
    _fields_ = [("N", C.c_int32), ("L", C.c_int32), ("T", C.c_int32), ("G", C.c_int32), ("S", C.c_int32),
                ("ptdf", _dp), ("f_max", _dp), ("demand", _dp),
                ("gen_mc", _dp), ("gen_pmax", _dp), ("gen_node", _ip),
                ("sto_mc", _dp), ("sto_pmax", _dp), ("sto_emax", _dp), ("sto_node", _ip)]


class DopfConfig(C.Structure):
    _fields_ = [("gamma", C.c_double), ("flow_weight", C.c_double), ("prox_weight", C.c_double),
                ("slack_mask_tol", C.c_double), ("eps", C.c_double),
                ("device", C.c_int32), ("hinge_capacity", C.c_int32), ("use_graph", C.c_int32), ("debug_flags", C.c_int32),
                ("n_scenarios", C.c_int32), ("gemm_ksplit", C.c_int32)]


class DopfStatus(C.Structure):
    _fields_ = [("iteration", C.c_int32), ("converged", C.c_int32),
                ("conv_lambda", C.c_int32), ("conv_mue", C.c_int32), ("conv_rho", C.c_int32),
                ("iterations_done", C.c_int32),
                ("res_lambda", C.c_double), ("res_mue", C.c_double), ("res_rho", C.c_double),
                ("gen_corrected", C.c_int32), ("sto_corrected", C.c_int32),
                ("tight_rows", C.c_int32), ("wide_rows", C.c_int32),
                ("launches_per_iteration", C.c_int32), ("sto_cold", C.c_int32), ("last_step_ms", C.c_double),
                ("fix_sequential", C.c_int32), ("reserved3", C.c_int32)]


EXPORTS = ["dopf_version", "dopf_default_config", "dopf_create", "dopf_destroy", "dopf_step", "dopf_get_status",
           "dopf_get_iterate", "dopf_get_duals", "dopf_set_state", "dopf_get_nodal_price", "dopf_get_total_costs",
           "dopf_nodal_price_from", "dopf_get_unit_penalty", "dopf_get_penalty_totals",
           "dopf_set_partition", "dopf_set_stream", "dopf_step_phase", "dopf_exchange_buffer", "dopf_last_error", "dopf_profile_iteration", "dopf_debug_counters", "dopf_get_scenario_status",
           "dopf_calculate_ptdf", "dopf_ptdf_last_error", "dopf_comm_get_unique_id", "dopf_comm_init"]


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> decentralopf.jl_b200/libdopf.so (in-tree)."""
    deps = _SRCS + _HDRS
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "--shared", "-cudart", "static", "-o", LIB_PATH] + _SRCS + ["-ldl"]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    if os.environ.get("DOPF_NVCC_FLAGS"):
        cmd[1:1] = os.environ["DOPF_NVCC_FLAGS"].split()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def load():
    """Load libdopf.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback for the ADMM iteration")
    lib = C.CDLL(LIB_PATH)
    lib.dopf_version.restype = C.c_char_p
    lib.dopf_last_error.restype = C.c_char_p
    lib.dopf_last_error.argtypes = [C.c_void_p]
    lib.dopf_default_config.argtypes = [C.POINTER(DopfConfig)]
    lib.dopf_create.argtypes = [C.POINTER(DopfProblem), C.POINTER(DopfConfig), C.POINTER(C.c_void_p)]
    lib.dopf_destroy.argtypes = [C.c_void_p]
    lib.dopf_step.argtypes = [C.c_void_p, C.c_int32, C.POINTER(DopfStatus)]
    lib.dopf_get_status.argtypes = [C.c_void_p, C.POINTER(DopfStatus)]
    lib.dopf_get_scenario_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dopf_get_iterate.argtypes = [C.c_void_p] + [C.c_void_p] * 8
    lib.dopf_get_duals.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 3
    lib.dopf_set_state.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
    lib.dopf_get_nodal_price.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
    lib.dopf_nodal_price_from.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    lib.dopf_get_unit_penalty.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5
    lib.dopf_get_penalty_totals.argtypes = [C.c_void_p] + [C.c_void_p] * 3
    lib.dopf_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32]
    lib.dopf_get_total_costs.argtypes = [C.c_void_p, C.c_void_p]
    lib.dopf_calculate_ptdf.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.dopf_ptdf_last_error.restype = C.c_char_p
    lib.dopf_profile_iteration.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]
    lib.dopf_set_partition.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    lib.dopf_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    lib.dopf_step_phase.argtypes = [C.c_void_p, C.c_int32]
    lib.dopf_exchange_buffer.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    lib.dopf_comm_get_unique_id.argtypes = [C.c_void_p]
    lib.dopf_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
    _lib = lib
    return lib
