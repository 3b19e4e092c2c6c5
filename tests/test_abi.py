"""The C ABI without a GPU: the library builds for sm_100a, loads, exports every symbol that
include/dopf.h declares, mirrors the header's struct layouts, and refuses to run without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "dopf.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dopf_[a-z_]+)\s*\(", hdr)))


def test_library_builds_and_exports_header_symbols(pkg):
    from dopf_b200 import _lib
    _lib.build()
    lib = _lib.load()
    names = _declared_functions()
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.dopf_version()


def test_struct_layouts_match_header(pkg):
    from dopf_b200 import _lib
    assert C.sizeof(_lib.DopfProblem) == 5 * 4 + 4 + 10 * 8            # 5 int32 + pad + 10 pointers
    assert C.sizeof(_lib.DopfConfig) == 5 * 8 + 6 * 4
    assert C.sizeof(_lib.DopfStatus) == 6 * 4 + 3 * 8 + 6 * 4 + 8 + 8
    cfg = _lib.DopfConfig()
    _lib.load().dopf_default_config(C.byref(cfg))
    assert (cfg.gamma, cfg.flow_weight, cfg.prox_weight, cfg.slack_mask_tol, cfg.eps) == (0.3, 10.0, 1.0, 1e-2, 1e-3)


def test_sass_contains_fp64_tensor_and_async_copy(pkg):
    """the PTDF products run on the fp64 tensor pipe (DMMA) fed by cp.async (LDGSTS)"""
    import shutil, subprocess
    from dopf_b200 import _lib
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA" in sass and "LDGSTS" in sass


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from dopf_b200.device import DeviceADMM, DopfError
    prob = pkg.Problem.from_structs(*pkg.cases.three_node())
    with pytest.raises(DopfError, match="no CUDA device"):
        DeviceADMM(prob)


def test_invalid_arguments_are_reported(pkg):
    from dopf_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.dopf_create(None, None, C.byref(h)) == -1
    assert b"null" in lib.dopf_last_error(None)
    assert lib.dopf_step(None, 1, None) == -1


def test_product_package_does_not_import_oracle():
    pkgdir = os.path.join(ROOT, "decentralopf.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src, (dirpath, f)


def test_julia_shim_matches_the_abi(pkg):
    """The Julia side cannot be executed here (no Julia in the image), so at least keep it consistent statically: the three C
    structs of dopf_imports.jl have the fields of the ctypes mirrors (same names, order and C types), and every symbol the
    shim `ccall`s is exported by libdopf.so."""
    import ctypes as C
    import re
    from dopf_b200 import _lib
    src = open(os.path.join(os.path.dirname(_lib.__file__), "julia", "dopf_imports.jl")).read()
    ctype = {"Cint": C.c_int32, "Cdouble": C.c_double, "Ptr{Cdouble}": _lib._dp, "Ptr{Cint}": _lib._ip}
    for name, mirror in (("DopfProblem", _lib.DopfProblem), ("DopfConfig", _lib.DopfConfig), ("DopfStatus", _lib.DopfStatus)):
        body = re.search(r"struct %s\n(.*?)\nend" % name, src, re.S).group(1)
        fields = re.findall(r"(\w+)::([\w{}]+)", body)
        assert [(f, ctype[t]) for f, t in fields] == list(mirror._fields_), name
    called = set(re.findall(r"ccall\(\(:(\w+), libdopf\)", src))
    assert called and called <= set(_lib.EXPORTS), called - set(_lib.EXPORTS)
    assert {"dopf_create", "dopf_step", "dopf_get_iterate", "dopf_get_duals", "dopf_comm_init", "dopf_comm_get_unique_id"} <= called
    # every ccall passes as many arguments as the ctypes binding (which mirrors the header) declares
    lib = _lib.load()
    for m in re.finditer(r"ccall\(\(:(\w+), libdopf\), (\w+), \(([^)]*)\)", src):
        fn, _, args = m.groups()
        n = len([a for a in args.split(",") if a.strip()])
        assert getattr(lib, fn).argtypes is not None and n == len(getattr(lib, fn).argtypes), (fn, n)
