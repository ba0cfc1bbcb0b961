"""CPU: the C-ABI library loads and exports every function include/sy_env.h declares
(no compute calls without a GPU), and the product refuses to run without CUDA."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_functions(header="sy_env.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sy_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from student_mechanism_design_b200 import _cabi

    assert _declared_functions() == sorted(_cabi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from student_mechanism_design_b200 import _cabi

    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), name
    lib.sy_abi_version.restype = ctypes.c_int
    assert lib.sy_abi_version() == _cabi.SY_ABI_VERSION
    assert ctypes.sizeof(_cabi.SyConfig) == 160


def test_policy_header_and_library_agree():
    import __graft_entry__ as ge

    ge.build()
    from student_mechanism_design_b200 import _policy_cabi

    declared = _declared_functions("sy_policy.h")
    assert declared == sorted(_policy_cabi.SIGNATURES)
    lib = ctypes.CDLL(_policy_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.sy_policy_abi_version.restype = ctypes.c_int
    assert lib.sy_policy_abi_version() == _policy_cabi.SY_POLICY_ABI_VERSION
    lib.sy_gnn_param_count.restype = ctypes.c_int32
    assert lib.sy_gnn_param_count(7) == 2 * (2 * 64 + 8) + 8 + 4 and lib.sy_gnn_param_count(17) == 0


def test_no_cpu_fallback():
    import torch

    import student_mechanism_design_b200 as pkg

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.SyError):
        pkg.BatchedScotlandYardEnv(4, 2, 10)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: the product must not import, include or dlopen it."""
    pkg_dir = os.path.join(ROOT, "student_mechanism_design_b200")
    pat_py = re.compile(r"^\s*(from|import)\s+(sy_oracle|ref_loader|oracle)\b|libsy_oracle|oracle/_", re.M)
    pat_c = re.compile(r"#include[^\n]*oracle|libsy_oracle")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            src_path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                assert not pat_py.search(open(src_path).read()), f
            elif f.endswith((".cu", ".h", ".cuh", ".cpp")):
                assert not pat_c.search(open(src_path).read()), f
