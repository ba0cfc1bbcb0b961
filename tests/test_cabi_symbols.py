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
    assert ctypes.sizeof(_cabi.SyConfig) == 168  # ABI 4: + reveal_skip_prob, reserved0


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


def _header_struct_fields(name):
    """member names of `typedef struct <name> { ... } <name>;` in include/sy_env.h, in order"""
    import re

    text = open(os.path.join(ROOT, "include", "sy_env.h")).read()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    return [re.search(r"(\w+)(\[[^\]]*\])?\s*$", stmt.strip()).group(1) for stmt in body.split(";") if stmt.strip()]


def test_struct_layouts_of_binding_and_integration_doc_match_the_header():
    """the ctypes structs of _cabi.py AND the `_fields_` lists a maintainer would copy from INTEGRATION.md section 3 name
    exactly the header's members, in order (round-1 finding: the document was two ABI revisions behind, so a binding
    written from it made sy_step read pointers past the caller's structs)"""
    import re

    from student_mechanism_design_b200 import _cabi

    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "SY_ABI_VERSION" in doc and f"sy_abi_version() == {_cabi.SY_ABI_VERSION}" in doc
    for name in ("SyConfig", "SyState", "SyObs", "SyOut", "SyHostOut"):
        want = _header_struct_fields(name)
        assert [f[0] for f in getattr(_cabi, name)._fields_] == want, name
        stmt = re.search(r"class %s\(C\.Structure\):(.*?)(?=\nclass |\n# |\nhandle|\nP = )" % name, doc, re.S).group(1)
        assert re.findall(r'"(\w+)"', stmt) == want, (name, "INTEGRATION.md")
    # sizes: 8 bytes of struct_bytes + one pointer per member
    for name in ("SyState", "SyObs", "SyOut", "SyHostOut"):
        import ctypes as C

        assert C.sizeof(getattr(_cabi, name)) == 8 * len(getattr(_cabi, name)._fields_)
