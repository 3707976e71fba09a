"""CPU: the C-ABI library builds, loads and exports every symbol include/dgdm_b200.h declares (no compute calls)."""
import os
import re

from dgdm_histopath_lab_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "dgdm_b200.h")).read()
    names = sorted(set(re.findall(r"\b(dgdm_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dgdm_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert lib.dgdm_abi_version() == 1


def test_size_queries_are_pure_host_functions():
    lib = _lib.load()
    assert lib.dgdm_csr_build_work_ints(100, 1000) >= 2 * 101 + 1000
    assert lib.dgdm_topk_pool_work_ints(1000) >= 1000
    assert lib.dgdm_spatial_attention_work_bytes(2000, 2000, 128, 8, 1, 0) > 2000 * 128 * 2 * 6
    assert lib.dgdm_attention_pool_work_floats(1000, 128, 8, 2) >= 8000


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "dgdm_histopath_lab_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
