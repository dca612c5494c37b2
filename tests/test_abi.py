"""CPU-side checks of the boundary: the CUDA library builds, loads without a GPU and exports every symbol that
include/yaik_b200.h declares; compute entry points are not called here."""
import os
import re

from yaik_b200 import build as ykbuild
from yaik_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "yaik_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(yk_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_builds_loads_and_exports_everything():
    path = ykbuild.build()
    lib = capi.load_library(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.yk_abi_version() == 1
    assert lib.yk_error_string(-4).decode().startswith("sample")


def test_no_cpu_fallback_in_product_sources():
    """The product must not reach into oracle/ or the emulation shim."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "yaik_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", "") or f in ("synth.py",), (dirpath, f)
                assert "cuda_emu" not in src, (dirpath, f)
