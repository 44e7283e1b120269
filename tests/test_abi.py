"""The C-ABI library loads here (no GPU) and exports every symbol include/b200vo.h declares."""
import ctypes
import os
import re

from monocular_visual_odometry_va4mr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200vo.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    assert os.path.exists(_lib.SO_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in include/b200vo.h but not exported"


def test_python_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()


def test_version_probe():
    lib = _lib.load()
    arch = ctypes.c_int(0)
    assert lib.b200vo_version(ctypes.byref(arch)) >= 100
    assert arch.value == 100


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's cpu_baseline leg may touch oracle/."""
    pkg = os.path.join(ROOT, "monocular_visual_odometry_va4mr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "vo_oracle" not in src, f
