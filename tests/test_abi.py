"""CPU-only: the C-ABI library loads and exports every symbol include/vitk.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vitk.h")).read()
    return sorted(set(re.findall(r"VITK_API\s+[\w\s\*]+?\b(vitk_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import chest_x_ray_vit_b200 as pkg
    if not os.path.exists(pkg._lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _declared()
    assert len(names) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", pkg._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vitk_\w+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(pkg._lib.exported_symbols())
    L = pkg._lib.lib()          # dlopen + prototype binding; no CUDA call
    assert L.vitk_version() == 100


def test_missing_library_fails_loudly(monkeypatch):
    import chest_x_ray_vit_b200 as pkg
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", "/nonexistent/libvitk.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        pkg._lib.lib()


def test_sass_uses_blackwell_tensor_and_tma_paths():
    import chest_x_ray_vit_b200 as pkg
    if not os.path.exists("/usr/local/cuda/bin/cuobjdump"):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")   # no legacy mma.sync path
