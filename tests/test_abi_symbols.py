"""The C-ABI library builds, loads and exports every symbol include/canvasgpu.h declares (no GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "canvasgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cg_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from canvas_b200 import build, native
    lib_path = build.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 12
    for name in names:
        assert hasattr(lib, name), f"{name} declared in canvasgpu.h but not exported"
        assert name in native.SIGNATURES, f"{name} has no ctypes signature in canvas_b200/native.py"
    assert set(native.SIGNATURES) == set(names)


def test_no_cpu_fallback_without_device():
    import pytest
    import torch
    from canvas_b200 import native
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for CPU boxes")
    with pytest.raises(native.CanvasGpuError):
        native.Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "canvas_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "libcanvas_oracle" not in text and "oracle/" not in text, f
