"""CPU: the C-ABI shared library loads without a GPU and exports every symbol that
include/takzero_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from takzero_b200 import build as tz_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "takzero_b200.h")).read()
    return sorted(set(re.findall(r"^TZ_API [^;(]*?\b(tz_\w+)\s*\(", text, flags=re.M)))


def test_library_exports_every_declared_symbol():
    path = tz_build.build()
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_no_undeclared_exports():
    import subprocess

    out = subprocess.check_output(["nm", "-D", "--defined-only", tz_build.build()], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_symbols()


def test_create_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA tz_create returns an error (on a GPU box it succeeds)."""
    import torch

    from takzero_b200 import capi

    if torch.cuda.is_available():
        m = capi.BatchedMCTS(4, 4, 2, arena_slots=4096)
        m.close()
    else:
        try:
            capi.BatchedMCTS(4, 4, 2, arena_slots=4096)
        except capi.TakzeroError as e:
            assert "no CUDA device" in str(e) or "CUDA" in str(e)
        else:
            raise AssertionError("tz_create must fail without a CUDA device")
