"""CPU-side checks of the drop-in boundary: the library loads, exports exactly what include/nwb.h
declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

import nao_wholebody_planning_b200 as nwb
from nao_wholebody_planning_b200 import _abi

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "nwb.h").read_text()
    return sorted(set(re.findall(r"NWB_API\s+[\w\s\*]+?\b(nwb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nwb.load_library()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nwb.h but not exported by libnwb.so"
    out = subprocess.run(["nm", "-D", "--defined-only", str(_abi.LIB_PATH)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (nwb_\w+)", out)))
    assert exported == names, "exported symbols and header declarations differ"
    # every declared symbol also has a ctypes signature in the Python mirror
    assert set(lib._nwb_sigs) == set(names)


def test_params_struct_matches_header_defaults():
    lib = nwb.load_library()
    p = nwb.NwbParams()
    assert lib.nwb_params_default(C.byref(p)) == 0
    d = nwb.default_params()
    for name, _ in nwb.NwbParams._fields_:
        a, b = getattr(p, name), getattr(d, name)
        if hasattr(a, "__len__"):
            a, b = list(a), list(b)
        assert a == b, name
    assert p.scale_sp == 0.8 and p.ik_maxiter == 100 and p.planner_max_iterations == 8000
    assert lib.nwb_abi_version() == 1


def test_frame_names():
    lib = nwb.load_library()
    assert lib.nwb_frame_name(0) == b"r_sole" and lib.nwb_frame_name(28) == b"l_sole"
    assert lib.nwb_frame_name(7) == b"torso" and lib.nwb_frame_name(29) is None


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nwb.NwbError, match="no usable CUDA device|CUDA"):
        nwb.Context()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(nwb.NwbError, match="not found"):
        _abi.load_library(tmp_path / "libnwb.so")
