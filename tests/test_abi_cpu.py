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


def test_numpy_views_of_the_service_structs_match_the_c_layout():
    """context._struct_dtype: the batch wrappers fill nwb_motion_plan_request / read nwb_motion_plan_response arrays
    through numpy views; the structured dtype must reproduce the ctypes (= C) layout field by field."""
    import ctypes as C

    import numpy as np

    from nao_wholebody_planning_b200 import _abi
    from nao_wholebody_planning_b200.context import _struct_dtype

    for ct in (_abi.Point3, _abi.Trajectory, _abi.MotionPlanRequest, _abi.MotionPlanResponse, _abi.ManipulationPlanResponse):
        dt = _struct_dtype(ct)
        assert dt.itemsize == C.sizeof(ct)
        for name, _ in ct._fields_:
            assert dt.fields[name][1] == getattr(ct, name).offset, (ct.__name__, name)
    n = 5
    reqs = (_abi.MotionPlanRequest * n)()
    rq = np.frombuffer(reqs, dtype=_struct_dtype(_abi.MotionPlanRequest))
    rq["right_hand_direction"]["y"] = np.arange(n) + 0.5
    rq["wb_start_config"] = np.arange(n, dtype=np.uint64) * 168 + 4096
    rq["wb_start_config_len"] = 21
    rq["execute_trajectory"] = 1
    assert reqs[3].right_hand_direction.y == 3.5 and reqs[4].wb_start_config_len == 21 and reqs[0].execute_trajectory == 1
    assert C.cast(reqs[2].wb_start_config, C.c_void_p).value == 4096 + 2 * 168
    resps = (_abi.MotionPlanResponse * n)()
    rs = np.frombuffer(resps, dtype=_struct_dtype(_abi.MotionPlanResponse))
    resps[1].success = 1
    resps[1].num_nodes_generated = 77
    resps[1].wb_goal_config[20] = -0.25
    resps[1].motion_plan.n_points = 9
    assert rs["success"][1] == 1 and rs["num_nodes_generated"][1] == 77 and rs["wb_goal_config"][1, 20] == -0.25
    assert rs["motion_plan"]["n_points"][1] == 9 and rs["success"][0] == 0
