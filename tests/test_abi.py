"""CPU-only checks of the C-ABI boundary: the header, the ctypes mirror and the built
library agree, every declared symbol is exported, and the pure-host entry points work
(no compute call is made here -- there is no GPU in this container)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from gmpe_b200 import abi, default_args, make_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gmpe.h")


def header_text():
    return open(HEADER).read()


def test_header_declares_exactly_the_mirrored_symbols():
    txt = header_text()
    decl = set(re.findall(r"^(?:int64_t|int|const char\*)\s+(gmpe_\w+)\s*\(", txt, flags=re.M))
    assert decl == set(abi.EXPORTED_SYMBOLS), decl ^ set(abi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    for sym in abi.EXPORTED_SYMBOLS:
        assert hasattr(lib, sym), sym
    assert lib.gmpe_abi_version() == abi.ABI_VERSION
    m = re.search(r"#define GMPE_ABI_VERSION (\d+)", header_text())
    assert int(m.group(1)) == abi.ABI_VERSION


def test_struct_sizes_match_the_c_compiler(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "gmpe.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", '
                   'sizeof(GmpeConfig), sizeof(GmpeLayout), sizeof(GmpeState), sizeof(GmpeStepOut), sizeof(GmpeStepInject));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.GmpeConfig), C.sizeof(abi.GmpeLayout), C.sizeof(abi.GmpeState), C.sizeof(abi.GmpeStepOut),
            C.sizeof(abi.GmpeStepInject)]
    assert got == want


def test_header_constants_match_python_mirror():
    txt = header_text()
    for name, val in re.findall(r"#define (GMPE_[A-Z0-9_]+) \(?(-?\d+)\)?\s", txt):
        py = name[len("GMPE_"):]
        if hasattr(abi, py):
            assert getattr(abi, py) == int(val), name
    for name, shift in re.findall(r"#define (GMPE_FLAG_[A-Z_]+) \(1u << (\d+)\)", txt):
        assert getattr(abi, name[len("GMPE_"):]) == 1 << int(shift), name


@pytest.mark.parametrize("kw", [
    dict(scenario_name="nav_metered_one_goal_graph_rotate_tube_july", dynamics_type="air_taxi", world_size=3),
    dict(scenario_name="navigation_graph", num_agents=16, num_landmarks=16, num_obstacles=8, num_walls=4, world_size=4),
])
def test_gmpe_layout_matches_python_arithmetic(kw):
    lib = abi.load_library()
    cfg = make_config(default_args(**kw), 77, precision="fp64")
    lay = abi.GmpeLayout()
    assert lib.gmpe_layout(C.byref(cfg), C.byref(lay)) == abi.OK
    py = abi.layout_for(cfg)
    for f, _ in abi.GmpeLayout._fields_:
        assert getattr(lay, f) == getattr(py, f), f


def test_invalid_config_is_rejected_with_error_code_not_exception():
    lib = abi.load_library()
    cfg = make_config(default_args(scenario_name="navigation_graph"), 4)
    cfg.num_agents = 0
    lay = abi.GmpeLayout()
    assert lib.gmpe_layout(C.byref(cfg), C.byref(lay)) == abi.E_INVALID
    h = C.c_void_p()
    rc = lib.gmpe_create(C.byref(cfg), 0, C.byref(h))
    assert rc == abi.E_INVALID
    assert b"num_agents" in lib.gmpe_last_error(None)
    assert lib.gmpe_destroy(None) == abi.E_NULL
    assert lib.gmpe_step(None, None, None, None, None, None) == abi.E_NULL
    # rot-inv family beyond N = 16: refused by the library itself with the reason, not only by make_config
    cfg = make_config(default_args(scenario_name="two_phase_graph", dynamics_type="air_taxi", num_agents=16, num_landmarks=16), 4)
    assert lib.gmpe_layout(C.byref(cfg), C.byref(lay)) == abi.OK
    cfg.num_agents = cfg.num_landmarks = 17
    assert lib.gmpe_layout(C.byref(cfg), C.byref(lay)) == abi.E_INVALID
    assert lib.gmpe_create(C.byref(cfg), 0, C.byref(h)) == abi.E_INVALID
    assert b"<= 16" in lib.gmpe_last_error(None)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(abi.GmpeError, match="no CPU fallback"):
        abi.load_library(str(tmp_path / "nope.so"))


def test_engine_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from gmpe_b200.engine import BatchedGraphMPE

    with pytest.raises(abi.GmpeError, match="no CPU fallback"):
        BatchedGraphMPE(make_config(default_args(), 4))
