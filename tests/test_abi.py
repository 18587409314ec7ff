"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, and its structs have the layout the bindings assume.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dvo_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build_cuda()
    from dvo_slam_b200 import engine
    return engine.load_library()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvo_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from dvo_slam_b200 import engine
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dvo_b200.h but not exported"
    assert sorted(engine.ABI_SYMBOLS) == syms
    assert lib.dvo_b200_abi_version() == 1


def test_struct_layout_matches_header(lib, tmp_path):
    from dvo_slam_b200 import engine
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dvo_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(dvo_b200_config),sizeof(dvo_b200_iteration_stats),sizeof(dvo_b200_level_stats),sizeof(dvo_b200_result),'
                    'offsetof(dvo_b200_result,levels),offsetof(dvo_b200_config,precision),offsetof(dvo_b200_level_stats,last_increment_log_likelihood));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(engine.Config), C.sizeof(engine.IterationStats), C.sizeof(engine.LevelStats), C.sizeof(engine.CResult),
            engine.CResult.levels.offset, engine.Config.precision.offset, engine.LevelStats.last_increment_log_likelihood.offset]
    assert got == want


def test_default_config_matches_reference_defaults(lib):
    from dvo_slam_b200 import engine
    c = engine.Config()
    d = engine.Config(first_level=0, last_level=0, max_iterations_per_level=0, precision=0, mu=1)
    lib.dvo_b200_config_default(C.byref(d))
    # dense_tracking_config.cpp:27-42
    for f in ("first_level", "last_level", "max_iterations_per_level", "use_initial_estimate", "precision", "mu",
              "intensity_derivative_threshold", "depth_derivative_threshold"):
        assert getattr(c, f) == getattr(d, f)
    assert (d.first_level, d.last_level, d.max_iterations_per_level, d.precision, d.mu) == (3, 1, 100, 5e-7, 0.0)


def test_create_fails_loudly_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ctx = C.c_void_p()
    rc = lib.dvo_b200_create(0, None, C.byref(ctx))
    assert rc == -2 and not ctx.value      # DVO_B200_ERR_CUDA, no CPU fallback
    from dvo_slam_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under dvo_slam_b200/ or include/ may reference it."""
    bad = []
    for base in ("dvo_slam_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".cc")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle_py|liboracle|dvo_oracle\.h|orc_[a-z]+\(", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
