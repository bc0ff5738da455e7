"""CPU: the C-ABI library loads without a GPU, exports every symbol include/b200_planner.h declares, struct
layouts match the Python mirrors, and compute calls fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_planner.h")


def declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_python_abi_tables_agree(b200):
    names = declared_functions()
    assert len(names) >= 30
    assert sorted(n for n, _, _ in b200.ABI) == names


def test_library_exports_every_declared_symbol(b200):
    assert os.path.exists(b200.LIB_PATH), "libb200planner.so missing: run __graft_entry__.build()"
    out = subprocess.check_output(["nm", "-D", "--defined-only", b200.LIB_PATH], text=True)
    exported = set(re.findall(r" T (b200_[a-z0-9_]+)", out))
    missing = [n for n in declared_functions() if n not in exported]
    assert not missing, missing
    lib = b200.lib()
    for n in declared_functions():
        assert hasattr(lib, n)
    assert b"sm_100a" in lib.b200_version()


def test_struct_layouts(b200, tmp_path):
    # compile a C (not C++) program against the header: proves it is a plain C header and gives the true layout
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "b200_planner.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(b200_map_params_t), sizeof(b200_planner_params_t),'
                   'sizeof(b200_path_result_t), offsetof(b200_path_result_t, g_final), offsetof(b200_path_result_t, goal_coords),'
                   'offsetof(b200_map_params_t, device), sizeof(b200_depth_params_t), offsetof(b200_depth_params_t, use_depth_filter));'
                   'return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes[0] == C.sizeof(b200.MapParams) == 104
    assert sizes[1] == C.sizeof(b200.PlannerParams) == 80
    assert sizes[2] == b200.RESULT_DTYPE.itemsize == 128
    assert sizes[3] == b200.RESULT_DTYPE.fields["g_final"][1]
    assert sizes[4] == b200.RESULT_DTYPE.fields["goal_coords"][1]
    assert sizes[5] == b200.MapParams.device.offset
    assert sizes[6] == C.sizeof(b200.DepthParams) == 128
    assert sizes[7] == b200.DepthParams.use_depth_filter.offset == 112
    from oracle import pyoracle as po  # the oracle mirrors the same struct (test infrastructure talking to test infrastructure)
    assert C.sizeof(po.DepthParams) == 128 and [f[0] for f in po.DepthParams._fields_] == [f[0] for f in b200.DepthParams._fields_]
    hdr = open(HEADER).read()
    for field in ("solved", "status", "n_path", "path_offset", "g_final", "path_length", "explored_nodes", "iter_num",
                  "los_checks", "los_samples", "goal_coords", "start_coords", "time_us"):
        assert field in hdr and field in b200.RESULT_DTYPE.names


def test_binary_is_sm100a_only():
    lib = os.path.join(ROOT, "3d-astar-path-planners_b200", "libb200planner.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(b200):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.B200Error):
        b200.GridMap((1.0, 1.0, 1.0), 0.1)
    p = b200.Planner("thetastar", resolution=0.1)          # planner handles are host-only until a map is bound
    with pytest.raises(b200.B200Error):
        p.findPaths(np.zeros((1, 3)), np.ones((1, 3)))      # no map -> B200_ERR_NO_MAP, never a CPU search


def test_planner_parameter_validation(b200):
    with pytest.raises(b200.B200Error):
        b200.Planner("astar", resolution=0.0005)            # stencil loop of AStar.cpp:131 needs r > 1e-3
    with pytest.raises(b200.B200Error):
        b200.Planner("astar", resolution=0.1, allocate_num=1)
    assert b200.Planner("no-such-algorithm", resolution=0.1).algorithm == "astar"   # planner_ros_node.cpp:280-284


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "3d-astar-path-planners_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "pyoracle" not in txt and "oracle/" not in txt.replace("oracle/oracle.cpp", ""), f
