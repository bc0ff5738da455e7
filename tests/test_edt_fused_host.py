"""CPU (-m "not gpu"): the per-thread building blocks of the fused EDT kernels (csrc/edt_fused.cuh: band hull scan, zipper
merge, prefix, hull walk, packed (dy, dz) intermediate, bit-word z distances) compiled for the host and run phase by phase
against a brute-force separable transform — every (bands, lanes, entry type) family the kernels instantiate."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_edt_building_blocks_match_brute_force(tmp_path):
    exe = str(tmp_path / "edt_fused_host_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe,
                           os.path.join(ROOT, "tests", "cpu", "edt_fused_host_test.cpp")])
    out = subprocess.run([exe, "7"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "ALL OK" in out.stdout
