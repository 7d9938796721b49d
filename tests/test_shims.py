"""The C++ drop-in shims (xrsfm_b200/shim/*.h) compile against stand-ins of the reference
types and link against libxrsfm_b200.so (no GPU needed: the binary only constructs objects)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shims_compile_link_and_run(tmp_path):
    exe = tmp_path / "shim_check"
    cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "mock", "shim_check.cpp"), "-o", str(exe),
           "-L", os.path.join(ROOT, "xrsfm_b200"), "-lxrsfm_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "xrsfm_b200")]
    subprocess.check_call(cmd)
    out = subprocess.check_output([str(exe)], text=True)
    assert "flatten: 0 cams" in out and "VerifyContextGL=" in out
    assert "ftr roundtrip: ok" in out
    # LBA window: frames 0..3 (set order), 11 observations, gauge frame 0 fixed; point 0 free (seen by the
    # new frame, small angle), point 1 constant (angle > 5), point 2 constant (not seen by the new frame)
    assert "lba: cams=4 obs=11 pts=3 fixed_t=1000 fixed_pts=011" in out
    assert "lba2: fixed_t=0011" in out
