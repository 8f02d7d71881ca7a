"""Host-side code of the library checked on the CPU (tests/native/plan_check.cu, compiled with nvcc as a
host program and linked against libmqcb200.so): invariants of the launch planners and the packed layout,
the unit decode the accumulation kernel shares with its planner, the lower-triangle gather of set_tensor
between guard pages, and the DIIS solve of the any-size device SCF against the reference's own DIIS test
(test/test_mqc_diis.f90: its filler, sizes, seeds and comparison algorithm)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_planner_and_layout_invariants(tmp_path):
    if not shutil.which("nvcc"):
        pytest.skip("nvcc not on PATH")
    from metalquicha_b200 import _lib
    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "plan_check")
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    os.path.join(ROOT, "tests", "native", "plan_check.cu"), "-o", exe,
                    "-L" + libdir, "-lmqcb200", "-Xlinker", "-rpath=" + libdir, "-lcudart"],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "invariants hold" in out.stdout
