"""The C++ host-side mirror of the reference interface (include/mqcb200.hpp) builds with a plain host
compiler, warnings as errors, links against libmqcb200.so, and behaves on a box without a GPU: the work
queue works (it is host code), the engine refuses to exist (tests/native/hpp_check.cpp)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_cpp_mirror_compiles_links_and_refuses_without_a_device(tmp_path):
    if not shutil.which("g++"):
        pytest.skip("g++ not on PATH")
    from metalquicha_b200 import _lib
    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "hpp_check")
    build = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-pthread",
                            os.path.join(ROOT, "tests", "native", "hpp_check.cpp"), "-o", exe,
                            "-L" + libdir, "-lmqcb200", "-Wl,-rpath," + libdir],
                           capture_output=True, text=True)
    assert build.returncode == 0, build.stderr[-3000:]
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "C++ mirror checks hold" in out.stdout


def test_cpp_mirror_wraps_only_declared_symbols_and_names_the_reference_routines():
    import re
    hpp = open(os.path.join(ROOT, "include", "mqcb200.hpp")).read()
    header = open(os.path.join(ROOT, "include", "mqcb200.h")).read()
    declared = set(re.findall(r"\b(mqcb200_\w+)\s*\(", header))
    used = set(re.findall(r"\b(mqcb200_\w+)\s*\(", hpp))
    assert used and used <= declared, sorted(used - declared)
    for routine in ("build_fock_df", "metric_inverse_sqrt", "build_df_tensor", "response_operator_df",
                    "fitted_potential_general", "df_gradient_densities", "electronic_energy", "assemble_fock"):
        assert re.search(r"\b%s\s*\(" % routine, hpp), routine
