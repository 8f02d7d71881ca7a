"""The product never imports, links or executes anything under oracle/."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_and_csrc_do_not_reference_the_oracle():
    """The package (kernels, C ABI, Python mirror, Fortran sources) and the public headers (C and C++)."""
    offenders = []
    for top in (os.path.join(ROOT, "metalquicha_b200"), os.path.join(ROOT, "include")):
        for dirpath, _, files in os.walk(top):
            if os.path.basename(dirpath) == "build":
                continue
            for f in files:
                if not f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".f90", "Makefile")):
                    continue
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|df_fock_oracle|df_fock_ref|gto_integrals|scf_oracle|oracle/", text, re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert offenders == []


def test_numpy_is_not_used_for_contractions_in_the_engine_mirror():
    """engine.py may diagonalise the guess density (as the reference's caller does with
    dsyev) but must not contract B with anything on the host."""
    text = open(os.path.join(ROOT, "metalquicha_b200", "engine.py")).read()
    assert "einsum" not in text and "tensordot" not in text and "matmul" not in text
    # no host matrix product at all since round 2: metric^(-1/2) is formed on the device too
    assert text.count(" @ ") == 0
    assert text.count("linalg.eigh") == 1          # density_pseudo_orbitals: the reference's CALLER does this dsyev
