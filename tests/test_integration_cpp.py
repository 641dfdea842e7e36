"""Drop-in check: the reference's own mm/ headers (unmodified, from /root/reference) plus
include/g4s_b200.hpp in place of mkl_mult.h compile and link against libg4s_b200.so with the reference's
mkl(A, B, C, timing) call shape (oracle/dropin_mkl.cpp -> oracle/_ref/dropin_mkl, built by oracle/Makefile where the
reference tree exists).  The binary travels to the GPU box, where the run half executes it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "dropin_mkl")


def test_reference_call_site_builds_against_the_library():
    if not os.path.exists("/root/reference/mm/inc/all.h"):
        pytest.skip("reference tree absent")
    subprocess.run(["make", "-C", os.path.join(ROOT, "g4s_b200", "csrc")], check=True, capture_output=True)
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-B", "_ref/dropin_mkl"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_reference_call_site_runs_on_gpu():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/dropin_mkl was not built (reference tree absent at build time)")
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "equal 1" in out.stdout
