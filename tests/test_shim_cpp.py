"""The header-only C++ host side (include/q3dr/occupied_tree_cuda.hpp): compiles everywhere, runs on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_shim.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "test_shim")


def _build():
    cmd = ["/usr/bin/g++", "-std=c++11", "-O1", "-Wall", "-Wextra", "-ffp-contract=off", "-o", EXE, SRC,
           "-L" + os.path.join(ROOT, "quad3dr_b200"), "-lq3dr_raycast", "-L" + os.path.join(ROOT, "oracle"), "-lq3dr_oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "quad3dr_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "warning" not in r.stderr, r.stderr


def test_shim_compiles_as_cxx11_and_links():
    """The reference is C++11: the shim must not need anything newer."""
    _build()
    r = subprocess.run([EXE, "--no-run"], capture_output=True, text=True)
    assert r.returncode == 0 and "compiled and linked" in r.stdout


@pytest.mark.gpu
def test_shim_parity_against_oracle():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shim parity OK" in r.stdout
