"""Runs the C++ facade tests (tests/cpp/test_facade.cpp): the reference's planning-layer API on the CUDA library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_facade(gpu):
    exe = os.path.join(ROOT, "tests", "cpp", "test_facade")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(p.stdout, p.stderr)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "0 failed" in p.stdout


def test_cpp_facade_compiles():
    """Host-only check: the facade headers and the C ABI header compile as plain C++17 / C."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "cpp")])
    subprocess.check_call(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "mgodpl_b200.h")])
