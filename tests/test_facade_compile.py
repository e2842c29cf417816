"""The C++ facade classes (reference signatures over the C ABI) must compile against mock reference types."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_facade_headers_compile():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    r = subprocess.run([gxx, "-std=c++14", "-fsyntax-only", "-Wall", os.path.join(ROOT, "tests", "facade_check.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
