"""The C++ drop-in layer (include/simd_scan.hpp + libsimd_scan_host.so): the reference's own unit
tests, restated in tests/cpp/dropin_test.cpp with reference names only, run on the GPU box."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("binary", ["dropin_test_w09", "dropin_test_w17"])
def test_reference_unit_tests_through_dropin_header(binary):
    exe = os.path.join(HERE, "cpp", binary)
    assert os.path.exists(exe), "build it first: python -c 'import __graft_entry__ as g; g.build()'"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "All tests passed" in r.stdout
