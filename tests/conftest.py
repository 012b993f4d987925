import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as o  # oracle/oracle.py (test infrastructure)
    o.build()
    return o


@pytest.fixture(scope="session")
def kmath():
    """Host build of finch_rs_b200/csrc/common.cuh's inline arithmetic (test helper)."""
    import ctypes as C
    import subprocess
    src = os.path.join(ROOT, "tests", "kernel_math_host.cpp")
    hdr = os.path.join(ROOT, "finch_rs_b200", "csrc", "common.cuh")
    out = os.path.join(ROOT, "tests", "_kernel_math_host.so")
    if (not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr))):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    L = C.CDLL(out)
    L.km_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.km_stream.restype = C.c_size_t
    L.km_classify.argtypes = [C.c_uint8]
    L.km_classify.restype = C.c_uint8
    L.km_murmur_bytes.argtypes = [C.c_char_p, C.c_uint32, C.c_uint64]
    L.km_murmur_bytes.restype = C.c_uint64
    L.km_codes_to_ascii.argtypes = [C.c_uint64, C.c_int, C.c_void_p]
    return L


def have_gpu():
    try:
        import finch_rs_b200 as fb
        return fb.lib().fb2_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def synth():
    """tools/synth.py: synthetic-input generators (test / bench infrastructure, not in the product library)."""
    import synth as s
    s.lib()
    return s


@pytest.fixture(scope="session")
def fb():
    import finch_rs_b200 as m
    m.lib()
    return m
