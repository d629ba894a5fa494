import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def restatement():
    from oracle import oracle_py
    return oracle_py.restatement()


@pytest.fixture(scope="session")
def reference_binary():
    """The reference's shipped solver (oracle/_ref); None where it has not been built."""
    from oracle import oracle_py
    return oracle_py.reference()


@pytest.fixture(scope="session")
def emu_lib():
    """CPU build of the DEVICE solver body under the fiber warp emulator (tests/emu)."""
    import ctypes as C
    emu = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(emu, "libmgqp_emu.so")
    srcs = [os.path.join(emu, "emu_small.cc"), os.path.join(emu, "warp_emu.h")]
    csrc = os.path.join(ROOT, "motion-generation-using-quadratic-programs_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DMGQP_EMULATE", "-shared", "-fPIC",
                               "-pthread", "-o", so, os.path.join(emu, "emu_small.cc")])
    lib = C.CDLL(so)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.emu_solve_batch.restype = C.c_int
    lib.emu_solve_batch.argtypes = [C.c_int] * 3 + [C.c_longlong] + [dp] * 8 + [ip] * 4 + [C.c_int] * 4
    return lib


def default_max_iter(p, m):
    return 100 + 10 * (p + m)
