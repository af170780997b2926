"""The C++ host mirror (wgmath_b200/host/wgebra_b200.hpp): it must compile and link against the C ABI
on CPU (not gpu), and on the GPU box it replays the reference's four unit tests against the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "replay_reference_tests")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build_exe():
    from oracle import oracle
    from wgmath_b200 import build as b
    b.build()
    oracle.build()
    src = EXE + ".cpp"
    deps = [src, os.path.join(ROOT, "wgmath_b200", "host", "wgebra_b200.hpp"), os.path.join(ROOT, "include", "wgb200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call([CXX, "-std=c++17", "-O2", "-Wall", "-o", EXE, src,
                               "-L" + os.path.join(ROOT, "wgmath_b200"), "-lwgebra_b200",
                               "-L" + os.path.join(ROOT, "oracle"), "-lwgsl_oracle",
                               "-Wl,-rpath," + os.path.join(ROOT, "wgmath_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return EXE


def test_cpp_mirror_compiles_and_links():
    exe = build_exe()
    assert os.path.exists(exe)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    # without a GPU the program must fail loudly with the library's NO_DEVICE error, never compute on the CPU
    if r.returncode != 0:
        assert r.returncode == 2 and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_mirror_replays_reference_tests():
    exe = build_exe() if not os.path.exists(EXE) else EXE
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
