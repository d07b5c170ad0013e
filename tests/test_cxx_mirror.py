"""The C++ host mirror (tensor_b200/cxx/tensor_b200.hpp): type-indexed Batch<T, dims...>
classes over the C ABI, standing in for the Haskell class instances (no GHC in the image).
CPU: the header and its test program compile and link.  GPU: the program runs and every
operation is bit-identical to the oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "test_host_mirror")


def compile_mirror_test():
    import oracle
    oracle.build()
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "cxx", "test_host_mirror.cpp"),
           "-L" + os.path.join(ROOT, "tensor_b200"), "-lt5b200", "-L" + os.path.join(ROOT, "oracle"), "-lt5oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "tensor_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    subprocess.check_call(cmd)


def test_cxx_mirror_compiles_and_links():
    compile_mirror_test()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_cxx_mirror_runs_bit_exact():
    compile_mirror_test()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout and "MISMATCH" not in out.stdout
