"""The C++ host-side mirror of the crate API (include/entropy_coders.hpp): it compiles against the C ABI
(CPU check), and the crate's tests re-expressed in C++ pass on the GPU with the oracle as checker."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"
SRC = os.path.join(ROOT, "tests", "cpp", "test_crate_api.cpp")


def _compile(out, extra):
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), "-I", CUDA_INC,
           SRC] + extra
    subprocess.check_call(cmd)
    return out


def test_cpp_mirror_compiles(tmp_path):
    _compile(None, ["-fsyntax-only"])


@pytest.mark.gpu
def test_cpp_mirror_runs_reference_tests(tmp_path):
    import oracle_lib
    from entropy_coders_b200 import build
    oracle_lib.build()
    build.build()
    exe = str(tmp_path / "test_crate_api")
    lib_dir = os.path.join(ROOT, "entropy_coders_b200")
    ora_dir = os.path.join(ROOT, "oracle", "_build")
    _compile(exe, ["-o", exe, "-L", lib_dir, "-lfse_b200", "-L", ora_dir, "-lfse_oracle", "-L", "/usr/local/cuda/lib64", "-lcudart",
                   "-Wl,-rpath," + lib_dir, "-Wl,-rpath," + ora_dir, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "cpp crate-API tests ok" in out.stdout
