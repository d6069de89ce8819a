"""Builds entropy_coders_b200/libfse_b200.so in-tree with nvcc for sm_100a (no JIT cache, no CPU path)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "csrc", "fse_b200.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "fse_bitio.cuh"), os.path.join(HERE, "csrc", "fse_zstd_norm.cuh"), os.path.join(HERE, "csrc", "fse_shared_enc.cuh"), os.path.join(HERE, "csrc", "fse_shared_dec.cuh"), os.path.join(HERE, "csrc", "fse_tps.cuh"), os.path.join(HERE, "csrc", "fse_decode128c.cuh"), os.path.join(HERE, "csrc", "fse_encode128.cuh"), os.path.join(HERE, "csrc", "fse_kernels128.cuh"), os.path.join(HERE, "csrc", "fse_decode64w.cuh"), os.path.join(HERE, "csrc", "fse_hist16.cuh"), os.path.join(HERE, "csrc", "fse_decode64c.cuh"), os.path.join(HERE, "csrc", "fse_kernels64.cuh"), os.path.join(HERE, "csrc", "fse_kernels.cuh"), os.path.join(HERE, "csrc", "fse_device.cuh"),
        os.path.join(ROOT, "include", "fse_b200.h")]
OUT = os.path.join(HERE, "libfse_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC, "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return OUT


def build_variant(name, defines=()):
    """development builds (tools/bin/lib<name>.so, selected with FSE_B200_LIB): -DFSE_DEV enables the environment overrides"""
    out = os.path.join(ROOT, "tools", "bin", "lib%s.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DFSE_DEV",
           "-Xcompiler", "-fPIC", "-shared", "-o", out, SRC, "-lcudart"] + ["-D" + d for d in defines]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
