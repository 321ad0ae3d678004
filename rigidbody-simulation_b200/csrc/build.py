"""Build librbsim_b200.so in-tree with nvcc for sm_100a (B200).

    python rigidbody-simulation_b200/csrc/build.py [--force] [--verbose]

The library is self-contained (static cudart) and exports only the C ABI of include/rbsim_b200.h.
It is git-ignored (*.so) but travels to the GPU box with the gpurun snapshot.
-fmad=false: the kernels keep the reference's NumPy rounding sequence (see rbs_kernels.cuh).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "lib", "librbsim_b200.so")
SOURCES = [os.path.join(HERE, "rbs_capi.cu")]
DEPS = SOURCES + [os.path.join(HERE, "rbs_kernels.cuh"), os.path.join(os.path.dirname(PKG), "include", "rbsim_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-cudart", "static",
]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found")


def up_to_date():
    return os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS)


def build(force=False, verbose=False, extra=()):
    if not force and up_to_date():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = OUT + ".building"                 # linked next to the target and renamed: a snapshot never sees half a library
    cmd = [nvcc_path()] + NVCC_FLAGS + list(extra) + ["-o", tmp] + SOURCES
    env = dict(os.environ)
    # the image's CC wrapper is fine as nvcc's host compiler; keep PATH as is
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed building librbsim_b200.so")
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    extra = ["-Xptxas", "-v"] if "--ptxas" in sys.argv else []
    print(build(force="--force" in sys.argv or bool(extra), verbose="--verbose" in sys.argv or bool(extra), extra=extra))
