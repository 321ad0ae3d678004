"""Build the CPU oracle shared library (test infrastructure) -> oracle/_build/librb_oracle.so.

Usage: python oracle/build_oracle.py
Flags: -O2 -ffp-contract=off (no FMA contraction: the reference's NumPy arithmetic rounds every
operation) and OpenMP when the compiler has it (the env CC in this image lacks libgomp.spec, so
/usr/bin/gcc is tried first).  *.so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "librb_oracle.so")
SRC = [os.path.join(HERE, "rb_oracle.c"), os.path.join(HERE, "rb_oracle_body.h")]


def build(force=False, verbose=False):
    if not force and os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    base = ["-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-Wall", "-Wno-unused-function"]
    errors = []
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            cmd = [cc] + base + omp + ["-o", OUT, SRC[0], "-lm"]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True)
            except FileNotFoundError as exc:
                errors.append(str(exc))
                continue
            if r.returncode == 0:
                if verbose:
                    print("built", OUT, "with", " ".join(cmd))
                return OUT
            errors.append(r.stderr.strip()[-400:])
    raise RuntimeError("could not build the CPU oracle:\n" + "\n".join(errors))


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
