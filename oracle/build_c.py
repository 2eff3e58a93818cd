"""TEST INFRASTRUCTURE ONLY -- compile oracle/asso_c.c into oracle/libasso_oracle.so with gcc.

    python oracle/build_c.py [--force]

The reference is pure Python, so there is nothing to compile into oracle/_ref/; this builds the C
restatement only.  The .so is git-ignored but travels to the GPU box with the snapshot.  No -march=native:
the box's host CPU may differ, the AVX-512 VPOPCNTQ path is selected at run time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "asso_c.c")
LIB = os.path.join(HERE, "libasso_oracle.so")
FLAGS = ["-O3", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-mpopcnt", "-std=gnu11", "-Wall"]


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    subprocess.run(["gcc", *FLAGS, SRC, "-o", LIB], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
