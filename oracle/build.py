"""Build recipe for the oracle's C restatement (test infrastructure).

`python -m oracle.build` compiles oracle/b2d_oracle.c into oracle/libb2d_oracle.so with gcc.
The reference itself (Rust + OpenGL) cannot be built here -- no rustc/cargo, crates not vendored,
no GL context -- so there is no oracle/_ref; DESIGN.md records this.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "b2d_oracle.c")
OUT = os.path.join(HERE, "libb2d_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-std=c11", "-fopenmp", "-fPIC", "-shared", "-fwrapv",
           "-Wall", "-Wextra", "-Wno-unused-parameter", "-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
