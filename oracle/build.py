"""Build the C oracle (test infrastructure) into oracle/libqldpc_oracle.so.

Strict IEEE flags on purpose: the reference relies on inf/NaN conventions
(src/decoding/kernels.py:327-333) that -ffast-math would break."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "qldpc_oracle.c")
OUT = os.path.join(HERE, "libqldpc_oracle.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
