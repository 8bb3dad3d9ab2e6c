"""Build the C restatement (oracle/c_oracle.c) into oracle/_build/.

TEST INFRASTRUCTURE (see oracle/__init__.py).  -ffp-contract=off keeps every
multiply and add separately rounded, like NumPy.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = os.path.join(HERE, "c_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUTPUT = os.path.join(OUT_DIR, "libmtoracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.isfile(OUTPUT)
            and os.path.getmtime(OUTPUT) >= os.path.getmtime(SOURCE)):
        return OUTPUT
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp",
           "-ffp-contract=off", "-fno-fast-math", "-o", OUTPUT, SOURCE, "-lm"]
    subprocess.run(cmd, check=True)
    return OUTPUT


if __name__ == "__main__":
    print(build(force=True))
