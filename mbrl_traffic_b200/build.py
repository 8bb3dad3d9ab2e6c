"""Build libmtraffic.so in-tree with nvcc for sm_100a.

    python -m mbrl_traffic_b200.build

nvcc cross-compiles without a GPU.  The shared object lands next to this
package (git-ignored, but it travels with the repo snapshot to the GPU box).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", "mtraffic.cu")]
HEADERS = sorted(
    os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))
    if f.endswith((".cuh", ".h"))) + [os.path.join(ROOT, "include", "mtraffic.h")]
OUTPUT = os.path.join(HERE, "libmtraffic.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    # one translation unit, ~290 kernels: let ptxas work on them in parallel
    "-split-compile", str(max(1, min(8, os.cpu_count() or 1))),
    "-I", os.path.join(ROOT, "include"),
]


def extra_flags():
    """MT_NVCC_FLAGS: extra compiler flags (experiment builds, e.g. -DMT_TRACE)."""
    return os.environ.get("MT_NVCC_FLAGS", "").split()


def needs_build():
    if not os.path.isfile(OUTPUT):
        return True
    built = os.path.getmtime(OUTPUT)
    return any(os.path.getmtime(p) > built for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile the CUDA library; returns the path of the shared object."""
    if not force and not needs_build():
        return OUTPUT
    nvcc = os.environ.get("NVCC", "nvcc")
    # compile next to the target and rename: a reader (or a snapshot of the
    # tree) never sees a half-written library
    tmp = OUTPUT + ".tmp.so"
    cmd = [nvcc] + NVCC_FLAGS + extra_flags() + (["-Xptxas", "-v"] if verbose else []) + \
        ["-o", tmp] + SOURCES
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          universal_newlines=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (%d): %s" % (proc.returncode,
                                                     " ".join(cmd)))
    os.replace(tmp, OUTPUT)
    return OUTPUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
