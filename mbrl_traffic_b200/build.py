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
# translation units (compiled in parallel, linked into one shared object)
SOURCES = [os.path.join(HERE, "csrc", "mtraffic.cu"),
           os.path.join(HERE, "csrc", "mt_pipe_arz64.cu")]
HEADERS = sorted(
    os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))
    if f.endswith((".cuh", ".h", ".inc"))) + [os.path.join(ROOT, "include", "mtraffic.h")]
OUTPUT = os.path.join(HERE, "libmtraffic.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    # kernels of one translation unit are launched from another (extern template)
    "-static-global-template-stub=false",
    "-I", os.path.join(ROOT, "include"),
]
# No `-split-compile`: it halves the build time of the big translation unit but
# changes the code of individual kernels, differently for every thread count
# (measured on one box, split 8 vs none: headline step 25.0 vs 23.0 us, ARZ f32
# EXACT 0.71 vs 0.84 of the HBM peak, ARZ f32 FAST 0.87 vs 0.94, non-local 0.56
# vs 0.62, fused 100-step rollout 1.96e11 vs 2.10e11): the library must not
# depend on the number of cores of the machine that built it.
SPLIT_FLAGS = {}


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
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + extra_flags() + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        cmd = [nvcc] + flags + SPLIT_FLAGS.get(os.path.basename(src), []) + ["-c", "-o", obj, src]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              universal_newlines=True)
        return obj, cmd, proc

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    for obj, cmd, proc in results:
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed (%d): %s" % (proc.returncode, " ".join(cmd)))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + \
        [obj for obj, _, _ in results]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          universal_newlines=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed (%d): %s" % (proc.returncode, " ".join(cmd)))
    os.replace(tmp, OUTPUT)
    return OUTPUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
