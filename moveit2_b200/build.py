"""Build recipe for libb200sv.so (in-tree, so the built library travels to the GPU box with the snapshot).

    python -m moveit2_b200.build [--force] [-v]

nvcc cross-compiles for sm_100a without a GPU.  -lineinfo keeps the ncu source page mapped to our code.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200sv.so")
SOURCES = ["api.cu", "check_kernels.cu", "fast_check.cu", "edt_kernels.cu", "propagation.cu", "motion_kernels.cu", "host_model.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-Wall,-Wno-unknown-pragmas", "-shared", "--expt-relaxed-constexpr", "-lz"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200sv.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: build an experimental variant (e.g. -DB200SV_SPHERE_CHUNK=8) next to the default library.
    Every source is compiled by its own nvcc process (in parallel), then linked."""
    from concurrent.futures import ThreadPoolExecutor
    out = out or LIB
    if not force and out == LIB and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build", os.path.basename(out) + ".obj")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f not in ("-shared", "-lz")] + ["-D" + d for d in defines] + \
             (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        subprocess.check_call([nvcc] + cflags + ["-c", "-o", obj, os.path.join(CSRC, src)])
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["-lz"])
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="-v" in sys.argv, defines=defs,
                out=outs[0] if outs else None))
