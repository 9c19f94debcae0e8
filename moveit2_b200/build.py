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
SOURCES = ["api.cu", "check_kernels.cu", "fast_check.cu", "edt_kernels.cu", "motion_kernels.cu", "host_model.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler",
              "-fPIC,-Wall,-Wno-unknown-pragmas", "-shared", "--expt-relaxed-constexpr", "-lz"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b200sv.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines / out: build an experimental variant (e.g. -DB200SV_SPHERE_CHUNK=8) next to the default library."""
    out = out or LIB
    if not force and out == LIB and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv or bool(defs), verbose="-v" in sys.argv, defines=defs,
                out=outs[0] if outs else None))
