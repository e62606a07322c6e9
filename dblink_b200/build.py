"""Builds libdblink_b200.so in-tree with nvcc for sm_100a (explicit nvcc -shared; no JIT cache)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libdblink_b200.so")
SOURCES = ["dbl_engine.cu", "dbl_host.cpp"]
HEADERS = ["dbl_internal.h", os.path.join(ROOT, "include", "dblink_b200.h")]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [
        nvcc_path(), "-shared", "-o", SO,
        "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
        # the draw protocol is defined over individually rounded binary64 operations: no FMA contraction
        "-fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2,-pthread",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-cudart", "static",
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libdblink_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
