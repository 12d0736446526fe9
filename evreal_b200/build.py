"""Builds evreal_b200/libevreal_b200.so (plus the oracle's C pieces if any) in-tree with nvcc for sm_100a.

Usage: python -m evreal_b200.build [--force]
The library is a plain C-ABI shared object (no torch headers): it is loaded with
ctypes by evreal_b200._lib.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libevreal_b200.so")
SOURCES = ["api.cu", "voxelize.cu", "metrics.cu", "conv_simt.cu", "conv_tc.cu", "poly.cu", "hyper.cu", "spade.cu", "etnet.cu", "model.cu", "lpips.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "evreal_b200.h"),
                                                               os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    builddir = os.path.join(HERE, "build")
    os.makedirs(builddir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(builddir, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (src, out))
        failed |= p.returncode != 0
    with open(os.path.join(builddir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed (see above)")
    if verbose:
        print("\n".join(log))
    cmd = [nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + ["-lcudart", "-lpthread"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
