"""Builds libresynthesizer_b200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
# RS_LIB_VARIANT=<suffix> with RS_NVCC_EXTRA=<-D...> builds a parameter-sweep variant beside the product library
LIB = os.path.join(LIBDIR, "libresynthesizer_b200%s.so" % os.environ.get("RS_LIB_VARIANT", ""))
SOURCES = ["rs_kernels.cu", "host_engine.cpp", "host_prep.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "--extended-lambda", "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=default", "-shared",
         "-cudart", "static"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = os.environ.get("RS_NVCC_EXTRA", "").split()
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed building %s" % LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
