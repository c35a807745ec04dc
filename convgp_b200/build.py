"""Build libconvgp_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libconvgp_b200.so")
SOURCES = ["api.cu", "probe.cu", "comm.cu", "linalg.cu", "likelihood.cu"]
HEADERS = ["common.cuh", "fast.cuh", "kuu_fast.cuh", "mma64.cuh", "mma32.cuh", "umma32.cuh", "umma16.cuh", "generic.cuh", "linalg.cuh", os.path.join("..", "..", "include", "convgp_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "--extended-lambda"] + os.environ.get("CGP_NVCC_EXTRA", "").split()


def _obj(src):
    return os.path.join(CSRC, src.replace(".cu", ".o"))


_ABI = os.path.join("..", "..", "include", "convgp_b200.h")
SOURCE_HEADERS = {  # which headers each translation unit includes (api.cu: every kernel family)
    "api.cu": [h for h in HEADERS if h != "linalg.cuh"],
    "probe.cu": ["common.cuh", _ABI],
    "comm.cu": ["common.cuh", _ABI],
    "linalg.cu": ["common.cuh", "linalg.cuh", _ABI],
    "likelihood.cu": ["common.cuh", _ABI],
}


def _deps(src):
    return [os.path.join(CSRC, f) for f in SOURCE_HEADERS.get(src, HEADERS)] + [os.path.abspath(__file__)]


def _stale_sources():
    """Sources whose object is older than the source itself, any header or this script."""
    out = []
    for src in SOURCES:
        obj, path = _obj(src), os.path.join(CSRC, src)
        if not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in [path] + _deps(src)):
            out.append(src)
    return out


def build(force=False, verbose=False):
    """Compile every stale .cu of the package (in parallel) and link one shared library.  Returns the library path."""
    todo = list(SOURCES) if force else _stale_sources()
    if not todo and os.path.exists(LIB) and all(os.path.getmtime(_obj(s)) <= os.path.getmtime(LIB) for s in SOURCES):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs = []
    for src in todo:
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", _obj(src)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % src)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [_obj(s) for s in SOURCES] + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
