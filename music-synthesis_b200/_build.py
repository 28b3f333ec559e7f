"""Builds csrc/*.cu into csrc/libmsb200.so for sm_100a with nvcc (in-tree, so the
shared object travels to the GPU box with the repo snapshot)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libmsb200.so")
SOURCES = ["runtime.cu", "conv_gemm.cu", "layout.cu", "generator.cu", "audio2mel.cu",
           "resstack.cu", "upstack.cu", "conv_gemm2.cu", "direct_conv.cu", "losses.cu", "fft_bands.cu",
           "wgrad.cu", "backward.cu", "datafeed.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
]


def _extra_flags():
    """MSB_NVCC_EXTRA="-DMSB_STACK_TRACE" enables the clock64 trace points."""
    return os.environ.get("MSB_NVCC_EXTRA", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(CSRC), "..", "include", "msb200.h"))
    nvcc = _nvcc()
    objs, jobs = [], []
    # a change of flags (MSB_NVCC_EXTRA debug builds) must not leave stale objects behind
    flags = " ".join(NVCC_FLAGS + _extra_flags())
    stamp = os.path.join(CSRC, ".build_flags")
    if not os.path.exists(stamp) or open(stamp).read() != flags:
        force = True
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = s[:-3] + ".o"
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + _extra_flags() + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            sys.stderr.write(r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs)
    with open(stamp, "w") as f:
        f.write(flags)
    return LIB


def build_debug_library(force=False):
    """tools/native/libmsb200_dbg.so: the measurement kernels of tools/microbench.py (not part of
    the product library)."""
    root = os.path.join(os.path.dirname(CSRC), "..", "tools", "native")
    src = os.path.join(root, "microbench.cu")
    lib = os.path.join(root, "libmsb200_dbg.so")
    deps = [src, os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "runtime.cuh"),
            os.path.join(CSRC, "runtime.cu")]
    if force or _stale(lib, deps):
        r = subprocess.run([_nvcc()] + NVCC_FLAGS + ["-shared", "-o", lib, src,
                                                     os.path.join(CSRC, "runtime.cu")],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed (debug library):\n%s%s" % (r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
