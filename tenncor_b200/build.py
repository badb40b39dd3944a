"""Build recipes for the native pieces (run by __graft_entry__.build()).

Everything is compiled in-tree with explicit nvcc / g++ commands for sm_100a only:
  tenncor_b200/lib/libtcr_b200.so   CUDA kernels + C-ABI (include/tcr_b200.h)
  tenncor_b200/_tenncor*.so         C++ host (teq/eteq/layr mirror) + pybind11 module
  oracle/_build/libtcr_oracle.so    CPU restatement used ONLY by tests / bench cpu_baseline
"""
import hashlib
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tenncor_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
INCLUDE = os.path.join(ROOT, "include")
BUILD = os.path.join(ROOT, "build")
LIBDIR = os.path.join(PKG, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-I" + INCLUDE]

LIB_PATH = os.path.join(LIBDIR, "libtcr_b200.so")
ORACLE_SRC = os.path.join(ROOT, "oracle", "tcr_oracle.c")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libtcr_oracle.so")


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX") or ".so"


HOST_LIB = os.path.join(PKG, "_tenncor" + _ext_suffix())


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def _digest(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    return h.hexdigest()


def _stale(target, sources, extra=""):
    stamp = target + ".stamp"
    d = _digest(sources, extra)
    if os.path.exists(target) and os.path.exists(stamp) and open(stamp).read() == d:
        return None
    return d


def _headers(d):
    return [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".hpp", ".cuh"))]


def build_cuda(verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    cus = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = _headers(CSRC) + _headers(INCLUDE)
    d = _stale(LIB_PATH, cus + hdrs, " ".join(NVCC_FLAGS))
    if d is None:
        return LIB_PATH

    def one(cu):
        obj = os.path.join(BUILD, os.path.basename(cu)[:-3] + ".o")
        od = _stale(obj, [cu] + hdrs, " ".join(NVCC_FLAGS))
        if od is not None:
            _run([NVCC] + NVCC_FLAGS + ["-c", cu, "-o", obj])
            open(obj + ".stamp", "w").write(od)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(cus))) as ex:
        objs = list(ex.map(one, cus))
    _run([NVCC] + ARCH + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart", "-lcuda", "-ldl"])
    open(LIB_PATH + ".stamp", "w").write(d)
    return LIB_PATH


def build_host(verbose=False):
    if not os.path.isdir(HOST):
        return None
    import pybind11

    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    if not srcs:
        return None
    hdrs = _headers(HOST) + _headers(INCLUDE)
    flags = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-I" + INCLUDE, "-I" + HOST,
             "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]
    d = _stale(HOST_LIB, srcs + hdrs, " ".join(flags))
    if d is None:
        return HOST_LIB

    def one(src):
        obj = os.path.join(BUILD, "host_" + os.path.basename(src)[:-4] + ".o")
        od = _stale(obj, [src] + hdrs, " ".join(flags))
        if od is not None:
            _run(["g++"] + flags + ["-c", src, "-o", obj])
            open(obj + ".stamp", "w").write(od)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(one, srcs))
    _run(["g++", "-shared", "-o", HOST_LIB] + objs +
         ["-L" + LIBDIR, "-ltcr_b200", "-Wl,-rpath,$ORIGIN/lib", "-ldl"])
    open(HOST_LIB + ".stamp", "w").write(d)
    return HOST_LIB


def build_oracle(verbose=False):
    """CPU restatement (test infrastructure only; never linked into the product)."""
    if not os.path.exists(ORACLE_SRC):
        return None
    os.makedirs(os.path.dirname(ORACLE_LIB), exist_ok=True)
    flags = ["-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-I" + INCLUDE]
    d = _stale(ORACLE_LIB, [ORACLE_SRC], " ".join(flags))
    if d is None:
        return ORACLE_LIB
    _run(["gcc"] + flags + [ORACLE_SRC, "-o", ORACLE_LIB, "-lm"])
    open(ORACLE_LIB + ".stamp", "w").write(d)
    return ORACLE_LIB


def build_all(verbose=False):
    out = {"cuda": build_cuda(verbose), "host": build_host(verbose), "oracle": build_oracle(verbose)}
    return out


if __name__ == "__main__":
    print(build_all(verbose=True))
