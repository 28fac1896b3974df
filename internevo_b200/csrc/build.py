"""In-tree build of the sm_100a extension (``internevo_b200/_C.so``).

Kernels are plain ``.cu`` translation units without torch headers (nvcc cross-compiles them in seconds, no GPU needed);
``bindings.cpp`` is the only torch-dependent file and is compiled with the host compiler.  The resulting shared object
registers its ops under ``torch.ops.b200`` and is loaded with ``torch.ops.load_library`` (see ``internevo_b200/ops/_lib.py``).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
BUILD_DIR = os.path.join(HERE, "build")
SO_PATH = os.path.join(PKG, "_C.so")

CU_SOURCES = ["gemm_sm100.cu", "elementwise.cu", "attention_sm100.cu", "comm_kernels.cu", "moe_comm.cu", "layernorm.cu"]
CPP_SOURCES = ["bindings.cpp"]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _cuda_home() -> str:
    for cand in (os.environ.get("CUDA_HOME"), os.environ.get("CUDA_PATH"), "/usr/local/cuda"):
        if cand and os.path.exists(os.path.join(cand, "bin", "nvcc")):
            return cand
    nvcc = shutil.which("nvcc")
    if nvcc:
        return os.path.dirname(os.path.dirname(nvcc))
    raise RuntimeError("nvcc not found; set CUDA_HOME")


def _digest(paths, extra=()):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    for e in extra:
        h.update(str(e).encode())
    return h.hexdigest()[:16]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


DATAIO_SO = os.path.join(PKG, "_dataio.so")


def build_dataio(force: bool = False) -> str:
    """Host-only helper library of the data pipeline (``dataio.cpp``: corpus scan, token-line parser; plain C ABI, loaded with
    ctypes) -> ``internevo_b200/_dataio.so``."""
    src = os.path.join(HERE, "dataio.cpp")
    os.makedirs(BUILD_DIR, exist_ok=True)
    tag = os.path.join(BUILD_DIR, "dataio." + _digest([src]))
    if force or not os.path.exists(DATAIO_SO) or not os.path.exists(tag):
        _run(["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", DATAIO_SO, src])
        for f in os.listdir(BUILD_DIR):
            if f.startswith("dataio."):
                os.remove(os.path.join(BUILD_DIR, f))
        open(tag, "w").close()
    return DATAIO_SO


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile everything that is out of date and link ``_C.so``; returns its path."""
    build_dataio(force)
    import torch
    from torch.utils import cpp_extension

    os.makedirs(BUILD_DIR, exist_ok=True)
    cuda = _cuda_home()
    nvcc = os.path.join(cuda, "bin", "nvcc")
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    includes = ["-I" + HERE, "-I" + os.path.join(cuda, "include")]
    torch_includes = ["-I" + p for p in cpp_extension.include_paths()] + ["-I" + sysconfig.get_paths()["include"]]

    jobs = []
    objs = []
    for src in CU_SOURCES:
        sp = os.path.join(HERE, src)
        tag = _digest([sp] + headers, ARCH_FLAGS)
        obj = os.path.join(BUILD_DIR, src.replace(".cu", "") + "." + tag + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append([nvcc, *ARCH_FLAGS, "-lineinfo", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC",
                         *includes, "-c", sp, "-o", obj])
    cxx11 = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    for src in CPP_SOURCES:
        sp = os.path.join(HERE, src)
        tag = _digest([sp] + headers, [torch.__version__])
        obj = os.path.join(BUILD_DIR, src.replace(".cpp", "") + "." + tag + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", "-D_GLIBCXX_USE_CXX11_ABI=%d" % cxx11, *includes,
                         *torch_includes, "-c", sp, "-o", obj])

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(_run, jobs):
                if verbose and out.strip():
                    print(out)
    link_tag = os.path.join(BUILD_DIR, "linked." + _digest([], objs))
    if force or jobs or not os.path.exists(SO_PATH) or not os.path.exists(link_tag):
        cudart_dirs = [os.path.join(cuda, "lib64")]
        cmd = ["g++", "-shared", "-o", SO_PATH, *objs, "-L" + torch_lib, "-lc10", "-lc10_cuda", "-ltorch_cpu",
               "-ltorch_cuda", "-ltorch", *["-L" + d for d in cudart_dirs], "-lcudart",
               "-Wl,-rpath," + torch_lib, "-Wl,--no-as-needed"]
        _run(cmd)
        for f in os.listdir(BUILD_DIR):
            if f.startswith("linked."):
                os.remove(os.path.join(BUILD_DIR, f))
        open(link_tag, "w").close()
        # drop stale objects
        keep = {os.path.basename(o) for o in objs}
        for f in os.listdir(BUILD_DIR):
            if f.endswith(".o") and f not in keep:
                os.remove(os.path.join(BUILD_DIR, f))
    return SO_PATH


if __name__ == "__main__":
    path = build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print("built", path)
