"""Build the reference's OWN native helper extensions (unmodified sources from its pinned flash-attention submodule:
csrc/rotary, csrc/fused_dense_lib, csrc/xentropy) for sm_100 into baseline/_ref, so that `bench.py --impl reference`
can run the reference's stock `use_flash_attn=True` code path on a B200.  Nothing from internevo_b200 is involved.

    python baseline/build_ref_ext.py
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

import torch
from torch.utils import cpp_extension

SRC = "/root/reference/third_party/flash-attention/csrc"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
EXTS = {
    "rotary_emb": ("rotary", ["rotary.cpp", "rotary_cuda.cu"]),
    "fused_dense_lib": ("fused_dense_lib", ["fused_dense.cpp", "fused_dense_cuda.cu"]),
    "xentropy_cuda_lib": ("xentropy", ["interface.cpp", "xentropy_kernel.cu"]),
}


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(" ".join(cmd), "\n", r.stdout[-3000:], r.stderr[-3000:])
        raise SystemExit(1)


def main():
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I" + p for p in cpp_extension.include_paths()] + ["-I" + sysconfig.get_paths()["include"],
                                                                "-I/usr/local/cuda/include"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    for name, (sub, files) in EXTS.items():
        so = os.path.join(OUT, name + ".so")
        if os.path.exists(so):
            print("have", so)
            continue
        with tempfile.TemporaryDirectory() as tmp:
            objs = []
            for f in files:
                src = os.path.join(tmp, f)
                shutil.copy(os.path.join(SRC, sub, f), src)
                obj = src + ".o"
                common = [f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
                          f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc]
                if f.endswith(".cu"):
                    run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100,code=sm_100", "--expt-relaxed-constexpr",
                         "--expt-extended-lambda", "-Xcompiler", "-fPIC", *common, "-c", src, "-o", obj])
                else:
                    run(["g++", "-O3", "-std=c++17", "-fPIC", *common, "-c", src, "-o", obj])
                objs.append(obj)
            run(["g++", "-shared", "-o", so, *objs, "-L" + torch_lib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda",
                 "-ltorch", "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart", "-lcublas", "-lcublasLt",
                 "-Wl,-rpath," + torch_lib])
        print("built", so)


if __name__ == "__main__":
    main()
    sys.exit(0)
