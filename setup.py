"""Packaging for internevo_b200.

    pip install -e .            # builds internevo_b200/_C.so in-tree with nvcc (sm_100a) and installs the package + the
                                # `internlm` import alias + the `huggingface` model code

The CUDA extension is NOT built through torch.utils.cpp_extension: the kernels are plain .cu translation units compiled
straight for `-gencode arch=compute_100a,code=sm_100a` by `internevo_b200/csrc/build.py` (seconds per file, no GPU needed),
and only `bindings.cpp` sees torch headers.  Set INTERNEVO_B200_SKIP_BUILD=1 to install the Python side only (CPU use).
"""
import os
import sys

from setuptools import find_packages, setup
from setuptools.command.build_py import build_py
from setuptools.command.develop import develop

HERE = os.path.dirname(os.path.abspath(__file__))


def _build_extension():
    if os.environ.get("INTERNEVO_B200_SKIP_BUILD") == "1":
        return
    sys.path.insert(0, HERE)
    from internevo_b200.csrc.build import build

    print("built", build(verbose=False))


class BuildPy(build_py):
    def run(self):
        _build_extension()
        super().run()


class Develop(develop):
    def run(self):
        _build_extension()
        super().run()


def _read(name):
    with open(os.path.join(HERE, name)) as f:
        return f.read()


def _requirements(name):
    return [ln.split("#")[0].strip() for ln in _read(os.path.join("requirements", name)).splitlines()
            if ln.split("#")[0].strip()]


setup(
    name="internevo_b200",
    version=_read("version.txt").strip(),
    description="Blackwell-native (B200, sm_100a) hybrid-parallel LLM training framework with the capability set of InternEvo",
    long_description=_read("README.md"),
    long_description_content_type="text/markdown",
    packages=find_packages(include=["internevo_b200", "internevo_b200.*", "internlm", "internlm.*", "huggingface",
                                    "huggingface.*"]),
    package_data={"internevo_b200": ["_C.so", "csrc/*.cu", "csrc/*.cuh", "csrc/*.h", "csrc/*.cpp"]},
    python_requires=">=3.10",
    install_requires=_requirements("runtime.txt"),
    extras_require={"all": _requirements("optional.txt")},
    cmdclass={"build_py": BuildPy, "develop": Develop},
    zip_safe=False,
)
