"""Build / load the native dataset index builders (``relora_b200/_data_helpers.so``, pybind11).

Host-only C++ (reference: ``megatron_dataset/Makefile:1-8`` builds ``helpers.cpp`` with g++ -O3).
"""
from __future__ import annotations

import hashlib
import importlib
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(PKG, "csrc", "data_helpers.cpp")
TARGET = os.path.join(PKG, "_data_helpers.so")
STAMP = os.path.join(PKG, "csrc", "_build", "data_helpers.stamp")


def build(force: bool = False) -> str:
    import pybind11

    digest = hashlib.sha256(open(SRC, "rb").read() + sys.version.encode()).hexdigest()
    if not force and os.path.exists(TARGET) and os.path.exists(STAMP) and open(STAMP).read() == digest:
        return TARGET
    os.makedirs(os.path.dirname(STAMP), exist_ok=True)
    cxx = os.environ.get("CXX", "g++")
    cmd = [cxx, "-O3", "-Wall", "-shared", "-std=c++17", "-fPIC", "-fvisibility=hidden",
           f"-I{pybind11.get_include()}", f"-I{sysconfig.get_paths()['include']}", SRC, "-o", TARGET + ".tmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("failed to build data helpers:\n" + res.stderr)
    os.replace(TARGET + ".tmp", TARGET)
    open(STAMP, "w").write(digest)
    return TARGET


def load():
    """Import the helpers, compiling them on first use (single process; callers barrier afterwards)."""
    try:
        return importlib.import_module("relora_b200._data_helpers")
    except ImportError:
        build()
        importlib.invalidate_caches()
        return importlib.import_module("relora_b200._data_helpers")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
