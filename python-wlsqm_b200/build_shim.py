#!/usr/bin/env python
"""Build the Cython C-ABI shim (wlsqm_b200/_shim.pyx -> wlsqm_b200/_shim.*.so) IN-TREE, next to libwlsqm_b200.so.

    python python-wlsqm_b200/build_shim.py [--force]

cythonize + gcc, linked against the CUDA library of the package (run-time search path $ORIGIN).  The library itself is
built first by python-wlsqm_b200/csrc/Makefile (__graft_entry__.build() runs both).
"""
import os
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
PKG = HERE / "wlsqm_b200"
SRC = PKG / "_shim.pyx"
INC = HERE.parent / "include"


def target() -> Path:
    return PKG / ("_shim" + sysconfig.get_config_var("EXT_SUFFIX"))


def up_to_date() -> bool:
    t = target()
    deps = [SRC, INC / "wlsqm_b200.h", PKG / "libwlsqm_b200.so"]
    return t.exists() and all(d.exists() and d.stat().st_mtime <= t.stat().st_mtime for d in deps[:2])


def build(force: bool = False) -> Path:
    if up_to_date() and not force:
        return target()
    if not (PKG / "libwlsqm_b200.so").exists():
        raise RuntimeError("build libwlsqm_b200.so first (make -C python-wlsqm_b200/csrc)")
    import numpy
    work = HERE / "csrc" / "_obj" / "shim"
    work.mkdir(parents=True, exist_ok=True)
    c_file = work / "_shim.c"
    subprocess.check_call([sys.executable, "-m", "cython", "-3", "--fast-fail", "-o", str(c_file), str(SRC)], cwd=str(HERE))
    cc = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"
    py_inc = sysconfig.get_paths()["include"]
    cmd = [cc, "-O2", "-fPIC", "-shared", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I", py_inc, "-I", numpy.get_include(), "-I", str(INC), str(c_file), "-o", str(target()),
           "-L", str(PKG), "-l:libwlsqm_b200.so", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return target()


if __name__ == "__main__":
    print("shim built:", build(force="--force" in sys.argv))
