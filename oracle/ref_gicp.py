"""Loader of oracle/_ref/fast_gicp/pygicp*.so — the REFERENCE's own tracker: fast_gicp's unmodified sources and its own
pybind11 module (src/python/main.cpp), built by oracle/Makefile against the reference's vendored Eigen and
oracle/pcl_shim (our stand-in for the PCL/boost slice fast_gicp needs; README there).  TEST INFRASTRUCTURE: parity
tests and bench.py's reference arm / cpu_baseline only.

The module is loaded under the name `ref_pygicp` so it never shadows this repo's `pygicp` drop-in."""
import glob
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref", "fast_gicp")
_mod = None


def path():
    c = sorted(glob.glob(os.path.join(_DIR, "pygicp*.so")))
    return c[0] if c else None


def available():
    return path() is not None


def load():
    """The reference pybind11 module (its PyInit symbol is `PyInit_pygicp`, so the loader is given that name but the
    module object is registered as `ref_pygicp`)."""
    global _mod
    if _mod is None:
        p = path()
        if p is None:
            raise ImportError(f"{_DIR}/pygicp*.so missing: run `make -C oracle ref` where /root/reference is present")
        loader = importlib.machinery.ExtensionFileLoader("pygicp", p)
        spec = importlib.util.spec_from_loader("pygicp", loader, origin=p)
        m = importlib.util.module_from_spec(spec)
        loader.exec_module(m)
        sys.modules["ref_pygicp"] = m
        _mod = m
    return _mod


def FastGICP():
    return load().FastGICP()
