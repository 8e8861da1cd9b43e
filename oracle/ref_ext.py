"""Loader of the REFERENCE's own PyTorch extensions installed by oracle/build_ref_ext.sh into oracle/_ref/site/
(`diff_gaussian_rasterization` with its `_C`, `simple_knn._C`): unmodified sources built through their own setup.py — the
stock path (rasterize_points.cu + ext.cpp, torch's caching allocator, no extra synchronisation).  TEST INFRASTRUCTURE:
the rasterizer parity tests, the per-kernel "kernel to beat" profiles and bench.py's reference arm only.

The packages are loaded under the aliases `ref_diff_gaussian_rasterization` / `ref_simple_knn` so they never shadow this
repo's drop-in packages of the same names (which sit first on sys.path)."""
import glob
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SITE = os.path.join(_HERE, "_ref", "site")


def available():
    return (os.path.exists(os.path.join(SITE, "diff_gaussian_rasterization", "__init__.py"))
            and bool(glob.glob(os.path.join(SITE, "diff_gaussian_rasterization", "_C*.so"))))


def _load_package(alias, name):
    if alias in sys.modules:
        return sys.modules[alias]
    import torch  # noqa: F401  (the extensions link against libtorch; it must be loaded first)

    pkg_dir = os.path.join(SITE, name)
    init = os.path.join(pkg_dir, "__init__.py")
    if os.path.exists(init):
        spec = importlib.util.spec_from_file_location(alias, init, submodule_search_locations=[pkg_dir])
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        try:
            spec.loader.exec_module(mod)
        except Exception:
            del sys.modules[alias]
            raise
        return mod
    # namespace-style package holding only the extension (simple_knn ships no __init__.py)
    import types

    mod = types.ModuleType(alias)
    mod.__path__ = [pkg_dir]
    sys.modules[alias] = mod
    return mod


def diff_gaussian_rasterization():
    """The reference's diff_gaussian_rasterization package (GaussianRasterizationSettings, GaussianRasterizer, _C)."""
    if not available():
        raise ImportError(f"{SITE} missing: run oracle/build_ref_ext.sh where /root/reference is present")
    return _load_package("ref_diff_gaussian_rasterization", "diff_gaussian_rasterization")


def simple_knn_C():
    """The reference's simple_knn._C extension module (distCUDA2)."""
    pkg = _load_package("ref_simple_knn", "simple_knn")
    name = "ref_simple_knn._C"
    if name in sys.modules:
        return sys.modules[name]
    so = glob.glob(os.path.join(SITE, "simple_knn", "_C*.so"))
    if not so:
        raise ImportError(f"{SITE}/simple_knn/_C*.so missing: run oracle/build_ref_ext.sh")
    spec = importlib.util.spec_from_file_location(name, so[0])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    pkg._C = mod
    return mod
