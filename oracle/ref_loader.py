"""Load the compiled, unmodified reference `yahmm` extension from oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

The extension is built by oracle/build_ref.sh from /root/reference/yahmm/yahmm.pyx.  It is loaded
under the private module name it was compiled with (`yahmm`, because CPython looks up
`PyInit_yahmm`) but is *not* left in `sys.modules`, so it never shadows the product's drop-in
`yahmm` package.  While it is being imported, `networkx` and `matplotlib` resolve to the shims in
oracle/stubs/ (networkx-1.11 API with insertion-ordered dicts; see that file's docstring).
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")
_STUBS = os.path.join(_HERE, "stubs")
_cached = None


def reference_so_path():
    hits = sorted(glob.glob(os.path.join(_REF_DIR, "yahmm*.so")))
    return hits[0] if hits else None


def available():
    return reference_so_path() is not None


def load_reference_yahmm():
    """Return the compiled reference module (cached). Raises ImportError if oracle/_ref is not built."""
    global _cached
    if _cached is not None:
        return _cached
    so = reference_so_path()
    if so is None:
        raise ImportError("oracle/_ref/yahmm*.so not built; run oracle/build_ref.sh")
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k == "yahmm" or k.split(".")[0] in ("networkx", "matplotlib")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, _STUBS)
    try:
        loader = importlib.machinery.ExtensionFileLoader("yahmm", so)
        spec = importlib.util.spec_from_file_location("yahmm", so, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    finally:
        sys.path.remove(_STUBS)
        for k in [k for k in sys.modules if k == "yahmm" or k.split(".")[0] in ("networkx", "matplotlib")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
    _cached = mod
    return mod


_cached_cparsers = None


def cparsers_available():
    return bool(glob.glob(os.path.join(_REF_DIR, "cparsers*.so")))


def load_reference_cparsers():
    """The compiled, unmodified PyPore/cparsers.pyx (FastStatSplit).  While it is imported `core` resolves to the
    stub in oracle/stubs/pypore_core/ and itertools gets the Python-2 name `izip` (cparsers.pyx:16-17)."""
    global _cached_cparsers
    if _cached_cparsers is not None:
        return _cached_cparsers
    hits = sorted(glob.glob(os.path.join(_REF_DIR, "cparsers*.so")))
    if not hits:
        raise ImportError("oracle/_ref/cparsers*.so not built; run oracle/build_ref.sh")
    import itertools
    had_izip = hasattr(itertools, "izip")
    if not had_izip:
        itertools.izip = zip
    saved_core = sys.modules.pop("core", None)
    stub = os.path.join(_STUBS, "pypore_core")
    sys.path.insert(0, stub)
    try:
        loader = importlib.machinery.ExtensionFileLoader("cparsers", hits[0])
        spec = importlib.util.spec_from_file_location("cparsers", hits[0], loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    finally:
        sys.path.remove(stub)
        sys.modules.pop("core", None)
        if saved_core is not None:
            sys.modules["core"] = saved_core
        if not had_izip:
            del itertools.izip
    _cached_cparsers = mod
    return mod
