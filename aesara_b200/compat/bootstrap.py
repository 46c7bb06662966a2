"""Make the reference front-end (``aesara``) importable next to this backend.

The B200 linker plugs in *behind* Aesara's own graph builder and rewriter
(everything above L3 is reused unchanged), so the host side
needs ``import aesara`` to work.  The reference tree as shipped does not
import under Python 3.12 / NumPy 2.  This module
applies the smallest possible overlay, entirely from outside the (read-only)
reference tree:

1. a generated-version stub (``aesara/version.py:1-9`` wants ``aesara._version``),
2. NumPy-2 aliases for the removed NumPy-1 names the reference still calls,
3. in-repo stand-ins for the un-vendored pure-Python deps ``cons``,
   ``etuples`` and ``unification`` (``aesara/graph/rewriting/unify.py:17-23``),
4. ``AESARA_FLAGS`` so the reference C linker compiles with g++ 13 / NumPy 2
   headers (used only as the *oracle* / CPU baseline, never by the product path),
5. the Cython Scan extension disabled (its shipped C does not build on 3.12;
   ``aesara/scan/op.py:1648`` falls back to the Python loop, same numerics).

Nothing here is needed (or used) at run time on a machine where only the
lowered programs are executed: ``aesara_b200.runtime`` has no Aesara import.
"""

import os
import sys
import types

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_CXXFLAGS = (
    "-DNPY_PY3K=1 -DPyInt_AsLong=PyLong_AsLong -DPyInt_FromLong=PyLong_FromLong "
    "-DPyInt_Check=PyLong_Check -DNPY_TARGET_VERSION=NPY_2_0_API_VERSION "
    "-DPyArray_MoveInto=PyArray_CopyInto -Wno-deprecated-declarations"
)


def reference_path():
    """Directory named by ``$AESARA_B200_REFERENCE`` that contains an ``aesara`` package
    which is not installed (a source checkout), or None: then ``import aesara`` must work
    on its own, as for any linker plugin."""
    cand = os.environ.get("AESARA_B200_REFERENCE")
    if cand and os.path.isdir(os.path.join(cand, "aesara")):
        return cand
    return None


def available():
    if "aesara" in sys.modules and sys.modules["aesara"] is not None:
        return True
    if reference_path() is not None:
        return True
    import importlib.util

    try:
        return importlib.util.find_spec("aesara") is not None
    except (ImportError, ValueError):
        return False


def _numpy2_aliases():
    import numpy as np

    if not hasattr(np, "obj2sctype"):  # aesara/tensor/type.py:104
        np.obj2sctype = lambda rep, default=None: np.dtype(rep).type
    if not hasattr(np, "sctype2char"):  # aesara/tensor/elemwise.py:708
        np.sctype2char = lambda t: np.dtype(t).char
    if not hasattr(np, "AxisError"):  # aesara/tensor/basic.py:223
        np.AxisError = np.exceptions.AxisError
    if not hasattr(np, "cast"):  # aesara/scalar/basic.py:3136 …

        class _Cast(dict):
            def __missing__(self, key):
                dt = np.dtype(key)
                fn = lambda x, _dt=dt: np.asarray(x).astype(_dt)  # noqa: E731
                self[key] = fn
                return fn

        np.cast = _Cast()
    if not hasattr(np, "MAXDIMS"):  # aesara/tensor/special.py (Softmax axis=None), = NPY_MAXDIMS
        np.MAXDIMS = 64
    for name, val in (
        ("bool8", np.bool_),
        ("float_", np.float64),
        ("complex_", np.complex128),
        ("product", np.prod),
        ("alltrue", np.all),
        ("sometrue", np.any),
        ("Inf", np.inf),
        ("NaN", np.nan),
        ("infty", np.inf),
        ("unicode_", np.str_),
    ):
        if not hasattr(np, name):
            setattr(np, name, val)


def _default_flags(compiledir):
    flags = {
        "blas__ldflags": "",
        "base_compiledir": compiledir,
        "gcc__cxxflags": _CXXFLAGS,
    }
    user = os.environ.get("AESARA_FLAGS", "")
    user_keys = {kv.split("=", 1)[0].strip() for kv in user.split(",") if "=" in kv}
    parts = [f"{k}={v}" for k, v in flags.items() if k not in user_keys]
    if user:
        parts.append(user)
    return ",".join(parts)


def load_aesara(compiledir=None):
    """Import and return the ``aesara`` module with the overlay applied."""
    mod = sys.modules.get("aesara")
    if mod is not None:
        return mod
    import warnings

    _numpy2_aliases()
    if _SHIMS not in sys.path:
        # appended: installed cons / etuples / unification packages win over the stand-ins
        sys.path.append(_SHIMS)
    ref = reference_path()
    if ref is not None and ref not in sys.path:
        sys.path.append(ref)
    if "aesara._version" not in sys.modules:
        stub = types.ModuleType("aesara._version")
        stub.__version__ = "2.9.3"
        stub.version = "2.9.3"
        sys.modules["aesara._version"] = stub
    # Python Scan loop instead of the un-buildable Cython one
    sys.modules.setdefault("aesara.scan.scan_perform_ext", None)
    if compiledir is None:
        compiledir = os.environ.get(
            "AESARA_B200_COMPILEDIR",
            os.path.join(os.path.expanduser("~"), ".cache", "aesara_b200", "ref_compiledir"),
        )
    os.makedirs(compiledir, exist_ok=True)
    os.environ["AESARA_FLAGS"] = _default_flags(compiledir)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import aesara  # noqa: F401
        import aesara.tensor  # noqa: F401
        import aesara.scan  # noqa: F401
    return sys.modules["aesara"]
