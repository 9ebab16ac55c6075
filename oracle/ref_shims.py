"""Import shims that let the REAL cosmogrb hot-path code run from /root/reference.

TEST INFRASTRUCTURE ONLY (never imported by the product package).  This module exists
so that ``oracle/make_golden.py`` can execute the unmodified reference kernels in the
authoring container and mint golden vectors under ``tests/golden/``, and so that the CPU arm
of ``bench.py`` (``--impl reference``, ``oracle/ref_numba.py``) can TIME them.  It reads the
reference tree named by ``COSMOGRB_REFERENCE`` (default ``/root/reference``; on the GPU box
the git-ignored copy staged under ``oracle/_ref`` by ``oracle/stage_ref.py``) at run time;
no ``-m gpu`` test, ``smoke()`` or product code uses it.

What is shimmed:
  * ``numba_scipy``   -> registers ``scipy.special.gammaincc`` / ``gamma`` for nopython
                         mode through scipy's ``cython_special`` C entry points.
  * ``interpolation`` -> ``interp(x, y, v)``; two variants selectable with
                         ``install(interp_extrap="linear"|"clamp")``.  ``linear`` clamps the
                         bracket index but not the barycentric weight (EconForge
                         ``interpolation.py`` behaviour, not installable here), ``clamp`` is
                         ``np.interp``.
  * ``cosmogrb.utils.numba_array`` -> the reference file executed with the
                         ``__array__`` method dropped (numba >= 0.6x rejects it).
  * heavy ``__init__``s of the ``cosmogrb`` packages are bypassed with stub packages whose
    ``__path__`` points into the reference tree; plotting / HDF5 / FITS helpers are
    empty stubs.
"""
from __future__ import annotations

import ctypes
import importlib
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("COSMOGRB_REFERENCE", "/root/reference")

_installed = {"done": False, "interp_extrap": None}


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _pkg(name, relpath):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REFERENCE_ROOT, relpath)]
    sys.modules[name] = m
    return m


def _install_numba_scipy():
    import numba
    import scipy.special as sc
    from numba import types as nbt
    from numba.extending import get_cython_function_address, overload

    def fptr(name, nargs):
        addr = get_cython_function_address("scipy.special.cython_special", name)
        proto = ctypes.CFUNCTYPE(ctypes.c_double, *([ctypes.c_double] * nargs))
        return proto(addr)

    gammaincc_fn = fptr("gammaincc", 2)
    gamma_fn = fptr("__pyx_fuse_1gamma", 1)

    @overload(sc.gammaincc)
    def _ov_gammaincc(a, x):
        if isinstance(a, (nbt.Float, nbt.Integer)) and isinstance(x, (nbt.Float, nbt.Integer)):
            def impl(a, x):
                return gammaincc_fn(float(a), float(x))
            return impl

    @overload(sc.gamma)
    def _ov_gamma(a):
        if isinstance(a, (nbt.Float, nbt.Integer)):
            def impl(a):
                return gamma_fn(float(a))
            return impl

    _stub("numba_scipy", __version__="shim")
    return numba


def _install_interpolation(interp_extrap):
    import numba as nb
    import numpy as np

    if interp_extrap == "clamp":
        @nb.njit
        def interp(x, y, v):
            return np.interp(v, x, y)
    elif interp_extrap == "linear":
        @nb.njit
        def interp(x, y, v):
            n = x.shape[0]
            out = np.empty(v.shape[0])
            for k in range(v.shape[0]):
                i = int(np.searchsorted(x, v[k])) - 1
                i = min(max(i, 0), n - 2)
                lam = (v[k] - x[i]) / (x[i + 1] - x[i])
                out[k] = (1.0 - lam) * y[i] + lam * y[i + 1]
            return out
    else:
        raise ValueError(interp_extrap)
    _stub("interpolation", interp=interp)


def _install_numba_array():
    src_path = os.path.join(REFERENCE_ROOT, "cosmogrb", "utils", "numba_array.py")
    with open(src_path) as f:
        src = f.read()
    # drop the __array__ method (in memory only; nothing is written anywhere)
    src = re.sub(
        r"\n        def __array__\(self\):.*?\n(?=        def _expand)", "\n", src, flags=re.S
    )
    mod = types.ModuleType("cosmogrb.utils.numba_array")
    mod.__file__ = src_path
    sys.modules["cosmogrb.utils.numba_array"] = mod
    exec(compile(src, src_path, "exec"), mod.__dict__)


def install(interp_extrap="linear"):
    """Install the shims. Idempotent for one ``interp_extrap``; a second variant in the same
    process is refused because numba caches the compiled kernels."""
    if _installed["done"]:
        if _installed["interp_extrap"] != interp_extrap:
            raise RuntimeError("shims already installed with another interp_extrap; use a new process")
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    _install_numba_scipy()
    _install_interpolation(interp_extrap)

    _pkg("cosmogrb", "cosmogrb")
    for sub in ("sampler", "utils", "response", "lightcurve", "instruments", "io", "grb"):
        _pkg(f"cosmogrb.{sub}", os.path.join("cosmogrb", sub))
    _pkg("cosmogrb.instruments.gbm", os.path.join("cosmogrb", "instruments", "gbm"))

    _install_numba_array()

    _stub("cosmogrb.utils.response_file", RSP=None)
    _stub("cosmogrb.utils.tte_file", TTEFile=None)
    _stub("cosmogrb.utils.array_to_cmap", array_to_cmap=None)
    _stub("cosmogrb.utils.plotting", channel_plot=None, step_plot=None)
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            mpl = _stub("matplotlib")
            mpl.pyplot = _stub("matplotlib.pyplot")
    try:
        importlib.import_module("IPython.display")
    except Exception:
        ip = _stub("IPython")
        ip.display = _stub("IPython.display", display=print)
    try:
        importlib.import_module("h5py")
    except Exception:
        _stub("h5py")

    _installed["done"] = True
    _installed["interp_extrap"] = interp_extrap


def load():
    """Return a namespace with the reference's hot-path callables (real code)."""
    import numba as nb
    import numpy as np

    cf = importlib.import_module("cosmogrb.sampler.cpl_functions")
    ccf = importlib.import_module("cosmogrb.sampler.cpl_constant_functions")
    tf = importlib.import_module("cosmogrb.sampler.temporal_functions")
    rsp = importlib.import_module("cosmogrb.response.response")
    bkg = importlib.import_module("cosmogrb.sampler.background")
    src = importlib.import_module("cosmogrb.sampler.source")
    cplsrc = importlib.import_module("cosmogrb.sampler.cpl_source")
    ccplsrc = importlib.import_module("cosmogrb.sampler.constant_cpl")
    gbmlc = importlib.import_module("cosmogrb.instruments.gbm.gbm_lightcurve")

    @nb.njit
    def numba_seed(s):
        np.random.seed(s)

    ns = types.SimpleNamespace(
        cpl_functions=cf,
        cpl_constant_functions=ccf,
        temporal_functions=tf,
        response=rsp,
        background=bkg,
        source=src,
        cpl_source=cplsrc,
        constant_cpl=ccplsrc,
        gbm_lightcurve=gbmlc,
        numba_seed=numba_seed,
    )
    return ns
