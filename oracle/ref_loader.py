"""Import the compiled pyDVS reference from ``oracle/_ref`` (test infrastructure only).

``load()`` returns ``(generate_spikes, generate_spikes_uncoded, decode_spikes)`` or ``None`` when the
reference has not been built (``oracle/build_ref.py``).  The modules are the reference's own code,
compiled unmodified; the only shim is ``numpy.float = float`` which the uncoded/decode modules read at
import time (``generate_spikes_uncoded.pyx:32``, ``decode_spikes.pyx:34``) and NumPy >= 1.24 removed.
"""
import importlib
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_ROOT = os.path.join(_HERE, "_ref")
_cache = None


def available():
    from . import build_ref
    return build_ref.is_built()


def load():
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        return None
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float  # shim for the reference's `DTYPE_FLOAT = np.float`
    # the compiled modules are named pydvs.<mod>; expose oracle/_ref as a path root for `pydvs` under
    # a private alias so that it cannot shadow (or be shadowed by) a user-visible `pydvs` package.
    saved = sys.modules.pop("pydvs", None)
    saved_sub = {k: sys.modules.pop(k) for k in list(sys.modules) if k.startswith("pydvs.")}
    sys.path.insert(0, _REF_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gs = importlib.import_module("pydvs.generate_spikes")
            gsu = importlib.import_module("pydvs.generate_spikes_uncoded")
            dec = importlib.import_module("pydvs.decode_spikes")
    finally:
        sys.path.remove(_REF_ROOT)
        for k in [k for k in sys.modules if k == "pydvs" or k.startswith("pydvs.")]:
            sys.modules["_pydvs_ref" + k[5:]] = sys.modules.pop(k)
        if saved is not None:
            sys.modules["pydvs"] = saved
        sys.modules.update(saved_sub)
    _cache = (gs, gsu, dec)
    return _cache


_dvs_op = None


def load_dvs_op():
    """``ref_dvs_op(src, diff, ref, thr, relax, up, down)``: the reference's own ``DVSOperator::operator()``
    (``pydvs/cpp/dvs_op.cpp``, compiled unchanged; ``oracle/build_ref.py``) in place on float32 ``[H, W]`` arrays,
    or ``None`` when it has not been built."""
    global _dvs_op
    if _dvs_op is not None:
        return _dvs_op
    from . import build_ref
    if not build_ref.dvs_op_is_built():
        return None
    import ctypes as C
    import numpy as np
    lib = C.CDLL(build_ref.DVS_OP_LIB)
    fp = C.POINTER(C.c_float)
    lib.ref_dvs_op_f32.restype = C.c_int
    lib.ref_dvs_op_f32.argtypes = [fp, fp, fp, fp, fp, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int]

    def ref_dvs_op(src, diff, ref, thr, relax, up, down):
        for a in (src, diff, ref, thr):
            assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == src.shape and a.ndim == 2
        ev = np.zeros_like(src)   # the operator advances a pointer into `ev` but never writes it
        rc = lib.ref_dvs_op_f32(*(a.ctypes.data_as(fp) for a in (src, diff, ref, thr, ev)), relax, up, down,
                                src.shape[0], src.shape[1])
        assert rc == 0
    _dvs_op = ref_dvs_op
    return _dvs_op
