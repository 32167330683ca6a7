"""ctypes access to oracle/liblcr_oracle.so — test infrastructure only."""
import ctypes as C
import os

import numpy as np

from longcallr_b200 import abi
from longcallr_b200.host import ResultView

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB = os.path.join(ROOT, "oracle", "liblcr_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(ORACLE_LIB)
        L.lcr_oracle_run.argtypes = [C.POINTER(abi.Params), C.POINTER(abi.Batch), C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.POINTER(abi.Result))]
        L.lcr_oracle_free.argtypes = [C.POINTER(abi.Result)]
        L.lcr_oracle_free.restype = None
        for f in ("lcr_oracle_log10", "lcr_oracle_exp10", "lcr_oracle_log"):
            getattr(L, f).argtypes = [C.c_double]
            getattr(L, f).restype = C.c_double
        L.lcr_oracle_sor.argtypes = [C.c_int] * 4
        L.lcr_oracle_sor.restype = C.c_float
        L.lcr_oracle_binom.argtypes = [C.c_uint32, C.c_uint32]
        L.lcr_oracle_uniform.argtypes = [C.c_uint64, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.lcr_oracle_uniform.restype = C.c_double
        L.lcr_oracle_f64_as_i32.argtypes = [C.c_double]
        _lib = L
    return _lib


def run(params, batch, ref_seqs, mode=0, threads=1, raw=False):
    """Run the oracle over a BatchView.  ref_seqs: list indexed by tid of uint8 arrays (or None)."""
    n = len(ref_seqs)
    keep = [np.ascontiguousarray(s, dtype=np.uint8) if s is not None else None for s in ref_seqs]
    ptrs = (C.c_void_p * n)(*[k.ctypes.data if k is not None else None for k in keep])
    lens = np.array([k.size if k is not None else 0 for k in keep], dtype="<u8")
    out = C.POINTER(abi.Result)()
    rc = lib().lcr_oracle_run(C.byref(params), C.byref(batch.c), ptrs, lens.ctypes.data, n, mode, threads, C.byref(out))
    if rc:
        raise RuntimeError(f"oracle status {rc}")
    if raw:
        return out
    try:
        return ResultView(out)
    finally:
        lib().lcr_oracle_free(out)
