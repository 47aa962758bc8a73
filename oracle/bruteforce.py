"""ctypes front end of oracle/bruteforce.cpp (TEST INFRASTRUCTURE ONLY): an unbanded full-matrix affine-gap DP and a
unit-cost edit distance, written independently of oracle/mfsc_oracle.cpp, used by tests/test_bruteforce.py to check the
oracle's DP engine (end cell, score of the path it reports, error counts) on windows of up to 2 kbp."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmfsc_bruteforce.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "bruteforce.cpp")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = ctypes.CDLL(_SO)
        L.bf_global.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        L.bf_extend.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.bf_score_ops.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p]
        L.bf_edit_distance.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        _lib = L
    return _lib


def global_score(a, b):
    return lib().bf_global(bytes(a), len(a), bytes(b), len(b))


def extend(a, b, breaklen=200):
    out = np.zeros(4, dtype=np.int64)
    lib().bf_extend(bytes(a), len(a), bytes(b), len(b), breaklen, out.ctypes.data)
    return dict(fi=int(out[0]), fj=int(out[1]), score=int(out[2]), reached_end=bool(out[3]))


def score_ops(a, b, ops):
    out = np.zeros(5, dtype=np.int64)
    lib().bf_score_ops(bytes(a), bytes(b), ops.encode(), out.ctypes.data)
    return dict(score=int(out[0]), errors=int(out[1]), a_used=int(out[2]), b_used=int(out[3]), cols=int(out[4]))


def edit_distance(a, b):
    return lib().bf_edit_distance(bytes(a), len(a), bytes(b), len(b))
