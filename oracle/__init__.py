"""oracle/ -- TEST INFRASTRUCTURE ONLY (the checker and the CPU baseline); nothing here ships in, is imported by, or
is reachable from the product package `imageworsener_b200`.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.

  load_reference()   oracle/_ref/libiwref.so: the UNMODIFIED reference core compiled by oracle/Makefile from
                     /root/reference/src, behind the same ctypes binding class the product uses
  load_port()        oracle/libiworacle.so: the plain-C restatement (iw_oracle.c) + the synthetic-input helpers
  synth_xorshift()   the survey's noise generator, without touching the product library
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_SO = os.path.join(HERE, "_ref", "libiwref.so")
PORT_SO = os.path.join(HERE, "libiworacle.so")

_ref = None
_port = None


def load_reference():
    """The unmodified reference build as a `Lib` (same binding class as the product: identical calls)."""
    global _ref
    if _ref is None:
        from imageworsener_b200.libimagew import Lib   # the binding class only; no product code is loaded
        _ref = Lib(REFERENCE_SO)
    return _ref


def load_port() -> C.CDLL:
    global _port
    if _port is None:
        _port = C.CDLL(PORT_SO, mode=os.RTLD_LOCAL)
        _port.iwo_synth_xorshift.restype = None
        _port.iwo_synth_xorshift.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        _port.iwo_fnv1a64.restype = C.c_uint64
        _port.iwo_fnv1a64.argtypes = [C.c_void_p, C.c_size_t]
    return _port


def synth_xorshift(nbytes: int, seed: int) -> np.ndarray:
    buf = np.empty(nbytes, np.uint8)
    load_port().iwo_synth_xorshift(buf.ctypes.data, buf.size, seed)
    return buf


def fnv1a64(arr: np.ndarray) -> int:
    arr = np.ascontiguousarray(arr)
    return int(load_port().iwo_fnv1a64(arr.ctypes.data, arr.nbytes))
