"""ctypes front end of the CPU oracle (oracle/vdl_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED (the reference ships no executor) -- see the C file's header.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvdl_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vdl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.orc_env_new.restype = C.c_void_p
        L.orc_env_free.argtypes = [C.c_void_p]
        L.orc_bind_column.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_int64]
        L.orc_run.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_seconds.argtypes = [C.c_void_p]
        L.orc_last_seconds.restype = C.c_double
        L.orc_num_statements.argtypes = [C.c_void_p]
        L.orc_num_outputs.argtypes = [C.c_void_p]
        L.orc_output_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_output_name.restype = C.c_char_p
        L.orc_output_len.argtypes = [C.c_void_p, C.c_int]
        L.orc_output_len.restype = C.c_int64
        L.orc_output_data.argtypes = [C.c_void_p, C.c_int]
        L.orc_output_data.restype = C.POINTER(C.c_int64)
        L.orc_gen_column.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64, C.c_int,
                                     C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int]
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


class OracleError(RuntimeError):
    pass


class Oracle:
    """Dense-vector interpreter of a VDL plan over host columns (numpy int32/int64 arrays)."""

    def __init__(self):
        self._env = lib().orc_env_new()
        self._keep = {}
        self.seconds = 0.0

    def __del__(self):
        if getattr(self, "_env", None):
            lib().orc_env_free(self._env)
            self._env = None

    def bind(self, name: str, arr: np.ndarray):
        if arr.dtype not in (np.int32, np.int64) and not (arr.dtype == np.uint8 and name.endswith(".heap")):
            raise TypeError(f"{name}: columns are int32 or int64 (Types.hs:84-87) -- or uint8 for a string heap --, got {arr.dtype}")
        arr = np.ascontiguousarray(arr)
        self._keep[name] = arr
        if lib().orc_bind_column(self._env, name.encode(), arr.ctypes.data, arr.dtype.itemsize, arr.shape[0]):
            raise OracleError(lib().orc_last_error(self._env).decode())

    def run(self, plan_text: str, threads: int = 0) -> dict:
        """Execute; returns {output name: int64 array} in MaterializeCompact order."""
        L = lib()
        if L.orc_run(self._env, plan_text.encode(), threads):
            raise OracleError(L.orc_last_error(self._env).decode())
        self.seconds = L.orc_last_seconds(self._env)
        out = {}
        for i in range(L.orc_num_outputs(self._env)):
            n = L.orc_output_len(self._env, i)
            p = L.orc_output_data(self._env, i)
            out[L.orc_output_name(self._env, i).decode()] = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.int64)
        return out

    @property
    def num_statements(self) -> int:
        return lib().orc_num_statements(self._env)


def gen_column(spec, rows: int, row_offset: int, seed: int, threads: int = 0) -> np.ndarray:
    """Host-side instance of the synthetic recipe (mplan2vdl_b200/synth.py ColumnSpec)."""
    arr = np.empty(rows, dtype=np.int32 if spec.width == 4 else np.int64)
    lib().orc_gen_column(arr.ctypes.data, spec.width, rows, row_offset, seed & (2**64 - 1), spec.stream, spec.kind,
                         spec.vmin, spec.stride, spec.p0, spec.p1, threads)
    return arr


def max_threads() -> int:
    return lib().orc_max_threads()
