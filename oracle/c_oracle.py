"""ctypes face of oracle/bwt_oracle.c  --  TEST INFRASTRUCTURE ONLY (see that file's header).

``COracle`` wraps one index; ``build_lib()`` compiles the C file with the Makefile beside it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libbwt_oracle.so")


def build_lib(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("bwt_oracle.c", "seeds_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class OrcCand(C.Structure):
    _fields_ = [("S", C.c_int64), ("D", C.c_int64), ("pos", C.c_int64), ("score", C.c_int64),
                ("status", C.c_int32), ("cigar_len", C.c_int32), ("cigar", C.c_char * 96)]


class OrcStats(C.Structure):
    _fields_ = [("searches", C.c_uint64), ("lf_steps", C.c_uint64), ("located", C.c_uint64),
                ("sw_calls", C.c_uint64), ("cells", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_lib())
        L.orc_index_new.restype = C.c_void_p
        L.orc_index_new.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p,
                                    C.c_uint64, C.c_int]
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_find_seeds.restype = C.c_int
        L.orc_find_seeds.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                     C.POINTER(C.c_int)]
        L.orc_sga.restype = C.c_int64
        L.orc_sga.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_seed_ends.restype = C.c_int64
        L.orc_seed_ends.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int64]
        L.orc_sw.restype = C.c_int
        L.orc_sw.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_char_p,
                             C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
        L.orc_align_batch.restype = C.c_int
        L.orc_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.POINTER(OrcStats)]
        L.orc_fs_scan_patterns.restype = C.c_int64
        L.orc_fs_scan_patterns.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_fs_scan_dp.restype = C.c_int64
        L.orc_fs_scan_dp.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_void_p]
        L.orc_sizeof_cand.restype = C.c_size_t
        assert L.orc_sizeof_cand() == C.sizeof(OrcCand)
        _lib = L
    return _lib


def c_fs_scan_patterns(patterns, length: int, ref: bytes) -> List[int]:
    """fs:find_seeds/1 with the literal clause list (oracle/seeds_oracle.c) -> descending 1-based starts"""
    blob = b"".join(patterns)
    cap = len(ref) // max(1, length) + 2
    out = np.zeros(cap, dtype=np.int64)
    n = lib().orc_fs_scan_patterns(blob, len(patterns), length, ref, len(ref), out.ctypes.data, cap)
    assert n >= 0
    return out[:n].tolist()


def c_fs_scan_dp(q: bytes, length: int, defects: int, ref: bytes, want_vec: bool = False):
    """fs:find_seeds/1 with matches/Len as a reachability DP over gen_pattern/5 -> descending 1-based starts"""
    cap = len(ref) // max(1, length) + 2
    out = np.zeros(cap, dtype=np.int64)
    vec = np.zeros(max(0, len(ref) - length + 1), dtype=np.uint8) if want_vec else None
    n = lib().orc_fs_scan_dp(q, len(q), length, defects, ref, len(ref), out.ctypes.data, cap,
                             vec.ctypes.data if want_vec else None)
    assert n >= 0, "query shorter than Len + Defects (or Defects > 8)"
    return (out[:n].tolist(), vec) if want_vec else out[:n].tolist()


def c_sw(q: bytes, r: bytes, force_double: bool = False):
    """-> 'no_match' | (score, cigar)"""
    score = C.c_int64()
    cl = C.c_int()
    cells = C.c_uint64()
    cap = 4 * (len(q) + len(r)) + 16
    buf = C.create_string_buffer(cap)
    st = lib().orc_sw(q, len(q), r, len(r), int(force_double), C.byref(score), buf, cap, C.byref(cl), C.byref(cells))
    if st != 0:
        return "no_match"
    return int(score.value), buf.raw[:cl.value].decode()


class COracle:
    """One index in the C oracle.  ``mode`` 0 = literal scan with the stored I column, 1 = closed form."""

    def __init__(self, L: np.ndarray, SA: np.ndarray, pointers, shift: int, ref: bytes, mode: int = 0):
        self._L = np.ascontiguousarray(L, dtype=np.uint8)
        self._SA = np.ascontiguousarray(SA, dtype=np.uint32)
        self._ptr = np.asarray(pointers, dtype=np.uint32)
        self._ref = np.frombuffer(ref, dtype=np.uint8) if not isinstance(ref, np.ndarray) else ref
        self.shift = int(shift)
        self.h = lib().orc_index_new(self._L.ctypes.data, self._SA.ctypes.data, len(self._L), self._ptr.ctypes.data,
                                     self.shift, self._ref.ctypes.data, len(self._ref), mode)

    def close(self):
        if self.h:
            lib().orc_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def find_seeds(self, sub: bytes) -> Tuple[int, int, int, int]:
        sp, ep, n = C.c_uint64(), C.c_uint64(), C.c_int()
        st = lib().orc_find_seeds(self.h, sub, len(sub), C.byref(sp), C.byref(ep), C.byref(n))
        return st, int(sp.value), int(ep.value), int(n.value)

    def sga(self, read: bytes) -> List[Tuple[int, int]]:
        cap = 4096
        S = np.zeros(cap, dtype=np.int64)
        D = np.zeros(cap, dtype=np.int64)
        c = lib().orc_sga(self.h, read, len(read), S.ctypes.data, D.ctypes.data, cap)
        if c < 0:
            raise RuntimeError("reference crash")
        return [(int(S[k]), int(D[k])) for k in range(c)]

    def seed_ends(self, read: bytes) -> np.ndarray:
        cap = 1 << 16
        out = np.zeros(cap, dtype=np.int64)
        c = lib().orc_seed_ends(self.h, read, len(read), out.ctypes.data, cap)
        if c < 0:
            raise RuntimeError("reference crash")
        assert c <= cap
        return out[:c].copy()

    def align_batch(self, reads: np.ndarray, offsets: np.ndarray, nthreads: int = 1, with_sw: bool = True,
                    cap_per_read: int = 64):
        """reads: uint8 concatenation, offsets: int64[n+1].  -> (ncand int64[n], cand OrcCand[n*cap], stats)"""
        n = len(offsets) - 1
        reads = np.ascontiguousarray(reads, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ncand = np.zeros(n, dtype=np.int64)
        cand = (OrcCand * (n * cap_per_read))()
        st = OrcStats()
        lib().orc_align_batch(self.h, reads.ctypes.data, offsets.ctypes.data, n, nthreads, int(with_sw), cap_per_read,
                              ncand.ctypes.data, C.addressof(cand), C.byref(st))
        return ncand, cand, st
