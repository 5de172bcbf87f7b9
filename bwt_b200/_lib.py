"""ctypes binding of libbwt_b200.so (include/bwt_b200.h).  No fallback: a missing library or a missing
GPU raises -- the product has no CPU path."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BWT_B200_LIB picks another build of the same library (A/B timing of two builds inside one gpurun call)
LIB_PATH = os.environ.get("BWT_B200_LIB") or os.path.join(_HERE, "libbwt_b200.so")


class BwtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("bwt_b200 error %d: %s" % (code, msg))
        self.code = code


class NoSpace(BwtError):
    """BWT_ERR_NOSPACE: caller capacity too small; ``needed`` holds what the call reported."""

    def __init__(self, msg, needed):
        super().__init__(-6, msg)
        self.needed = needed


class BwtHit(C.Structure):
    _fields_ = [("pos", C.c_int64), ("s", C.c_int64), ("cigar_off", C.c_int64), ("d", C.c_int32),
                ("score", C.c_int32), ("status", C.c_int32), ("cigar_len", C.c_int32)]


class BwtHit16(C.Structure):
    _fields_ = [("s", C.c_uint32), ("d_flags", C.c_uint32), ("cigar_off", C.c_uint32), ("score", C.c_int16), ("cigar_len", C.c_uint16)]


class BwtPackedReads(C.Structure):
    _fields_ = [("n", C.c_int64), ("fixed_len", C.c_int32), ("reserved", C.c_int32), ("lens", C.c_void_p), ("words", C.c_void_p),
                ("n_words", C.c_int64), ("n_exc", C.c_int64), ("exc_idx", C.c_void_p), ("exc_off", C.c_void_p),
                ("exc_ascii", C.c_void_p), ("word_index", C.c_void_p), ("maxlen_index", C.c_void_p)]


class BwtIndexInfo(C.Structure):
    _fields_ = [("n_rows", C.c_uint64), ("pointers", C.c_uint32 * 4), ("shift", C.c_int64), ("ref_len", C.c_uint64),
                ("dollar_row", C.c_uint64), ("device_bytes", C.c_uint64), ("device", C.c_int32), ("kmer_k", C.c_int32),
                ("fused", C.c_int32), ("reserved", C.c_int32)]


class BwtStats(C.Structure):
    _fields_ = [("reads", C.c_uint64), ("searches", C.c_uint64), ("lf_steps", C.c_uint64), ("hbm_lf_steps", C.c_uint64),
                ("located_rows", C.c_uint64), ("clusters", C.c_uint64), ("sw_calls", C.c_uint64),
                ("sw_cells", C.c_uint64), ("sw_general", C.c_uint64),
                ("ms_total", C.c_float), ("ms_seed", C.c_float), ("ms_sw", C.c_float), ("ms_emit", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_sga_kernel", C.c_float), ("ms_sw_kernel", C.c_float),
                ("launches", C.c_uint32), ("reserved", C.c_uint32),
                ("sw_shortcut", C.c_uint64), ("sw_shortcut_cells", C.c_uint64),
                ("sw_band", C.c_uint64), ("sw_band_cells", C.c_uint64), ("sw_band_filled", C.c_uint64)]


class BwtFsStats(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("nodes", C.c_uint64), ("lf_steps", C.c_uint64), ("raw_hits", C.c_uint64),
                ("seeds", C.c_uint64), ("spills", C.c_uint64), ("ms_kernel", C.c_float), ("launches", C.c_uint32),
                ("settled", C.c_uint64)]


class BwtSampleOpts(C.Structure):
    _fields_ = [("n", C.c_int64), ("read_len", C.c_int32), ("device", C.c_int32), ("seed", C.c_uint64),
                ("edit_rate", C.c_double), ("ins_share", C.c_double), ("del_share", C.c_double), ("n_rate", C.c_double),
                ("synth_len", C.c_uint64), ("synth_seed", C.c_uint64)]


# every symbol include/bwt_b200.h declares: (restype, argtypes)
P = C.c_void_p
SIGNATURES = {
    "bwt_last_error": (C.c_char_p, []),
    "bwt_abi_version": (C.c_int, []),
    "bwt_device_count": (C.c_int, []),
    "bwt_open": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(P)]),
    "bwt_open_arrays": (C.c_int, [P, C.c_uint64, P, P, C.c_int64, P, C.c_uint64, C.c_int, C.POINTER(P)]),
    "bwt_build": (C.c_int, [P, C.c_uint64, C.c_int64, C.c_int, C.POINTER(P)]),
    "bwt_build_n": (C.c_int, [P, C.c_uint64, C.c_int64, C.c_int, P, C.POINTER(P)]),
    "bwt_build_synthetic": (C.c_int, [C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.POINTER(P)]),
    "bwt_build_synthetic_part": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.POINTER(P)]),
    "bwt_sample_reads_edits": (C.c_int, [P, C.POINTER(BwtSampleOpts), P, C.c_int64, P, P, C.POINTER(C.c_int64)]),
    "bwt_save": (C.c_int, [P, C.c_char_p, C.c_char_p]),
    "bwt_export": (C.c_int, [P, P, P, P]),
    "bwt_info": (C.c_int, [P, C.POINTER(BwtIndexInfo)]),
    "bwt_fm_inspect": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint64), P, C.POINTER(C.c_int64)]),
    "bwt_set_kmer_depth": (C.c_int, [P, C.c_int]),
    "bwt_close": (C.c_int, [P]),
    "bwt_search_batch": (C.c_int, [P, P, P, C.c_int64, P, P, P, P]),
    "bwt_locate": (C.c_int, [P, C.c_uint32, C.c_uint32, P]),
    "bwt_sga_batch": (C.c_int, [P, P, P, C.c_int64, P, P, P, P, C.c_int64, C.POINTER(C.c_int64)]),
    "bwt_sw_batch": (C.c_int, [C.c_int, P, P, P, P, C.c_int64, P, P, P, P, P, C.c_int64, C.POINTER(C.c_int64)]),
    "bwt_align_batch": (C.c_int, [P, P, P, C.c_int64, P, P, P, C.c_int64, C.POINTER(C.c_int64), P, C.c_int64,
                                  C.POINTER(C.c_int64)]),
    "bwt_batch_create": (C.c_int, [P, C.c_int64, C.c_int64, C.POINTER(P)]),
    "bwt_batch_destroy": (C.c_int, [P]),
    "bwt_batch_upload": (C.c_int, [P, P, P, C.c_int64]),
    "bwt_batch_run": (C.c_int, [P]),
    "bwt_batch_result_size": (C.c_int, [P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bwt_batch_download": (C.c_int, [P, P, P, P, C.c_int64, P, C.c_int64]),
    "bwt_batch_stats": (C.c_int, [P, C.POINTER(BwtStats)]),
    "bwt_pack_words": (C.c_int64, [P, C.c_int64]),
    "bwt_pack_reads": (C.c_int, [P, P, C.c_int64, C.c_int, P, C.c_int64, P, P, C.c_int64, P, P, C.c_int64,
                                 C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "bwt_pack_index": (C.c_int, [P, C.c_int64, P, P]),
    "bwt_batch_upload_packed": (C.c_int, [P, C.POINTER(BwtPackedReads)]),
    "bwt_batch_run_compact": (C.c_int, [P]),
    "bwt_batch_download_compact": (C.c_int, [P, P, P, P, C.c_int64, P, C.c_int64]),
    "bwt_aligner_run_packed": (C.c_int, [P, C.POINTER(BwtPackedReads), P, P, P, C.c_int64, C.POINTER(C.c_int64), P, C.c_int64,
                                         C.POINTER(C.c_int64)]),
    "bwt_batch_device_result": (C.c_int, [P, C.POINTER(P), C.POINTER(P), C.POINTER(P)]),
    "bwt_aligner_create": (C.c_int, [P, C.c_int64, C.c_int64, C.c_int, C.POINTER(P)]),
    "bwt_aligner_run": (C.c_int, [P, P, P, C.c_int64, P, P, P, C.c_int64, C.POINTER(C.c_int64), P, C.c_int64,
                                  C.POINTER(C.c_int64)]),
    "bwt_aligner_stats": (C.c_int, [P, C.POINTER(BwtStats)]),
    "bwt_comm_unique_id": (C.c_int, [P]),
    "bwt_comm_create": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.POINTER(P)]),
    "bwt_comm_destroy": (C.c_int, [P]),
    "bwt_partition_gather": (C.c_int, [P, P, C.c_int, P, P, P, C.c_int64, P, C.c_int64, P, P, C.POINTER(C.c_float),
                                       C.POINTER(C.c_int64)]),
    "bwt_aligner_destroy": (C.c_int, [P]),
    "bwt_fs_find_seeds_batch": (C.c_int, [P, P, P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, P, P, P, C.c_int64,
                                          C.POINTER(C.c_int64), C.POINTER(BwtFsStats)]),
    "bwt_fastq_parse": (C.c_int, [P, C.c_int64, C.c_int, C.c_int64, P, C.c_int64, P, P, P, C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]),
    "bwt_search_resident": (C.c_int, [P, P, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "bwt_sample_reads": (C.c_int, [P, C.c_int64, C.c_int32, C.c_uint64, C.c_double, P, P]),
    "bwt_peak_gather_gbs": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(C.c_double)]),
    "bwt_peak_dpx_gops": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def lib():
    """Load libbwt_b200.so; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("bwt_b200: %s is missing -- build it with `make -C bwt_b200/csrc` "
                              "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here = the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.bwt_abi_version() != 5:
            raise ImportError("bwt_b200: ABI version mismatch")
        _lib = L
    return _lib


def check(code: int):
    if code != 0:
        msg = lib().bwt_last_error().decode("utf-8", "replace")
        raise BwtError(code, msg)
