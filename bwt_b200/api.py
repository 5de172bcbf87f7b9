"""numpy-level face of the C ABI (include/bwt_b200.h): what tests, bench.py and the Erlang-shaped
mirror modules (fm_index, bwa, sga, sw, worker_bwt) call.  Every method is a thin ctypes call; all
compute happens in libbwt_b200.so on the GPU.  No fallback of any kind.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from ._lib import (BwtError, BwtFsStats, BwtHit, BwtHit16, BwtIndexInfo, BwtPackedReads, BwtSampleOpts, BwtStats, NoSpace,
                   check, lib)

HIT_DTYPE = np.dtype([("pos", "<i8"), ("s", "<i8"), ("cigar_off", "<i8"), ("d", "<i4"), ("score", "<i4"),
                      ("status", "<i4"), ("cigar_len", "<i4")])
assert HIT_DTYPE.itemsize == C.sizeof(BwtHit) == 40

HIT16_DTYPE = np.dtype([("s", "<u4"), ("d_flags", "<u4"), ("cigar_off", "<u4"), ("score", "<i2"), ("cigar_len", "<u2")])
assert HIT16_DTYPE.itemsize == C.sizeof(BwtHit16) == 16
HIT16_FULL_M = 0x40000000

AS183_SEED0 = (3172, 9814, 20125)      # random:seed0/0 of OTP's stdlib: the state of a process that never seeded

SEEDS, NO_SEEDS, TOO_MANY_SEEDS, SEARCH_CRASH = 0, 1, 2, 3
READ_OK, READ_CRASH, READ_LIMIT = 0, 1, 2
HIT_MATCH, HIT_NO_MATCH, HIT_WINDOW_OOR = 0, 1, 2
FS_OK, FS_SHORT, FS_LIMIT = 0, 1, 2


def _u8(a) -> np.ndarray:
    if isinstance(a, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(a), dtype=np.uint8)
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def pack(seqs: Sequence[bytes]) -> Tuple[np.ndarray, np.ndarray]:
    """list of byte strings -> (concatenated uint8, int64 offsets[n+1])"""
    offs = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    data = np.frombuffer(b"".join(seqs), dtype=np.uint8) if seqs else np.zeros(0, np.uint8)
    return np.ascontiguousarray(data), offs


def device_count() -> int:
    return lib().bwt_device_count()


def fm_inspect(fm_path: str):
    """(n_rows, (Pc,Pg,Pt,Last), shift) of a <Chrom>_pJ.fm, read on the host (no GPU)"""
    n, sh = C.c_uint64(), C.c_int64()
    ptr = (C.c_uint32 * 4)()
    check(lib().bwt_fm_inspect(fm_path.encode(), C.byref(n), ptr, C.byref(sh)))
    return int(n.value), tuple(int(x) for x in ptr), int(sh.value)


class Index:
    """One FM-index partition + its padded .ref resident in HBM (opaque ``bwt_index``)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)

    # -- constructors ----------------------------------------------------------------------------
    @classmethod
    def open(cls, fm_path: str, ref_path: str, device: int = 0) -> "Index":
        h = C.c_void_p()
        check(lib().bwt_open(fm_path.encode(), ref_path.encode(), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_arrays(cls, L, SA, pointers, shift: int, ref, device: int = 0) -> "Index":
        L = _u8(L)
        SA = np.ascontiguousarray(SA, dtype=np.uint32)
        ptr = np.ascontiguousarray(pointers, dtype=np.uint32)
        ref = _u8(ref)
        h = C.c_void_p()
        check(lib().bwt_open_arrays(L.ctypes.data, len(L), SA.ctypes.data, ptr.ctypes.data, int(shift), ref.ctypes.data,
                                    len(ref), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build(cls, chunk, shift: int = 1, device: int = 0, n_state: Optional[np.ndarray] = None) -> "Index":
        """fm_index:make_index/4 on the device.  n_state: uint32[3] AS183 state (``AS183_SEED0`` for a fresh Erlang process),
        updated in place: every 'N' of the chunk is replaced in the indexed text by random:uniform(4)'s pick
        (fm_index.erl:131); None: a chunk with 'N' is rejected."""
        chunk = _u8(chunk)
        h = C.c_void_p()
        if n_state is not None:
            assert n_state.dtype == np.uint32 and len(n_state) == 3
            check(lib().bwt_build_n(chunk.ctypes.data, len(chunk), int(shift), device, n_state.ctypes.data, C.byref(h)))
        else:
            check(lib().bwt_build(chunk.ctypes.data, len(chunk), int(shift), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_synthetic(cls, length: int, seed: int = 42, shift: int = 1, device: int = 0) -> "Index":
        h = C.c_void_p()
        check(lib().bwt_build_synthetic(int(length), int(seed), int(shift), device, C.byref(h)))
        return cls(h.value)

    @classmethod
    def build_synthetic_part(cls, start: int, length: int, seed: int = 42, shift: int = 1, device: int = 0) -> "Index":
        """bases [start, start+length) of the synthetic reference ``seed``: one ``_pJ`` partition (fm_index.erl:113-124)"""
        h = C.c_void_p()
        check(lib().bwt_build_synthetic_part(int(start), int(length), int(seed), int(shift), device, C.byref(h)))
        return cls(h.value)

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if self._h:
            lib().bwt_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def info(self) -> BwtIndexInfo:
        i = BwtIndexInfo()
        check(lib().bwt_info(self._h, C.byref(i)))
        return i

    def set_kmer_depth(self, k: int):
        """0/8 = the 4^8 table only; 9..15 = deeper interval table (results never depend on it)."""
        check(lib().bwt_set_kmer_depth(self._h, int(k)))

    def export(self, want_L=True, want_SA=True, want_ref=True):
        i = self.info()
        L = np.empty(i.n_rows, np.uint8) if want_L else None
        SA = np.empty(i.n_rows, np.uint32) if want_SA else None
        ref = np.empty(i.ref_len, np.uint8) if want_ref else None
        check(lib().bwt_export(self._h, _ptr(L), _ptr(SA), _ptr(ref)))
        return L, SA, ref

    def save(self, fm_path: str, ref_path: str):
        check(lib().bwt_save(self._h, fm_path.encode(), ref_path.encode()))

    # -- bwa:find_seeds/6 ---------------------------------------------------------------------------
    def search(self, subs, offsets=None):
        """-> (status int32[n], Sp uint32[n], Ep uint32[n], N int32[n])"""
        if offsets is None:
            subs, offsets = pack(subs)
        subs = _u8(subs)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        st = np.zeros(n, np.int32)
        sp = np.zeros(n, np.uint32)
        ep = np.zeros(n, np.uint32)
        nm = np.zeros(n, np.int32)
        check(lib().bwt_search_batch(self._h, subs.ctypes.data, offsets.ctypes.data, n, st.ctypes.data, sp.ctypes.data,
                                     ep.ctypes.data, nm.ctypes.data))
        return st, sp, ep, nm

    def locate(self, first_row: int, count: int) -> np.ndarray:
        out = np.zeros(count, np.uint32)
        if count:
            check(lib().bwt_locate(self._h, int(first_row), int(count), out.ctypes.data))
        return out

    # -- sga:sga/6 ------------------------------------------------------------------------------------
    def sga(self, reads, offsets=None):
        """-> (read_status int32[n], cl_offsets int64[n+1], S int64[total], D int32[total]); reference order."""
        if offsets is None:
            reads, offsets = pack(reads)
        reads = _u8(reads)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        st = np.zeros(n, np.int32)
        clo = np.zeros(n + 1, np.int64)
        cap = max(16, 4 * n)
        total = C.c_int64()
        while True:
            S = np.zeros(cap, np.int64)
            D = np.zeros(cap, np.int32)
            rc = lib().bwt_sga_batch(self._h, reads.ctypes.data, offsets.ctypes.data, n, st.ctypes.data, clo.ctypes.data,
                                     S.ctypes.data, D.ctypes.data, cap, C.byref(total))
            if rc == -6:
                cap = int(total.value) + 16
                continue
            check(rc)
            return st, clo, S[:total.value].copy(), D[:total.value].copy()

    # -- worker_bwt inner loop ------------------------------------------------------------------------
    def align(self, reads, offsets=None):
        """-> (read_status int32[n], hit_offsets int64[n+1], hits HIT_DTYPE[total], cigars bytes)"""
        if offsets is None:
            reads, offsets = pack(reads)
        reads = _u8(reads)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        st = np.zeros(n, np.int32)
        ho = np.zeros(n + 1, np.int64)
        hcap, ccap = max(16, 4 * n), max(256, 64 * n)
        nh, nc = C.c_int64(), C.c_int64()
        while True:
            hits = np.zeros(hcap, HIT_DTYPE)
            cig = np.zeros(ccap, np.uint8)
            rc = lib().bwt_align_batch(self._h, reads.ctypes.data, offsets.ctypes.data, n, st.ctypes.data, ho.ctypes.data,
                                       hits.ctypes.data, hcap, C.byref(nh), cig.ctypes.data, ccap, C.byref(nc))
            if rc == -6:
                hcap, ccap = int(nh.value) + 16, int(nc.value) + 16
                continue
            check(rc)
            return st, ho, hits[:nh.value].copy(), cig[:nc.value].tobytes()

    # -- seeds:generate_fs/3 + fs:find_seeds/1 (seeds.erl:11-62) ------------------------------------------
    def fs_find_seeds(self, queries, offsets=None, length: int = 20, defects: int = 2, max_raw_hits: int = 1024):
        """-> (status int32[n], seed_offsets int64[n+1], seeds int64[total] (1-based window starts, descending per
        query, as fs:find_seeds/1 returns them), BwtFsStats)"""
        if offsets is None:
            queries, offsets = pack(queries)
        queries = _u8(queries)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        st = np.zeros(n, np.int32)
        so = np.zeros(n + 1, np.int64)
        cap = max(16, 2 * n)
        total = C.c_int64()
        stats = BwtFsStats()
        while True:
            seeds = np.zeros(cap, np.int64)
            rc = lib().bwt_fs_find_seeds_batch(self._h, queries.ctypes.data, offsets.ctypes.data, n, int(length),
                                               int(defects), int(max_raw_hits), st.ctypes.data, so.ctypes.data,
                                               seeds.ctypes.data, cap, C.byref(total), C.byref(stats))
            if rc == -6:
                cap = int(total.value) + 16
                continue
            check(rc)
            return st, so, seeds[:total.value].copy(), stats


def sw_batch(q, q_offsets, r, r_offsets, device: int = 0):
    """sw:sw/2 for n pairs -> (status int32[n], score int32[n], [cigar bytes or None per pair])"""
    q = _u8(q)
    r = _u8(r)
    qo = np.ascontiguousarray(q_offsets, dtype=np.int64)
    ro = np.ascontiguousarray(r_offsets, dtype=np.int64)
    n = len(qo) - 1
    st = np.zeros(n, np.int32)
    sc = np.zeros(n, np.int32)
    co = np.zeros(n, np.int64)
    cl = np.zeros(n, np.int32)
    cap = max(256, 32 * n)
    tot = C.c_int64()
    while True:
        cig = np.zeros(cap, np.uint8)
        rc = lib().bwt_sw_batch(device, _ptr(q) if len(q) else None, qo.ctypes.data, _ptr(r) if len(r) else None,
                                ro.ctypes.data, n, st.ctypes.data, sc.ctypes.data, co.ctypes.data, cl.ctypes.data,
                                cig.ctypes.data, cap, C.byref(tot))
        if rc == -6:
            cap = int(tot.value) + 16
            continue
        check(rc)
        raw = cig.tobytes()
        cigars = [raw[co[k]:co[k] + cl[k]] if st[k] == HIT_MATCH else None for k in range(n)]
        return st, sc, cigars


class PackedReads:
    """An ASCII batch in the 2-bit packed layout of ``bwt_packed_reads`` (``bwt_pack_reads``); keeps its arrays alive.
    ``pinned=True`` puts them into page-locked memory (torch) so that uploads run at full PCIe speed."""

    def __init__(self, reads, offsets=None, threads: int = 8, pinned: bool = False, fixed_len: int = 0):
        if offsets is None:
            reads, offsets = pack(reads)
        reads = _u8(reads)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        nw = lib().bwt_pack_words(offsets.ctypes.data, n)
        if nw < 0:
            raise BwtError(-1, "offsets must be non-decreasing")

        def buf(count, dtype):
            if pinned:
                import torch
                t = torch.empty(max(1, count), dtype={np.uint32: torch.int32, np.uint16: torch.int16, np.int64: torch.int64,
                                                      np.uint8: torch.uint8}[dtype]).pin_memory()
                self._keep.append(t)
                return t.numpy().view(dtype)
            return np.empty(max(1, count), dtype)
        self._keep = []
        self.n = n
        self.words = buf(nw + 1, np.uint32)
        self.lens = buf(n, np.uint16)
        ne, nb, nwo = C.c_int64(), C.c_int64(), C.c_int64()
        ecap, bcap = 0, 0
        self.exc_idx = self.exc_off = self.exc_ascii = None
        while True:
            rc = lib().bwt_pack_reads(reads.ctypes.data if len(reads) else None, offsets.ctypes.data, n, int(threads),
                                      self.words.ctypes.data, len(self.words), self.lens.ctypes.data, _ptr(self.exc_idx), ecap,
                                      _ptr(self.exc_off), _ptr(self.exc_ascii), bcap, C.byref(nwo), C.byref(ne), C.byref(nb))
            if rc == -6 and (ne.value > ecap or nb.value > bcap):
                ecap, bcap = int(ne.value), int(nb.value)
                self.exc_idx = buf(ecap, np.uint32)
                self.exc_off = buf(ecap + 1, np.int64)
                self.exc_ascii = buf(bcap, np.uint8)
                continue
            check(rc)
            break
        self.n_words, self.n_exc, self.exc_bytes = int(nwo.value), int(ne.value), int(nb.value)
        self.fixed_len = int(fixed_len)
        if self.fixed_len:
            assert n == 0 or (self.lens[:n] == self.fixed_len).all(), "fixed_len given for a ragged batch"
        nblk = n // 65536 + 2
        self.word_index = np.zeros(nblk, np.int64)
        self.maxlen_index = np.zeros(nblk, np.uint16)
        check(lib().bwt_pack_index(self.lens.ctypes.data, n, self.word_index.ctypes.data, self.maxlen_index.ctypes.data))
        self.struct = BwtPackedReads(n, self.fixed_len, 0, self.lens.ctypes.data, self.words.ctypes.data, self.n_words, self.n_exc,
                                     _ptr(self.exc_idx), _ptr(self.exc_off), _ptr(self.exc_ascii), self.word_index.ctypes.data,
                                     self.maxlen_index.ctypes.data)

    def upload_bytes(self) -> int:
        """bytes a run moves to the device for this batch"""
        return 4 * (self.n_words + 1) + (0 if self.fixed_len else 2 * self.n) + self.n_exc * 12 + self.exc_bytes


def expand_compact(read_status8, hit_off32, hits16, cigars, read_lens, shift: int):
    """compact result (uint8 status, uint32 offsets, HIT16 records, arena) -> the wide layout of ``Index.align``:
    (int32 status, int64 offsets, HIT_DTYPE hits, cigar bytes) with "<Lq>M" CIGARs written out and Pos = S - (Lq+D) + Shift"""
    n = len(read_status8)
    ho = hit_off32.astype(np.int64)
    nh = int(ho[n]) if n else 0
    h16 = hits16[:nh]
    out = np.zeros(nh, HIT_DTYPE)
    lens = np.repeat(np.asarray(read_lens, np.int64), np.diff(ho))
    d = (h16["d_flags"] & 0x0FFFFFFF).astype(np.int64)
    out["s"] = h16["s"]
    out["d"] = d
    out["status"] = (h16["d_flags"] >> 28) & 3
    out["score"] = h16["score"]
    out["pos"] = h16["s"].astype(np.int64) - (lens + d) + int(shift)
    full = (h16["d_flags"] & HIT16_FULL_M) != 0
    texts = []
    cig = bytes(cigars)
    off = 0
    clen = np.zeros(nh, np.int32)
    coff = np.zeros(nh, np.int64)
    for k in range(nh):
        if full[k]:
            t = b"%dM" % lens[k]
        else:
            t = cig[int(h16["cigar_off"][k]):int(h16["cigar_off"][k]) + int(h16["cigar_len"][k])]
        texts.append(t)
        coff[k] = off
        clen[k] = len(t)
        off += len(t)
    out["cigar_off"] = coff
    out["cigar_len"] = clen
    return read_status8.astype(np.int32), ho, out, b"".join(texts)


class Batch:
    """Resident batch: upload / run / download as separate calls (``bwt_batch``)."""

    def __init__(self, index: Index, max_reads: int, max_read_bytes: int):
        self.index = index
        self._b = C.c_void_p()
        check(lib().bwt_batch_create(index.handle, int(max_reads), int(max_read_bytes), C.byref(self._b)))

    def close(self):
        if self._b:
            lib().bwt_batch_destroy(self._b)
            self._b = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, reads: np.ndarray, offsets: np.ndarray):
        self._n = len(offsets) - 1
        check(lib().bwt_batch_upload(self._b, reads.ctypes.data, offsets.ctypes.data, self._n))

    def upload_ptr(self, reads_ptr: int, offsets_ptr: int, n: int):
        self._n = n
        check(lib().bwt_batch_upload(self._b, reads_ptr, offsets_ptr, n))

    def upload_packed(self, packed: "PackedReads"):
        self._n = packed.n
        check(lib().bwt_batch_upload_packed(self._b, C.byref(packed.struct)))

    def run(self):
        check(lib().bwt_batch_run(self._b))

    def run_compact(self):
        check(lib().bwt_batch_run_compact(self._b))

    def download_compact(self):
        """-> (uint8 status[n], uint32 offsets[n+1], HIT16 hits, uint8 cigars)"""
        nh, nc = self.result_size()
        n = self._n
        st = np.zeros(n, np.uint8)
        ho = np.zeros(n + 1, np.uint32)
        hits = np.zeros(max(nh, 1), HIT16_DTYPE)
        cig = np.zeros(max(nc, 1), np.uint8)
        check(lib().bwt_batch_download_compact(self._b, st.ctypes.data, ho.ctypes.data, hits.ctypes.data, len(hits), cig.ctypes.data, len(cig)))
        return st, ho, hits[:nh], cig[:nc]

    def result_size(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        check(lib().bwt_batch_result_size(self._b, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def download(self, read_status=None, hit_offsets=None, hits=None, cigars=None):
        nh, nc = self.result_size()
        n = self._n
        if read_status is None:
            read_status = np.zeros(n, np.int32)
        if hit_offsets is None:
            hit_offsets = np.zeros(n + 1, np.int64)
        if hits is None:
            hits = np.zeros(max(nh, 1), HIT_DTYPE)
        if cigars is None:
            cigars = np.zeros(max(nc, 1), np.uint8)
        check(lib().bwt_batch_download(self._b, read_status.ctypes.data, hit_offsets.ctypes.data, hits.ctypes.data,
                                       len(hits), cigars.ctypes.data, len(cigars)))
        return read_status, hit_offsets, hits[:nh], cigars[:nc]

    def download_ptr(self, status_ptr, offs_ptr, hits_ptr, hits_cap, cig_ptr, cig_cap):
        check(lib().bwt_batch_download(self._b, status_ptr, offs_ptr, hits_ptr, hits_cap, cig_ptr, cig_cap))

    def stats(self) -> BwtStats:
        s = BwtStats()
        check(lib().bwt_batch_stats(self._b, C.byref(s)))
        return s

    def device_result(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().bwt_batch_device_result(self._b, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


class Comm:
    """NCCL communicator of the genome-partition mode (``bwt_comm``): the gather of the per-partition records straight out
    of the batches' device-resident results.  ``unique_id()`` on one rank, handed to the others by the host program."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(lib().bwt_comm_unique_id(buf))
        return bytes(buf)

    def __init__(self, uid: Optional[bytes], world: int, rank: int, device: int = 0):
        self.world, self.rank = int(world), int(rank)
        self._c = C.c_void_p()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(uid) if uid is not None else None
        check(lib().bwt_comm_create(idbuf, self.world, self.rank, int(device), C.byref(self._c)))

    def close(self):
        if self._c:
            lib().bwt_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def gather(self, batch: "Batch", root: int = 0, out=None):
        """Collective over the ranks' batches (run with ``run_compact`` on the same reads).  ``out`` = (status uint8[world*n],
        offsets uint32[world*(n+1)], hits HIT16[cap], cigars uint8[cap]) on the root (allocated when None; grown on NoSpace).
        -> (parts, ms, link_bytes): parts[p] = (status, offsets, hits, cigars) views of rank p's block on the root, None elsewhere."""
        n, W = batch._n, self.world
        ph = np.zeros(W, np.int64)
        pc = np.zeros(W, np.int64)
        ms = C.c_float()
        lb = C.c_int64()
        is_root = self.rank == root
        if is_root and out is None:
            nh, nc = batch.result_size()
            out = (np.zeros(W * n, np.uint8), np.zeros(W * (n + 1), np.uint32), np.zeros(max(16, 2 * W * nh), HIT16_DTYPE),
                   np.zeros(max(256, 2 * W * nc), np.uint8))
        st, ho, hits, cig = out if is_root else (None, None, None, None)
        rc = lib().bwt_partition_gather(self._c, batch._b, int(root), _ptr(st), _ptr(ho), _ptr(hits), len(hits) if is_root else 0,
                                        _ptr(cig), len(cig) if is_root else 0, ph.ctypes.data, pc.ctypes.data, C.byref(ms), C.byref(lb))
        if rc == -6:
            raise NoSpace(lib().bwt_last_error().decode(), (int(ph.sum()), int(pc.sum())))
        check(rc)
        if not is_root:
            return None, float(ms.value), int(lb.value)
        parts, hb, cb = [], 0, 0
        for p in range(W):
            parts.append((st[p * n:(p + 1) * n], ho[p * (n + 1):(p + 1) * (n + 1)], hits[hb:hb + ph[p]], cig[cb:cb + pc[p]]))
            hb += int(ph[p]); cb += int(pc[p])
        return parts, float(ms.value), int(lb.value)


class Aligner:
    """Pipelined ``bwt_align_batch``: persistent chunk buffers, copies overlapped with kernels (``bwt_aligner``)."""

    def __init__(self, index: Index, chunk_reads: int = 1 << 20, chunk_bytes: int = 256 << 20, depth: int = 5):
        self.index = index
        self._a = C.c_void_p()
        check(lib().bwt_aligner_create(index.handle, int(chunk_reads), int(chunk_bytes), int(depth), C.byref(self._a)))

    def close(self):
        if self._a:
            lib().bwt_aligner_destroy(self._a)
            self._a = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_ptr(self, reads_ptr, offs_ptr, n, status_ptr, hit_offs_ptr, hits_ptr, hits_cap, cig_ptr, cig_cap):
        """raw-pointer form (pinned buffers of the caller); -> (n_hits, cigar_bytes); raises NoSpace with the needed sizes"""
        nh, nc = C.c_int64(), C.c_int64()
        rc = lib().bwt_aligner_run(self._a, reads_ptr, offs_ptr, int(n), status_ptr, hit_offs_ptr, hits_ptr, int(hits_cap),
                                   C.byref(nh), cig_ptr, int(cig_cap), C.byref(nc))
        if rc == -6:
            raise NoSpace(lib().bwt_last_error().decode(), (int(nh.value), int(nc.value)))
        check(rc)
        return int(nh.value), int(nc.value)

    def run_packed_ptr(self, packed: "PackedReads", status_ptr, hit_offs_ptr, hits_ptr, hits_cap, cig_ptr, cig_cap):
        """packed reads in, compact records out (raw pointers of the caller); -> (n_hits, cigar_bytes)"""
        nh, nc = C.c_int64(), C.c_int64()
        rc = lib().bwt_aligner_run_packed(self._a, C.byref(packed.struct), status_ptr, hit_offs_ptr, hits_ptr, int(hits_cap),
                                          C.byref(nh), cig_ptr, int(cig_cap), C.byref(nc))
        if rc == -6:
            raise NoSpace(lib().bwt_last_error().decode(), (int(nh.value), int(nc.value)))
        check(rc)
        return int(nh.value), int(nc.value)

    def run_packed(self, packed: "PackedReads"):
        """-> (uint8 status[n], uint32 offsets[n+1], HIT16 hits, uint8 cigars)"""
        n = packed.n
        st = np.zeros(n, np.uint8)
        ho = np.zeros(n + 1, np.uint32)
        hcap, ccap = max(16, 2 * n), max(256, 16 * n)
        while True:
            hits = np.zeros(hcap, HIT16_DTYPE)
            cig = np.zeros(ccap, np.uint8)
            try:
                nh, nc = self.run_packed_ptr(packed, st.ctypes.data, ho.ctypes.data, hits.ctypes.data, hcap, cig.ctypes.data, ccap)
            except NoSpace as e:
                hcap, ccap = e.needed[0] + 16, e.needed[1] + 16
                continue
            return st, ho, hits[:nh].copy(), cig[:nc].copy()

    def run(self, reads, offsets=None):
        if offsets is None:
            reads, offsets = pack(reads)
        reads = _u8(reads)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        st = np.zeros(n, np.int32)
        ho = np.zeros(n + 1, np.int64)
        hcap, ccap = max(16, 4 * n), max(256, 64 * n)
        while True:
            hits = np.zeros(hcap, HIT_DTYPE)
            cig = np.zeros(ccap, np.uint8)
            try:
                nh, nc = self.run_ptr(reads.ctypes.data, offsets.ctypes.data, n, st.ctypes.data, ho.ctypes.data,
                                      hits.ctypes.data, hcap, cig.ctypes.data, ccap)
            except NoSpace as e:
                hcap, ccap = e.needed[0] + 16, e.needed[1] + 16
                continue
            return st, ho, hits[:nh].copy(), cig[:nc].tobytes()

    def stats(self) -> BwtStats:
        s = BwtStats()
        check(lib().bwt_aligner_stats(self._a, C.byref(s)))
        return s


def sample_reads(index: Index, n: int, read_len: int, seed: int = 42, sub_rate: float = 0.01, out: Optional[np.ndarray] = None,
                 want_starts: bool = False):
    """Synthetic fixed-length reads sampled on the device from the index's .ref -> (uint8[n*read_len], offsets, starts|None)"""
    if out is None:
        out = np.empty(n * read_len, np.uint8)
    starts = np.empty(n, np.int64) if want_starts else None
    check(lib().bwt_sample_reads(index.handle, int(n), int(read_len), int(seed), float(sub_rate), out.ctypes.data, _ptr(starts)))
    offs = np.arange(n + 1, dtype=np.int64) * read_len
    return out, offs, starts


def sample_reads_edits(index: Optional[Index], n: int, read_len: int, seed: int = 42, edit_rate: float = 0.01,
                       ins_share: float = 0.1, del_share: float = 0.1, n_rate: float = 0.0, synth: Optional[Tuple[int, int]] = None,
                       device: int = 0, out: Optional[np.ndarray] = None, offsets: Optional[np.ndarray] = None,
                       want_starts: bool = False):
    """SURVEY.md 8(d) edit model (substitutions / 1-base insertions / deletions; ragged lengths), sampled on the device
    from the index's .ref, or -- index None, synth = (length, seed) -- from the synthetic reference itself.
    -> (uint8[total], int64 offsets[n+1], starts|None)"""
    o = BwtSampleOpts(int(n), int(read_len), int(device), int(seed), float(edit_rate), float(ins_share), float(del_share),
                      float(n_rate), int(synth[0]) if synth else 0, int(synth[1]) if synth else 0)
    if offsets is None:
        offsets = np.empty(n + 1, np.int64)
    if out is None:
        out = np.empty(n * read_len + max(4096, n // 8), np.uint8)
    starts = np.empty(n, np.int64) if want_starts else None
    tot = C.c_int64()
    h = index.handle if index is not None else None
    rc = lib().bwt_sample_reads_edits(h, C.byref(o), out.ctypes.data, len(out), offsets.ctypes.data, _ptr(starts), C.byref(tot))
    if rc == -6:
        raise NoSpace(lib().bwt_last_error().decode(), (int(tot.value),))
    check(rc)
    return out[:tot.value], offsets, starts


def search_resident(index: Index, subs: np.ndarray, sub_len: int, repeats: int = 1):
    """-> (ms for all repeats, LF steps per repeat, LF steps per repeat not served by the 8-mer table)"""
    subs = _u8(subs)
    n = len(subs) // sub_len
    ms = C.c_float()
    lf, hbm = C.c_uint64(), C.c_uint64()
    check(lib().bwt_search_resident(index.handle, subs.ctypes.data, n, sub_len, repeats, C.byref(ms), C.byref(lf),
                                    C.byref(hbm)))
    return float(ms.value), int(lf.value), int(hbm.value)


def peak_gather_gbs(device: int = 0, nbytes: int = 8 << 30) -> float:
    g = C.c_double()
    check(lib().bwt_peak_gather_gbs(device, int(nbytes), C.byref(g)))
    return float(g.value)


def peak_dpx_gops(device: int = 0) -> float:
    g = C.c_double()
    check(lib().bwt_peak_dpx_gops(device, C.byref(g)))
    return float(g.value)


__all__ = ["Index", "Batch", "Aligner", "Comm", "PackedReads", "expand_compact", "sample_reads", "sample_reads_edits", "HIT16_DTYPE", "BwtError", "NoSpace", "pack", "sw_batch", "search_resident", "peak_gather_gbs",
           "peak_dpx_gops", "device_count", "HIT_DTYPE"]
