"""CPU restatement of the cloudozer/BWT read-alignment hot path  --  TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: a literal, slow, pure-Python/numpy restatement of the reference's
Erlang algorithm.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import it; the product (``bwt_b200``) never does and fails loudly without its CUDA library.

Parity status: PINNED for the index builder / on-disk format (the shipped
``fm_indices/GL000193.1_p1.fm`` is reproduced byte-for-byte from ``GL000193.1_p1.ref``; see
``tests/test_oracle_golden.py``).  The reference tree holds NO golden search / seed / SW outputs, so
for those functions the Erlang source is the only specification and this file follows it clause by
clause (citations are ``/root/reference`` file:line).  The reference cannot be executed in this
container (no erl/erlc/escript), so there is no ``oracle/_ref``.

Row numbering follows Erlang: FM rows are 1-based ``element(e, FM)``; the stored ``I`` field is the
0-based LF target that the reference (mis)uses as a 1-based row number.  That
off-by-one is reproduced, not fixed.
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

# ---------------------------------------------------------------------------------------------
# include/bwt.hrl:1-20 -- compile-time tunables of the reference
# ---------------------------------------------------------------------------------------------
TOLERANCE = 20          # bwt.hrl:4
MIN_LEN = 8             # bwt.hrl:5
MAX_RANGE = 45          # bwt.hrl:6
MAX_LEN = 15            # bwt.hrl:7
MATCH = 2               # bwt.hrl:14
UNKNOWN = 1.6           # bwt.hrl:15  (a float: any cell touching 'N' turns scores into doubles)
MISMATCH = -1           # bwt.hrl:16
GAP_PENALTY = -2        # bwt.hrl:17
GAP_EXT_PENALTY = -1    # bwt.hrl:18
REF_EXTENSION_LEN = 200  # worker_bwt.erl:17
OVERLAP = 2             # fm_index.erl:11

NO_SEEDS = "no_seeds"
TOO_MANY_SEEDS = "too_many_seeds"
NO_MATCH = "no_match"


class ReferenceCrash(Exception):
    """The reference would crash here (case_clause / badmatch / badarg); the message names the Erlang error."""


# ---------------------------------------------------------------------------------------------
# The FM index as the reference stores it: n rows {F, L, I, SA} + meta
# ---------------------------------------------------------------------------------------------
class FMIndex:
    """Columns of the reference's row tuple (fmi.erl:43-62) plus meta (fm_index.erl:139).

    ``F, L`` are uint8 ASCII, ``I`` / ``SA`` int64.  ``row(e)`` is Erlang's ``element(e, FM)`` (1-based).
    """

    def __init__(self, F, L, I, SA, pointers, shift):
        self.F = np.asarray(F, dtype=np.uint8)
        self.L = np.asarray(L, dtype=np.uint8)
        self.I = np.asarray(I, dtype=np.int64)
        self.SA = np.asarray(SA, dtype=np.int64)
        self.pointers = tuple(int(x) for x in pointers)  # (Pc, Pg, Pt, Last), 1-based
        self.shift = int(shift)
        self.n = len(self.L)

    def row(self, e: int):
        if e < 1 or e > self.n:
            raise ReferenceCrash("badarg: element(%d, FM)" % e)
        k = e - 1
        return int(self.F[k]), int(self.L[k]), int(self.I[k]), int(self.SA[k])


# ---------------------------------------------------------------------------------------------
# Offline builder (fm_index.erl:127-141, bwt.erl:84-120, dsa.erl:15-48, fmi.erl:43-84)
# ---------------------------------------------------------------------------------------------
_IUPAC_TO_ACGT = {  # fm_index.erl:131-135 ; 'N' -> random:uniform(4) is NOT restated (unpinned)
    ord("B"): ord("C"), ord("D"): ord("G"), ord("R"): ord("A"), ord("Y"): ord("C"), ord("K"): ord("T"),
    ord("M"): ord("A"), ord("S"): ord("C"), ord("W"): ord("A"), ord("V"): ord("A"),
    ord("A"): ord("A"), ord("C"): ord("C"), ord("G"): ord("G"), ord("T"): ord("T"),
}


AS183_SEED0 = (3172, 9814, 20125)    # random:seed0/0 (OTP stdlib random.erl; NOT in the reference tree: unpinned there)


def as183_uniform(state: list) -> float:
    """random:uniform/0 of OTP's `random` module (AS183); state = [A1, A2, A3], updated in place."""
    state[0] = (state[0] * 171) % 30269
    state[1] = (state[1] * 172) % 30307
    state[2] = (state[2] * 170) % 30323
    r = state[0] / 30269 + state[1] / 30307 + state[2] / 30323
    return r - int(r)


def map_chunk_to_acgt(chunk: bytes, n_state: Optional[list] = None) -> bytes:
    """fm_index.erl:131-136.  n_state: AS183 state carried across the chunks of one make_indices run ('N' ->
    element(random:uniform(4), {$A,$C,$G,$T})); None: a chunk with 'N' is refused."""
    out = bytearray(len(chunk))
    for k, c in enumerate(chunk):
        if c == ord("N"):
            if n_state is None:
                raise ValueError("chunk holds 'N': the reference substitutes random:uniform(4) (AS183): pass n_state")
            out[k] = b"ACGT"[int(as183_uniform(n_state) * 4)]
            continue
        if c not in _IUPAC_TO_ACGT:
            raise ReferenceCrash("error(%d): unknown symbol (fm_index.erl:136)" % c)
        out[k] = _IUPAC_TO_ACGT[c]
    return bytes(out)


def suffix_array(text: bytes) -> np.ndarray:
    """Suffix array of ``text`` (which must end in a unique smallest ``$``) by prefix doubling.

    dsa.erl:15-48 sorts suffixes under plain byte order with ``$`` (36) below ``ACGT``; every correct
    suffix array of X$ is that same array, so any construction is admissible here.
    """
    t = np.frombuffer(text, dtype=np.uint8).astype(np.int64)
    n = len(t)
    rank = t.copy()
    sa = np.argsort(rank, kind="stable")
    k = 1
    while True:
        r2 = np.full(n, -1, dtype=np.int64)
        r2[: n - k] = rank[k:]
        key = rank * (n + 257) + (r2 + 1)
        sa = np.argsort(key, kind="stable")
        ks = key[sa]
        newrank = np.empty(n, dtype=np.int64)
        newrank[sa] = np.concatenate(([0], np.cumsum(ks[1:] != ks[:-1])))
        rank = newrank
        if rank.max() == n - 1:
            break
        k *= 2
    return sa


def build_fm(chunk: bytes, shift: int) -> FMIndex:
    """fm_index:make_index/4 minus file I/O: ``bwt:fm(st:append($$, Ref_seq))`` + pointers + shift."""
    seq = map_chunk_to_acgt(chunk) + b"$"           # fm_index.erl:131-138
    sa = suffix_array(seq)                          # bwt.erl:87 (dsa:dsa)
    t = np.frombuffer(seq, dtype=np.uint8)
    n = len(t)
    F = t[sa]                                       # dsa.erl:19,28 : F = first char of the suffix
    L = t[(sa - 1) % n]                             # dsa.erl:19,28 : L = char before it ($ row wraps)
    # bwt.erl:105-120 queues the 0-based row number K of every row by its F; fmi.erl:43-62 hands the
    # next queued K of character L to each row in order => I = 0-based LF target.
    I = np.zeros(n, dtype=np.int64)
    for c in (ord("$"), ord("A"), ord("C"), ord("G"), ord("T")):
        rows_with_F = np.nonzero(F == c)[0]
        rows_with_L = np.nonzero(L == c)[0]
        I[rows_with_L] = rows_with_F
    # fmi.erl:66-84 : first 1-based row whose F is C / G / T (searching on from the previous hit), Last = n
    def find_pointer(ch, start):
        e = start
        while True:
            if e > n:
                raise ReferenceCrash("badarg: find_pointer ran off the index")
            if F[e - 1] == ch:
                return e
            e += 1
    pc = find_pointer(ord("C"), 1)
    pg = find_pointer(ord("G"), pc + 1)
    pt = find_pointer(ord("T"), pg + 1)
    return FMIndex(F, L, I, sa, (pc, pg, pt, n), shift)


def chunk_plan(body_len: int, nbr: Optional[int] = None) -> List[Tuple[int, int]]:
    """fm_index.erl:113-124 : [(Start (1-based), length actually cut)] for J = 1..Nbr.

    ``lists:sublist(Body, Start, Chunk_size)`` silently clips at the end of Body, so the last chunk is
    shorter than the nominal Chunk_size."""
    if nbr is None:
        nbr = 1 + ((body_len + OVERLAP) >> 24)
    size = (body_len + nbr * OVERLAP) // nbr
    plan = []
    for j in range(1, nbr + 1):
        start = 1 + (j - 1) * (size - OVERLAP)
        plan.append((start, max(0, min(size, body_len - start + 1))))
    return plan


def fasta_contig_body(fasta: bytes, name: str) -> Tuple[int, bytes]:
    """fm_index.erl:71-116,172-177 : (Shift, Body) of one contig: N/CR/LF trimmed both ends, CR/LF stripped."""
    lines = fasta.split(b"\n")
    body = bytearray()
    found = False
    for ln in lines:
        if ln.startswith(b">"):
            if found:
                break
            toks = ln[1:].replace(b"\n", b" ").split(b" ")
            toks = [x for x in toks if x]
            found = bool(toks) and toks[0].decode() == name
            continue
        if found:
            body += ln + b"\n"
    if not found:
        raise KeyError(name)
    raw = bytes(body)
    skip = (ord("N"), 10, 13)
    hi = len(raw)
    while hi > 0 and raw[hi - 1] in skip:
        hi -= 1
    lo = 0
    shift = 0
    while lo < hi and raw[lo] in skip:
        if raw[lo] == ord("N"):
            shift += 1
        lo += 1
    return shift, bytes(c for c in raw[lo:hi] if c not in (10, 13))


# ---------------------------------------------------------------------------------------------
# Erlang external term format for {Meta, FM}  (fm_index.erl:139-141,154-156; layout: SURVEY.md App. B)
# ---------------------------------------------------------------------------------------------
def _etf_int(v: int) -> bytes:
    if 0 <= v <= 255:
        return b"\x61" + bytes([v])
    if -(1 << 31) <= v < (1 << 31):
        return b"\x62" + struct.pack(">i", v)
    mag = abs(v)
    digits = mag.to_bytes((mag.bit_length() + 7) // 8, "little")
    return b"\x6e" + bytes([len(digits), 1 if v < 0 else 0]) + digits


def write_fm(fm: FMIndex) -> bytes:
    out = bytearray(b"\x83\x68\x02\x6c\x00\x00\x00\x02")
    out += b"\x68\x02\x64\x00\x08pointers\x68\x04" + b"".join(_etf_int(p) for p in fm.pointers)
    out += b"\x68\x02\x64\x00\x05shift" + _etf_int(fm.shift) + b"\x6a"
    if fm.n <= 255:
        out += b"\x68" + bytes([fm.n])
    else:
        out += b"\x69" + struct.pack(">I", fm.n)
    for k in range(fm.n):
        out += b"\x68\x04" + _etf_int(int(fm.F[k])) + _etf_int(int(fm.L[k])) + _etf_int(int(fm.I[k])) + _etf_int(int(fm.SA[k]))
    return bytes(out)


class _Etf:
    def __init__(self, data: bytes):
        self.d = data
        self.p = 0

    def u8(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def take(self, k):
        v = self.d[self.p:self.p + k]
        if len(v) != k:
            raise ValueError("truncated ETF")
        self.p += k
        return v

    def term(self):
        tag = self.u8()
        if tag == 0x61:
            return self.u8()
        if tag == 0x62:
            return struct.unpack(">i", self.take(4))[0]
        if tag == 0x6E:
            k = self.u8()
            sign = self.u8()
            v = int.from_bytes(self.take(k), "little")
            return -v if sign else v
        if tag == 0x64:
            k = struct.unpack(">H", self.take(2))[0]
            return ("atom", self.take(k).decode("latin1"))
        if tag == 0x68:
            return tuple(self.term() for _ in range(self.u8()))
        if tag == 0x69:
            return tuple(self.term() for _ in range(struct.unpack(">I", self.take(4))[0]))
        if tag == 0x6A:
            return []
        if tag == 0x6C:
            k = struct.unpack(">I", self.take(4))[0]
            items = [self.term() for _ in range(k)]
            tail = self.term()
            if tail != []:
                raise ValueError("improper list")
            return items
        if tag == 0x6B:
            k = struct.unpack(">H", self.take(2))[0]
            return list(self.take(k))
        raise ValueError("unsupported ETF tag %d at %d" % (tag, self.p - 1))


def parse_fm(data: bytes) -> FMIndex:
    """``binary_to_term`` of a ``_pJ.fm`` file (fm_index.erl:154-156) -> FMIndex."""
    r = _Etf(data)
    if r.u8() != 0x83:
        raise ValueError("not an external term")
    meta, rows = r.term()
    props = {k[1]: v for (k, v) in meta}
    arr = np.array(rows, dtype=np.int64).reshape(-1, 4)
    return FMIndex(arr[:, 0], arr[:, 1], arr[:, 2], arr[:, 3], props["pointers"], props["shift"])


# ---------------------------------------------------------------------------------------------
# fmi.erl:25-39,89-100  -- LF step by linear scan, and locate
# ---------------------------------------------------------------------------------------------
def se_down(sp: int, ep: int, c2: int, fm: FMIndex):
    """fmi.erl:25-31 : I of the first row in [Sp, Ep) whose L is C2, else not_found (None)."""
    while True:
        if sp == ep:
            return None
        _, l, p, _ = fm.row(sp)
        if l == c2:
            return p
        sp += 1


def se_up(ep: int, c2: int, fm: FMIndex) -> int:
    """fmi.erl:35-39 : I of the last row at or above Ep whose L is C2 (no lower bound)."""
    while True:
        _, l, p, _ = fm.row(ep)
        if l == c2:
            return p
        ep -= 1


def get_seed_ends(fm: FMIndex, sp: int, ep: int, n: int):
    """fmi.erl:89-100 : [SA(Ep)+N, ..., SA(Sp)+N] or no_seeds when fewer than MIN_LEN chars matched."""
    if n < MIN_LEN:
        return NO_SEEDS
    acc: List[int] = []
    while True:
        if sp == ep:
            return [fm.row(ep)[3] + n] + acc
        acc = [fm.row(sp)[3] + n] + acc
        sp += 1
        if sp > fm.n + 1:  # the Erlang recursion would crash on element/2 first
            raise ReferenceCrash("badarg: get_seed_ends with Sp > Ep")


# ---------------------------------------------------------------------------------------------
# bwa.erl:12-45
# ---------------------------------------------------------------------------------------------
def find_seeds(fm: FMIndex, subseq: bytes):
    pc, pg, pt, last = fm.pointers
    rev = subseq[::-1]
    h = rev[0]
    if h == ord("A") or h == ord("N"):      # bwa.erl:16,21-23
        sp, ep = 1, pc - 1
    elif h == ord("C"):
        sp, ep = pc, pg - 1
    elif h == ord("G"):
        sp, ep = pg, pt - 1
    elif h == ord("T"):
        sp, ep = pt, last
    else:
        raise ReferenceCrash("case_clause: %d (bwa.erl:15)" % h)
    tail = rev[1:]
    n = 1
    k = 0
    while True:                              # bwa.erl:31-45, clause order preserved
        if n == MAX_LEN:
            if ep - sp > MAX_RANGE:
                return TOO_MANY_SEEDS
            return get_seed_ends(fm, sp, ep, MAX_LEN)
        if k < len(tail) and tail[k] == ord("N"):
            if ep - sp > MAX_RANGE:
                return TOO_MANY_SEEDS
            return get_seed_ends(fm, sp, ep, n)
        if k < len(tail):
            c2 = tail[k]
            sp1 = se_down(sp, ep, c2, fm)
            if sp1 is None:
                return get_seed_ends(fm, sp, ep, n)
            ep1 = se_up(ep, c2, fm)
            sp, ep, k, n = sp1, ep1, k + 1, n + 1
            continue
        if ep - sp > MAX_RANGE:
            return TOO_MANY_SEEDS
        return get_seed_ends(fm, sp, ep, n)


def find_seeds_interval(fm: FMIndex, subseq: bytes):
    """Same walk as find_seeds but returns (status, Sp, Ep, N) -- what the search kernel emits.

    status: 0 = seeds (N >= 8), 1 = no_seeds, 2 = too_many_seeds.
    """
    pc, pg, pt, last = fm.pointers
    rev = subseq[::-1]
    h = rev[0]
    if h in (ord("A"), ord("N")):
        sp, ep = 1, pc - 1
    elif h == ord("C"):
        sp, ep = pc, pg - 1
    elif h == ord("G"):
        sp, ep = pg, pt - 1
    elif h == ord("T"):
        sp, ep = pt, last
    else:
        raise ReferenceCrash("case_clause")
    tail = rev[1:]
    n, k = 1, 0
    while True:
        checked = False
        if n == MAX_LEN:
            checked = True
        elif k < len(tail) and tail[k] == ord("N"):
            checked = True
        elif k < len(tail):
            sp1 = se_down(sp, ep, tail[k], fm)
            if sp1 is not None:
                sp, ep, k, n = sp1, se_up(ep, tail[k], fm), k + 1, n + 1
                continue
        else:
            checked = True
        if checked and ep - sp > MAX_RANGE:
            return 2, sp, ep, n
        return (0 if n >= MIN_LEN else 1), sp, ep, n


# ---------------------------------------------------------------------------------------------
# sga.erl:17-76
# ---------------------------------------------------------------------------------------------
def skip_Ns(q: bytes):
    """sga.erl:69-76 : (Skipped, trimmed string) or None (no_more_subseqs)."""
    if len(q) < MIN_LEN:
        return None
    j = 0
    skipped = 0
    for ch in q[::-1]:
        if j == MIN_LEN:
            break
        if ch == ord("N"):
            skipped += j
            j = 1
        else:
            j += 1
    else:
        if j != MIN_LEN:
            return None
    if skipped == 0:
        return 0, q
    return skipped, q[: len(q) - skipped]


def get_similar(qty: int, ls: Sequence[int]) -> List[Tuple[int, int]]:
    """sga.erl:42-64 : vote sorted seed ends into chains; result is DESCENDING by position."""
    if qty < 2:
        return []
    s = sorted(ls)
    need = 2 + qty // 7
    acc: List[Tuple[int, int]] = []
    p1, count, dist = s[0], 1, 0
    for p2 in s[1:]:
        if p2 - p1 <= TOLERANCE:
            count, dist, p1 = count + 1, dist + p2 - p1, p2
        elif count >= need:
            acc.insert(0, (p1, dist))
            p1, count, dist = p2, 1, 0
        else:
            p1, count, dist = p2, 1, 0
    if count >= need:
        acc.insert(0, (p1, dist))
    return acc


def sga_trace(fm: FMIndex, qseq: bytes):
    """sga.erl:17-34 with its per-iteration decisions exposed (used by tests of the search kernel).

    Returns (result, steps) where steps = [(subseq_end, shift_added, find_seeds_result)].
    """
    acc: List[int] = []
    qty = 0
    shift = 0
    q = qseq
    steps = []
    while True:
        r = skip_Ns(q)
        if r is None:
            return get_similar(qty, acc), steps
        skip, q1 = r
        res = find_seeds(fm, q1)
        steps.append((len(q1), shift + skip, res))
        if isinstance(res, list):
            for s in res:                    # sga.erl:37-39 add_seeds
                acc.insert(0, s + shift + skip)
            qty += 1
        q = q1[: max(len(q1) - MIN_LEN, 0)]  # lists:sublist(Qseq1, Qlen-8)
        shift += MIN_LEN + skip


def sga(fm: FMIndex, qseq: bytes) -> List[Tuple[int, int]]:
    return sga_trace(fm, qseq)[0]


# ---------------------------------------------------------------------------------------------
# sw.erl:47-202
# ---------------------------------------------------------------------------------------------
_DIR_ORDER = {"left": 0, "up": 1, "up_left": 2}   # Erlang atom order: left < up < up_left


def sigma(q: int, r: int):
    """sw.erl:164-202, first matching clause wins.  q = read char, r = reference char."""
    N = ord("N")
    if q == N and r == N:
        return UNKNOWN
    if q == r:
        return MATCH
    if q == N or r == N:
        return UNKNOWN
    qc, rc = chr(q), chr(r)
    if rc == "R":
        return MATCH if qc in "AG" else MISMATCH
    if rc == "Y":
        return MATCH if qc in "CT" else MISMATCH
    if rc == "B":
        return MISMATCH if qc == "A" else MATCH
    if rc == "D":
        return MISMATCH if qc == "C" else MATCH
    if rc == "K":
        return MATCH if qc in "GT" else MISMATCH
    if rc == "M":
        return MATCH if qc in "AC" else MISMATCH
    if rc == "S":
        return MATCH if qc in "GC" else MISMATCH
    if rc == "W":
        return MATCH if qc in "AT" else MISMATCH
    if rc == "V":
        return MISMATCH if qc == "T" else MATCH
    return MISMATCH


def _term_key(v, d):
    # lists:max/1 on {V, Atom}: numbers compare by value (2 == 2.0), then the atom decides.
    return (v, _DIR_ORDER[d])


def sw_literal(qseq: bytes, ref: bytes):
    """sw:sw/2 with the list-of-columns table and the clause-machine traceback (sw.erl:47-160).

    Slow (pure Python lists); use for small cases and to validate ``sw`` below.
    Returns no_match or (Score, CigarString).
    """
    if len(qseq) == 0 or len(ref) == 0:
        return NO_MATCH                                   # sw.erl:54 (non-list / empty arguments)
    lq = len(qseq)
    tab = [[(0, "undef")] for _ in range(len(ref) + 1)]   # sw.erl:61, each column newest-first
    thr = -MATCH * (lq // 2)                              # sw.erl:62
    vmax = 0
    for s in qseq:                                        # build_tab, sw.erl:69-80
        if vmax < thr:
            return NO_MATCH
        first = tab[0]
        v, _ = first[0]
        v1 = v + (GAP_PENALTY if len(first) == 1 else GAP_EXT_PENALTY)
        acct = [[(v1, "up")] + first]
        rowmax = v1
        for col, sr in zip(tab[1:], ref):                 # add_row2tab, sw.erl:88-100
            vup, dup = col[0]
            vleft, dleft = acct[-1][0]
            vul = acct[-1][1][0]
            cands = [
                (vul + sigma(s, sr), "up_left"),
                (vleft + (GAP_EXT_PENALTY if dleft == "left" else GAP_PENALTY), "left"),
                (vup + (GAP_EXT_PENALTY if dup == "up" else GAP_PENALTY), "up"),
            ]
            best = cands[0]
            for c in cands[1:]:                           # lists:max/1
                if _term_key(*c) > _term_key(*best):
                    best = c
            acct.append([best] + col)
            if best[0] > rowmax:                          # erlang:max/2 keeps the first on ==
                rowmax = best[0]
        tab = acct
        vmax = rowmax
        thr += MATCH
    if vmax < thr:
        return NO_MATCH
    # get_starting_point, sw.erl:110-112 : exact (=:=) match on the head of each column, left to right
    start = None
    for j, col in enumerate(tab):
        hv = col[0][0]
        if hv == vmax and type(hv) is type(vmax):
            start = j
            break
    if start is None:
        raise ReferenceCrash("function_clause: get_starting_point")
    left_cols = tab[:start][::-1]                         # nearest-left column first
    _, d0 = tab[start][0]
    column = tab[start][1:]

    def get_str(n, d):
        return str(n) + {"up_left": "M", "left": "D", "up": "I"}[d]

    # get_CIGAR/5,6 clause machine, sw.erl:116-151
    cigar = ""
    tabl = left_cols
    if d0 == "up":
        skip, d, n = 1, d0, 1
        col = column
    else:
        skip, d, n = 1, d0, 1
        col = tabl[0][skip:]
        tabl = tabl[1:]
    while True:
        cell_d = col[0][1]
        if len(col) == 1 and cell_d == "undef":
            cigar = get_str(n, d) + cigar
            break
        if cell_d == "up":
            if d == "up":
                n += 1
            else:
                cigar = get_str(n, d) + cigar
                d, n = "up", 1
            col = col[1:]
            skip += 1
            continue
        if cell_d == d:
            n += 1
        else:
            cigar = get_str(n, d) + cigar
            d, n = cell_d, 1
        if cell_d == "up_left":
            skip += 1
        col = tabl[0][skip:]
        tabl = tabl[1:]
    return int(vmax), cigar                               # trunc/1 (truncates toward zero)


def sw(qseq: bytes, ref: bytes):
    """Same function as ``sw_literal`` on a dense matrix with a plain traceback (numpy-free, still O(Lq*Lr))."""
    lq, lr = len(qseq), len(ref)
    if lq == 0 or lr == 0:
        return NO_MATCH
    UNDEF, LEFT, UP, UL = 3, 0, 1, 2
    prev_v = [0] * (lr + 1)
    prev_d = [UNDEF] * (lr + 1)
    dirs = [prev_d]
    thr = -MATCH * (lq // 2)
    vmax = 0
    for i in range(1, lq + 1):
        if vmax < thr:
            return NO_MATCH
        s = qseq[i - 1]
        cur_v = [0] * (lr + 1)
        cur_d = [UP] * (lr + 1)
        cur_v[0] = prev_v[0] + (GAP_PENALTY if i == 1 else GAP_EXT_PENALTY)
        rowmax = cur_v[0]
        for j in range(1, lr + 1):
            bv, bd = prev_v[j - 1] + sigma(s, ref[j - 1]), UL
            cv = cur_v[j - 1] + (GAP_EXT_PENALTY if cur_d[j - 1] == LEFT else GAP_PENALTY)
            if (cv, LEFT) > (bv, bd):
                bv, bd = cv, LEFT
            cv = prev_v[j] + (GAP_EXT_PENALTY if prev_d[j] == UP else GAP_PENALTY)
            if (cv, UP) > (bv, bd):
                bv, bd = cv, UP
            cur_v[j], cur_d[j] = bv, bd
            if bv > rowmax:
                rowmax = bv
        dirs.append(cur_d)
        prev_v, prev_d = cur_v, cur_d
        vmax = rowmax
        thr += MATCH
    if vmax < thr:
        return NO_MATCH
    j = next(k for k in range(lr + 1) if prev_v[k] == vmax)
    i = lq
    ops: List[str] = []
    while i > 0:
        d = dirs[i][j]
        if d == UL:
            ops.append("M"); i -= 1; j -= 1
        elif d == UP:
            ops.append("I"); i -= 1
        else:
            ops.append("D"); j -= 1
    ops.reverse()
    out = []
    k = 0
    while k < len(ops):
        m = k
        while m < len(ops) and ops[m] == ops[k]:
            m += 1
        out.append("%d%s" % (m - k, ops[k]))
        k = m
    return int(vmax), "".join(out)


# ---------------------------------------------------------------------------------------------
# worker_bwt.erl:73-108 -- per-read assembly
# ---------------------------------------------------------------------------------------------
def pad_ref(ref: bytes) -> bytes:
    ext = b"N" * REF_EXTENSION_LEN                         # worker_bwt.erl:53-55
    return ext + ref + ext


def read_candidates(fm: FMIndex, padded_ref: bytes, qseq: bytes, sw_fn=sw):
    """Returns (sga_result, Cigars) with Cigars as the reference builds it: ascending position,
    each entry ((Score, Cigar), Pos, Ref_window).  worker_bwt.erl:84-97."""
    seeds = sga(fm, qseq)
    cigars = []
    for (s, d) in seeds:
        ref_len = len(qseq) + d
        start = s - ref_len + REF_EXTENSION_LEN
        if start < 0 or start + ref_len > len(padded_ref):
            raise ReferenceCrash("badmatch: window outside the padded reference (worker_bwt.erl:89)")
        window = padded_ref[start:start + ref_len]
        r = sw_fn(qseq, window)
        if r != NO_MATCH:
            cigars.insert(0, (r, s - ref_len + fm.shift, window))
    return seeds, cigars


# ---------------------------------------------------------------------------------------------
# fastq.erl:5-40, master.erl:59,149-178 -- ingest, batch framing and the result sink (rows F2/F3)
# ---------------------------------------------------------------------------------------------
WORKLOAD_SIZE = 200     # master.erl:59
WATERLINE = 10          # worker_bwt.erl:16
SCORE_FILTER = 265      # master.erl:149


class LineDev:
    """file:read_line/1 over bytes: returns the next line INCLUDING its newline (if any), or None at eof."""

    def __init__(self, data: bytes):
        self.d = data
        self.p = 0

    def read_line(self):
        if self.p >= len(self.d):
            return None
        k = self.d.find(b"\n", self.p)
        k = len(self.d) if k < 0 else k + 1
        ln = self.d[self.p:k]
        self.p = k
        return ln


def fastq_read_seq(dev: LineDev):
    """fastq.erl:5-18 -> (name, seq) | None (eof).  lists:droplast/1 drops the LAST character, newline or not."""
    while True:
        ln = dev.read_line()
        if ln is None:
            return None
        if ln[:1] == b"@":
            data = dev.read_line()
            if data is None:
                raise ReferenceCrash("badmatch: {ok,SeqData} = eof (fastq.erl:8)")
            dev.read_line()
            dev.read_line()
            return ln[1:][:-1], data[:-1]


def fastq_read_seqs(dev: LineDev, n: int):
    """fastq.erl:20-40 -> ("ok", Acc) | ("eof", Acc) | "eof"; Acc is in REVERSE file order."""
    acc = []
    while n > 0:
        s = fastq_read_seq(dev)
        if s is None:
            return "eof" if not acc else ("eof", acc)
        acc.insert(0, s)
        n -= 1
    return ("ok", acc)


def produce_workload(dev: LineDev, n: int, state: dict):
    """master.erl:162-178 -> list of lists (newest first) | "stop".  state: {"fastq_eof": bool}"""
    acc = []
    while True:
        if n == 0:
            return acc
        if not state["fastq_eof"]:
            r = fastq_read_seqs(dev, WORKLOAD_SIZE)
            if r == "eof":
                state["fastq_eof"] = True
                return acc
            acc.insert(0, r[1])
            n -= 1
            continue
        if not acc:
            return "stop"
        raise ReferenceCrash("function_clause: produce_workload with fastq_eof and a non-empty accumulator")


def master_sink_line(msg, chromosome: str):
    """master.erl:149-158: None for no_cigar and for scores below 265, else the printed line (io:format
    "~n~s      ~s      ~b      ~s      ~b      ~s~n")."""
    if msg == "no_cigar":
        return None
    _, name, (score, cigar), pos, refseq = msg
    if score < SCORE_FILTER:
        return None
    return "\n%s      %s      %d      %s      %d      %s\n" % (name.decode("latin1"), chromosome, pos, cigar, score,
                                                                refseq.decode("latin1"))


def worker_messages(fm: FMIndex, padded_ref: bytes, qseq_list):
    """worker_bwt.erl:73-108 for one list [(Qname, Qseq)] -> messages in emission order.  Equal top scores: the
    LAST candidate of the ascending list wins (OTP lists:sort/2 with a strict comparator; recalled, unpinned)."""
    seeds = []
    for (name, q) in qseq_list:
        r = sga(fm, q)
        if r:
            seeds.insert(0, ((name, q), r))
    out = []
    for ((name, q), sd) in seeds:
        cigars = []
        for (s, d) in sd:
            ref_len = len(q) + d
            start = s - ref_len + REF_EXTENSION_LEN
            if start < 0 or start + ref_len > len(padded_ref):
                raise ReferenceCrash("badmatch (worker_bwt.erl:89)")
            w = padded_ref[start:start + ref_len]
            c = sw(q, w)
            if c != NO_MATCH:
                cigars.insert(0, (c, s - ref_len + fm.shift, w))
        if not cigars:
            out.append("no_cigar")
        elif len(cigars) == 1:
            out.append(("cigar", name) + cigars[0])
        else:
            best = max(c[0][0] for c in cigars)
            out.append(("cigar", name) + [c for c in cigars if c[0][0] == best][-1])
    return out
