"""CPU restatement of src/seeds.erl (seeds:generate_fs/3 and the fs module it writes)  --  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path
(bwt_b200/csrc/inexact.cu behind bwt_fs_find_seeds_batch) never does.

Parity status: **parity unpinned**.  seeds.erl has no caller, no test and no fixture anywhere in the reference tree
(SURVEY.md 8a row A9), and no Erlang runtime exists in this image to run it.  What follows is a clause-for-clause
restatement of the source, which is the only specification; two independent formulations (the literal pattern list
and an alignment DP, `fs_matches_dp`) are held against each other by tests/test_seeds_oracle.py.

seeds:generate_fs(Qsec, Len, Defects) (seeds.erl:11-47) writes a module `fs` with
  * matches/Len: one `matches(P1,..,PLen) -> true` clause per pattern printed by gen_pattern/5 (a pattern element is a
    character literal or `_`), then a catch-all `-> false`;
  * find_seeds(Ref_Seq) (seeds.erl:23-43): slide a Len-wide window over the list from N = 1; on a match record N and
    jump Len ahead, else advance by one; a tail shorter than Len ends the scan.  The result is the accumulator, i.e.
    1-based window starts in DESCENDING order.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

WILD = ord("*")


class FsCompileError(Exception):
    """The generated fs.erl would not compile: a printed clause has fewer than Len arguments (head mismatch), which
    happens when Qsec is shorter than Len + Defects and a deletion path runs out of query (seeds.erl:51-54: the Tail
    returned by lists:sublist/2 is shorter than N)."""


def gen_pattern(acc: List[int], sub: Sequence[int], length: int, defects: int, out: List[bytes]) -> None:
    """seeds.erl:51-62, clause for clause.  `acc` is the reversed pattern so far, `sub` the unread query."""
    if defects == 0:                                         # :51-55
        n = length - len(acc)
        tail = list(sub[:n])                                 # lists:sublist/2 returns a shorter list without raising
        seq = tail[::-1] + acc                               # still reversed; print_pattern's foldl reverses it back
        out.append(bytes(seq[::-1]))
        return
    if len(sub) > 0 and len(acc) < length:                   # :57
        s, rest = sub[0], sub[1:]
        gen_pattern([WILD] + acc, rest, length, defects - 1, out)   # Mutation   :58
        gen_pattern([WILD] + acc, sub, length, defects - 1, out)    # Insertion  :59
        gen_pattern(acc, rest, length, defects - 1, out)            # Deletion   :60
        gen_pattern([s] + acc, rest, length, defects, out)          # Match      :61
        return
    return                                                   # :62


def generate_patterns(qsec: bytes, length: int, defects: int) -> List[bytes]:
    """The `matches` clauses of fs.erl in the order generate_fs/3 prints them (seeds.erl:45-46)."""
    if length < 2:
        # [_,_,_|Args_str1] = Args_str (seeds.erl:35) needs "X1," in front of a second argument
        raise FsCompileError("Len < 2: badmatch at seeds.erl:35")
    out: List[bytes] = []
    gen_pattern([], list(qsec[:length + defects]), length, defects, out)
    for p in out:
        if len(p) != length:
            raise FsCompileError("clause with %d arguments in matches/%d" % (len(p), length))
    return out


def fs_matches(patterns: Sequence[bytes], window: bytes) -> bool:
    """fs:matches/Len: first clause whose every element is `_` or equals the argument."""
    for p in patterns:
        if all(a == WILD or a == b for a, b in zip(p, window)):
            return True
    return False


def fs_find_seeds(patterns: Sequence[bytes], length: int, ref_seq: bytes) -> List[int]:
    """fs:find_seeds/1 (seeds.erl:23-43), literal."""
    acc: List[int] = []
    n = 1
    rest = 0                                                 # index of the head of Ref_Seq
    total = len(ref_seq)
    while True:
        left = total - rest
        if left == length:                                   # :26-30
            if fs_matches(patterns, ref_seq[rest:rest + length]):
                acc.insert(0, n)
            return acc
        if left > length:                                    # :36-40
            if fs_matches(patterns, ref_seq[rest:rest + length]):
                acc.insert(0, n)
                n += length
                rest += length
            else:
                n += 1
                rest += 1
            continue
        return acc                                           # :42


def match_vector(patterns: Sequence[bytes], length: int, ref_seq: bytes, chunk: int = 1 << 14) -> np.ndarray:
    """bool[len(ref)-length+1]: fs:matches on every window (the literal pattern list, vectorised with numpy)."""
    ref = np.frombuffer(ref_seq, dtype=np.uint8)
    nwin = len(ref) - length + 1
    if nwin <= 0:
        return np.zeros(0, dtype=bool)
    pats = np.frombuffer(b"".join(sorted(set(patterns))), dtype=np.uint8).reshape(-1, length)
    res = np.zeros(nwin, dtype=bool)
    win = np.lib.stride_tricks.sliding_window_view(ref, length)
    for lo in range(0, nwin, chunk):
        w = win[lo:lo + chunk]
        alive = np.zeros(len(w), dtype=bool)
        for p in pats:
            lit = np.nonzero(p != WILD)[0]
            alive |= (w[:, lit] == p[lit]).all(axis=1)
        res[lo:lo + chunk] = alive
    return res


def greedy_scan(match: np.ndarray, length: int) -> List[int]:
    """The window walk of fs:find_seeds/1 over a precomputed match vector -> descending 1-based starts."""
    acc: List[int] = []
    cur = 0
    for m in np.nonzero(match)[0]:
        if m >= cur:
            acc.append(int(m) + 1)
            cur = int(m) + length
    return acc[::-1]


def fs_matches_dp(qsec: bytes, length: int, defects: int, window: bytes) -> bool:
    """Second, independent formulation: is there a path of gen_pattern/5 whose pattern matches `window`?
    State after emitting a pattern elements having read i query characters with d defects left (left to right)."""
    sub = qsec[:length + defects]
    if len(window) != length:
        return False
    frontier = {(0, 0, defects)}
    seen = set()
    while frontier:
        a, i, d = frontier.pop()
        if (a, i, d) in seen:
            continue
        seen.add((a, i, d))
        if d == 0:
            n = length - a
            tail = sub[i:i + n]
            if len(tail) != n:
                raise FsCompileError("short clause")
            if tail == window[a:]:
                return True
            continue
        if i < len(sub) and a < length:
            frontier.add((a + 1, i + 1, d - 1))              # mutation: `_` matches anything
            frontier.add((a + 1, i, d - 1))                  # insertion
            frontier.add((a, i + 1, d - 1))                  # deletion
            if sub[i] == window[a]:
                frontier.add((a + 1, i + 1, d))              # match
    return False
