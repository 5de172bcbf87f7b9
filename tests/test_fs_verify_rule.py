"""The window verification of the bounded-edit front end (fs_verify_window_t in bwt_b200/csrc/inexact.cu), restated in
Python bit for bit -- all states (diff = q - p + D, d, m) of a pattern position in one integer, a transition = mask +
shift -- and held against the oracle's reachability DP on every window of random and repetitive texts.  CPU only: the
kernel itself is compared with the same oracle in tests/test_gpu_inexact.py; this pins the RULE."""
import random

import numpy as np

from oracle.c_oracle import c_fs_scan_dp


def verify_window(W: bytes, Q: bytes, length: int, D: int) -> bool:
    DD, NDI = D + 1, 2 * D + 1
    GRP = DD * DD

    def bit(di, d, m):
        return 1 << ((di * DD + d) * DD + m)

    M_LT = M_DGE1 = M_DIGE1 = M_DILT = M_ACC = 0
    for di in range(NDI):
        for d in range(D + 1):
            for m in range(D + 1):
                b = bit(di, d, m)
                if m < D: M_LT |= b
                if d >= 1: M_DGE1 |= b
                if di >= 1: M_DIGE1 |= b
                if di + 1 < NDI: M_DILT |= b
                q = di - D
                if q >= 0 and q <= d and d - q <= m: M_ACC |= b
    GMASK = (1 << GRP) - 1
    S = 0
    for sft in range(NDI):
        if length - D + sft >= 0:
            S |= bit(sft, D, 0)
    for p in range(length, 0, -1):
        t = W[p - 1]
        HASQ = EQ = 0
        for di in range(NDI):
            q = p + di - D
            if q >= 1:
                HASQ |= GMASK << (di * GRP)
                if Q[q - 1] == t and Q[q - 1] in b"ACGT":
                    EQ |= GMASK << (di * GRP)
        if p < length:
            for _ in range(D):
                S |= (S & M_DGE1 & M_DIGE1 & HASQ) >> (GRP + DD)
        Sm = S & EQ
        N = ((Sm & M_LT) << 1) | (Sm & ~M_LT)
        N |= (S & HASQ & ~EQ & M_DGE1) >> DD
        N |= (S & M_DGE1 & M_DILT) << (GRP - DD)
        S = N
        if not S:
            return False
    return (S & M_ACC) != 0


def matching_starts(q: bytes, length: int, D: int, ref: bytes):
    """all 1-based window starts the oracle says match some pattern (before the greedy walk)"""
    _, vec = c_fs_scan_dp(q, length, D, ref, want_vec=True)
    return set((np.nonzero(vec)[0] + 1).tolist())


def test_bit_parallel_rule_equals_the_oracle_dp():
    rng = random.Random(3)
    checked = hits = 0
    for t in range(60):
        n = rng.choice([60, 200, 500])
        if t % 3 == 0:
            unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 5)))
            ref = (unit * n)[:n]
        else:
            ref = bytes(rng.choice(b"ACGT") for _ in range(n))
        for length, D in ((3, 1), (6, 2), (9, 1), (14, 2), (25, 2), (40, 1)):
            if length + D + 2 > n:
                continue
            for _ in range(4):
                s = rng.randrange(0, n - length - D - 1)
                q = bytearray(ref[s:s + length + D])
                for _k in range(rng.randrange(0, 3)):
                    q[rng.randrange(len(q))] = rng.choice(b"ACGTN")
                if rng.random() < 0.3:
                    del q[rng.randrange(len(q))]
                    q.append(rng.choice(b"ACGT"))
                q = bytes(q)
                want = matching_starts(q, length, D, ref)
                got = {s0 + 1 for s0 in range(0, n - length + 1) if verify_window(ref[s0:s0 + length], q, length, D)}
                assert got == want, (ref, q, length, D, sorted(got ^ want)[:6])
                checked += n - length + 1
                hits += len(want)
    assert checked > 50000 and hits > 500


def test_pigeonhole_completeness():
    """every matching window holds, in one of Defects+1 equal pieces, a verbatim copy of the query at a shift of at most
    +-Defects: the candidate generation of the front end (fs_pigeonhole) cannot miss a window"""
    rng = random.Random(8)
    windows = 0
    for t in range(80):
        n = rng.choice([80, 300])
        if t % 4 == 0:
            unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randrange(1, 4)))
            ref = (unit * n)[:n]
        else:
            ref = bytes(rng.choice(b"ACGT") for _ in range(n))
        for length, D in ((4, 1), (7, 2), (12, 2), (20, 1), (31, 2)):
            if length + D + 2 > n:
                continue
            for _ in range(5):
                s = rng.randrange(0, n - length - D - 1)
                q = bytearray(ref[s:s + length + D])
                for _k in range(rng.randrange(0, 3)):
                    q[rng.randrange(len(q))] = rng.choice(b"ACGT")
                q = bytes(q)
                for n1 in matching_starts(q, length, D, ref):
                    W = ref[n1 - 1:n1 - 1 + length]
                    found = False
                    for j in range(D + 1):
                        a, b = j * length // (D + 1), (j + 1) * length // (D + 1)
                        for sft in range(-D, D + 1):
                            if a + sft >= 0 and b + sft <= length + D and b > a and q[a + sft:b + sft] == W[a:b]:
                                found = True
                    assert found, (ref, q, length, D, n1)
                    windows += 1
    assert windows > 2000
