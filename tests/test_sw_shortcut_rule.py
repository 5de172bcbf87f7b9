"""The decision rule behind the no-DP shortcut of sw_classify_kernel (bwt_b200/csrc/sw.cu), restated in Python and held
against the oracle's sw:sw/2 on strings built to break it (low complexity, tandem repeats, indels next to mismatches).

The rule: run a wavefront search with PLAIN affine gaps (mismatch 3, window-side gap 2 + 1/extra, query-side gap 4 + 3/extra,
relative to 2 per query character); if the leftmost diagonal on which the cheapest penalty p* reaches the last row is a
gapless diagonal of the window (its mismatch count is p*/3), the reference's result is {2*Lq - p*, "<Lq>M"}.
This file is CPU-only and checks the RULE; tests/test_gpu_parity.py checks the kernel (with the shortcut on and off)."""
import random

from oracle.c_oracle import c_sw

ACGT = set(b"ACGT")
NEG = -10 ** 6


def wfa_shortcut(q: bytes, r: bytes, pmax: int = 15, maxd: int = 16):
    lq, lr = len(q), len(r)
    dd = lr - lq
    if dd < 0 or dd > maxd or lq < 4:
        return None
    pmax = min(pmax, 2 * (lq // 2))               # the row threshold -2*(Lq div 2) + 2i holds on the diagonal up to here
    # (the kernel also narrows the budget by read length -- 9 up to 128 bp, 12 up to 192, 15 beyond: any budget is sound)
    if any(c not in ACGT for c in q) or any(c not in ACGT for c in r):
        return None
    k = max(0, (pmax - 1) // 3)
    lo, hi = -k, dd + k

    def ext(o, i):
        while i < lq and 0 <= o + i < lr and q[i] == r[o + i]:
            i += 1
        return i
    M, I, D = {}, {}, {}
    for s in range(0, pmax + 1):
        reached = []
        for o in range(lo, hi + 1):
            d = max(M.get((s - 2, o - 1), NEG), D.get((s - 1, o - 1), NEG))      # skip a window character
            if d > NEG and not (0 <= o + d <= lr):
                d = NEG
            iv = max(M.get((s - 4, o + 1), NEG), I.get((s - 3, o + 1), NEG))     # skip a query character
            if iv > NEG:
                iv += 1
                if iv > lq or o + iv < 0:
                    iv = NEG
            m = NEG
            if s == 0:
                if 0 <= o <= lr:
                    m = 0                                                        # row 0 is free
            else:
                pm = M.get((s - 3, o), NEG)
                if pm > NEG and pm < lq and 0 <= o + pm < lr:
                    m = pm + 1                                                   # mismatch
                m = max(m, d, iv)
            if m > NEG:
                m = ext(o, m)
                M[(s, o)] = m
                if m == lq and 0 <= o + lq <= lr:
                    reached.append(o)
            if d > NEG:
                D[(s, o)] = d
            if iv > NEG:
                I[(s, o)] = iv
        if reached:
            omin = min(reached)
            if omin < 0 or omin > dd:
                return None
            mm = sum(1 for t in range(lq) if q[t] != r[omin + t])
            return (2 * lq - s, "%dM" % lq) if 3 * mm == s else None
    return None


def _cases(rng, n):
    def rs(k, alpha):
        return bytes(rng.choice(alpha) for _ in range(k))
    for _ in range(n):
        alpha = rng.choice([b"ACGT", b"AC", b"ACGT", b"AAC", b"ACGT"])
        lq, dd = rng.randint(4, 48), rng.randint(0, 7)
        mode = rng.random()
        if mode < 0.55:
            base = rs(lq + dd + 4, alpha)
            s0 = rng.randint(0, dd + 2)
            q = bytearray(base[s0:s0 + lq])
            q += rs(lq - len(q), alpha)
            for _k in range(rng.choice([0, 1, 1, 2, 2, 2, 3, 3, 4, 5, 5, 6])):
                u = rng.random()
                if u < 0.7:
                    q[rng.randrange(lq)] = rng.choice(alpha)
                elif u < 0.85 and len(q) > 5:
                    del q[rng.randrange(len(q))]
                    q.append(rng.choice(alpha))
                else:
                    q.insert(rng.randrange(len(q)), rng.choice(alpha))
            yield bytes(q[:lq]), base[:lq + dd]
        elif mode < 0.85:
            unit = rs(rng.randint(1, 4), alpha)
            base = (unit * 80)[:lq + dd]
            r = bytearray(base)
            for _k in range(rng.randint(0, 3)):
                r[rng.randrange(len(r))] = rng.choice(b"ACGT")
            s0 = rng.randint(0, dd)
            q = bytearray(base[s0:s0 + lq])
            for _k in range(rng.randint(0, 5)):
                q[rng.randrange(lq)] = rng.choice(b"ACGT")
            yield bytes(q), bytes(r)
        else:
            yield rs(lq, alpha), rs(lq + dd, alpha)


def test_rule_agrees_with_sw_on_adversarial_pairs():
    rng = random.Random(20141)
    settled = {}
    for q, r in _cases(rng, 25000):
        sc = wfa_shortcut(q, r)
        if sc is None:
            continue
        assert c_sw(q, r) == sc, (q, r, sc)
        m = (2 * len(q) - sc[0]) // 3
        settled[m] = settled.get(m, 0) + 1
    assert all(settled.get(m, 0) > 200 for m in (0, 1, 2, 3, 4, 5)), settled


def test_rule_declines_what_it_cannot_prove():
    q = b"ACGTACGTACGTTTGACCA"
    assert wfa_shortcut(q, q) == (2 * len(q), "%dM" % len(q))
    assert wfa_shortcut(q, q[:7] + q[8:] + b"A") is None           # the best path has a gap
    assert wfa_shortcut(q, b"ACGN" + q[4:]) is None                # 'N' scores 1.6: the double path decides
    assert wfa_shortcut(q, q[:-1]) is None                         # window shorter than the read
    assert wfa_shortcut(b"ACG", b"ACG") is None                    # below the length the argument covers
