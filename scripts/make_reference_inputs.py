#!/usr/bin/env python
"""Writes tests/golden/reference_inputs.txt: the fixed input set scripts/make_reference_golden.escript feeds to the
UNMODIFIED reference (bwa:find_seeds/6, sga:sga/6 + the worker's fold, sw:sw/2) on its shipped GL000193.1 index.
Seeded; inputs only -- no oracle is involved in producing them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bwt_b200.synth import sample_reads          # noqa: E402  (host-side sampler, numpy only)

ref = open(os.path.join(ROOT, "tests", "golden", "GL000193.1_p1.ref"), "rb").read()
rng = np.random.default_rng(20261018)
lines = []

# -- subsequences for bwa:find_seeds/6: 15-mers and shorter from the reference, mutated ones, with 'N', a crash case
for _ in range(400):
    ln = int(rng.choice([8, 9, 12, 14, 15, 15, 15, 20, 31]))
    p = int(rng.integers(0, len(ref) - ln))
    s = bytearray(ref[p:p + ln])
    u = rng.random()
    if u < 0.3:
        s[int(rng.integers(0, ln))] = b"ACGT"[int(rng.integers(0, 4))]
    elif u < 0.4:
        s[int(rng.integers(0, ln))] = ord("N")
    lines.append(b"S " + bytes(s))
lines += [b"S ACGTACG", b"S A", b"S " + ref[36222:36237], b"S " + ref[67864:67879], b"S " + ref[118831:118846],
          b"S " + ref[1000:1014] + b"N", b"S " + ref[1000:1014] + b"X"]

# -- reads: config-1 style (100 bp, 1 % edits), 2 % edits, with 'N', 150 / 250 bp, from the tandem-repeat region
#    (79-102 kb: a dozen equal-scoring loci -> the top-1 pick of lists:sort/2 with a strict comparator), chunk ends
def add_reads(n, ln, seed, edit, n_rate=0.0, lo=None, hi=None):
    sub = ref if lo is None else ref[lo:hi]
    reads, offs, _ = sample_reads(sub, n, ln, seed=seed, edit_rate=edit, n_rate=n_rate)
    for k in range(n):
        r = reads[offs[k]:offs[k + 1]].tobytes()
        if r:
            lines.append(b"R " + r)

add_reads(300, 100, 1, 0.01)
add_reads(100, 100, 2, 0.02)
add_reads(60, 100, 3, 0.01, n_rate=0.01)
add_reads(40, 150, 4, 0.01)
add_reads(30, 250, 5, 0.01)
add_reads(120, 100, 6, 0.0, lo=79000, hi=102000)
add_reads(60, 100, 7, 0.01, lo=79000, hi=102000)
for p in list(range(0, 30, 5)) + list(range(len(ref) - 130, len(ref) - 100, 5)):
    lines.append(b"R " + ref[p:p + 100])

# -- pairs for sw:sw/2: plain, IUPAC codes and 'N' in the window, 'N' in the query, row-threshold aborts, ties
iupac = b"ACGTNRYKMSWBDV"
def rnd(n, alpha):
    return bytes(alpha[int(x)] for x in rng.integers(0, len(alpha), n))
for it in range(400):
    lq = int(rng.integers(4, 140))
    q = bytearray(rnd(lq, b"ACGT"))
    w = bytearray(q)
    for _ in range(int(rng.integers(0, 5))):
        p = int(rng.integers(0, len(w)))
        u = rng.random()
        if u < 0.5:
            w[p] = (iupac if it % 3 == 0 else b"ACGT")[int(rng.integers(0, 14 if it % 3 == 0 else 4))]
        elif u < 0.75 and len(w) > 2:
            del w[p]
        else:
            w.insert(p, b"ACGT"[int(rng.integers(0, 4))])
    if it % 7 == 0:
        q[int(rng.integers(0, lq))] = ord("N")
    w = rnd(int(rng.integers(0, 6)), b"ACGT") + bytes(w) + rnd(int(rng.integers(0, 6)), b"ACGT")
    lines.append(b"P " + bytes(q) + b" " + w)
for it in range(60):                                     # low complexity: many equal-scoring paths
    unit = rnd(int(rng.integers(1, 4)), b"AC")
    lq = int(rng.integers(10, 80))
    lines.append(b"P " + (unit * 100)[:lq] + b" " + (unit * 100)[:lq + int(rng.integers(0, 9))])
for it in range(40):                                     # unrelated strings
    lines.append(b"P " + rnd(int(rng.integers(5, 90)), b"ACGT") + b" " + rnd(int(rng.integers(5, 120)), b"ACGT"))

# -- appended in round 2 (indices of everything above stay as they were): pairs aimed at the tie rules of sw:sw/2 -- low-complexity
#    and tandem-repeat strings (competing diagonals, deletions between matching flanks), slack 0..20, substitutions and
#    indels next to each other; the generator of tests/test_gpu_parity.py::test_sw_shortcut_adversarial, shorter strings
rng2 = np.random.default_rng(20261019)
for it in range(600):
    alpha = [b"ACGT", b"AC", b"ACGT", b"AAC", b"ACGT"][it % 5]
    lq, dd = int(rng2.integers(4, 70)), int(rng2.integers(0, 21))
    pick = lambda n: bytes(alpha[int(x)] for x in rng2.integers(0, len(alpha), n))   # noqa: E731
    mode = rng2.random()
    if mode < 0.55:
        base = pick(lq + dd + 4)
        s0 = int(rng2.integers(0, dd + 3))
        q = bytearray(base[s0:s0 + lq]) + bytearray(pick(max(0, lq - len(base[s0:s0 + lq]))))
        for _ in range(int(rng2.choice([0, 0, 1, 1, 1, 2, 3, 4, 5, 6]))):
            u = rng2.random()
            if u < 0.7:
                q[int(rng2.integers(0, lq))] = alpha[int(rng2.integers(0, len(alpha)))]
            elif u < 0.85 and len(q) > 5:
                del q[int(rng2.integers(0, len(q)))]
                q += pick(1)
            else:
                q.insert(int(rng2.integers(0, len(q))), alpha[int(rng2.integers(0, len(alpha)))])
        lines.append(b"P " + bytes(q[:lq]) + b" " + base[:lq + dd])
    elif mode < 0.85:
        unit = pick(int(rng2.integers(1, 5)))
        base = (unit * 400)[:lq + dd]
        r = bytearray(base)
        for _ in range(int(rng2.integers(0, 3))):
            r[int(rng2.integers(0, len(r)))] = b"ACGT"[int(rng2.integers(0, 4))]
        s0 = int(rng2.integers(0, dd + 1))
        q = bytearray(base[s0:s0 + lq])
        for _ in range(int(rng2.integers(0, 6))):
            q[int(rng2.integers(0, lq))] = b"ACGT"[int(rng2.integers(0, 4))]
        lines.append(b"P " + bytes(q) + b" " + bytes(r))
    else:
        lines.append(b"P " + pick(lq) + b" " + pick(lq + dd))

# -- pairs aimed at the banded table of sw_band_kernel (swb_core.cuh): the best alignment has a gap.  Indels of one to
#    three characters in small alphabets and tandem repeats (an indel can sit anywhere in a run: co-optimal paths, the
#    reference's tie order decides; longer gaps are what the certificate must decline), indels next to mismatches and next
#    to the ends, and config-1 style 100 bp reads with one or two indels against their own window
rng3 = np.random.default_rng(20261020)
for it in range(900):
    alpha = [b"ACGT", b"AC", b"ACGT", b"AAC", b"ACGT", b"AACCGT"][it % 6]
    lq, dd = int(rng3.integers(8, 64)), int(rng3.integers(0, 6))
    pick = lambda n: bytes(alpha[int(x)] for x in rng3.integers(0, len(alpha), n))   # noqa: E731
    if it % 3 != 2:
        base = pick(lq + dd + 6)
        s0 = int(rng3.integers(0, dd + 3))
        q = bytearray(base[s0:s0 + lq + 3])
    else:
        unit = pick(int(rng3.integers(1, 6)))
        base = bytearray((unit * 80)[:lq + dd + 4])
        for _ in range(int(rng3.integers(0, 3))):
            base[int(rng3.integers(0, len(base)))] = b"ACGT"[int(rng3.integers(0, 4))]
        base = bytes(base)
        s0 = int(rng3.integers(0, dd + 1))
        q = bytearray(base[s0:s0 + lq + 2])
    for _ in range(int(rng3.choice([1, 1, 1, 2, 2, 3]))):
        u = rng3.random()
        at = int(rng3.integers(0, len(q)))
        if u < 0.35:
            q[at] = alpha[int(rng3.integers(0, len(alpha)))]
        elif u < 0.7 and len(q) > 6:
            del q[at:at + int(rng3.choice([1, 1, 1, 2, 3]))]
        else:
            q[at:at] = pick(int(rng3.choice([1, 1, 1, 2])))
    q = bytes(q[:lq])
    if len(q) >= 4:
        lines.append(b"P " + q + b" " + base[:len(q) + dd])
for it in range(240):
    p0 = int(rng3.integers(1000, len(ref) - 400))
    src, pos, q, nind = ref, p0, bytearray(), 0
    while pos - p0 < 100:
        u = rng3.random()
        if u < 0.012:
            v = rng3.random()
            if v < 0.4:
                q.append(b"ACGT"[int(rng3.integers(0, 4))]); pos += 1
            elif v < 0.7:
                q.append(b"ACGT"[int(rng3.integers(0, 4))]); nind += 1
            else:
                pos += 1; nind += 1
        else:
            q.append(src[pos]); pos += 1
    if nind == 0:
        q.insert(int(rng3.integers(5, len(q) - 5)), b"ACGT"[int(rng3.integers(0, 4))]); nind = 1
    q = bytes(q)
    w = src[pos - (len(q) + nind):pos]
    if all(c in b"ACGT" for c in w):
        lines.append(b"P " + q + b" " + w)

out = os.path.join(ROOT, "tests", "golden", "reference_inputs.txt")
with open(out, "wb") as f:
    f.write(b"\n".join(lines) + b"\n")
print("%d lines -> %s" % (len(lines), out))
