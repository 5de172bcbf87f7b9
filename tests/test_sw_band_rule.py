"""The banded table with a certificate (bwt_b200/csrc/swb_core.cuh, run on the GPU by sw_band_kernel in swb.cu) held
against the oracle's sw:sw/2 on the CPU: tests/swb_host.cpp compiles the SAME header with g++, so the row update, the end
test (budget, taint bit, value guard) and the walk back are the very code the kernel executes; only the kernel's loads of
the packed strings are GPU-only (tests/test_gpu_parity.py covers those, in four routings).

The strings are built to break the certificate: alphabets of two or three letters, tandem repeats (an indel can sit
anywhere in a run: co-optimal paths, the reference's tie order decides), gaps of two and three characters (tainted:
must be declined or, if certified, still equal the oracle), indels next to mismatches and next to the ends.
Development runs of this generator: 5.9 M pairs through this harness (3.6 M certified, 1.8 M of them with a gap in the
CIGAR, every certified pair also at the wider bands) and 3.2 M through a Python restatement of the rule: 0 differences.
A prototype without the taint test differed on 0.3 % of the pairs it certified -- the generator reaches the cases the
bit is for.
"""
import ctypes as C
import os
import random
import subprocess

import pytest

from oracle.c_oracle import c_sw

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def host():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libswb_host.so")
    src = [os.path.join(HERE, "swb_host.cpp"), os.path.join(ROOT, "bwt_b200", "csrc", "swb_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "bwt_b200", "csrc"),
                               src[0], "-o", so])
    lib = C.CDLL(so)
    lib.swb_host.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_char_p, C.c_int]

    def run(q, r, wb=0, p_short=9):
        sc = C.c_int()
        buf = C.create_string_buffer(1024)
        rc = lib.swb_host(q, len(q), r, len(r), wb, p_short, C.byref(sc), buf, 1024)
        assert rc >= 0
        return (sc.value, buf.value.decode()) if rc == 1 else None
    return run


def cases(rng, n, lmax=48):
    def rs(k, alpha):
        return bytes(rng.choice(alpha) for _ in range(k))
    for _ in range(n):
        alpha = rng.choice([b"ACGT", b"AC", b"ACGT", b"AAC", b"ACGT", b"AACCGT"])
        lq, dd = rng.randint(4, lmax), rng.randint(0, 7)
        mode = rng.random()
        if mode < 0.6:
            base = rs(lq + dd + 6, alpha)
            s0 = rng.randint(0, dd + 2)
            q = bytearray(base[s0:s0 + lq + 3])
            for _k in range(rng.choice([1, 1, 1, 2, 2, 2, 3, 3, 4])):
                u = rng.random()
                if u < 0.4:
                    q[rng.randrange(len(q))] = rng.choice(alpha)
                elif u < 0.7 and len(q) > 5:
                    at = rng.randrange(len(q))
                    del q[at:at + rng.choice([1, 1, 1, 2, 3])]
                else:
                    at = rng.randrange(len(q))
                    q[at:at] = rs(rng.choice([1, 1, 1, 2]), alpha)
            q = bytes(q[:lq])
            if len(q) < 4:
                continue
            yield q, base[:len(q) + dd]
        elif mode < 0.9:
            unit = rs(rng.randint(1, 5), alpha)
            base = (unit * 80)[:lq + dd + 4]
            r = bytearray(base)
            for _k in range(rng.randint(0, 3)):
                r[rng.randrange(len(r))] = rng.choice(b"ACGT")
            s0 = rng.randint(0, dd)
            q = bytearray(base[s0:s0 + lq + 2])
            for _k in range(rng.randint(0, 3)):
                u = rng.random()
                if u < 0.5:
                    q[rng.randrange(len(q))] = rng.choice(b"ACGT")
                elif u < 0.75:
                    del q[rng.randrange(len(q))]
                else:
                    q.insert(rng.randrange(len(q)), rng.choice(b"ACGT"))
            q = bytes(q[:lq])
            yield q, bytes(r[:len(q) + dd])
        else:
            yield rs(lq, alpha), rs(lq + dd, alpha)


@pytest.mark.parametrize("seed,n,lmax,p_short", [(9001, 60000, 48, 9), (9002, 25000, 130, 9), (9003, 6000, 256, 9),
                                                 (9004, 40000, 48, 12), (9005, 20000, 130, 12), (9006, 30000, 60, 15)])
def test_band_agrees_with_sw_on_adversarial_pairs(host, seed, n, lmax, p_short):
    """p_short = the penalty budget of reads up to 128 bp (BWT_SW_BAND_P of the kernel): every budget must be exact"""
    rng = random.Random(seed)
    settled = gapped = 0
    for q, r in cases(rng, n, lmax):
        got = host(q, r, 0, p_short)
        if got is None:
            continue
        want = c_sw(q, r)
        assert want != "no_match" and (want[0], want[1]) == got, (q, r, got, want)
        settled += 1
        gapped += "I" in got[1] or "D" in got[1]
    assert settled > 0.45 * n and gapped > 0.2 * n, (settled, gapped)


def test_wider_bands_say_the_same(host):
    """the kernel runs a warp at the widest band any of its pairs needs: a pair must come out the same at 6, 8, 12, 16"""
    rng = random.Random(9100)
    seen = 0
    for q, r in cases(rng, 20000, 60):
        base = host(q, r)
        for wb in (8, 12, 16):
            other = host(q, r, wb)
            if base is not None and other is not None:
                assert other == base, (q, r, wb, base, other)
                seen += 1
            elif base is not None:
                # a forced width below the pair's need is refused; anything wider must certify what the narrow one did
                assert other is None and wb < (len(r) - len(q)) + 2 * 4 + 1, (q, r, wb)
    assert seen > 20000


def test_band_declines_what_it_cannot_prove(host):
    q = b"ACGTACGTACGTTTGACCATTGACA"
    assert host(q, q) == (2 * len(q), "%dM" % len(q))
    assert host(q, q[:9] + b"G" + q[9:]) == (2 * len(q) - 2, "9M1D16M")        # one window character skipped
    assert host(q[:9] + b"G" + q[9:], q + b"A")[1] == "9M1I16M"                   # one query character skipped
    assert host(q, q[:9] + b"GG" + q[9:]) is None                                  # the best path extends a gap
    assert host(q, b"ACGN" + q[4:]) is None                                        # 'N' scores 1.6: the double path decides
    assert host(q, q[:-1]) is None                                                 # window shorter than the read
    assert host(b"ACG", b"ACG") is None                                            # below the length the argument covers
    assert host(b"AAAAAAAAAAAAAAAAAAAAAAAA", b"CCCCCCCCCCCCCCCCCCCCCCCCC") is None  # penalty beyond the budget
    # an indel inside a homopolymer run: every placement scores the same, the reference's tie order picks one
    q2, r2 = b"ACGTTTTTTGACCATGCA", b"ACGTTTTTTTGACCATGCA"
    assert host(q2, r2) == tuple(c_sw(q2, r2)[:2])


def test_band_equals_the_interpreted_reference(host):
    """every sw:sw/2 pair of tests/golden/reference_inputs.txt (2,240; 1,140 of them built for the banded table) against
    what the reference's UNMODIFIED sw.erl returns for it when oracle/mini_erl.py evaluates the source text
    (tests/golden/reference_outputs_interpreted.json): whatever the band certifies, at any budget, is that result"""
    import json
    with open(os.path.join(HERE, "golden", "reference_outputs_interpreted.json")) as f:
        items = [it for it in json.load(f)["items"] if it["kind"] == "P"]
    assert len(items) >= 2200
    certified = gapped = 0
    for it in items:
        q, r = it["in"][0].encode(), it["in"][1].encode()
        for p_short in (9, 12, 15):
            got = host(q, r, 0, p_short)
            if got is None:
                continue
            assert it["out"] != "no_match" and tuple(it["out"]) == got, (q, r, p_short, got, it["out"])
            if p_short == 9:
                certified += 1
                gapped += "I" in got[1] or "D" in got[1]
    assert certified > 1200 and gapped > 600, (certified, gapped)


def test_band_exhaustive_on_short_strings(host):
    """EVERY pair over {A,C} with Lq = 4..6, Dd = 0..2 and over {A,C,G} with Lq = 4, Dd = 0..1 (every tie the rule can
    meet at these sizes), at the smallest and the largest budget: what the band certifies is the oracle's result.
    (Development run of the same loop: Lq up to 8 over {A,C}, Lq = 4..5 over {A,C,G}, Lq = 4 over ACGT with Dd up to 3:
    7.2 M pairs, 2.1 M certified, 0 differences.)"""
    import itertools
    certified = 0
    for alpha, lqs, dds in ((b"AC", (4, 5, 6), (0, 1, 2)), (b"ACG", (4,), (0, 1))):
        for lq in lqs:
            for dd in dds:
                windows = [bytes(t) for t in itertools.product(alpha, repeat=lq + dd)]
                for qt in itertools.product(alpha, repeat=lq):
                    q = bytes(qt)
                    for r in windows:
                        want = None
                        for p_short in (9, 15):
                            got = host(q, r, 0, p_short)
                            if got is None:
                                continue
                            if want is None:
                                want = c_sw(q, r)
                            assert want != "no_match" and (want[0], want[1]) == got, (q, r, p_short, got, want)
                            certified += p_short == 9
    assert certified > 30000, certified
