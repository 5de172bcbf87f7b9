"""seeds.erl restatement (oracle/seeds_oracle.py + .c): the literal clause list, the numpy match vector, the C scan of
the clause list and the C reachability DP are four formulations of fs:matches/Len + fs:find_seeds/1; they must agree
with each other and with the committed known answers (tests/golden/seeds_kat.json, written by
scripts/make_seeds_golden.py).  CPU only."""
import json
import os
import random

import pytest

from oracle import seeds_oracle as S
from oracle.c_oracle import c_fs_scan_dp, c_fs_scan_patterns

from conftest import GOLDEN


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "seeds_kat.json")) as f:
        return json.load(f)


def test_gen_pattern_small_known_answers():
    # seeds.erl:51-62 by hand.  Defects = 0: the first Len characters.
    assert S.generate_patterns(b"ACGTAC", 4, 0) == [b"ACGT"]
    # Len 2, one defect over "ACG": order = mutation, insertion, deletion, match of the first character, recursively
    assert S.generate_patterns(b"ACG", 2, 1) == [b"*C", b"*A", b"CG", b"A*", b"A*", b"AG"]
    # every pattern has exactly Len elements and at most `Defects` wildcards
    pats = S.generate_patterns(b"ACGTACGTAC", 7, 3)
    assert all(len(p) == 7 and p.count(b"*") <= 3 for p in pats)


def test_short_query_is_compile_error_or_empty():
    for length in range(2, 8):
        for defects in range(0, 4):
            for lq in range(0, length + defects):
                try:
                    assert S.generate_patterns(b"ACGTACGTACGT"[:lq], length, defects) == []
                except S.FsCompileError:
                    pass
    with pytest.raises(S.FsCompileError):
        S.generate_patterns(b"ACGT", 1, 0)


def test_known_answers_all_formulations(kat):
    refs = {k: v.encode() for k, v in kat["refs"].items()}
    for c in kat["cases"]:
        ref, q = refs[c["ref"]], c["q"].encode()
        pats = S.generate_patterns(q, c["len"], c["defects"])
        assert len(pats) == c["n_patterns"]
        assert S.fs_find_seeds(pats, c["len"], ref) == c["seeds"]
        assert c_fs_scan_patterns(pats, c["len"], ref) == c["seeds"]
        assert c_fs_scan_dp(q, c["len"], c["defects"], ref) == c["seeds"]
        assert S.greedy_scan(S.match_vector(pats, c["len"], ref), c["len"]) == c["seeds"]


def test_random_cross_check():
    rng = random.Random(7)
    for _ in range(150):
        length, defects = rng.randint(2, 9), rng.randint(0, 3)
        n = rng.randint(0, 300)
        ref = bytes(rng.choice(b"ACGT") for _ in range(n))
        if rng.random() < 0.6 and n > length + defects + 1:
            s = rng.randint(0, n - length - defects)
            q = bytearray(ref[s:s + length + defects + 2])
            for _k in range(rng.randint(0, 3)):
                q[rng.randrange(len(q))] = rng.choice(b"ACGT")
            q = bytes(q)
        else:
            q = bytes(rng.choice(b"ACGTN") for _ in range(length + defects + rng.randint(0, 3)))
        pats = S.generate_patterns(q, length, defects)
        a = S.fs_find_seeds(pats, length, ref)
        assert a == c_fs_scan_patterns(pats, length, ref)
        got, vec = c_fs_scan_dp(q, length, defects, ref, True)
        assert a == got
        mv = S.match_vector(pats, length, ref)
        assert (mv == vec.astype(bool)).all()
        for w in range(0, max(0, n - length + 1), 5):
            assert S.fs_matches_dp(q, length, defects, ref[w:w + length]) == bool(mv[w])


def test_walk_semantics():
    # a match jumps Len ahead (seeds.erl:38), so overlapping occurrences are skipped; the last full window is tested
    pats = S.generate_patterns(b"AAAA", 4, 0)
    assert S.fs_find_seeds(pats, 4, b"AAAAAAAAAA") == [5, 1]          # windows 1 and 5; 9 would need 12 characters
    assert S.fs_find_seeds(pats, 4, b"CAAAA") == [2]                  # seeds.erl:26-30: exactly Len left
    assert S.fs_find_seeds(pats, 4, b"AAA") == []
    assert c_fs_scan_dp(b"AAAA", 4, 0, b"AAAAAAAAAA") == [5, 1]


def test_kats_equal_the_interpreted_reference():
    """tests/golden/seeds_kat_interpreted.json: the UNMODIFIED src/seeds.erl run by oracle/mini_erl.py (generate_fs/3 writes
    fs.erl, the evaluator loads that text and runs fs:find_seeds/1) on the KAT inputs.  The oracle's answers equal it."""
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "seeds_kat.json")) as f:
        kat = json.load(f)
    with open(os.path.join(here, "golden", "seeds_kat_interpreted.json")) as f:
        interp = json.load(f)
    assert len(interp["cases"]) == len(kat["cases"]) == 60
    for c, r in zip(kat["cases"], interp["cases"]):
        assert (c["ref"], c["q"], c["len"], c["defects"]) == (r["ref"], r["q"], r["len"], r["defects"])
        assert c["seeds"] == r["seeds"], (c["ref"], c["q"], c["len"], c["defects"])
        assert c["n_patterns"] >= 1 and r["true_clauses"] >= 1
