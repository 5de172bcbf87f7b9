"""Parity of the bounded-edit FM search (bwt_fs_find_seeds_batch, inexact.cu) with the restatement of seeds.erl in
oracle/: same window starts, same order, for the committed known answers, for random queries on small references
(literal clause list) and for GL000193.1 (reachability DP).  Runs on the B200 box only (``-m gpu``)."""
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import seeds_oracle as S                   # noqa: E402  (the checker, never the product)
from oracle.c_oracle import c_fs_scan_dp, c_fs_scan_patterns   # noqa: E402
from bwt_b200 import api, seeds                        # noqa: E402
from bwt_b200.errors import ReferenceCrash             # noqa: E402
from bwt_b200.synth import random_reference            # noqa: E402

from conftest import GOLDEN                            # noqa: E402


@pytest.fixture(autouse=True, params=["pigeonhole", "search_only"])
def fs_mode(request):
    """Every test runs twice: with the pigeonhole front end (exact piece searches + window verification, Defects <= 2,
    falling back to the depth-first search where it does not apply) and with BWT_FS_PIGEON=0 (the search alone)."""
    if request.param == "search_only":
        os.environ["BWT_FS_PIGEON"] = "0"
    else:
        os.environ.pop("BWT_FS_PIGEON", None)
    yield request.param
    os.environ.pop("BWT_FS_PIGEON", None)


def test_front_end_settles_and_falls_back(gl_ref, fs_mode):
    """unique loci are settled by the front end, a tandem repeat hands the query to the search; same answers"""
    rng = random.Random(9)
    with api.Index.build(gl_ref, shift=1) as ix:
        qs = []
        for _ in range(200):
            s = rng.randrange(0, len(gl_ref) - 110)
            q = bytearray(gl_ref[s:s + 104])
            if rng.random() < 0.5:
                q[rng.randrange(100)] = rng.choice(b"ACGT")
            if rng.random() < 0.3:
                del q[rng.randrange(90)]
            qs.append(bytes(q[:102]))
        st, so, sd, stats = ix.fs_find_seeds(qs, None, 98, 2, 4096)
        assert (st == api.FS_OK).all()
        for got, q in zip(_split(so, sd), qs):
            assert got == c_fs_scan_dp(q, 98, 2, gl_ref), q
        if fs_mode == "pigeonhole":
            assert 100 <= stats.settled <= 200
        else:
            assert stats.settled == 0
    ref = b"ACGGT" * 400 + random_reference(500, 3)
    with api.Index.build(ref, shift=1) as ix:
        q = (b"ACGGT" * 30)[:62]
        st, so, sd, stats = ix.fs_find_seeds([q], None, 60, 2, 1 << 14)
        assert st.tolist() == [api.FS_OK] and stats.settled == 0        # every piece occurs hundreds of times
        assert _split(so, sd)[0] == c_fs_scan_dp(q, 60, 2, ref)


def _split(so, sd):
    return [sd[so[r]:so[r + 1]].tolist() for r in range(len(so) - 1)]


def test_known_answers():
    with open(os.path.join(GOLDEN, "seeds_kat.json")) as f:
        kat = json.load(f)
    for name, text in kat["refs"].items():
        ref = text.encode()
        with api.Index.build(ref, shift=1) as ix:
            groups = {}
            for c in kat["cases"]:
                if c["ref"] == name:
                    groups.setdefault((c["len"], c["defects"]), []).append(c)
            for (length, defects), cs in groups.items():
                st, so, sd, stats = ix.fs_find_seeds([c["q"].encode() for c in cs], None, length, defects, 4096)
                assert (st == api.FS_OK).all()
                for got, c in zip(_split(so, sd), cs):
                    assert got == c["seeds"], (name, c["q"], length, defects)
                assert stats.queries == len(cs) and stats.seeds == len(sd)


def test_random_small_references_vs_clause_list():
    rng = random.Random(11)
    for t in range(12):
        n = rng.choice([40, 300, 2000])
        ref = random_reference(n, 100 + t) if t % 3 else (random_reference(37, t) * (n // 37 + 1))[:n]
        with api.Index.build(ref, shift=1) as ix:
            for length, defects in ((2, 0), (3, 1), (6, 2), (9, 3), (14, 2), (7, 4)):
                if length + defects + 2 > n:
                    continue
                qs = []
                for _ in range(6):
                    if rng.random() < 0.7:
                        s = rng.randrange(0, n - length - defects - 1)
                        q = bytearray(ref[s:s + length + defects + rng.randrange(0, 3)])
                        for _k in range(rng.randrange(0, 3)):
                            q[rng.randrange(len(q))] = rng.choice(b"ACGTN")
                    else:
                        q = bytearray(rng.choice(b"ACGT") for _ in range(length + defects))
                    qs.append(bytes(q))
                st, so, sd, _ = ix.fs_find_seeds(qs, None, length, defects, 1 << 14)
                assert (st == api.FS_OK).all()
                for got, q in zip(_split(so, sd), qs):
                    pats = S.generate_patterns(q, length, defects)
                    want = c_fs_scan_patterns(pats, length, ref)
                    assert got == want, (t, q, length, defects, got[:8], want[:8])


def test_gl000193_vs_dp(gl_ref):
    rng = random.Random(5)
    with api.Index.build(gl_ref, shift=1) as ix:
        for length, defects, nq in ((20, 2, 10), (30, 1, 8), (48, 2, 6), (98, 2, 4), (248, 2, 3), (16, 3, 4)):
            qs = []
            for _ in range(nq):
                s = rng.randrange(0, len(gl_ref) - length - defects - 4)
                q = bytearray(gl_ref[s:s + length + defects + 2])
                for _k in range(rng.randrange(0, 4)):          # up to 3 substitutions: some queries lose their locus
                    q[rng.randrange(len(q))] = rng.choice(b"ACGT")
                if rng.random() < 0.3:
                    del q[rng.randrange(len(q) - 3)]           # a dropped base
                if rng.random() < 0.3:
                    q.insert(rng.randrange(len(q)), rng.choice(b"ACGT"))
                qs.append(bytes(q[:length + defects + 2]))
            st, so, sd, stats = ix.fs_find_seeds(qs, None, length, defects, 1 << 15)
            assert (st == api.FS_OK).all(), st
            assert stats.nodes > 0 and stats.lf_steps > 0
            for got, q in zip(_split(so, sd), qs):
                want = c_fs_scan_dp(q, length, defects, gl_ref)
                assert got == want, (q, length, defects, got[:8], want[:8])


def test_statuses_limits_and_mirror(gl_ref):
    with api.Index.build(gl_ref[:20000], shift=1) as ix:
        q = gl_ref[100:130]
        # short query: the generated module has no usable clause / does not compile
        st, so, sd, _ = ix.fs_find_seeds([q[:21], q[:22], b"", q], None, 20, 2)
        assert st.tolist() == [api.FS_SHORT, api.FS_OK, api.FS_SHORT, api.FS_OK]
        assert so[1] == so[0] and so[3] == so[2]
        # too many matching windows for the caller's bound
        st, so, sd, _ = ix.fs_find_seeds([b"ACGT"], None, 2, 2, 64)
        assert st.tolist() == [api.FS_LIMIT] and len(sd) == 0
        # empty batch, bad parameters
        st, so, sd, _ = ix.fs_find_seeds([], None, 20, 2)
        assert len(st) == 0 and so.tolist() == [0]
        with pytest.raises(api.BwtError) as e:
            ix.fs_find_seeds([q], None, 1, 0)
        assert e.value.code == -8
        with pytest.raises(api.BwtError):
            ix.fs_find_seeds([q], None, 20, 9)
        # Erlang-shaped mirror
        fs = seeds.generate_fs(q, 20, 2)
        assert fs.find_seeds(ix) == c_fs_scan_dp(q, 20, 2, gl_ref[:20000])
        with pytest.raises(ReferenceCrash):
            seeds.generate_fs(q[:21], 20, 2).find_seeds(ix)
        with pytest.raises(ReferenceCrash):
            seeds.generate_fs(q, 1, 0)


def test_stack_spill_and_many_hits():
    # a periodic text keeps every branch of the search alive: deep stacks (HBM spill) and wide intervals
    ref = b"AC" * 300 + b"G" + b"CA" * 50 + b"TTTTTTTTTT"
    with api.Index.build(ref, shift=1) as ix:
        # (a) the pattern's left end cannot match: the whole tree is walked and dies at the very end
        q = b"GGGGGGGGGG" + b"AC" * 40
        st, so, sd, stats = ix.fs_find_seeds([q], None, 60, 4, 1 << 16)
        assert st.tolist() == [api.FS_OK] and len(sd) == 0
        print("dead-end walk: nodes %d lf %d spills %d" % (stats.nodes, stats.lf_steps, stats.spills))
        assert stats.nodes > 10000                             # Defects = 4: the front end does not apply
        os.environ["BWT_FS_STACK"] = "32"                       # tiny shared-memory stack: the HBM spill path must run
        try:
            st2, so2, sd2, stats2 = ix.fs_find_seeds([q, b"AC" * 40], None, 60, 4, 1 << 16)
        finally:
            del os.environ["BWT_FS_STACK"]
        print("tiny stack: nodes %d spills %d" % (stats2.nodes, stats2.spills))
        assert st2.tolist() == [api.FS_OK, api.FS_OK] and stats2.spills > 0
        assert _split(so2, sd2) == [[], c_fs_scan_dp(b"AC" * 40, 60, 4, ref)]
        # (b) wide intervals: thousands of matching windows before the walk
        for length, defects in ((12, 2), (9, 3)):
            qs = [(b"AC" * 20)[:length + defects], (b"CA" * 20)[:length + defects], b"T" * (length + defects)]
            st, so, sd, stats = ix.fs_find_seeds(qs, None, length, defects, 1 << 18)
            assert (st == api.FS_OK).all(), st
            print("wide: windows %d seeds %d spills %d" % (stats.raw_hits, stats.seeds, stats.spills))
            for got, qq in zip(_split(so, sd), qs):
                assert got == c_fs_scan_dp(qq, length, defects, ref)
