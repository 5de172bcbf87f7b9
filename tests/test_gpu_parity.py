"""Parity of the CUDA path (through the C ABI, libbwt_b200.so) with the oracle, on the same seeded inputs.

Bit-exact everywhere: SA intervals, seed clusters, scores, CIGAR strings, positions, statuses.
Runs on the B200 box only (``-m gpu``).  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bwt_oracle as O                      # noqa: E402  (the checker, never the product)
from oracle.c_oracle import COracle, c_sw               # noqa: E402
from bwt_b200 import api                                # noqa: E402
from bwt_b200.synth import random_reference, sample_reads, sample_reads_fast  # noqa: E402

from test_oracle_golden import (KAT_2_CIGARS, KAT_2_SGA, READ_KAT_1, READ_KAT_2, SEARCH_KATS, SHIPPED_FM_SHA256,  # noqa: E402
                                _rand_query, _rand_seq)


@pytest.fixture(autouse=True, params=["shortcut", "dp_only", "small", "band_only"])
def sw_mode(request):
    """Every test of this file runs four times: with the no-DP shortcut of sw_classify_kernel / sw_wfa_kernel, the banded
    table of sw_band_kernel and the packed-integer full-table tiers behind them (the default routing); with
    BWT_SW_SHORTCUT=0 BWT_SW_BAND=0, which sends every pair through the full-table kernels; in the routing of
    worker-sized batches (BWT_SW_SMALL: every pair that needs the table takes the warp-per-pair kernel in doubles, the
    fast tiers are never launched), forced here for batches of any size; and with BWT_SW_SHORTCUT=0 alone, which hands
    every plain pair -- gapless ones too -- to the banded table first."""
    os.environ.pop("BWT_SW_SHORTCUT", None)
    os.environ.pop("BWT_SW_BAND", None)
    os.environ["BWT_SW_SMALL"] = "0"
    if request.param == "dp_only":
        os.environ["BWT_SW_SHORTCUT"] = "0"
        os.environ["BWT_SW_BAND"] = "0"
    elif request.param == "small":
        os.environ["BWT_SW_SMALL"] = "2000000000"
    elif request.param == "band_only":
        os.environ["BWT_SW_SHORTCUT"] = "0"
    yield request.param
    os.environ.pop("BWT_SW_SHORTCUT", None)
    os.environ.pop("BWT_SW_BAND", None)
    os.environ.pop("BWT_SW_SMALL", None)


@pytest.fixture(scope="module")
def gl_index(gl_fm, gl_ref):
    ix = api.Index.from_arrays(gl_fm.L, gl_fm.SA, gl_fm.pointers, gl_fm.shift, gl_ref)
    yield ix
    ix.close()


# ---------------------------------------------------------------------------------------------------------
# index: loader, builder, writer
# ---------------------------------------------------------------------------------------------------------
def test_open_reference_format(tmp_path, gl_fm, gl_ref):
    fm_path, ref_path = str(tmp_path / "GL000193.1_p1.fm"), str(tmp_path / "GL000193.1_p1.ref")
    with open(fm_path, "wb") as f:
        f.write(O.write_fm(gl_fm))          # byte-identical to the shipped file (test_oracle_golden.py)
    with open(ref_path, "wb") as f:
        f.write(gl_ref)
    with api.Index.open(fm_path, ref_path) as ix:
        i = ix.info()
        assert tuple(i.pointers) == (53511, 94127, 134735, 189790) and i.shift == 1
        assert i.n_rows == 189790 and i.ref_len == 189789 and i.dollar_row == 97538
        L, SA, ref = ix.export()
        assert (L == gl_fm.L).all() and (SA == gl_fm.SA).all() and ref.tobytes() == gl_ref


def test_open_rejects_garbage(tmp_path, gl_ref):
    p = str(tmp_path / "bad.fm")
    with open(p, "wb") as f:
        f.write(b"\x83\x68\x02not an index")
    r = str(tmp_path / "bad.ref")
    with open(r, "wb") as f:
        f.write(gl_ref[:100])
    with pytest.raises(api.BwtError) as e:
        api.Index.open(p, r)
    assert e.value.code == -3
    with pytest.raises(api.BwtError) as e:
        api.Index.open(str(tmp_path / "missing.fm"), r)
    assert e.value.code == -2


def test_device_builder_reproduces_shipped_index(tmp_path, gl_fm, gl_ref):
    import hashlib
    with api.Index.build(gl_ref, shift=1) as ix:
        i = ix.info()
        assert tuple(i.pointers) == gl_fm.pointers and i.dollar_row == 97538
        L, SA, _ = ix.export()
        assert (SA == gl_fm.SA).all()
        assert (L == gl_fm.L).all()
        fm_path, ref_path = str(tmp_path / "x_p1.fm"), str(tmp_path / "x_p1.ref")
        ix.save(fm_path, ref_path)
        with open(fm_path, "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == SHIPPED_FM_SHA256
        with open(ref_path, "rb") as f:
            assert f.read() == gl_ref


def test_device_builder_small_and_repetitive():
    for text in (b"TGTGTGATGATGACCCCCCATGA", b"AACCATGG", b"ACGT" * 700 + b"A", b"A" * 300 + b"CGT" + b"A" * 300,
                 random_reference(5000, 3), (random_reference(300, 4) * 40)):
        fm = O.build_fm(text, 1)
        with api.Index.build(text, shift=1) as ix:
            L, SA, ref = ix.export()
            assert (SA == fm.SA).all(), text[:30]
            assert (L == fm.L).all()
            assert tuple(ix.info().pointers) == fm.pointers


def test_builder_rejects_n():
    with pytest.raises(api.BwtError):
        api.Index.build(b"ACGTNACGTACGT", shift=1)


def test_builder_iupac_mapping():
    text = b"ACGTRYKMSWBDVACGTTGCA" * 5
    fm = O.build_fm(text, 7)
    with api.Index.build(text, shift=7) as ix:
        L, SA, ref = ix.export()
        assert (SA == fm.SA).all() and (L == fm.L).all() and ref.tobytes() == text and ix.info().shift == 7


# ---------------------------------------------------------------------------------------------------------
# bwa:find_seeds/6
# ---------------------------------------------------------------------------------------------------------
def test_search_kats(gl_index, gl_fm):
    st, sp, ep, n = gl_index.search([k for k, _ in SEARCH_KATS])
    for k, (sub, want) in enumerate(SEARCH_KATS):
        assert (int(st[k]), int(sp[k]), int(ep[k]), int(n[k])) == O.find_seeds_interval(gl_fm, sub)
        sa = gl_index.locate(int(sp[k]), int(ep[k] - sp[k] + 1))
        assert [int(x) + int(n[k]) for x in sa[::-1]] == want      # fmi.erl:89-100: SA[Ep]+N ... SA[Sp]+N
    assert (int(sp[0]), int(ep[0])) == (168455, 168457)


def test_search_parity_random(gl_index, gl_ref, gl_coracle):
    rng = np.random.default_rng(5)
    subs = [_rand_query(rng, gl_ref, 1, 48) for _ in range(30000)]
    st, sp, ep, n = gl_index.search(subs)
    kinds = set()
    for k, q in enumerate(subs):
        w = gl_coracle.find_seeds(q)
        kinds.add(w[0])
        if w[0] == 3:
            assert st[k] == 3, q
        else:
            assert (int(st[k]), int(sp[k]), int(ep[k]), int(n[k])) == w, q
    assert kinds >= {0, 1, 2}


def test_search_empty_and_single(gl_index, gl_coracle):
    st, sp, ep, n = gl_index.search([])
    assert len(st) == 0
    st, sp, ep, n = gl_index.search([b"A", b"N", b"T", b"", b"X", b"ACGTACG", b"GATTACAX"])
    # one character: the interval of a whole first letter is wider than 45 rows -> too_many_seeds (bwa.erl:44)
    assert list(st) == [2, 2, 2, 3, 3, 1, 3]
    for k, q in enumerate([b"A", b"N", b"T"]):
        assert (int(st[k]), int(sp[k]), int(ep[k]), int(n[k])) == gl_coracle.find_seeds(q)


# ---------------------------------------------------------------------------------------------------------
# sga:sga/6
# ---------------------------------------------------------------------------------------------------------
def _check_sga(ix, orc, reads, offs):
    st, clo, S, D = ix.sga(reads, offs)
    n = len(offs) - 1
    nonempty = 0
    for r in range(n):
        q = reads[offs[r]:offs[r + 1]].tobytes()
        try:
            want = orc.sga(q)
        except RuntimeError:
            assert st[r] == api.READ_CRASH
            assert clo[r + 1] == clo[r]
            continue
        assert st[r] == api.READ_OK
        got = list(zip(S[clo[r]:clo[r + 1]].tolist(), D[clo[r]:clo[r + 1]].tolist()))
        assert got == want, (r, q)
        nonempty += bool(want)
    return nonempty


def test_sga_kats(gl_index):
    st, clo, S, D = gl_index.sga([READ_KAT_1, READ_KAT_2])
    assert list(zip(S[clo[0]:clo[1]].tolist(), D[clo[0]:clo[1]].tolist())) == [(127408, 1)]
    assert list(zip(S[clo[1]:clo[2]].tolist(), D[clo[1]:clo[2]].tolist())) == KAT_2_SGA


def test_sga_parity_c1(gl_index, gl_ref, gl_coracle):
    reads, offs, _ = sample_reads(gl_ref, 10000, 100, seed=42, edit_rate=0.01)
    assert _check_sga(gl_index, gl_coracle, reads, offs) > 2000


def test_sga_parity_with_n_and_ragged(gl_index, gl_ref, gl_coracle):
    reads, offs, _ = sample_reads(gl_ref, 3000, 100, seed=9, edit_rate=0.01, n_rate=0.02)
    assert _check_sga(gl_index, gl_coracle, reads, offs) > 100
    rng = np.random.default_rng(3)
    ragged = [_rand_query(rng, gl_ref, 1, 60) for _ in range(2000)] + [b"", b"ACGTACG", b"NNNNNNNNNNNN", b"ACGTACGT"]
    data, offs = api.pack(ragged)
    _check_sga(gl_index, gl_coracle, data, offs)


def test_sga_parity_long_reads(gl_index, gl_ref, gl_coracle):
    for ln, cnt, seed in ((150, 2000, 1), (250, 1500, 2), (1000, 200, 3)):
        reads, offs, _ = sample_reads(gl_ref, cnt, ln, seed=seed, edit_rate=0.01)
        assert _check_sga(gl_index, gl_coracle, reads, offs) > cnt // 10


def test_sga_many_seeds_slow_path():
    # a tandem repeat makes every 15-mer hit hundreds of rows... but >45 rows is too_many_seeds; intervals that stop
    # early (not_found) skip the width check, so craft a reference where that happens with wide intervals
    unit = random_reference(37, 21)
    ref = unit * 40 + random_reference(2000, 22) + unit * 40
    fm = O.build_fm(ref, 1)
    orc = COracle(fm.L, fm.SA, fm.pointers, 1, ref, mode=0)
    rng = np.random.default_rng(8)
    reads = []
    for _ in range(600):
        p = int(rng.integers(0, len(ref) - 120))
        q = bytearray(ref[p:p + 120])
        for _ in range(int(rng.integers(0, 12))):
            q[int(rng.integers(0, 120))] = b"ACGT"[int(rng.integers(0, 4))]
        reads.append(bytes(q))
    data, offs = api.pack(reads)
    with api.Index.from_arrays(fm.L, fm.SA, fm.pointers, 1, ref) as ix:
        _check_sga(ix, orc, data, offs)
        b = api.Batch(ix, len(reads), len(data))
        b.upload(data, offs)
        b.run()
        b.close()


# ---------------------------------------------------------------------------------------------------------
# sw:sw/2
# ---------------------------------------------------------------------------------------------------------
def _mutated_window(rng, q, iupac, max_edits=4, flank=4):
    r = bytearray(q)
    for _ in range(int(rng.integers(0, max_edits + 1))):
        p = int(rng.integers(0, len(r) + 1))
        k = rng.integers(0, 3)
        if k == 0 and len(r) > 1:
            del r[min(p, len(r) - 1)]
        elif k == 1:
            r.insert(p, _rand_seq(rng, 1, iupac)[0])
        elif len(r):
            r[min(p, len(r) - 1)] = _rand_seq(rng, 1, iupac)[0]
    return _rand_seq(rng, int(rng.integers(0, flank)), iupac) + bytes(r) + _rand_seq(rng, int(rng.integers(0, flank)), iupac)


def _check_sw(pairs):
    q, qo = api.pack([p[0] for p in pairs])
    r, ro = api.pack([p[1] for p in pairs])
    st, sc, cig = api.sw_batch(q, qo, r, ro)
    matched = 0
    for k, (a, b) in enumerate(pairs):
        want = c_sw(a, b)
        if want == "no_match":
            assert st[k] == api.HIT_NO_MATCH, (a, b)
        else:
            assert st[k] == api.HIT_MATCH, (a, b)
            assert (int(sc[k]), cig[k].decode()) == want, (a, b)
            matched += 1
    return matched


def test_sw_parity_acgt():
    rng = np.random.default_rng(21)
    pairs = []
    for it in range(6000):
        lq = int(rng.integers(1, 160))
        q = _rand_seq(rng, lq, False)
        pairs.append((q, _mutated_window(rng, q, False, max_edits=6, flank=8)))
    for it in range(500):                                   # unrelated strings: row-threshold aborts
        pairs.append((_rand_seq(rng, int(rng.integers(1, 120)), False), _rand_seq(rng, int(rng.integers(1, 160)), False)))
    assert _check_sw(pairs) > 4000


def test_sw_shortcut_adversarial(sw_mode):
    """Pairs aimed at the shortcut's decision: low-complexity and tandem-repeat strings (competing diagonals and
    deletions between matching flanks), slack 0..20, zero to two substitutions or an indel."""
    rng = np.random.default_rng(29)
    pairs = []
    for it in range(38000):
        alpha = [b"ACGT", b"AC", b"ACGT", b"AAC", b"ACGT"][it % 5]
        # the last 8000: 130..260 bp (penalty budgets 12 and 15, up to nine diagonals beside the window's, the 256 bp limit)
        lq, dd = (int(rng.integers(4, 130)) if it < 30000 else int(rng.integers(130, 261))), int(rng.integers(0, 21))
        pick = lambda n: bytes(alpha[int(x)] for x in rng.integers(0, len(alpha), n))   # noqa: E731
        mode = rng.random()
        if mode < 0.55:
            base = pick(lq + dd + 4)
            s0 = int(rng.integers(0, dd + 3))
            q = bytearray(base[s0:s0 + lq]) + bytearray(pick(max(0, lq - len(base[s0:s0 + lq]))))
            for _ in range(int(rng.choice([0, 0, 1, 1, 1, 2, 3, 4, 5, 6]))):
                u = rng.random()
                if u < 0.7:
                    q[int(rng.integers(0, lq))] = alpha[int(rng.integers(0, len(alpha)))]
                elif u < 0.85 and len(q) > 5:
                    del q[int(rng.integers(0, len(q)))]
                    q += pick(1)
                else:
                    q.insert(int(rng.integers(0, len(q))), alpha[int(rng.integers(0, len(alpha)))])
            pairs.append((bytes(q[:lq]), base[:lq + dd]))
        elif mode < 0.85:
            unit = pick(int(rng.integers(1, 5)))
            base = (unit * 400)[:lq + dd]
            r = bytearray(base)
            for _ in range(int(rng.integers(0, 3))):
                r[int(rng.integers(0, len(r)))] = b"ACGT"[int(rng.integers(0, 4))]
            s0 = int(rng.integers(0, dd + 1))
            q = bytearray(base[s0:s0 + lq])
            for _ in range(int(rng.integers(0, 6))):
                q[int(rng.integers(0, lq))] = b"ACGT"[int(rng.integers(0, 4))]
            pairs.append((bytes(q), bytes(r)))
        else:
            pairs.append((pick(lq), pick(lq + dd)))
    assert _check_sw(pairs) > 18000


def test_shortcut_is_taken_and_counted(gl_index, gl_ref, sw_mode):
    reads, offs, _ = sample_reads(gl_ref, 4000, 100, seed=77, edit_rate=0.01)
    b = api.Batch(gl_index, 4000, len(reads))
    try:
        b.upload(reads, offs)
        b.run()
        s = b.stats()
    finally:
        b.close()
    assert s.sw_calls > 1000
    if sw_mode in ("shortcut", "small"):
        assert 0.1 * s.sw_calls < s.sw_shortcut <= s.sw_calls and 0 < s.sw_shortcut_cells <= s.sw_cells
    else:
        assert s.sw_shortcut == 0 and s.sw_shortcut_cells == 0
    # the banded table: reads with an indel in the default routing, nearly every pair when the shortcut is off
    if sw_mode == "shortcut":
        assert 0 < s.sw_band < 0.5 * s.sw_calls and s.sw_band_filled < s.sw_band_cells <= s.sw_cells
    elif sw_mode == "band_only":
        assert s.sw_band > 0.25 * s.sw_calls and s.sw_band_filled < 0.25 * s.sw_band_cells
    else:
        assert s.sw_band == 0 and s.sw_band_cells == 0 and s.sw_band_filled == 0


def test_sw_parity_iupac_and_n():
    rng = np.random.default_rng(22)
    pairs = []
    for it in range(1500):
        lq = int(rng.integers(1, 120))
        q = _rand_seq(rng, lq, it % 2 == 0)
        pairs.append((q, _mutated_window(rng, q, True)))
    assert _check_sw(pairs) > 500


def test_sw_parity_lengths_and_strip_edges():
    rng = np.random.default_rng(23)
    pairs = []
    for lq in list(range(1, 45)) + [99, 100, 101, 149, 150, 151, 249, 250, 251, 400, 1000, 1001, 1100]:
        for _ in range(3):
            q = _rand_seq(rng, lq, False)
            pairs.append((q, _mutated_window(rng, q, False, max_edits=3)))
            pairs.append((q, q))
    # wide windows (large D), beyond the shared-memory tier of the fast kernel
    q = _rand_seq(rng, 100, False)
    pairs.append((q, _rand_seq(rng, 300, False) + q + _rand_seq(rng, 300, False)))
    pairs.append((q, _rand_seq(rng, 60, False) + q + _rand_seq(rng, 60, False)))
    assert _check_sw(pairs) > 100


def test_sw_literal_small(gl_ref):
    # the clause-machine traceback of sw.erl:116-160 itself (pure Python), on a few pairs
    rng = np.random.default_rng(24)
    pairs = []
    for _ in range(60):
        q = _rand_seq(rng, int(rng.integers(1, 30)), False)
        pairs.append((q, _mutated_window(rng, q, False)))
    q, qo = api.pack([p[0] for p in pairs])
    r, ro = api.pack([p[1] for p in pairs])
    st, sc, cig = api.sw_batch(q, qo, r, ro)
    for k, (a, b) in enumerate(pairs):
        want = O.sw_literal(a, b)
        got = "no_match" if st[k] != 0 else (int(sc[k]), cig[k].decode())
        assert got == want


def test_sw_empty():
    st, sc, cig = api.sw_batch(*api.pack([b"", b"ACGT", b"ACGT"]), *api.pack([b"ACGT", b"", b"ACGT"]))
    assert list(st) == [1, 1, 0] and sc[2] == 8 and cig[2] == b"4M"
    st, sc, cig = api.sw_batch(*api.pack([]), *api.pack([]))
    assert len(st) == 0


# ---------------------------------------------------------------------------------------------------------
# worker_bwt inner loop
# ---------------------------------------------------------------------------------------------------------
def _check_align(ix, orc, reads, offs, min_hits=0):
    st, ho, hits, cig = ix.align(reads, offs)
    n = len(offs) - 1
    ncand, cand, stats = orc.align_batch(reads, offs, nthreads=8, cap_per_read=64)
    total = 0
    for r in range(n):
        c = int(ncand[r])
        if c < 0:
            assert st[r] == api.READ_CRASH and ho[r + 1] == ho[r]
            continue
        assert st[r] == api.READ_OK
        assert ho[r + 1] - ho[r] == c, r
        assert c <= 64
        for k in range(c):
            w = cand[r * 64 + (c - 1 - k)]            # oracle order is sga's (descending S); Cigars is ascending
            h = hits[ho[r] + k]
            assert (int(h["s"]), int(h["d"]), int(h["status"])) == (w.S, w.D, w.status), (r, k)
            if w.status != api.HIT_WINDOW_OOR:
                assert int(h["pos"]) == w.pos
            if w.status == 0:
                assert int(h["score"]) == w.score
                assert int(h["cigar_len"]) == w.cigar_len
                got = cig[h["cigar_off"]:h["cigar_off"] + h["cigar_len"]]
                if w.cigar_len < 96:
                    assert got == w.cigar[:w.cigar_len], (r, k)
                total += 1
    assert total >= min_hits
    return stats


def test_align_kats(gl_index):
    st, ho, hits, cig = gl_index.align([READ_KAT_1, READ_KAT_2])
    h = hits[0]
    assert (int(h["score"]), cig[h["cigar_off"]:h["cigar_off"] + h["cigar_len"]], int(h["pos"])) == (192, b"60M1I40M", 127307)
    got = [(int(h["score"]), cig[h["cigar_off"]:h["cigar_off"] + h["cigar_len"]].decode(), int(h["pos"]))
           for h in hits[ho[1]:ho[2]] if h["status"] == 0]
    assert got == KAT_2_CIGARS


def test_align_parity_c1(gl_index, gl_ref, gl_coracle):
    """BASELINE.json configs[0]: GL000193.1, 10k synthetic 100 bp reads with 1 % edits."""
    reads, offs, _ = sample_reads(gl_ref, 10000, 100, seed=42, edit_rate=0.01)
    _check_align(gl_index, gl_coracle, reads, offs, min_hits=4000)


def test_align_parity_n_reads_and_edges(gl_index, gl_ref, gl_coracle):
    reads, offs, _ = sample_reads(gl_ref, 2000, 100, seed=6, edit_rate=0.02, n_rate=0.01)
    _check_align(gl_index, gl_coracle, reads, offs, min_hits=100)
    # reads from the first / last bases of the chunk: windows that reach into (or beyond) the 'N' padding
    edge = []
    for p in list(range(0, 40, 3)) + list(range(len(gl_ref) - 140, len(gl_ref) - 100, 3)):
        edge.append(gl_ref[p:p + 100])
        edge.append(gl_ref[p:p + 50] + b"ACGTTGCAAC" + gl_ref[p + 50:p + 100])
    data, offs = api.pack(edge + [b"", b"ACGT"])
    _check_align(gl_index, gl_coracle, data, offs)
    st, ho, hits, cig = gl_index.align([])
    assert len(st) == 0 and list(ho) == [0] and len(hits) == 0


def test_align_parity_150_250(gl_index, gl_ref, gl_coracle):
    for ln, cnt, seed in ((150, 3000, 4), (250, 1500, 5)):
        reads, offs, _ = sample_reads(gl_ref, cnt, ln, seed=seed, edit_rate=0.01)
        _check_align(gl_index, gl_coracle, reads, offs, min_hits=cnt // 4)


def test_align_batch_split_invariance(gl_index, gl_ref):
    """Read sharding: the result of a batch is the concatenation of the results of its shards."""
    reads, offs, _ = sample_reads_fast(np.frombuffer(gl_ref, np.uint8), 4000, 100, seed=12)
    st, ho, hits, cig = gl_index.align(reads, offs)
    mid = 1777
    a = gl_index.align(reads[:offs[mid]], offs[:mid + 1])
    b = gl_index.align(reads[offs[mid]:], offs[mid:] - offs[mid])
    assert (np.concatenate([a[0], b[0]]) == st).all()
    assert (np.concatenate([a[1][:-1], b[1] + a[1][-1]]) == ho).all()
    for f in ("pos", "s", "d", "score", "status", "cigar_len"):
        assert (np.concatenate([a[2][f], b[2][f]]) == hits[f]).all()
    assert a[3] + b[3] == cig


def test_synthetic_index_parity_midsize():
    """Device-generated reference + device builder at a size the oracle still handles (4 Mbp): export the index,
    run the oracle on the exported arrays, compare the whole path for 100/150 bp reads."""
    n = 4_000_000
    with api.Index.build_synthetic(n, seed=7, shift=1) as ix:
        L, SA, ref = ix.export()
        assert len(L) == n + 1 and set(np.unique(ref).tolist()) == {65, 67, 71, 84}
        # SA is a permutation and sorted: check suffix order on a sample of adjacent pairs
        assert (np.sort(SA) == np.arange(n + 1, dtype=np.uint32)).all()
        text = ref.tobytes() + b"$"
        rng = np.random.default_rng(1)
        for p in rng.integers(0, n, size=3000):
            a, b = int(SA[p]), int(SA[p + 1])
            assert text[a:a + 64] < text[b:b + 64] or text[a:] < text[b:]
        assert (L == np.frombuffer(text, np.uint8)[(SA.astype(np.int64) - 1) % (n + 1)]).all()
        orc = COracle(L, SA, tuple(ix.info().pointers), 1, ref.tobytes(), mode=1)
        for ln, cnt in ((100, 20000), (150, 5000)):
            reads, offs, _ = sample_reads_fast(ref, cnt, ln, seed=ln, edit_rate=0.01)
            _check_align(ix, orc, reads, offs, min_hits=cnt // 10)


# ---------------------------------------------------------------------------------------------------------
# k-mer interval table depth never changes a result
# ---------------------------------------------------------------------------------------------------------
def test_kmer_table_depths(gl_fm, gl_ref, gl_coracle):
    rng = np.random.default_rng(77)
    subs = [_rand_query(rng, gl_ref, 1, 40) for _ in range(20000)]
    want = [gl_coracle.find_seeds(q) for q in subs]
    reads, offs, _ = sample_reads(gl_ref, 3000, 100, seed=13, edit_rate=0.02, n_rate=0.003)
    base = None
    with api.Index.from_arrays(gl_fm.L, gl_fm.SA, gl_fm.pointers, gl_fm.shift, gl_ref) as ix:
        assert ix.info().kmer_k == 9                    # 4^9 <= 4 * 189790 < 4^10
        for depth in (0, 9, 11, 14, 15):
            ix.set_kmer_depth(depth)
            assert ix.info().kmer_k == max(depth, 8)
            st, sp, ep, n = ix.search(subs)
            for k, w in enumerate(want):
                if w[0] == 3:
                    assert st[k] == 3
                else:
                    assert (int(st[k]), int(sp[k]), int(ep[k]), int(n[k])) == w, (depth, subs[k])
            res = ix.align(reads, offs)
            if base is None:
                base = res
                _check_align(ix, gl_coracle, reads, offs, min_hits=500)
            else:
                assert (res[0] == base[0]).all() and (res[1] == base[1]).all() and res[3] == base[3]
                assert (res[2] == base[2]).all()
        with pytest.raises(api.BwtError):
            ix.set_kmer_depth(16)


def test_pipelined_aligner_equals_one_shot(gl_index, gl_ref, gl_coracle):
    reads, offs, _ = sample_reads(gl_ref, 5000, 100, seed=31, edit_rate=0.02, n_rate=0.002)
    want = gl_index.align(reads, offs)
    for chunk_reads, chunk_bytes, depth in ((700, 1 << 20, 3), (5000, 30_000, 2), (64, 1 << 20, 4), (10000, 1 << 22, 1)):
        al = api.Aligner(gl_index, chunk_reads, chunk_bytes, depth)
        got = al.run(reads, offs)
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and got[3] == want[3]
        assert (got[2] == want[2]).all()
        st = al.stats()
        assert st.reads == 5000 and st.launches > 0
        # too little room: the needed sizes come back
        nh, nc = len(want[2]), len(want[3])
        hits = np.zeros(max(1, nh // 2), api.HIT_DTYPE)
        cig = np.zeros(nc + 16, np.uint8)
        stt = np.zeros(5000, np.int32)
        ho = np.zeros(5001, np.int64)
        with pytest.raises(api.NoSpace) as e:
            al.run_ptr(reads.ctypes.data, offs.ctypes.data, 5000, stt.ctypes.data, ho.ctypes.data, hits.ctypes.data, len(hits),
                       cig.ctypes.data, len(cig))
        assert e.value.needed == (nh, nc)
        al.close()
    _check_align(gl_index, gl_coracle, reads, offs, min_hits=500)
    # a read longer than the chunk byte limit is a documented error, not a hang
    al = api.Aligner(gl_index, 10, 50, 2)
    with pytest.raises(api.BwtError):
        al.run(reads, offs)
    al.close()


def test_fast_kernels_take_plain_reads(gl_index, gl_ref, sw_mode):
    """ACGT-only inputs must stay on the fast kernels (a silent detour through the double-precision kernel would
    still be correct, only slow): the work counters say which path ran."""
    if sw_mode == "small":
        pytest.skip("the small-batch routing sends DP pairs to the warp-per-pair kernel on purpose")
    reads, offs, _ = sample_reads_fast(np.frombuffer(gl_ref, np.uint8), 20000, 100, seed=3)
    b = api.Batch(gl_index, 20000, len(reads))
    b.upload(reads, offs)
    b.run()
    s = b.stats()
    # only windows that reach into the 'N' padding at the chunk ends may take the double-precision kernel
    assert s.reads == 20000 and s.sw_calls > 5000 and s.sw_general < 100
    assert s.searches == 20000 * 12 and s.lf_steps > 0 and s.sw_cells > 0 and s.launches >= 5
    b.close()


def test_partitions_parity_and_merge(gl_ref):
    """Config 5 geometry on one GPU: two _pJ partitions of GL000193.1 (builder formula, 2 bp overlap), every read
    against both; each partition bit-identical to the oracle run on that partition's index, merge = concatenation."""
    from bwt_b200 import partition
    reads, offs, _ = sample_reads(gl_ref, 3000, 100, seed=8, edit_rate=0.01)
    sl = partition.partition_slices(len(gl_ref), 2)
    parts = []
    for j, (start0, length, add) in enumerate(sl, start=1):
        chunk = gl_ref[start0:start0 + length]
        fm = O.build_fm(chunk, add)
        orc = COracle(fm.L, fm.SA, fm.pointers, add, chunk, mode=0)
        with partition.build_partition(np.frombuffer(gl_ref, np.uint8), 2, j) as ix:
            assert ix.info().shift == add and tuple(ix.info().pointers) == fm.pointers
            _check_align(ix, orc, reads, offs, min_hits=400)
            st, ho, hits, cig = ix.align(reads, offs)
            parts.append((st, ho, hits, np.frombuffer(cig, np.uint8)))
    status, mho, mhits, mcig = partition.merge_partitions(parts)
    assert (np.diff(mho) == np.diff(parts[0][1]) + np.diff(parts[1][1])).all()
    # a read drawn from the first half aligns (if at all) in partition 0 with the same coordinates as un-partitioned
    assert set(np.unique(mhits["part"]).tolist()) == {0, 1}


def test_master_run_end_to_end(gl_index, gl_fm, gl_ref):
    """FASTQ text -> workloads of 200 in the reference's (reverse) order -> GPU path -> top-1 pick -> 265 filter -> the
    lines master.erl prints; against the literal Python oracle of the same chain (150 bp reads: scores reach 265)."""
    from bwt_b200 import master
    reads, offs, _ = sample_reads(gl_ref, 260, 150, seed=77, edit_rate=0.01)
    recs = [(b"read_%d/1 len=%d" % (k, offs[k + 1] - offs[k]), reads[offs[k]:offs[k + 1]].tobytes()) for k in range(260)]
    text = b"".join(b"@" + n + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n" for n, s in recs)
    padded = O.pad_ref(gl_ref)
    got, dealt = master.run(gl_index, padded, text, "GL000193.1")
    dev, state, want = O.LineDev(text), {"fastq_eof": False}, []
    while True:
        wl = O.produce_workload(dev, O.WATERLINE, state)
        if wl == "stop":
            break
        for lst in wl:
            for msg in O.worker_messages(gl_fm, padded, lst):
                ln = O.master_sink_line(msg, "GL000193.1")
                if ln is not None:
                    want.append(ln)
    assert dealt == 260 and got == want and len(got) > 20


def test_make_indices_from_fasta(tmp_path):
    """Row F1: fm_index:t/0 builds contigs "2" and "4" of bwt_files/test_genome.fasta (fm_index.erl:15-17); the files
    must be what the oracle's builder + ETF writer produce (Appendix D-2 tables)."""
    from bwt_b200 import fm_index
    from test_oracle_golden import TOY_FASTA
    fa = tmp_path / "test_genome.fasta"
    fa.write_bytes(TOY_FASTA)
    made = fm_index.make_indices(str(fa), ["2", "4"], str(tmp_path / "fm"))
    assert [(m[0], m[1], m[2]) for m in made] == [("2", 1, 1), ("4", 1, 1)]
    for name, j, shift, fm_path, ref_path in made:
        sh, body = O.fasta_contig_body(TOY_FASTA, name)
        fm = O.build_fm(body, sh + 1)
        assert open(fm_path, "rb").read() == O.write_fm(fm)
        assert open(ref_path, "rb").read() == body
        meta, ix = fm_index.get_index(name, 1, str(tmp_path / "fm"))
        assert dict(meta) == {"pointers": fm.pointers, "shift": sh + 1}
        ix.close()
    assert fm_index.make_indices(str(fa), ["3"], str(tmp_path / "fm"))[0][2] == 14      # 13 N trimmed on the left + Start 1
