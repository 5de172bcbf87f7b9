"""The packed-reads / compact-records face of the C ABI (bwt_pack_reads, bwt_batch_upload_packed, bwt_batch_run_compact,
bwt_aligner_run_packed) against the ASCII / wide face, which test_gpu_parity.py holds against the oracle; the two stages
of the no-DP classification one at a time; two pairs of different query lengths in one thread of the DP kernel.
Runs on the B200 box only (``-m gpu``)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.c_oracle import c_sw                        # noqa: E402  (the checker, never the product)
from bwt_b200 import api                                # noqa: E402
from bwt_b200.synth import sample_reads                 # noqa: E402

from test_oracle_golden import _rand_seq                # noqa: E402


@pytest.fixture(scope="module")
def gl_index(gl_fm, gl_ref):
    ix = api.Index.from_arrays(gl_fm.L, gl_fm.SA, gl_fm.pointers, gl_fm.shift, gl_ref)
    yield ix
    ix.close()


def _mixed_reads(gl_ref):
    """plain reads, reads with 'N' and other bytes, ragged lengths, empty, very short, longer than the packed limit"""
    reads, offs, _ = sample_reads(gl_ref, 3000, 100, seed=91, edit_rate=0.02, n_rate=0.003)
    rs = [reads[offs[k]:offs[k + 1]].tobytes() for k in range(3000)]
    rs += [b"", b"A", b"ACGTACG", gl_ref[500:508], gl_ref[1000:1015], gl_ref[1000:1016], gl_ref[1000:1017], gl_ref[7000:7031],
           gl_ref[2000:2033], gl_ref[30000:31031], gl_ref[30000:31040], gl_ref[30000:31041], gl_ref[40000:41500],
           gl_ref[9000:9050] + b"R" + gl_ref[9051:9100], gl_ref[9000:9099] + b"x", b"N" * 100, gl_ref[600:850], gl_ref[12000:12150]]
    return rs


def _same_result(want, got_wide):
    assert (got_wide[0] == want[0]).all()
    assert (got_wide[1] == want[1]).all()
    for f in ("pos", "s", "d", "score", "status", "cigar_len", "cigar_off"):
        assert (got_wide[2][f] == want[2][f]).all(), f
    assert got_wide[3] == want[3]


def test_packed_equals_ascii(gl_index, gl_ref):
    rs = _mixed_reads(gl_ref)
    data, offs = api.pack(rs)
    want = gl_index.align(data, offs)
    assert (want[2]["status"] == 0).sum() > 800
    pk = api.PackedReads(data, offs, threads=3)
    assert pk.n == len(rs) and pk.n_exc > 5
    shift = int(gl_index.info().shift)
    lens = np.diff(offs)
    # pipelined aligner, several chunk shapes (chunks cut the flagged-read list at different places)
    for chunk_reads, depth in ((700, 3), (len(rs) + 5, 1), (64, 4)):
        al = api.Aligner(gl_index, chunk_reads, 1 << 20, depth)
        got = al.run_packed(pk)
        _same_result(want, api.expand_compact(*got, read_lens=lens, shift=shift))
        # "<Lq>M" CIGARs are flagged, not stored
        h = got[2]
        full = (h["d_flags"] & api.HIT16_FULL_M) != 0
        assert full.sum() > 300 and (h["cigar_len"][full] == 0).all()
        # too little room: the needed sizes come back
        with pytest.raises(api.NoSpace) as e:
            st = np.zeros(pk.n, np.uint8); ho = np.zeros(pk.n + 1, np.uint32)
            hits = np.zeros(max(1, len(h) // 2), api.HIT16_DTYPE); cig = np.zeros(len(got[3]) + 16, np.uint8)
            al.run_packed_ptr(pk, st.ctypes.data, ho.ctypes.data, hits.ctypes.data, len(hits), cig.ctypes.data, len(cig))
        assert e.value.needed == (len(h), len(got[3]))
        al.close()
    # resident batch, both input faces x both output faces
    b = api.Batch(gl_index, len(rs), len(data))
    try:
        b.upload_packed(pk)
        b.run_compact()
        _same_result(want, api.expand_compact(*b.download_compact(), read_lens=lens, shift=shift))
        b.run()                                                    # packed in, wide out
        got = b.download()
        _same_result(want, (got[0], got[1], got[2], got[3].tobytes()))
        b.upload(data, offs)                                       # ASCII in, compact out
        b.run_compact()
        _same_result(want, api.expand_compact(*b.download_compact(), read_lens=lens, shift=shift))
    finally:
        b.close()


def test_packed_fixed_length(gl_index, gl_ref):
    ref = np.frombuffer(gl_ref, np.uint8)
    rng = np.random.default_rng(5)
    starts = rng.integers(0, len(ref) - 100, 5000)
    reads = np.ascontiguousarray(ref[starts[:, None] + np.arange(100)[None, :]]).copy()
    reads[rng.random(reads.shape) < 0.01] = ord("C")
    reads[17, 40] = ord("N")                                        # one flagged read in a fixed-length batch
    data = reads.reshape(-1)
    offs = np.arange(5001, dtype=np.int64) * 100
    want = gl_index.align(data, offs)
    pk = api.PackedReads(data, offs, fixed_len=100)
    assert pk.n_exc == 1 and pk.upload_bytes() < 5000 * 29 + 200
    al = api.Aligner(gl_index, 1024, 1 << 20, 2)
    got = al.run_packed(pk)
    _same_result(want, api.expand_compact(*got, read_lens=np.diff(offs), shift=int(gl_index.info().shift)))
    al.close()


def _check_pairs(pairs):
    q, qo = api.pack([p[0] for p in pairs])
    r, ro = api.pack([p[1] for p in pairs])
    st, sc, cig = api.sw_batch(q, qo, r, ro)
    n = 0
    for k, (a, b) in enumerate(pairs):
        want = c_sw(a, b)
        if want == "no_match":
            assert st[k] == api.HIT_NO_MATCH, (a, b)
        else:
            assert st[k] == api.HIT_MATCH and (int(sc[k]), cig[k].decode()) == want, (a, b)
            n += 1
    return n


@pytest.mark.parametrize("mode", ["1", "2"])
def test_classification_stages_one_at_a_time(mode):
    """BWT_SW_SHORTCUT=1: only the quick test (perfect diagonal / one mismatch at Dd = 0) settles pairs; =2: only the
    wavefront rule.  Low-complexity strings, competing diagonals, slack 0..17."""
    os.environ["BWT_SW_SHORTCUT"] = mode
    try:
        rng = np.random.default_rng(int(mode) + 40)
        pairs = []
        for it in range(12000):
            alpha = [b"ACGT", b"AC", b"AAC", b"ACGT"][it % 4]
            lq, dd = int(rng.integers(4, 200)), int(rng.choice([0, 0, 0, 1, 2, 3, 8, 16, 17]))
            pick = lambda n: bytes(alpha[int(x)] for x in rng.integers(0, len(alpha), n))   # noqa: E731
            base = pick(lq + dd)
            s0 = int(rng.integers(0, dd + 1))
            q = bytearray(base[s0:s0 + lq])
            for _ in range(int(rng.choice([0, 0, 1, 1, 2, 3]))):
                q[int(rng.integers(0, lq))] = alpha[int(rng.integers(0, len(alpha)))]
            if rng.random() < 0.15 and lq > 6:
                del q[int(rng.integers(0, len(q)))]
                q += pick(1)
            pairs.append((bytes(q), base))
        assert _check_pairs(pairs) > 9000
    finally:
        os.environ.pop("BWT_SW_SHORTCUT", None)


@pytest.mark.parametrize("band", ["0", "1"])
def test_dp_pairs_of_different_lengths(band):
    """Every pair through the DP kernels (BWT_SW_SHORTCUT=0), neighbouring pairs of very different query lengths: the two
    halves of a thread's packed registers then start their first real row in different strips (band 0: the full-table
    kernel alone); with the banded table in front (band 1) the lanes of a warp run bands of different heights and the
    full-table kernel gets what is left."""
    os.environ["BWT_SW_SHORTCUT"] = "0"
    os.environ["BWT_SW_BAND"] = band
    try:
        rng = np.random.default_rng(77)
        pairs = []
        for it in range(5000):
            lq = int(rng.choice([1, 2, 9, 10, 11, 19, 20, 21, 35, 98, 99, 100, 101, 102, 150, 249, 250, 251, 300, 399, 400]))
            q = _rand_seq(rng, lq, False)
            r = bytearray(q)
            for _ in range(int(rng.integers(0, 4))):
                p = int(rng.integers(0, len(r)))
                u = rng.random()
                if u < 0.5:
                    r[p] = b"ACGT"[int(rng.integers(0, 4))]
                elif u < 0.75 and len(r) > 2:
                    del r[p]
                else:
                    r.insert(p, b"ACGT"[int(rng.integers(0, 4))])
            pairs.append((q, _rand_seq(rng, int(rng.integers(0, 6)), False) + bytes(r) + _rand_seq(rng, int(rng.integers(0, 6)), False)))
        assert _check_pairs(pairs) > 3500
    finally:
        os.environ.pop("BWT_SW_SHORTCUT", None)
        os.environ.pop("BWT_SW_BAND", None)


def test_indel_reads_sampler_and_depth_invariance(gl_index, gl_ref):
    """bwt_sample_reads_edits: ragged reads (every tail depth 8..15 occurs), results independent of the table depth."""
    reads, offs, starts = api.sample_reads_edits(gl_index, 6000, 100, seed=9, edit_rate=0.02, want_starts=True)
    lens = np.diff(offs)
    assert lens.min() < 100 < lens.max() and abs(int(lens.sum()) - 600000) < 3000
    ref = np.frombuffer(gl_ref, np.uint8)
    same = [bytes(reads[offs[k]:offs[k + 1]]) == bytes(ref[starts[k]:starts[k] + 100]) for k in range(300)]
    assert 10 < sum(same) < 120                                      # (1 - 0.02)^100 = 13 % of the reads carry no edit
    a = gl_index.align(reads, offs)
    assert (a[2]["status"] == 0).sum() > 1500
    n_rate_reads, n_offs, _ = api.sample_reads_edits(gl_index, 500, 100, seed=9, edit_rate=0.0, n_rate=0.05)
    assert (n_rate_reads == ord("N")).sum() > 1500
    # synthetic source: the same bases as an index built from that generator
    with api.Index.build_synthetic(300_000, seed=5, shift=1) as sx:
        r1 = api.sample_reads_edits(sx, 200, 80, seed=4, edit_rate=0.01)
        r2 = api.sample_reads_edits(None, 200, 80, seed=4, edit_rate=0.01, synth=(300_000, 5))
        assert (r1[1] == r2[1]).all() and (r1[0] == r2[0]).all()
        with api.Index.build_synthetic_part(1000, 50_000, seed=5, shift=1001) as px:
            assert (px.export()[2] == sx.export()[2][1000:51000]).all() and px.info().shift == 1001


def test_partition_gather_single_rank(gl_index, gl_ref):
    """bwt_partition_gather with one rank (no NCCL): the root's block is the batch's own compact result."""
    reads, offs, _ = sample_reads(gl_ref, 1500, 100, seed=13, edit_rate=0.01)
    b = api.Batch(gl_index, 1500, len(reads))
    comm = api.Comm(None, 1, 0)
    try:
        b.upload(reads, offs)
        b.run_compact()
        want = b.download_compact()
        parts, ms, link = comm.gather(b, 0)
        assert len(parts) == 1 and link == 0 and ms >= 0
        got = parts[0]
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all() and (got[2] == want[2]).all() and (got[3] == want[3]).all()
        with pytest.raises(api.NoSpace) as e:
            comm.gather(b, 0, (np.zeros(1500, np.uint8), np.zeros(1501, np.uint32), np.zeros(4, api.HIT16_DTYPE), np.zeros(1 << 16, np.uint8)))
        assert e.value.needed[0] == len(want[2])
    finally:
        comm.close()
        b.close()
