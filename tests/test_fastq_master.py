"""Rows F2/F3: FASTQ ingest, batch framing and the result sink against the oracle's restatement of
fastq.erl / master.erl.  The parser is host code: no GPU needed."""
import numpy as np
import pytest

from oracle import bwt_oracle as O
from bwt_b200 import fastq, master


def _fastq(records, junk_every=0, final_newline=True):
    out = []
    for k, (name, seq) in enumerate(records):
        if junk_every and k % junk_every == 0:
            out.append(b"stray line\n")
        out.append(b"@" + name + b"\n" + seq + b"\n+\n" + b"@" * len(seq) + b"\n")     # quality lines that start with '@'
    t = b"".join(out)
    return t if final_newline else t[:-1]


def _oracle_all(text):
    dev = O.LineDev(text)
    recs = []
    while True:
        r = O.fastq_read_seq(dev)
        if r is None:
            return recs
        recs.append(r)


@pytest.mark.parametrize("junk,final_nl", [(0, True), (3, True), (0, False), (5, False)])
def test_parser_equals_reference_rule(junk, final_nl):
    rng = np.random.default_rng(4)
    recs = [(b"read%d desc %d" % (k, k * 7), bytes(rng.choice(list(b"ACGTN"), size=int(rng.integers(0, 130))).astype(np.uint8)))
            for k in range(257)]
    text = _fastq(recs, junk, final_nl)
    want = _oracle_all(text)
    reads, offs, names, used = fastq.parse(text)
    got = [(names[k], reads[offs[k]:offs[k + 1]].tobytes()) for k in range(len(names))]
    assert got == want and used == len(text)
    if final_nl and not junk:
        assert got == recs


def test_parser_quirks():
    # header as the last line: badmatch in the reference
    with pytest.raises(Exception):
        fastq.parse(b"@r1\nACGT\n+\nIIII\n@r2\n")
    with pytest.raises(O.ReferenceCrash):
        _oracle_all(b"@r1\nACGT\n+\nIIII\n@r2\n")
    # record cut after the sequence line, sequence without newline: droplast eats a base
    for t in (b"@r1\nACGT", b"@r1\nACGT\n", b"@r1\nACGT\n+", b"@\n\n", b"", b"\n\n", b"no header\n"):
        reads, offs, names, used = fastq.parse(t)
        assert [(names[k], reads[offs[k]:offs[k + 1]].tobytes()) for k in range(len(names))] == _oracle_all(t)


def test_incremental_parse():
    recs = [(b"r%d" % k, b"ACGT" * (k % 9 + 1)) for k in range(100)]
    text = _fastq(recs)
    got, pos = [], 0
    while pos < len(text):
        chunk = text[pos:pos + 97]
        final = pos + 97 >= len(text)
        reads, offs, names, used = fastq.parse(chunk, final=final)
        got += [(names[k], reads[offs[k]:offs[k + 1]].tobytes()) for k in range(len(names))]
        assert used > 0 or final
        pos += used if used else len(chunk)
    assert got == recs


def test_batches_and_workloads_match_reference_order():
    recs = [(b"r%d" % k, b"ACGTACGTAC") for k in range(200 * 3 + 57)]
    text = _fastq(recs)
    rd, dev = fastq.Reader(text), O.LineDev(text)
    for n in (1, 7, 200, 200, 500):
        assert rd.read_seqs(n) == O.fastq_read_seqs(dev, n)
    assert rd.read_seqs(5) == "eof" and O.fastq_read_seqs(dev, 5) == "eof"
    rd, dev = fastq.Reader(text), O.LineDev(text)
    s1, s2 = {"fastq_eof": False}, {"fastq_eof": False}
    for _ in range(4):
        assert master.produce_workload(rd, 2, s1) == O.produce_workload(dev, 2, s2)
        assert s1 == s2
    assert master.produce_workload(rd, 2, s1) == "stop"


def test_sink_filter_and_format():
    msg = ("cigar", b"read1 x", (291, "150M"), 12345, b"ACGT")
    assert master.sink_line(msg, "21") == O.master_sink_line(msg, "21") == "\nread1 x      21      12345      150M      291      ACGT\n"
    assert master.sink_line(("cigar", b"r", (264, "1M"), 1, b"A"), "21") is None
    assert master.sink_line("no_cigar", "21") is None


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/src/fastq.erl"), reason="the reference tree only exists in the build container")
def test_oracle_equals_the_interpreted_fastq_erl():
    """src/fastq.erl, unmodified, run by oracle/mini_erl.py (file:read_line/1 over an in-memory device): read_seqs/2 in
    batches, every quirk of the record reader -- against the oracle's restatement"""
    from oracle import mini_erl as E
    m = E.Machine("/root/reference/src", "/root/reference/include")
    m.load("fastq")

    def conv(term):
        if isinstance(term, E.Atom):
            return str(term)
        if isinstance(term, tuple) and term and isinstance(term[0], E.Atom) and str(term[0]) in ("ok", "eof"):
            return (str(term[0]), [(bytes(E.to_pylist(a)), bytes(E.to_pylist(b))) for (a, b) in E.to_pylist(term[1])])
        raise AssertionError(term)

    rng = np.random.default_rng(9)
    recs = [(b"r%d x" % k, bytes(rng.choice(list(b"ACGTN"), size=int(rng.integers(0, 40))).astype(np.uint8))) for k in range(23)]
    texts = [_fastq(recs, 0, True), _fastq(recs, 4, True), _fastq(recs, 0, False), _fastq(recs[:7], 3, False), b"", b"junk\n\n",
             b"@r1\nACGT", b"@r1\nACGT\n+", b"@\n\n"]
    for text in texts:
        for n in (1, 5, 200):
            dev_e, dev_o = E.ReadDevice(text), O.LineDev(text)
            while True:
                got = conv(m.call("fastq", "read_seqs", [dev_e, n]))
                want = O.fastq_read_seqs(dev_o, n)
                assert got == want, (text[:40], n)
                if got == "eof" or got[0] == "eof":
                    break
    with pytest.raises(E.ErlError):                                        # header as the last line: {ok,SeqData} = eof
        m.call("fastq", "read_seqs", [E.ReadDevice(b"@r1\nACGT\n+\nIIII\n@r2\n"), 5])
