"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/bwt_b200.h declares, refuses to
compute without a device, the NIF shim compiles against the mock erl_nif.h, and the host mirrors' pure logic."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "bwt_b200.h")) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(bwt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bwt_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 25
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table and the header agree
    L = _lib.lib()                                   # raises if a symbol is missing
    for n in names:
        assert getattr(L, n) is not None
    assert L.bwt_abi_version() == 5


def test_struct_layouts_match_header():
    from bwt_b200 import _lib, api
    assert C.sizeof(_lib.BwtHit) == 40 and api.HIT_DTYPE.itemsize == 40
    assert C.sizeof(_lib.BwtIndexInfo) == 72
    assert C.sizeof(_lib.BwtHit16) == 16 and api.HIT16_DTYPE.itemsize == 16
    assert C.sizeof(_lib.BwtPackedReads) == 88 and C.sizeof(_lib.BwtSampleOpts) == 72
    assert C.sizeof(_lib.BwtStats) == 9 * 8 + 8 * 4 + 2 * 4 + 5 * 8


def test_no_cpu_fallback_without_device(tmp_path):
    from bwt_b200 import api
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.BwtError) as e:
        api.Index.build(b"ACGTACGTACGTACGT", shift=1)
    assert e.value.code == -7                        # BWT_ERR_NODEVICE: there is no CPU path
    with pytest.raises(api.BwtError) as e:
        api.sw_batch(*api.pack([b"ACGT"]), *api.pack([b"ACGT"]))
    assert e.value.code == -7


def test_nif_shim_compiles_against_mock_header():
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "tests", "mock_erl"),
                           os.path.join(ROOT, "c_src", "bwt_nif.c")])


def test_product_does_not_import_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "bwt_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                with open(os.path.join(d, fn)) as f:
                    assert "oracle" not in f.read(), fn


def test_chunk_plan_matches_oracle():
    from bwt_b200 import fm_index
    from oracle import bwt_oracle as O
    for n, nbr in ((189789, None), ((1 << 24) + 5, None), (1000, 8), (3137454466, 8), (50_000_000, None)):
        assert fm_index.chunk_plan(n, nbr) == O.chunk_plan(n, nbr)


def test_top1_pick():
    from bwt_b200.worker_bwt import otp_sort_top
    c = [((185, "100M"), 1, b""), ((197, "100M"), 2, b""), ((176, "100M"), 3, b""), ((197, "99M"), 4, b"")]
    assert otp_sort_top(c)[1] == 4
    assert otp_sort_top(c[:3])[1] == 2


def test_pack():
    from bwt_b200 import api
    d, o = api.pack([b"AC", b"", b"GTT"])
    assert d.tobytes() == b"ACGTT" and o.tolist() == [0, 2, 2, 5]
    d, o = api.pack([])
    assert len(d) == 0 and o.tolist() == [0]


def test_fasta_contig_extraction_matches_oracle():
    from bwt_b200 import fm_index
    from oracle import bwt_oracle as O
    from test_oracle_golden import TOY_FASTA
    got = {name: fm_index.contig_body(TOY_FASTA[off:off + ln]) for name, off, ln in fm_index._contigs(TOY_FASTA)}
    assert sorted(got) == ["1", "2", "3", "4", "5"]
    for name in got:
        assert got[name] == O.fasta_contig_body(TOY_FASTA, name)
    assert got["3"][0] == 13 and got["2"] == (0, b"AACCATGG")
    crlf = b">x y z\r\nNNAC\r\nGT\r\nNN\r\n>y\nACGT"
    got = {name: fm_index.contig_body(crlf[off:off + ln]) for name, off, ln in fm_index._contigs(crlf)}
    assert got == {"x": (2, b"ACGT"), "y": (0, b"ACGT")}


def test_fm_container_atom_encodings(tmp_path):
    """The .fm reader takes ATOM_EXT (100: OTP 17, the shipped fixture) as well as the forms current OTP writes
    (SMALL_ATOM_UTF8_EXT 119, ATOM_UTF8_EXT 118, SMALL_ATOM_EXT 115); a truncated file is a format error, not an
    allocation of gigabytes."""
    from oracle import bwt_oracle as O
    from bwt_b200 import api
    fm = O.build_fm(b"TGTGTGATGATGACCCCCCATGA", 14)
    data = O.write_fm(fm)
    want = (len(fm.L), tuple(fm.pointers), 14)
    variants = {"atom_ext": data}
    for name, tag, wide in (("small_utf8", 119, False), ("utf8", 118, True), ("small", 115, False)):
        d = data
        for atom in (b"pointers", b"shift"):
            old = bytes([100, 0, len(atom)]) + atom
            new = (bytes([tag, 0, len(atom)]) if wide else bytes([tag, len(atom)])) + atom
            assert old in d
            d = d.replace(old, new)
        variants[name] = d
    for name, d in variants.items():
        p = tmp_path / (name + ".fm")
        p.write_bytes(d)
        assert api.fm_inspect(str(p)) == want, name
    # truncated: the tuple header promises more rows than the file holds
    p = tmp_path / "cut.fm"
    p.write_bytes(data[:len(data) // 2])
    with pytest.raises(api.BwtError) as e:
        api.fm_inspect(str(p))
    assert e.value.code == -3
    big = bytearray(data)
    at = data.index(bytes([106, 104, len(fm.L)]))            # NIL, SMALL_TUPLE n  ->  LARGE_TUPLE 0x7fffffff
    big[at + 1:at + 3] = bytes([105, 0x7F, 0xFF, 0xFF, 0xFF])
    p.write_bytes(bytes(big))
    with pytest.raises(api.BwtError) as e:
        api.fm_inspect(str(p))
    assert e.value.code == -3 and "row count" in str(e.value)


def test_read_limit_is_not_silently_dropped():
    """A read beyond this library's limits (BWT_READ_LIMIT: longer than 1031 bases) is neither a reference crash nor
    'no clusters': the mirror of worker_bwt's inner loop raises LibraryLimit, sga does too."""
    import numpy as np
    from bwt_b200 import api, sga, worker_bwt
    from bwt_b200.errors import LibraryLimit, ReferenceCrash

    class FakeIndex:
        def __init__(self, status):
            self.status = status

        def align(self, reads):
            n = len(reads)
            return (np.full(n, self.status, np.int32), np.zeros(n + 1, np.int64), np.zeros(0, api.HIT_DTYPE), b"")

        def sga(self, reads):
            n = len(reads)
            return np.full(n, self.status, np.int32), np.zeros(n + 1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.int32)

    long_read = [("r1", b"ACGT" * 300)]
    with pytest.raises(LibraryLimit):
        worker_bwt.process_list(FakeIndex(api.READ_LIMIT), b"", long_read)
    with pytest.raises(ReferenceCrash):
        worker_bwt.process_list(FakeIndex(api.READ_CRASH), b"", long_read)
    assert worker_bwt.process_list(FakeIndex(api.READ_OK), b"", long_read) == []
    with pytest.raises(LibraryLimit):
        sga.sga_batch(FakeIndex(api.READ_LIMIT), [b"ACGT" * 300])


def test_host_packing_matches_the_layout():
    """bwt_pack_reads (host side of the packed ABI): reversed 2-bit words, lengths, flagged reads with their text; any
    thread count gives the same bytes."""
    import numpy as np
    from bwt_b200 import api
    rng = np.random.default_rng(4)
    reads = []
    for k in range(9000):
        ln = int(rng.integers(0, 140)) if k % 50 else int(rng.integers(1030, 1100))
        r = bytearray(b"ACGT"[int(x)] for x in rng.integers(0, 4, ln))
        if k % 37 == 0 and ln:
            r[int(rng.integers(0, ln))] = b"NRx"[k % 3]
        reads.append(bytes(r))
    data, offs = api.pack(reads)
    pk = api.PackedReads(data, offs, threads=1)
    pk8 = api.PackedReads(data, offs, threads=7)
    assert pk.n_words == pk8.n_words and (pk.words[:pk.n_words + 1] == pk8.words[:pk8.n_words + 1]).all()
    assert (pk.lens == pk8.lens).all() and pk.n_exc == pk8.n_exc and (pk.exc_idx == pk8.exc_idx).all()
    assert bytes(pk.exc_ascii[:pk.exc_bytes]) == bytes(pk8.exc_ascii[:pk8.exc_bytes]) and pk.words[pk.n_words] == 0
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    w = 0
    flagged = []
    for k, r in enumerate(reads):
        assert pk.lens[k] == len(r)
        if len(r) > 1040:
            flagged.append(k)
            continue
        if any(c not in code for c in r):
            flagged.append(k)
        nw = (len(r) + 15) // 16
        if k % 97 == 0:                                   # spot-check the words
            for x in range(nw):
                want = 0
                for j in range(16 * x, min(len(r), 16 * x + 16)):
                    want |= code.get(r[len(r) - 1 - j], 0) << (2 * (j & 15))
                assert int(pk.words[w + x]) == want
        w += nw
    assert w == pk.n_words and list(pk.exc_idx[:pk.n_exc]) == flagged
    for k, r in zip(range(pk.n_exc), flagged):
        assert bytes(pk.exc_ascii[pk.exc_off[k]:pk.exc_off[k + 1]]) == reads[r]
