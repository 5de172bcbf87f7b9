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
    assert L.bwt_abi_version() == 3


def test_struct_layouts_match_header():
    from bwt_b200 import _lib, api
    assert C.sizeof(_lib.BwtHit) == 40 and api.HIT_DTYPE.itemsize == 40
    assert C.sizeof(_lib.BwtIndexInfo) == 72
    assert C.sizeof(_lib.BwtHit16) == 16 and api.HIT16_DTYPE.itemsize == 16
    assert C.sizeof(_lib.BwtPackedReads) == 72 and C.sizeof(_lib.BwtSampleOpts) == 72
    assert C.sizeof(_lib.BwtStats) == 9 * 8 + 8 * 4 + 2 * 4 + 2 * 8


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
