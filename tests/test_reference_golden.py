"""The route from "faithful restatement" to "pinned on the reference's own outputs" (VERDICT r1, item 8).

tests/golden/reference_inputs.txt is a fixed, seeded input set (scripts/make_reference_inputs.py; no oracle involved).
scripts/make_reference_golden.escript runs the UNMODIFIED reference (src/{fmi,bwa,sga,sw}.erl + the worker's fold copied
from worker_bwt.erl:84-106) over it on any box with Erlang/OTP and writes tests/golden/reference_outputs.json.  The build
image has no OTP, so that file may be absent; these tests then still run every input through the oracle (CPU) and
through the CUDA path (GPU) and compare the two, and they compare BOTH with the reference's file the moment it is
committed.

Since the image has no BEAM, scripts/make_reference_golden_interpreted.py ALSO runs the same unmodified source files through
oracle/mini_erl.py, a minimal generic Erlang evaluator (tree walker over the subset the four modules use; OTP's lists /
erlang functions re-implemented from their documented semantics), and commits tests/golden/reference_outputs_interpreted.json:
the oracle and the CUDA path are compared with it here.  That pins them to the reference's own source text -- clause
order, guards, term order of lists:max/1 -- but not to a BEAM run (OTP's lists:sort/2 tie order stays unpinned).
"""
import json
import os

import numpy as np
import pytest

from oracle import bwt_oracle as O
from oracle.c_oracle import c_sw

HERE = os.path.dirname(os.path.abspath(__file__))
INPUTS = os.path.join(HERE, "golden", "reference_inputs.txt")
OUTPUTS = os.path.join(HERE, "golden", "reference_outputs.json")


def load_inputs():
    items = []
    with open(INPUTS, "rb") as f:
        for ln in f.read().split(b"\n"):
            if ln:
                p = ln.split(b" ")
                items.append((p[0].decode(), p[1:]))
    return items


def top_of(cigars):
    """worker_bwt.erl:99-106 on the Cigars list (entries [score, cigar, pos])"""
    if not cigars:
        return "no_cigar"
    if len(cigars) == 1:
        return cigars[0]
    best = max(c[0] for c in cigars)
    return [c for c in cigars if c[0] == best][-1]      # OTP merge sort, strict comparator (recalled; unpinned)


def oracle_item(fm, padded, kind, args):
    """the oracle's answer in the shape the escript writes"""
    if kind == "S":
        try:
            r = O.find_seeds(fm, args[0])
        except O.ReferenceCrash:
            return "crash"
        return r if isinstance(r, str) else [int(x) for x in r]
    if kind == "P":
        r = c_sw(args[0], args[1])
        return r if r == "no_match" else [r[0], r[1]]
    q = args[0]
    try:
        seeds = O.sga(fm, q)
    except O.ReferenceCrash:
        return {"sga": "crash"}
    out = {"sga": [[int(s), int(d)] for s, d in seeds]}
    try:
        _, cigars = O.read_candidates(fm, padded, q, sw_fn=c_sw)
    except O.ReferenceCrash:
        out["cigars"] = out["top"] = "badmatch"
        return out
    out["cigars"] = [[int(c[0][0]), c[0][1], int(c[1])] for c in cigars]
    out["top"] = top_of(out["cigars"])
    return out


@pytest.fixture(scope="module")
def oracle_outputs(gl_fm, gl_ref):
    padded = O.pad_ref(gl_ref)
    return [oracle_item(gl_fm, padded, k, a) for k, a in load_inputs()]


def reference_outputs():
    if not os.path.exists(OUTPUTS):
        return None
    with open(OUTPUTS) as f:
        doc = json.load(f)
    items = []
    for it in doc["items"]:
        if it["kind"] in ("S", "P"):
            items.append(it["out"])
        else:
            items.append({k: v for k, v in it.items() if k in ("sga", "cigars", "top")})
    return items


INTERPRETED = os.path.join(HERE, "golden", "reference_outputs_interpreted.json")


def interpreted_outputs():
    """{input index: output} of the reference's own source files run by oracle/mini_erl.py
    (scripts/make_reference_golden_interpreted.py); None when the file is absent"""
    if not os.path.exists(INTERPRETED):
        return None
    with open(INTERPRETED) as f:
        doc = json.load(f)
    out = {}
    for it in doc["items"]:
        out[it["index"]] = it["out"] if it["kind"] in ("S", "P") else {k: v for k, v in it.items() if k in ("sga", "cigars", "top")}
    return out


def test_oracle_equals_the_interpreted_reference(oracle_outputs):
    """the hand-written oracle against the UNMODIFIED src/{fmi,bwa,sga,sw}.erl executed by a generic Erlang evaluator:
    every committed input (bwa:find_seeds/6, sga:sga/6 + the worker's fold, sw:sw/2)"""
    got = interpreted_outputs()
    assert got is not None, "tests/golden/reference_outputs_interpreted.json is part of the repository"
    items = load_inputs()
    assert len(got) == len(items)
    kinds = {"S": 0, "R": 0, "P": 0}
    for k, want in got.items():
        assert oracle_outputs[k] == want, (k, items[k])
        kinds[items[k][0]] += 1
    assert kinds["S"] >= 400 and kinds["R"] >= 700 and kinds["P"] >= 1100


@pytest.mark.skipif(not os.path.exists("/root/reference/src/sw.erl"), reason="the reference tree only exists in the build container")
def test_interpreted_file_is_what_the_evaluator_says_about_the_reference_sources():
    """re-run a sample of the inputs through oracle/mini_erl.py on /root/reference's own files: the committed JSON is
    reproducible here, not hand-edited"""
    import subprocess
    import sys
    import tempfile
    got = interpreted_outputs()
    script = os.path.join(os.path.dirname(HERE), "scripts", "make_reference_golden_interpreted.py")
    for first, last in ((0, 40), (407, 411), (1129, 1150), (1700, 1740)):
        with tempfile.TemporaryDirectory() as tmp:
            env = dict(os.environ, BWT_GOLDEN_OUT=os.path.join(tmp, "out.json"))
            subprocess.check_call([sys.executable, script, "/root/reference", str(first), str(last)], env=env, stdout=subprocess.DEVNULL)
            with open(os.path.join(tmp, "out.json")) as f:
                doc = json.load(f)
        for it in doc["items"]:
            want = it["out"] if it["kind"] in ("S", "P") else {k: v for k, v in it.items() if k in ("sga", "cigars", "top")}
            assert got[it["index"]] == want, it["index"]


def test_inputs_cover_every_kind_and_edge():
    items = load_inputs()
    kinds = {k for k, _ in items}
    assert kinds == {"S", "R", "P"}
    assert sum(k == "R" for k, _ in items) >= 700 and sum(k == "P" for k, _ in items) >= 500
    assert any(k == "R" and b"N" in a[0] for k, a in items)
    assert any(k == "P" and any(c in a[1] for c in b"RYKMSWBDV") for k, a in items)


def test_oracle_answers_every_input(oracle_outputs):
    """every clause family shows up in the oracle's answers, so the reference file will exercise them all"""
    items = load_inputs()
    s_out = [o for (k, _), o in zip(items, oracle_outputs) if k == "S"]
    assert "too_many_seeds" in s_out and "crash" in s_out and any(isinstance(o, list) and len(o) > 1 for o in s_out)
    r_out = [o for (k, _), o in zip(items, oracle_outputs) if k == "R"]
    assert sum(1 for o in r_out if isinstance(o.get("cigars"), list) and len(o["cigars"]) > 1) >= 20     # the top-1 pick matters
    ties = [o for o in r_out if isinstance(o.get("cigars"), list) and len(o["cigars"]) > 1
            and sum(c[0] == max(x[0] for x in o["cigars"]) for c in o["cigars"]) > 1]
    assert ties, "no equal-score top-1 tie in the input set"
    assert any("I" in c[1] or "D" in c[1] for o in r_out if isinstance(o.get("cigars"), list) for c in o["cigars"])
    p_out = [o for (k, _), o in zip(items, oracle_outputs) if k == "P"]
    assert "no_match" in p_out and any(o != "no_match" and "I" in o[1] for o in p_out)


def test_oracle_equals_reference_outputs(oracle_outputs):
    ref = reference_outputs()
    if ref is None:
        pytest.skip("tests/golden/reference_outputs.json not generated yet (needs Erlang/OTP: scripts/make_reference_golden.escript)")
    assert len(ref) == len(oracle_outputs)
    for k, (want, got) in enumerate(zip(ref, oracle_outputs)):
        assert got == want, (k, load_inputs()[k])


@pytest.mark.gpu
def test_cuda_path_equals_oracle_and_reference(gl_fm, gl_ref, oracle_outputs):
    """the same inputs through the C ABI: search + locate, align (sga + windows + sw + Cigars), sw_batch"""
    from bwt_b200 import api
    items = load_inputs()
    ref = reference_outputs()
    got = [None] * len(items)
    with api.Index.from_arrays(gl_fm.L, gl_fm.SA, gl_fm.pointers, gl_fm.shift, gl_ref) as ix:
        s_idx = [k for k, (kd, _) in enumerate(items) if kd == "S"]
        st, sp, ep, n = ix.search([items[k][1][0] for k in s_idx])
        for j, k in enumerate(s_idx):
            if st[j] == 3:
                got[k] = "crash"
            elif st[j] == 2:
                got[k] = "too_many_seeds"
            elif st[j] == 1:
                got[k] = "no_seeds"
            else:
                sa = ix.locate(int(sp[j]), int(ep[j] - sp[j] + 1))
                got[k] = [int(x) + int(n[j]) for x in sa[::-1]]              # fmi.erl:89-100
        r_idx = [k for k, (kd, _) in enumerate(items) if kd == "R"]
        reads = [items[k][1][0] for k in r_idx]
        rst, clo, S, D = ix.sga(reads)
        ast, ho, hits, cig = ix.align(reads)
        for j, k in enumerate(r_idx):
            if rst[j] == api.READ_CRASH:
                got[k] = {"sga": "crash"}
                continue
            assert rst[j] == api.READ_OK
            o = {"sga": [[int(s), int(d)] for s, d in zip(S[clo[j]:clo[j + 1]], D[clo[j]:clo[j + 1]])]}
            hs = hits[ho[j]:ho[j + 1]]
            if any(h["status"] == api.HIT_WINDOW_OOR for h in hs):
                o["cigars"] = o["top"] = "badmatch"
            else:
                o["cigars"] = [[int(h["score"]), cig[int(h["cigar_off"]):int(h["cigar_off"]) + int(h["cigar_len"])].decode(), int(h["pos"])]
                               for h in hs if h["status"] == api.HIT_MATCH]
                o["top"] = top_of(o["cigars"])
            got[k] = o
    p_idx = [k for k, (kd, _) in enumerate(items) if kd == "P"]
    q, qo = api.pack([items[k][1][0] for k in p_idx])
    r, ro = api.pack([items[k][1][1] for k in p_idx])
    pst, sc, pc = api.sw_batch(q, qo, r, ro)
    for j, k in enumerate(p_idx):
        got[k] = "no_match" if pst[j] == api.HIT_NO_MATCH else [int(sc[j]), pc[j].decode()]
    # the same pairs in the routing of large batches (a call this small takes the warp-per-pair kernel otherwise): quick
    # test, wavefront rule, banded table in two passes, full-table kernel
    os.environ["BWT_SW_SMALL"] = "0"
    try:
        pst2, sc2, pc2 = api.sw_batch(q, qo, r, ro)
    finally:
        os.environ.pop("BWT_SW_SMALL", None)
    for j, k in enumerate(p_idx):
        assert got[k] == ("no_match" if pst2[j] == api.HIT_NO_MATCH else [int(sc2[j]), pc2[j].decode()]), (k, items[k])
    interp = interpreted_outputs()
    for k in range(len(items)):
        assert got[k] == oracle_outputs[k], (k, items[k])
        if ref is not None:
            assert got[k] == ref[k], (k, items[k])
        if interp is not None and k in interp:
            assert got[k] == interp[k], (k, items[k])
