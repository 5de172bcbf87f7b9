#!/usr/bin/env python
"""Runs the reference's UNMODIFIED src/{fmi,bwa,sga,sw}.erl (+ include/bwt.hrl), read from the reference tree, through
oracle/mini_erl.py (a minimal Erlang evaluator; the build image has no OTP) on the shipped GL000193.1 index over
tests/golden/reference_inputs.txt, and writes tests/golden/reference_outputs_interpreted.json in the shape
scripts/make_reference_golden.escript writes on a box with a real BEAM.

    python scripts/make_reference_golden_interpreted.py /root/reference [first_item [last_item]]

The worker's fold around sga / sw (window cut, Cigars list, top-1 pick: worker_bwt.erl:84-106) is done here in Python, as
the escript does it in Erlang; everything inside bwa:find_seeds/6, sga:sga/6 and sw:sw/2 is the reference's own text.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bwt_oracle as O          # noqa: E402  (only parse_fm: the .fm container reader)
from oracle import mini_erl as E            # noqa: E402

ref_root = sys.argv[1]
inputs = os.path.join(ROOT, "tests", "golden", "reference_inputs.txt")
out_path = os.path.join(ROOT, "tests", "golden", "reference_outputs_interpreted.json")

with open(os.path.join(ref_root, "fm_indices", "GL000193.1_p1.fm"), "rb") as f:
    fm = O.parse_fm(f.read())                                            # fm_index.erl:154-156: binary_to_term
with open(os.path.join(ref_root, "fm_indices", "GL000193.1_p1.ref"), "rb") as f:
    ref0 = f.read()
FM = tuple((int(a), int(b), int(c), int(d)) for a, b, c, d in zip(fm.F, fm.L, fm.I, fm.SA))
Pc, Pg, Pt, Last = (int(x) for x in fm.pointers)
Shift = int(fm.shift)
Ref = b"N" * 200 + ref0 + b"N" * 200                                    # worker_bwt.erl:17,53-55

m = E.machine_for_reference(ref_root)
NO_SEEDS, TOO_MANY, NO_MATCH = E.atom("no_seeds"), E.atom("too_many_seeds"), E.atom("no_match")


def cigar_of(res):
    return [int(res[0]), bytes(E.to_pylist(res[1])).decode()]


def handle(kind, args):
    if kind == "S":
        try:
            r = m.call("bwa", "find_seeds", [FM, Pc, Pg, Pt, Last, E.from_bytes(args[0])])
        except (E.ErlError, IndexError):
            return "crash"
        if r is NO_SEEDS:
            return "no_seeds"
        if r is TOO_MANY:
            return "too_many_seeds"
        return [int(x) for x in E.to_pylist(r)]
    if kind == "P":
        try:
            r = m.call("sw", "sw", [E.from_bytes(args[0]), E.from_bytes(args[1])])
        except E.ErlError:
            return "crash"
        return "no_match" if r is NO_MATCH else cigar_of(r)
    q = args[0]
    try:
        seeds = E.to_pylist(m.call("sga", "sga", [FM, Pc, Pg, Pt, Last, E.from_bytes(q)]))
    except (E.ErlError, IndexError):
        return {"sga": "crash"}
    out = {"sga": [[int(s), int(d)] for (s, d) in seeds]}
    cigars = []
    try:
        for (s, d) in seeds:                                             # worker_bwt.erl:84-97
            ref_len = len(q) + d
            start = s - ref_len + 200
            if start < 0 or start + ref_len > len(Ref):
                raise E.ErlError("badmatch")
            window = Ref[start:start + ref_len]
            r = m.call("sw", "sw", [E.from_bytes(q), E.from_bytes(window)])
            if r is not NO_MATCH:
                cigars.insert(0, cigar_of(r) + [int(s - ref_len + Shift)])
    except E.ErlError:
        out["cigars"] = out["top"] = "badmatch"
        return out
    out["cigars"] = cigars
    if not cigars:
        out["top"] = "no_cigar"
    elif len(cigars) == 1:
        out["top"] = cigars[0]
    else:                                                                # lists:sort/2 with a strict comparator: OTP's order among equal scores, unpinned
        best = max(c[0] for c in cigars)
        out["top"] = [c for c in cigars if c[0] == best][-1]
    return out


lines = [ln for ln in open(inputs, "rb").read().split(b"\n") if ln]
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else len(lines)
items = []
t0 = time.time()
for k in range(first, last):
    p = lines[k].split(b" ")
    kind = p[0].decode()
    res = handle(kind, p[1:])
    if kind in ("S", "P"):
        items.append({"index": k, "kind": kind, "in": [x.decode() for x in p[1:]] if kind == "P" else p[1].decode(), "out": res})
    else:
        items.append(dict({"index": k, "kind": "R", "in": p[1].decode()}, **res))
    if (k - first) % 50 == 0:
        print("%d / %d items, %.0f s, %d evaluator steps" % (k - first, last - first, time.time() - t0, m.steps), flush=True)
doc = {"evaluator": "oracle/mini_erl.py over the unmodified src/{fmi,bwa,sga,sw}.erl + include/bwt.hrl of the reference (no BEAM in the build image)",
       "otp_release": None, "index": "fm_indices/GL000193.1_p1", "inputs": "tests/golden/reference_inputs.txt", "first": first, "last": last,
       "items": items}
suffix = "" if (first, last) == (0, len(lines)) else ".%d-%d" % (first, last)
if os.environ.get("BWT_GOLDEN_OUT"):                                     # tests: a sample range into a scratch file
    out_path, suffix = os.environ["BWT_GOLDEN_OUT"], ""
with open(out_path + suffix, "w") as f:
    json.dump(doc, f, separators=(",", ":"))
print("%d items -> %s in %.0f s" % (len(items), out_path + suffix, time.time() - t0))
