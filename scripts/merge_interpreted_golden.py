#!/usr/bin/env python
"""Merges the range files scripts/make_reference_golden_interpreted.py writes when it is run in parallel over slices of
tests/golden/reference_inputs.txt into tests/golden/reference_outputs_interpreted.json (one item per input, by index)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
base = os.path.join(ROOT, "tests", "golden", "reference_outputs_interpreted.json")
items, head = {}, None
if os.path.exists(base):                                    # an earlier merge: new ranges extend it
    with open(base) as f:
        head = json.load(f)
    items = {it["index"]: it for it in head["items"]}
for path in sorted(glob.glob(base + ".*-*")):
    with open(path) as f:
        doc = json.load(f)
    head = head or doc
    for it in doc["items"]:
        if it["index"] in items:
            assert items[it["index"]] == it, ("two runs disagree on input", it["index"])
        items[it["index"]] = it
n = sum(1 for ln in open(os.path.join(ROOT, "tests", "golden", "reference_inputs.txt"), "rb").read().split(b"\n") if ln)
missing = [k for k in range(n) if k not in items]
assert not missing, "inputs without an output: %s ..." % missing[:10]
head = {k: v for k, v in head.items() if k not in ("items", "first", "last")}
head["items"] = [items[k] for k in range(n)]
with open(base, "w") as f:
    json.dump(head, f, separators=(",", ":"))
print("%d items -> %s (%d bytes)" % (n, base, os.path.getsize(base)))
