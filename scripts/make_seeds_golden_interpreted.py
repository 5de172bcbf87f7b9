#!/usr/bin/env python
"""Row A9 / F4 pinned on the reference's source text: runs the UNMODIFIED src/seeds.erl through oracle/mini_erl.py --
seeds:generate_fs(Qsec, Len, Defects) WRITES the module fs.erl (seeds.erl:13-47); the evaluator then loads that text and
runs fs:find_seeds(Ref_Seq) -- on the inputs of tests/golden/seeds_kat.json, and writes
tests/golden/seeds_kat_interpreted.json.   python scripts/make_seeds_golden_interpreted.py /root/reference [first [last]]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mini_erl as E            # noqa: E402

ref_root = sys.argv[1]
m = E.Machine(os.path.join(ref_root, "src"), os.path.join(ref_root, "include"))
m.load("seeds")
with open(os.path.join(ROOT, "tests", "golden", "seeds_kat.json")) as f:
    kat = json.load(f)
first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
last = int(sys.argv[3]) if len(sys.argv) > 3 else len(kat["cases"])
out = []
for k in range(first, last):
    c = kat["cases"][k]
    m.devices.clear()
    m.call("seeds", "generate_fs", [E.from_bytes(c["q"].encode()), c["len"], c["defects"]])
    text = m.devices[-1].text()                                          # the fs.erl the reference would have written
    m.load_text("fs", text)
    seeds = [int(x) for x in E.to_pylist(m.call("fs", "find_seeds", [E.from_bytes(kat["refs"][c["ref"]].encode())]))]
    out.append({"index": k, "ref": c["ref"], "q": c["q"], "len": c["len"], "defects": c["defects"],
                "true_clauses": text.count("-> true"), "seeds": seeds})
    print(k, c["ref"], c["len"], c["defects"], len(seeds), flush=True)
path = os.environ.get("BWT_GOLDEN_OUT") or os.path.join(ROOT, "tests", "golden", "seeds_kat_interpreted.json")
with open(path, "w") as f:
    json.dump({"evaluator": "oracle/mini_erl.py over the unmodified src/seeds.erl and the fs.erl it writes (no BEAM in the build image)",
               "inputs": "tests/golden/seeds_kat.json", "cases": out}, f, indent=0)
print(len(out), "cases ->", path)
