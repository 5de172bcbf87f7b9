"""Writes tests/golden/seeds_kat.json: known answers of the seeds.erl restatement (oracle/seeds_oracle.py, the literal
clause list + literal window walk) on self-contained inputs.  The reference has no fixture for seeds.erl and no
Erlang runtime exists here, so these are ORACLE outputs ("parity unpinned"), kept so that the C DP formulation, and
the CUDA path on the GPU box, are held to the same vectors.   python scripts/make_seeds_golden.py"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import seeds_oracle as S  # noqa: E402


def main():
    rng = random.Random(20141)
    cases = []
    with open(os.path.join(ROOT, "tests", "golden", "GL000193.1_p1.ref"), "rb") as f:
        gl = f.read()
    refs = {"toy4": b"TGTGTGATGATGACCCCCCATGA", "gl_head": gl[:3000], "tandem": (gl[5000:5060] * 12) + gl[7000:7400],
            "polyA": b"A" * 80 + b"CGT" + b"A" * 80}
    for name, ref in refs.items():
        for length, defects in ((4, 0), (5, 1), (8, 2), (12, 2), (10, 3)):
            if length + defects + 1 > len(ref):
                continue
            for _ in range(3):
                s = rng.randrange(0, len(ref) - length - defects)
                q = bytearray(ref[s:s + length + defects + rng.randrange(0, 3)])
                for _k in range(rng.randrange(0, 3)):
                    q[rng.randrange(len(q))] = rng.choice(b"ACGTN")
                q = bytes(q)
                pats = S.generate_patterns(q, length, defects)
                cases.append({"ref": name, "q": q.decode(), "len": length, "defects": defects, "n_patterns": len(pats),
                              "seeds": S.fs_find_seeds(pats, length, ref)})
    out = {"refs": {k: v.decode() for k, v in refs.items()}, "cases": cases}
    with open(os.path.join(ROOT, "tests", "golden", "seeds_kat.json"), "w") as f:
        json.dump(out, f, indent=0)
    print(len(cases), "cases")


if __name__ == "__main__":
    main()
