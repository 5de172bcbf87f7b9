import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bwt_b200 import api
from oracle.c_oracle import c_fs_scan_dp
kat = json.load(open("tests/golden/seeds_kat.json"))
ref = kat["refs"]["toy4"].encode()
cs = [c for c in kat["cases"] if c["ref"] == "toy4" and c["len"] == 4]
qs = [c["q"].encode() for c in cs]
print(qs)
with api.Index.build(ref, shift=1) as ix:
    for sub in (qs, qs[:1], qs[1:2], qs[2:], qs[::-1], qs * 3):
        st, so, sd, s = ix.fs_find_seeds(sub, None, 4, 0, 4096)
        print(sub, "st", st.tolist(), "so", so.tolist(), "seeds", sd.tolist(), "want", [c_fs_scan_dp(q, 4, 0, ref) for q in sub],
              "nodes", s.nodes, "lf", s.lf_steps, "raw", s.raw_hits, "spills", s.spills)
