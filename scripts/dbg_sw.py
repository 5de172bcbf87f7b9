import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bwt_b200 import api
from oracle.c_oracle import c_sw
def run(pairs, tag):
    t0 = time.time()
    q, qo = api.pack([p[0] for p in pairs]); r, ro = api.pack([p[1] for p in pairs])
    st, sc, cig = api.sw_batch(q, qo, r, ro)
    bad = 0
    for k, (a, b) in enumerate(pairs):
        w = c_sw(a, b)
        got = "no_match" if st[k] == 1 else (int(sc[k]), cig[k].decode())
        bad += got != w
    print(tag, len(pairs), "pairs, mismatches", bad, "%.2fs" % (time.time() - t0), flush=True)
rng = np.random.default_rng(1)
rs = lambda n: bytes(b"ACGT"[int(x)] for x in rng.integers(0, 4, n))
q = rs(100)
run([(q, q)], "one perfect")
run([(q, q), (q, rs(3) + q)], "two")
run([(q, q)] * 4, "four")
run([(q, q)] * 5, "five")
run([(q, rs(k % 7) + q + rs(k % 3)) for k in range(64)], "64 shifted")
run([(rs(100), rs(104)) for k in range(64)], "64 random")
qs = [rs(int(rng.integers(4, 130))) for _ in range(3000)]
run([(a, rs(int(rng.integers(0, 9))) + a + rs(int(rng.integers(0, 9)))) for a in qs], "3000 mixed lengths")
