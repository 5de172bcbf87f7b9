"""bwt_align_batch latency for worker-sized calls (the Erlang worker sends 10 lists x 200 reads per workload):
GL000193.1, 2000 reads x 100 bp per call, 100 calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bwt_b200 import api
from bwt_b200.synth import sample_reads
ref = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "GL000193.1_p1.ref"), "rb").read()
with api.Index.build(ref, shift=1) as ix:
    reads, offs, _ = sample_reads(ref, 2000, 100, seed=3, edit_rate=0.01)
    ix.align(reads, offs)
    t0 = time.perf_counter()
    for _ in range(100):
        st, ho, hits, cig = ix.align(reads, offs)
    dt = (time.perf_counter() - t0) / 100
    print("bwt_align_batch, 2000 reads per call: %.3f ms per call, %.2f M reads/s, %d hits" % (dt * 1e3, 2000 / dt / 1e6, len(hits)))
