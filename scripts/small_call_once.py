"""One worker-sized call (200 reads) after warm-up, for an ncu launch list: which kernels a call launches and how long each runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bwt_b200 import api
from bwt_b200.synth import sample_reads
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = open(os.path.join(ROOT, "tests", "golden", "GL000193.1_p1.ref"), "rb").read()
with api.Index.build(ref, shift=1) as ix:
    reads, offs, _ = sample_reads(ref, int(sys.argv[1]) if len(sys.argv) > 1 else 200, 100, seed=3, edit_rate=0.01)
    for _ in range(3):
        st, ho, hits, cig = ix.align(reads, offs)
    print(len(hits))
