import os, sys
sys.path.insert(0, "/root/repo")
from bwt_b200 import api
from bwt_b200.synth import sample_reads
ref = open("/root/repo/tests/golden/GL000193.1_p1.ref","rb").read()
with api.Index.build(ref, shift=1) as ix:
    reads, offs, _ = sample_reads(ref, 2000, 100, seed=3, edit_rate=0.01)
    b = api.Batch(ix, 2000, len(reads)); b.upload(reads, offs)
    for _ in range(3): b.run()
    s = b.stats()
    print("ms_total %.3f seed %.3f sw %.3f emit %.3f sw_fast %.3f | pairs %d shortcut %d general %d" % (s.ms_total, s.ms_seed, s.ms_sw, s.ms_emit, s.ms_sw_kernel, s.sw_calls, s.sw_shortcut, s.sw_general))
    b.close()
