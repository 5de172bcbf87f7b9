"""Worker-sized calls through bwt_align_batch (the reference's list is 200 reads, master.erl:59; a workload is 10 lists):
wall time per call, stage times and launch count of one run, with the small-batch routing on (default) and off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bwt_b200 import api
from bwt_b200.synth import sample_reads
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ref = open(os.path.join(ROOT, "tests", "golden", "GL000193.1_p1.ref"), "rb").read()
with api.Index.build(ref, shift=1) as ix:
    for n in (200, 2000):
        reads, offs, _ = sample_reads(ref, n, 100, seed=3, edit_rate=0.01)
        for small in ("4096", "0"):
            os.environ["BWT_SW_SMALL"] = small
            for _ in range(5):
                ix.align(reads, offs)
            t0 = time.perf_counter()
            reps = 200
            for _ in range(reps):
                st, ho, hits, cig = ix.align(reads, offs)
            dt = (time.perf_counter() - t0) / reps
            b = api.Batch(ix, n, len(reads)); b.upload(reads, offs)
            for _ in range(3): b.run()
            s = b.stats()
            b.close()
            print("n=%d BWT_SW_SMALL=%s: %.3f ms per bwt_align_batch call (%.2f M reads/s, %d hits) | one run: total %.3f seed %.3f (sga %.3f) sw %.3f (fast %.3f) emit %.3f ms, %d launches, pairs %d shortcut %d general %d"
                  % (n, small, dt * 1e3, n / dt / 1e6, len(hits), s.ms_total, s.ms_seed, s.ms_sga_kernel, s.ms_sw, s.ms_sw_kernel, s.ms_emit, s.launches, s.sw_calls, s.sw_shortcut, s.sw_general))
