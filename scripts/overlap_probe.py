"""How much do two half batches on two streams gain over one batch?  (sga_kernel waits on HBM, the extension kernels on the ALU pipe.)"""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bwt_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ix = api.Index.build_synthetic(3_137_454_466, seed=42, shift=1)
reads, offs, _ = api.sample_reads_edits(ix, n, 100, seed=1042, edit_rate=0.01)
one = api.Batch(ix, n, int(offs[n]))
one.upload(reads, offs)
for _ in range(3):
    one.run()
t0 = time.perf_counter()
for _ in range(5):
    one.run()
t_one = (time.perf_counter() - t0) / 5
print("one batch of %d reads: %.2f ms wall, %.2f ms events" % (n, t_one * 1e3, one.stats().ms_total))
one.close()
for parts in (2, 3, 4):
    bs = []
    for p in range(parts):
        lo, hi = n * p // parts, n * (p + 1) // parts
        b = api.Batch(ix, hi - lo, int(offs[hi] - offs[lo]))
        b.upload(reads[offs[lo]:offs[hi]], np.ascontiguousarray(offs[lo:hi + 1] - offs[lo]))
        bs.append(b)
    def go(b, k):
        for _ in range(k):
            b.run()
    for b in bs:
        go(b, 2)
    t0 = time.perf_counter()
    th = [threading.Thread(target=go, args=(b, 5)) for b in bs]
    for t in th: t.start()
    for t in th: t.join()
    t_par = (time.perf_counter() - t0) / 5
    print("%d batches on %d threads: %.2f ms wall per %d reads (x%.2f)" % (parts, parts, t_par * 1e3, n, t_one / t_par))
    for b in bs:
        b.close()
