"""The 'N' workload in isolation (A/B of the double-precision kernel): 400 Mbp synthetic index, 1 M reads of 100 bp with the
8(d) edit model and 1 % of the bases turned into 'N'.  Prints stage times and the general-pair rate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bwt_b200 import api
n = 1_000_000
ix = api.Index.build_synthetic(400_000_000, seed=42, shift=1, device=0)
rbuf = np.empty(n * 100 + 200000, np.uint8); offs = np.empty(n + 1, np.int64)
reads, _, _ = api.sample_reads_edits(ix, n, 100, seed=7, edit_rate=0.01, ins_share=0.1, del_share=0.1, n_rate=0.01, device=0, out=rbuf, offsets=offs)
b = api.Batch(ix, n, int(offs[n]))
b.upload(reads, offs)
for _ in range(3):
    b.run()
s = b.stats()
print("%s: total %.2f ms seed %.2f (sga %.2f) sw %.2f emit %.2f | pairs %d general %d -> %.3f us per general pair, %.1f M reads/s"
      % (os.environ.get("BWT_B200_LIB", "default lib"), s.ms_total, s.ms_seed, s.ms_sga_kernel, s.ms_sw, s.ms_emit, s.sw_calls, s.sw_general,
         s.ms_sw * 1e3 / max(1, s.sw_general), n / s.ms_total / 1e3))
b.close(); ix.close()
