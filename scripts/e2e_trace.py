"""One packed e2e run with BWT_TRACE=1: the per-chunk host timeline of the pipelined aligner."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bwt_b200 import api
n = 10_000_000
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ix = api.Index.build_synthetic(3_137_454_466, seed=42, shift=1)
reads, offs, _ = api.sample_reads_edits(ix, n, 100, seed=1042, edit_rate=0.01)
pk = api.PackedReads(reads, offs, threads=16, pinned=True)
al = api.Aligner(ix, chunk, chunk * 116, depth)
st = torch.empty(n, dtype=torch.uint8).pin_memory(); ho = torch.empty(n + 1, dtype=torch.int32).pin_memory()
hits = torch.empty(16 * (2 * n), dtype=torch.uint8).pin_memory(); cig = torch.empty(4 * n, dtype=torch.uint8).pin_memory()
def step():
    return al.run_packed_ptr(pk, st.data_ptr(), ho.data_ptr(), hits.data_ptr(), 2 * n, cig.data_ptr(), 4 * n)
step(); step()
os.environ["BWT_TRACE"] = "1"
t0 = time.perf_counter(); step(); dt = time.perf_counter() - t0
os.environ.pop("BWT_TRACE")
print("traced step: %.2f ms" % (dt * 1e3), file=sys.stderr)
s = al.stats()
print("sum over chunks: total %.2f seed %.2f sw %.2f emit %.2f h2d %.2f d2h %.2f (ms)" % (s.ms_total, s.ms_seed, s.ms_sw, s.ms_emit, s.ms_h2d, s.ms_d2h), file=sys.stderr)
