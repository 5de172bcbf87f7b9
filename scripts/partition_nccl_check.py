"""2-GPU check of the genome-partition mode: rank r holds partition r+1, every rank aligns all reads, hit records
are gathered with NCCL; rank 0 compares the merged result with the two partitions aligned one after the other on
its own GPU.   torchrun --nproc-per-node 2 scripts/partition_nccl_check.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bwt_b200 import api, partition  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_ref, n_reads, rl = 64_000_000, 1_000_000, 250
# the same synthetic reference on every rank (device generator, then exported), partition r+1 indexed on rank r
with api.Index.build_synthetic(n_ref, seed=42, device=local) as whole:
    _, _, ref = whole.export(want_L=False, want_SA=False)
    reads, offs, _ = api.sample_reads(whole, n_reads, rl, seed=7, sub_rate=0.01)
ix = partition.build_partition(ref, world, rank + 1, device=local)
al = api.Aligner(ix, 1 << 19, (1 << 19) * rl, 2)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
merged = partition.align_partitioned(ix, reads, offs, device="cuda", aligner=al)
torch.cuda.synchronize(); dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    parts = []
    for j in range(1, world + 1):
        with partition.build_partition(ref, world, j, device=local) as pj:
            st, ho, hits, cig = pj.align(reads, offs)
            parts.append((st, ho, hits, np.frombuffer(cig, np.uint8)))
    want = partition.merge_partitions(parts)
    assert (merged[1] == want[1]).all() and (merged[2] == want[2]).all() and merged[3] == want[3] and (merged[0] == want[0]).all()
    print("partition mode ok: %d partitions, %d reads x %d bp, %d merged hits, %.1f ms (%.2f M reads/s incl. NCCL gather)" % (
        world, n_reads, rl, len(merged[2]), dt * 1e3, n_reads / dt / 1e6))
al.close(); ix.close()
dist.barrier(); dist.destroy_process_group()
