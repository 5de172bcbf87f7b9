"""BASELINE.json configs[4]: the reference split into N partitions (one per GPU, builder formula of fm_index.erl:113-124),
every GPU runs the bounded-edit search of seeds.erl (bwt_fs_find_seeds_batch: Len = read_len - Defects, Defects = 2) for
ALL reads against its partition, seeds merged on rank 0 by an NCCL all_gather (bwt_b200/partition.py).

  python scripts/bench_inexact.py --ref-bp 400000000 --reads 200000                       # one partition, one GPU
  torchrun --nproc-per-node 8 scripts/bench_inexact.py --ref-bp 3137454466 --reads 200000  # 8 partitions

Prints one JSON line on rank 0: queries/s (whole job, max over ranks, search + gather), per-rank kernel time, search
states and LF steps per query, and the Occ-sector bandwidth of the walk (64 B per LF step) against the random-gather
ceiling measured in the same run.  With --check the merged seeds are verified on rank 0: every seed is a window within
Defects edits of the read (CPU DP from oracle/, the checker) and every read's true locus is covered by a seed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bwt_b200 import api, partition  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bp", type=int, default=400_000_000)
    ap.add_argument("--reads", type=int, default=200_000)
    ap.add_argument("--read-len", type=int, default=250)
    ap.add_argument("--defects", type=int, default=2)
    ap.add_argument("--sub-rate", type=float, default=0.004)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--check", type=int, default=0, help="reads verified on rank 0 with the CPU DP")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    length = a.read_len - a.defects
    # the same synthetic reference and reads on every rank; partition rank+1 indexed on this GPU
    with api.Index.build_synthetic(a.ref_bp, seed=42, device=local) as whole:
        _, _, ref = whole.export(want_L=False, want_SA=False)
        reads, offs, starts = api.sample_reads(whole, a.reads, a.read_len, seed=7, sub_rate=a.sub_rate, want_starts=True)
    ix = partition.build_partition(ref, world, rank + 1, device=local) if world > 1 else api.Index.build(ref, shift=1, device=local)
    gather_gbs = api.peak_gather_gbs(local, 4 << 30) if rank == 0 else None

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        if world > 1:
            return partition.fs_find_seeds_partitioned(ix, reads, offs, length, a.defects, 1024, device="cuda")
        st, so, sd, stats = ix.fs_find_seeds(reads, offs, length, a.defects, 1024)
        return partition.merge_seed_partitions([(st, so, sd, int(ix.info().shift))]), stats
    for _ in range(a.warmup):
        step()
    sync()
    t0 = time.perf_counter()
    ms_k = nodes = lf = 0
    for _ in range(a.steps):
        merged, stats = step()
        ms_k += stats.ms_kernel; nodes += stats.nodes; lf += stats.lf_steps
    sync()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, ms_k], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        st, so, pos, part = merged
        nq = a.reads * a.steps
        out = {"metric": "inexact-search queries/sec (seeds.erl semantics)", "value": nq / float(t[0]), "unit": "queries/s",
               "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "config": {"workload": "%d bp synthetic reference in %d partition(s); %d reads x %d bp, %.1f %% substitutions; "
                                      "Len %d, Defects %d; seeds merged by NCCL all_gather" % (a.ref_bp, world, a.reads, a.read_len,
                                                                                                100 * a.sub_rate, length, a.defects)},
               "kernel_ms_per_step_max_rank": float(t[1]) / a.steps, "kernel_queries_per_s_rank0": nq / (ms_k * 1e-3),
               "states_per_query_rank0": nodes / nq, "lf_steps_per_query_rank0": lf / nq,
               "occ_gbs_rank0": 64.0 * lf / (ms_k * 1e-3) / 1e9, "random_gather_peak_gbs": gather_gbs,
               "frac_of_random_gather": 64.0 * lf / (ms_k * 1e-3) / 1e9 / gather_gbs,
               "seeds_per_query": len(pos) / a.reads, "limit_status": int((st == api.FS_LIMIT).sum())}
        if a.check:
            from oracle.c_oracle import c_fs_scan_dp          # the checker, never the product
            bad = covered = 0
            for r in range(min(a.check, a.reads)):
                q = reads[offs[r]:offs[r + 1]].tobytes()
                ps = pos[so[r]:so[r + 1]]
                covered += bool(((ps >= starts[r] + 1 - a.defects) & (ps <= starts[r] + 1 + a.defects)).any())
                for p in ps:                                  # each seed: its window alone must satisfy fs:matches
                    w = ref[p - 1:p - 1 + length].tobytes()
                    bad += c_fs_scan_dp(q, length, a.defects, w) != [1]
            out["check"] = {"reads": min(a.check, a.reads), "seeds_failing_matches": bad, "true_locus_covered": covered}
            assert bad == 0
        print(json.dumps(out), flush=True)
    ix.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
