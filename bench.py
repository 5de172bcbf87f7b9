#!/usr/bin/env python
"""bench.py -- aligned reads/s of the worker_bwt hot path (sga -> windows -> sw -> hit tuples) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU port of the reference (oracle/), all host cores

Workload (BASELINE.json configs[2], the largest single-GPU configuration): one un-chunked FM index of
a 3,137,454,466 bp synthetic reference (the bp total of bwt_files/human_g1k_v37_decoy.fasta.index, i.i.d. ACGT),
built on the device; 10 M synthetic 100 bp forward-strand reads per GPU with 1 % substitutions.  A step = one pass of
the whole path over all of the rank's reads.  At N > 1 the index is replicated and every rank gets its own 10 M reads
(read sharding, no data-path collective): weak scaling.

  value : reads/s with the reads already resident in HBM (bwt_batch_run only), CUDA-event time, max over ranks
  e2e   : reads/s through the C ABI with HOST buffers: pinned reads -> H2D -> run -> D2H of the hit tuples
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_BP = 3_137_454_466
METRIC = "aligned reads/sec"
UNIT = "reads/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-bp", type=int, default=REF_BP)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--sub-rate", type=float, default=0.01, help="per-base edit probability")
    ap.add_argument("--edits", default="indel", choices=["indel", "sub"],
                    help="indel: SURVEY.md 8(d) model, 80/10/10 substitution/insertion/deletion, ragged lengths; sub: substitutions only")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--ref-step-reads", type=int, default=200_000, help="reads per step of --impl reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the untimed peak / LF-walk side measurements")
    ap.add_argument("--chunk-reads", type=int, default=1 << 20, help="reads per pipelined chunk of the e2e leg")
    ap.add_argument("--depth", type=int, default=2, help="chunks in flight in the e2e leg")
    ap.add_argument("--inexact-queries", type=int, default=50_000, help="queries of the untimed bounded-edit side leg (0 = skip)")
    return ap.parse_args()


def workload_name(a):
    edits = ("%.1f %% edits (80 %% substitution / 10 %% insertion / 10 %% deletion, lengths %d +- few)" % (100 * a.sub_rate, a.read_len)
             if a.edits == "indel" else "%.1f %% substitutions" % (100 * a.sub_rate))
    return ("synthetic %.3f Gbp reference (human_g1k_v37_decoy bp total), un-chunked FM index; %d synthetic %d bp reads "
            "per GPU, %s; sga + sw + hit tuples" % (a.ref_bp / 1e9, a.reads, a.read_len, edits))


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def mark(self):
        """number of samples so far: call at the start and at the end of the timed region"""
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.t.join(timeout=2)
        sm, mx, reasons, pw = [], [], set(), []
        hi = len(self.rows) if hi is None else hi
        # samples taken inside the timed region; the loop period can exceed a short region: then the samples next to it
        rows = self.rows[lo:hi] if hi > lo else self.rows[max(0, lo - 1):hi + 1]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy bandwidth, of measured)"
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def load_traffic(kernel):
    """per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture of this same workload"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        e = t.get(kernel)
        if e:
            return e
    return None


# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_setup(ix, a):
    """Index columns back on the host for the CPU port (oracle/bwt_oracle.c, closed-form LF: same results as the literal scan)."""
    from oracle.c_oracle import COracle
    t0 = time.time()
    L, SA, ref = ix.export()
    info = ix.info()
    orc = COracle(L, SA, tuple(info.pointers), int(info.shift), ref, mode=1)
    return orc, time.time() - t0


def cpu_port_run(orc, reads, offs, lo, hi, threads, cap=16):
    """-> (seconds for reads [lo, hi), stats, ncand, candidates)"""
    sub_offs = np.ascontiguousarray(offs[lo:hi + 1])
    t0 = time.perf_counter()
    ncand, cand, st = orc.align_batch(reads, sub_offs, nthreads=threads, with_sw=True, cap_per_read=cap)
    return time.perf_counter() - t0, st, ncand, cand


ORC_DTYPE = np.dtype([("S", "<i8"), ("D", "<i8"), ("pos", "<i8"), ("score", "<i8"), ("status", "<i4"), ("cigar_len", "<i4"),
                      ("cigar", "u1", (96,))])


def parity_compare(m, ncand, cand, cap, st_g, ho, hits, cig):
    """The CPU port's candidate lists of reads [0, m) against the GPU hits of the same reads: counts, S, D, status and --
    for matches -- score, Pos and the CIGAR text.  Vectorised; -> number of reads + candidates that differ."""
    cand = np.frombuffer(cand, dtype=ORC_DTYPE)
    cnt_c = np.asarray(ncand[:m], np.int64)
    cnt_g = np.diff(ho[:m + 1])
    crash = cnt_c < 0
    bad_reads = int(np.count_nonzero(crash != (st_g[:m] == 1)))
    ok = ~crash
    bad_reads += int(np.count_nonzero(cnt_g[ok] != cnt_c[ok]))
    agree = ok & (cnt_g == cnt_c)
    use = np.where(agree, np.minimum(cnt_c, cap), 0)
    tot = int(use.sum())
    if tot == 0:
        return bad_reads, 0
    rr = np.repeat(np.arange(m, dtype=np.int64), use)
    k = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(use) - use, use)
    Cc = cand[rr * cap + k]                                   # the port lists sga's order (descending S) ...
    Gg = hits[ho[rr] + (cnt_c[rr] - 1 - k)]                   # ... the ABI the reference's Cigars order (ascending)
    bad = (Cc["S"] != Gg["s"]) | (Cc["D"] != Gg["d"]) | (Cc["status"] != Gg["status"])
    mt = (Cc["status"] == 0) & ~bad
    bad |= mt & ((Cc["score"] != Gg["score"]) | (Cc["pos"] != Gg["pos"]) | (Cc["cigar_len"] != Gg["cigar_len"]))
    mt &= ~bad
    idx = np.nonzero(mt)[0]
    if len(idx):
        L = np.minimum(Cc["cigar_len"][idx], 96)
        off = Gg["cigar_off"][idx]
        cg = Cc["cigar"][idx]
        cig = np.frombuffer(cig, np.uint8) if not isinstance(cig, np.ndarray) else cig
        for t in range(int(L.max())):
            sel = L > t
            bad[idx[sel]] |= cg[sel, t] != cig[off[sel] + t]
    return bad_reads + int(np.count_nonzero(bad)), tot


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and rank != 0:
        return 0

    from bwt_b200 import api          # raises if libbwt_b200.so is missing: no CPU fallback

    import torch
    dist = None
    if world > 1 and a.impl == "ours":
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = local if a.impl == "ours" else 0

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # -- setup (untimed): index on the device, reads in pinned host memory
    t0 = time.time()
    ix = api.Index.build_synthetic(a.ref_bp, seed=a.seed, shift=1, device=dev)
    build_s = time.time() - t0
    n = a.reads
    reads_t = torch.empty(n * a.read_len + max(4096, n // 8), dtype=torch.uint8).pin_memory()
    offs_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    offs = offs_t.numpy()
    if a.edits == "indel":
        reads, _, _ = api.sample_reads_edits(ix, n, a.read_len, seed=a.seed + 1000 * (rank + 1), edit_rate=a.sub_rate,
                                             out=reads_t.numpy(), offsets=offs)
    else:
        reads = reads_t.numpy()[:n * a.read_len]
        api.sample_reads(ix, n, a.read_len, seed=a.seed + 1000 * (rank + 1), sub_rate=a.sub_rate, out=reads)
        offs[:] = np.arange(n + 1, dtype=np.int64) * a.read_len
    read_bytes = int(offs[n])
    threads = os.cpu_count() or 1

    if a.impl == "reference":
        return run_reference(a, ix, reads, offs, threads)

    # -- untimed side measurements on rank 0: roofline denominators measured on this device, and the FM walk alone
    extra = {}
    if rank == 0 and not a.no_side:
        extra["gather_gbs"] = api.peak_gather_gbs(dev, 8 << 30)          # random 32-byte sectors over 8 GiB
        extra["dpx_gops"] = api.peak_dpx_gops(dev)
        # LF steps/s of bwa:find_seeds alone: 15-mers cut from the reads, 4^8 table only (every later step reads two
        # Occ sectors), the walk the deep table replaces in the aligned-reads path
        k_auto = ix.info().kmer_k
        ix.set_kmer_depth(8)
        m = min(n, 4_000_000)
        subs = np.ascontiguousarray(reads[(offs[1:m + 1] - 15)[:, None] + np.arange(15)[None, :]]).reshape(-1)   # the last 15 bases of each read
        ms, lf, hbm_lf = api.search_resident(ix, subs, 15, repeats=3)
        ix.set_kmer_depth(k_auto)
        gbs = 64.0 * hbm_lf * 3 / (ms * 1e-3) / 1e9
        extra["lf_walk"] = {"subsequences": m, "lf_steps_per_s": lf * 3 / (ms * 1e-3), "occ_steps_per_s": hbm_lf * 3 / (ms * 1e-3),
                            "achieved_gbs": round(gbs, 1), "frac_of_hbm_copy_peak": round(gbs / load_peaks()[0], 4),
                            "frac_of_random_gather": round(gbs / extra["gather_gbs"], 4), "kmer_depth": 8,
                            "note": "64 B (two Occ sectors) per step beyond the 8-mer table"}

        # the bounded-edit search of seeds.erl (row A9/F4) on the same index: Len = read_len - 2, Defects = 2 on a
        # sample of the reads; LF steps/s of a walk where every step reads two Occ sectors (no k-mer table)
        mq = min(n, a.inexact_queries)
        if mq > 0:
            t0 = time.perf_counter()
            fst, fso, fsd, fss = ix.fs_find_seeds(reads[:offs[mq]], offs[:mq + 1], a.read_len - 4, 2, 1024)
            wall = time.perf_counter() - t0
            extra["inexact"] = {"queries": mq, "len": a.read_len - 4, "defects": 2, "kernel_ms": fss.ms_kernel,
                                "queries_per_s_kernel": mq / (fss.ms_kernel * 1e-3), "queries_per_s_call": mq / wall,
                                "states_per_query": fss.nodes / mq, "lf_steps_per_s": fss.lf_steps / (fss.ms_kernel * 1e-3),
                                "occ_gbs": round(64.0 * fss.lf_steps / (fss.ms_kernel * 1e-3) / 1e9, 1),
                                "frac_of_random_gather": round(64.0 * fss.lf_steps / (fss.ms_kernel * 1e-3) / 1e9 / extra["gather_gbs"], 4),
                                "seeds_per_query": len(fsd) / mq, "limit_status": int((fst == api.FS_LIMIT).sum()),
                                "note": "fs_search_kernel: warp-per-query DFS over gen_pattern/5 (seeds.erl:51-62), 64 B per LF step"}

    batch = api.Batch(ix, n, read_bytes)
    batch.upload(reads, offs)

    # -- resident: warm-up, then K timed steps (CUDA events inside the library, on the batch's own stream)
    clocks = ClockSampler(dev)
    clocks.start()                    # before the warm-up: nvidia-smi needs a moment to deliver its first sample
    for _ in range(max(a.warmup, 0)):
        batch.run()
    barrier()
    c_lo = clocks.mark()
    w0 = time.perf_counter()
    ev_ms = 0.0
    agg = dict(lf=0, hbm_lf=0, located=0, searches=0, clusters=0, sw_calls=0, cells=0, general=0, launches=0, sc_calls=0, sc_cells=0,
               ms_sga=0.0, ms_sw=0.0, ms_seed=0.0, ms_swstage=0.0, ms_emit=0.0)
    for _ in range(a.steps):
        batch.run()
        s = batch.stats()
        ev_ms += s.ms_total
        agg["lf"] += s.lf_steps; agg["hbm_lf"] += s.hbm_lf_steps; agg["located"] += s.located_rows
        agg["searches"] += s.searches; agg["clusters"] += s.clusters; agg["sw_calls"] += s.sw_calls
        agg["cells"] += s.sw_cells; agg["general"] += s.sw_general; agg["launches"] += s.launches
        agg["sc_calls"] += s.sw_shortcut; agg["sc_cells"] += s.sw_shortcut_cells
        agg["ms_sga"] += s.ms_sga_kernel; agg["ms_sw"] += s.ms_sw_kernel
        agg["ms_seed"] += s.ms_seed; agg["ms_swstage"] += s.ms_sw; agg["ms_emit"] += s.ms_emit
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    time.sleep(0.12)                  # let the sample that covers the end of the region arrive
    clk = clocks.stop(c_lo, clocks.mark())
    n_hits, cig_bytes = batch.result_size()

    # -- end to end through the C ABI: pinned host reads -> H2D -> run -> D2H of the hit tuples, every step
    e2e = None
    if not a.no_e2e:
        st_t = torch.empty(n, dtype=torch.int32).pin_memory()
        ho_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
        hits_t = torch.empty((n_hits + n_hits // 8 + 1024) * 40, dtype=torch.uint8).pin_memory()
        cig_t = torch.empty(cig_bytes + cig_bytes // 8 + 4096, dtype=torch.uint8).pin_memory()

        aligner = api.Aligner(ix, chunk_reads=a.chunk_reads, chunk_bytes=a.chunk_reads * (a.read_len + 2), depth=a.depth)
        e2e_launches = 0
        e2e_copy = [0.0, 0.0]

        def e2e_step():
            # the reference-facing call: host buffers in, hit tuples out; copies overlap the kernels chunk by chunk
            return aligner.run_ptr(reads_t.data_ptr(), offs_t.data_ptr(), n, st_t.data_ptr(), ho_t.data_ptr(), hits_t.data_ptr(),
                                   hits_t.numel() // 40, cig_t.data_ptr(), cig_t.numel())
        e2e_step()
        barrier()
        e0 = time.perf_counter()
        for _ in range(a.steps):
            nh, nc = e2e_step()
            ast_ = aligner.stats()
            e2e_launches += ast_.launches
            e2e_copy[0] += ast_.ms_h2d; e2e_copy[1] += ast_.ms_d2h
        barrier()
        e2e_s = time.perf_counter() - e0
        assert (nh, nc) == (n_hits, cig_bytes), "pipelined and resident paths disagree"
        h2d = read_bytes + (n + 1) * 8
        d2h = n * 4 + (n + 1) * 8 + nh * 40 + nc
        e2e = [e2e_s, h2d, d2h, e2e_copy[0] / a.steps, e2e_copy[1] / a.steps]

    # -- max over ranks
    t_ev, t_wall = ev_ms, wall_ms
    if dist is not None:
        t = torch.tensor([ev_ms, wall_ms, e2e[0] if e2e else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_ev, t_wall = float(t[0]), float(t[1])
        if e2e:
            e2e[0] = float(t[2])
        cnt = torch.tensor([agg["launches"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(cnt)
        total_launches = int(cnt[0])
    else:
        total_launches = agg["launches"]

    if rank == 0:
        K = a.steps
        value = world * n * K / (t_ev * 1e-3)
        hbm_peak, peak_src = load_peaks()
        # algorithmic traffic / work of the two kernels that make up the step (DESIGN.md "units"):
        #   sga_kernel : per LF step not served by the k-mer table two 32-byte Occ blocks, per table lookup one 8-byte
        #                entry, per located row 4 bytes of SA, per read its bases once
        #   sw_fast    : per pair Lq + Lr bases in and a 16-byte result + its CIGAR out; cells x 4.25 ALU-pipe
        #                instructions (8.5 per cell pair: PRMT, 2 VIADDMNMX, VIMNMX, VIADD.16x2, 3.5 LOP3 -- from SASS)
        sga_bytes = (64.0 * agg["hbm_lf"] + 8.0 * agg["searches"] + 4.0 * agg["located"] + float(read_bytes * K))
        sga_gbs = sga_bytes / (agg["ms_sga"] * 1e-3) / 1e9 if agg["ms_sga"] > 0 else 0.0
        #   shortcut   : pairs whose result sw_classify_kernel proves without the DP (perfect / one-mismatch diagonal) are
        #                NOT in the DP kernel's cell count: gcups = cells the DP kernel computed / its own time
        dp_calls, dp_cells = agg["sw_calls"] - agg["sc_calls"], agg["cells"] - agg["sc_cells"]
        sw_bytes = dp_calls * (2.0 * a.read_len + 16.0) + float(cig_bytes * K) * (dp_calls / max(1, agg["sw_calls"]))
        sw_gbs = sw_bytes / (agg["ms_sw"] * 1e-3) / 1e9 if agg["ms_sw"] > 0 else 0.0
        gcups = dp_cells / (agg["ms_sw"] * 1e-3) / 1e9 if agg["ms_sw"] else 0.0
        dominant = "sga_kernel" if agg["ms_sga"] >= agg["ms_sw"] else "sw_fast_kernel"
        share = {"sga_kernel": round(agg["ms_sga"] / t_ev, 3), "sw_fast_kernel": round(agg["ms_sw"] / t_ev, 3),
                 "sw_classify_kernel (SW stage minus the DP kernels)": round((agg["ms_swstage"] - agg["ms_sw"]) / t_ev, 3)}
        tr = load_traffic(dominant)
        dom_bytes, dom_gbs, dom_ms = (sga_bytes, sga_gbs, agg["ms_sga"]) if dominant == "sga_kernel" else (sw_bytes, sw_gbs, agg["ms_sw"])
        roof = {"kernel": dominant, "bound": "hbm", "achieved": round(dom_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(dom_gbs / hbm_peak, 4), "traffic": tr["dram_bytes_per_launch"] if tr else None,
                "traffic_source": tr.get("source") if tr else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes / K, "ms_per_launch": dom_ms / K, "share_of_step": share,
                "dram_gbs_from_traffic": round(tr["dram_bytes_per_launch"] / (dom_ms / K * 1e-3) / 1e9, 1) if tr and dom_ms else None,
                "note": ("the dominant kernel is the integer DP (sw_fast_kernel): it is bound by ALU-pipe issue, not by HBM; "
                         "see roofline_alu for its real ceiling and roofline_sga for the memory-bound kernel") if dominant == "sw_fast_kernel"
                else ("the dominant kernel is sga_kernel: random 8-byte table entries and ~14-byte suffix-array runs over a 25 GB index. "
                      "`achieved` counts ALGORITHMIC bytes; `traffic` (ncu) is what DRAM really moves for them in 64-byte atoms -- "
                      "traffic / ms_per_launch is the DRAM rate the kernel runs at; frac_of_random_gather in roofline_sga compares the "
                      "algorithmic rate with the random 32-byte-sector gather ceiling measured in this run")}
        dpx_peak = extra.get("dpx_gops")
        roof_alu = {"kernel": "sw_fast_kernel", "bound": "alu-pipe issue (DPX VIADDMNMX.S16x2 / packed 16-bit)", "gcups": round(gcups, 1),
                    "achieved": round(gcups * 4.25, 1), "peak": dpx_peak, "unit": "Gop/s (ALU-pipe thread instructions)",
                    "frac": round(gcups * 4.25 / dpx_peak, 4) if dpx_peak else None,
                    "peak_source": "bwt_peak_dpx_gops: VIADDMNMX issue rate measured in this run"}
        roof_sga = {"kernel": "sga_kernel", "bound": "hbm random access", "achieved": round(sga_gbs, 1), "peak": hbm_peak,
                    "unit": "GB/s", "frac": round(sga_gbs / hbm_peak, 4), "ms_per_launch": agg["ms_sga"] / K,
                    "algorithmic_bytes_per_launch": sga_bytes / K,
                    "random_gather_peak_gbs": extra.get("gather_gbs"),
                    "frac_of_random_gather": round(sga_gbs / extra["gather_gbs"], 4) if extra.get("gather_gbs") else None}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": a.warmup,
            "ms_per_step": t_ev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "reads_per_gpu_per_step": n, "read_len": a.read_len, "ref_bp": a.ref_bp,
                       "parallelism": "read sharding, index replicated x%d" % world,
                       "l2": "inputs larger than L2 (index %.1f GB, reads %.2f GB per step); no flush needed" % (
                           ix.info().device_bytes / 1e9, read_bytes / 1e9),
                       "index_build_s": round(build_s, 2), "timer": "CUDA events on the batch stream, max over ranks",
                       "wall_ms_per_step": t_wall / K},
            "clocks": clk, "gpu_launches": total_launches,
            "roofline": roof, "roofline_alu": roof_alu, "roofline_sga": roof_sga,
            "lf_walk": extra.get("lf_walk"), "inexact": extra.get("inexact"),
            "work": {"lf_steps_per_s": agg["lf"] / (agg["ms_sga"] * 1e-3) if agg["ms_sga"] else None,
                     "hbm_lf_steps_per_read": agg["hbm_lf"] / (n * K), "lf_steps_per_read": agg["lf"] / (n * K),
                     "located_rows_per_read": agg["located"] / (n * K), "clusters_per_read": agg["clusters"] / (n * K),
                     "sw_calls_per_read": agg["sw_calls"] / (n * K), "cells_per_read": agg["cells"] / (n * K),
                     "sw_gcups": gcups,
                     "sw_shortcut_pairs_frac": agg["sc_calls"] / max(1, agg["sw_calls"]),
                     "sw_cells_equivalent_gcups": agg["cells"] / (agg["ms_swstage"] * 1e-3) / 1e9 if agg["ms_swstage"] else None,
                     "sw_general_pairs": agg["general"], "hits_per_step": n_hits,
                     "ms_sw_fast_kernels": agg["ms_sw"] / K, "ms_sga_kernel": agg["ms_sga"] / K,
                     "ms_stage": {"seed": agg["ms_seed"] / K, "sw": agg["ms_swstage"] / K, "emit": agg["ms_emit"] / K}},
        }
        if e2e:
            out["e2e"] = {"value": world * n * K / e2e[0], "unit": UNIT, "h2d_bytes_per_step": e2e[1], "d2h_bytes_per_step": e2e[2],
                          "timer": "wall clock around bwt_aligner_run (pinned host buffers in and out), barrier+synchronize both sides, max over ranks",
                          "chunk_reads": a.chunk_reads, "chunks_in_flight": a.depth,
                          "ms_h2d_per_step": e2e[3], "ms_d2h_per_step": e2e[4],
                          "h2d_gbs": e2e[1] / (e2e[3] * 1e-3) / 1e9 if e2e[3] else None,
                          "d2h_gbs": e2e[2] / (e2e[4] * 1e-3) / 1e9 if e2e[4] else None,
                          "note": "copy times are summed over the chunks (CUDA events on each chunk's stream); chunks overlap"}
        if world == 1 and a.cpu_seconds > 0:
            out["cpu_baseline"], out["parity_sample"] = cpu_baseline(a, ix, reads, offs, threads, batch.download())
        print(json.dumps(out), flush=True)
        if out.get("parity_sample", {}).get("mismatches", 0) != 0:
            print("bench.py: GPU hits differ from the CPU port on the sampled reads", file=sys.stderr)
            return 3
    batch.close()
    ix.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def cpu_baseline(a, ix, reads, offs, threads, gpu_result):
    """The CPU port timed on a bounded sample of the step's own reads, and -- with the candidates it produced anyway --
    the in-run parity check: they must equal the GPU hits of the same reads."""
    orc, setup_s = cpu_port_setup(ix, a)
    probe = min(a.reads, 20_000 * threads)
    t, _, _, _ = cpu_port_run(orc, reads, offs, 0, probe, threads)
    rate = probe / t
    m = int(min(a.reads, max(probe, rate * a.cpu_seconds)))
    t, st, ncand, cand = cpu_port_run(orc, reads, offs, 0, m, threads)
    orc.close()
    out = {"value": m / t, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "first %d of the step's %d reads, same index (exported from the device), %.1f s of CPU work on %d threads; "
                     "C port of the reference's algorithm (oracle/bwt_oracle.c), not BEAM" % (m, a.reads, t, threads),
           "lf_steps_per_read": st.lf_steps / m, "cells_per_read": st.cells / m, "setup_s": round(setup_s, 1)}
    st_g, ho, hits, cig = gpu_result
    t0 = time.perf_counter()
    mism, compared = parity_compare(m, ncand, cand, 16, st_g, ho, hits, cig)
    parity = {"reads": m, "candidates": compared, "mismatches": mism, "compare_s": round(time.perf_counter() - t0, 1),
              "what": "CPU port vs GPU hits of the same reads: candidate count, S, D, status, score, Pos, CIGAR text"}
    return out, parity


def run_reference(a, ix, reads, offs, threads):
    """--impl reference: the reference's CPU algorithm (C port in oracle/; the Erlang original needs OTP, absent here) on
    all host threads; each step = a bounded sample of the same workload."""
    orc, setup_s = cpu_port_setup(ix, a)
    ix.close()
    m = min(a.ref_step_reads, a.reads)
    pos = 0

    def step():
        nonlocal pos
        if pos + m > a.reads:
            pos = 0
        t = cpu_port_run(orc, reads, offs, pos, pos + m, threads)[0]
        pos += m
        return t
    for _ in range(a.warmup):
        step()
    tot = 0.0
    for _ in range(a.steps):
        tot += step()
    v = m * a.steps / tot
    sample = "each step = %d reads of the %d-read workload, %d threads; C port of the reference (oracle/bwt_oracle.c)" % (m, a.reads, threads)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": tot / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample, "index_export_s": round(setup_s, 1)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
    orc.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
