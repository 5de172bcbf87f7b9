#!/usr/bin/env python
"""bench.py -- aligned reads/s of the worker_bwt hot path (sga -> windows -> sw -> hit tuples) on B200.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU port of the reference (oracle/), all host cores
  python bench.py --config 4 | --config 5                  # BASELINE.json configs[3] / configs[4] as the main line

Main line (BASELINE.json configs[2], the largest single-GPU configuration): one un-chunked FM index of
a 3,137,454,466 bp synthetic reference (the bp total of bwt_files/human_g1k_v37_decoy.fasta.index, i.i.d. ACGT),
built on the device; 10 M synthetic ~100 bp forward-strand reads per GPU, 1 % per-base edits of which 80 % substitute,
10 % insert and 10 % delete a base (--edits sub = substitutions only, the round-1 workload).  A step = one
pass of the whole path over all of the rank's reads.  At N > 1 the index is replicated and every rank gets its own
10 M reads (read sharding, no data-path collective): weak scaling.

  value : reads/s with the reads (ASCII + offsets, as a worker holds them) already resident in HBM: bwt_batch_run only --
          2-bit packing, seeding, extension, hit records -- CUDA-event time, max over ranks
  e2e   : reads/s through the C ABI with HOST buffers: pinned packed reads -> H2D -> run -> D2H of the compact hit
          records, every step (bwt_aligner_run_packed); e2e_ascii = the same through the ASCII / 40-byte-record face

Legs (sub-objects of the same JSON line, each measured the same way on its own workload):
  config4 : configs[3]: the same index replicated, 80 M x ~150 bp reads sharded over the ranks (strong scaling)
  config5 : configs[4]: the reference in 8 _pJ partitions (fm_index.erl:113-124), every partition aligned against ALL
            reads (~250 bp), the per-partition records gathered to rank 0 over NCCL inside the timed region
            (bwt_partition_gather); + the bounded-edit search (seeds.erl) Len 244 / Defects 2 per partition
  config1 : configs[0]: GL000193.1, 10 k reads as 50 worker-sized lists of 200 through bwt_align_batch (N = 1 only)
  n_reads : 1 % 'N' per base: every pair through the double-precision warp kernel (N = 1 only)
"""
import argparse
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
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5], help="which BASELINE.json configuration is the main line")
    ap.add_argument("--ref-bp", type=int, default=REF_BP)
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU per step (config 3)")
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--sub-rate", type=float, default=0.01, help="per-base edit probability")
    ap.add_argument("--edits", default="indel", choices=["indel", "sub"],
                    help="indel: SURVEY.md 8(d) model, 80/10/10 substitution/insertion/deletion, ragged lengths; sub: substitutions only")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work of the cpu_baseline leg (0 = skip)")
    ap.add_argument("--ref-step-reads", type=int, default=200_000, help="reads per step of --impl reference")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the untimed peak / LF-walk side measurements")
    ap.add_argument("--no-legs", action="store_true", help="skip the config4 / config5 / config1 / n_reads legs")
    ap.add_argument("--chunk-reads", type=int, default=1 << 20, help="reads per pipelined chunk of the e2e leg")
    ap.add_argument("--depth", type=int, default=5, help="chunks in flight in the e2e leg (sweeps: profiles/r2b4_e2e_token_sweep.log, r2b6_e2e_token_depth_chunk_sweep.log; 8 ranks on 32 host cores: profiles/r2n8c_e2e_depth.log)")
    ap.add_argument("--inexact-queries", type=int, default=50_000, help="queries of the untimed bounded-edit side leg (0 = skip)")
    ap.add_argument("--c4-reads", type=int, default=80_000_000, help="config 4: reads in total (sharded over the ranks)")
    ap.add_argument("--c5-reads", type=int, default=2_000_000, help="config 5: reads (every partition sees all of them)")
    ap.add_argument("--c5-queries", type=int, default=20_000, help="config 5: queries of the bounded-edit search per partition")
    ap.add_argument("--partitions", type=int, default=8)
    return ap.parse_args()


def edits_text(a):
    return ("%.1f %% edits (80 %% substitution / 10 %% insertion / 10 %% deletion, ragged lengths)" % (100 * a.sub_rate)
            if a.edits == "indel" else "%.1f %% substitutions" % (100 * a.sub_rate))


def workload_name(a):
    return ("synthetic %.3f Gbp reference (human_g1k_v37_decoy bp total), un-chunked FM index; %d synthetic %d bp reads "
            "per GPU, %s; sga + sw + hit tuples" % (a.ref_bp / 1e9, a.reads, a.read_len, edits_text(a)))


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
    """per-launch DRAM bytes of a kernel from the committed ncu --set full capture of this same workload"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            t = json.load(f)
        e = t.get(kernel)
        if e:
            return e
    return None


# ---------------------------------------------------------------------------------------------------------------------
class Env:
    def __init__(self, a):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1 and a.impl == "ours":
            import torch.distributed as dist
            torch.cuda.set_device(self.local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        self.dev = self.local if a.impl == "ours" else 0
        self.threads = os.cpu_count() or 1

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor([float(v) for v in vals], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, vals):
        if self.dist is None:
            return [int(v) for v in vals]
        t = self.torch.tensor([int(v) for v in vals], dtype=self.torch.int64, device="cuda")
        self.dist.all_reduce(t)
        return [int(v) for v in t]

    def bcast_bytes(self, b, n):
        """rank 0's n bytes to every rank (plumbing: the NCCL unique id of the partition gather)"""
        if self.dist is None:
            return b
        t = self.torch.zeros(n, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            t.copy_(self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8))
        self.dist.broadcast(t, 0)
        return bytes(t.cpu().numpy())

    def pinned(self, count, dtype):
        torch = self.torch
        tt = {np.uint8: torch.uint8, np.int32: torch.int32, np.int64: torch.int64, np.uint32: torch.int32, np.uint16: torch.int16}[dtype]
        t = torch.empty(max(1, int(count)), dtype=tt).pin_memory()
        return t, t.numpy().view(dtype)


def sample(api, a, ix, n, read_len, seed, env, pinned=True, synth=None, n_rate=0.0):
    """-> (reads uint8 view, offs int64, keepalive): the SURVEY.md 8(d) edit model (or substitutions only) on the device"""
    cap = n * read_len + max(4096, n // 8)
    if pinned:
        rt, rbuf = env.pinned(cap, np.uint8)
        ot, offs = env.pinned(n + 1, np.int64)
    else:
        rt = ot = None
        rbuf = np.empty(cap, np.uint8)
        offs = np.empty(n + 1, np.int64)
    if a.edits == "indel" or synth is not None or n_rate > 0:
        ins = dele = 0.1 if a.edits == "indel" else 0.0
        reads, _, _ = api.sample_reads_edits(ix, n, read_len, seed=seed, edit_rate=a.sub_rate, ins_share=ins, del_share=dele,
                                             n_rate=n_rate, synth=synth, device=env.dev, out=rbuf, offsets=offs)
    else:
        reads = rbuf[:n * read_len]
        api.sample_reads(ix, n, read_len, seed=seed, sub_rate=a.sub_rate, out=reads)
        offs[:] = np.arange(n + 1, dtype=np.int64) * read_len
    return reads, offs, (rt, ot)


AGG_KEYS = ("lf", "hbm_lf", "located", "searches", "clusters", "sw_calls", "cells", "general", "launches", "sc_calls", "sc_cells", "band_calls", "band_cells", "band_filled")


def new_agg():
    agg = dict.fromkeys(AGG_KEYS, 0)
    agg.update(ms_sga=0.0, ms_sw=0.0, ms_seed=0.0, ms_swstage=0.0, ms_emit=0.0)
    return agg


def resident_steps(batch, steps, warmup, env, clocks=None):
    """W untimed + K timed bwt_batch_run on the resident batch -> (event ms summed, wall ms, agg, clock marks)"""
    for _ in range(max(warmup, 0)):
        batch.run()
    env.barrier()
    c_lo = clocks.mark() if clocks else 0
    w0 = time.perf_counter()
    ev_ms = 0.0
    agg = new_agg()
    for _ in range(steps):
        batch.run()
        s = batch.stats()
        ev_ms += s.ms_total
        agg["lf"] += s.lf_steps; agg["hbm_lf"] += s.hbm_lf_steps; agg["located"] += s.located_rows
        agg["searches"] += s.searches; agg["clusters"] += s.clusters; agg["sw_calls"] += s.sw_calls
        agg["cells"] += s.sw_cells; agg["general"] += s.sw_general; agg["launches"] += s.launches
        agg["sc_calls"] += s.sw_shortcut; agg["sc_cells"] += s.sw_shortcut_cells
        agg["band_calls"] += s.sw_band; agg["band_cells"] += s.sw_band_cells; agg["band_filled"] += s.sw_band_filled
        agg["ms_sga"] += s.ms_sga_kernel; agg["ms_sw"] += s.ms_sw_kernel
        agg["ms_seed"] += s.ms_seed; agg["ms_swstage"] += s.ms_sw; agg["ms_emit"] += s.ms_emit
    env.barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    return ev_ms, wall_ms, agg, (c_lo, clocks.mark() if clocks else 0)


class CompactOut:
    """pinned host buffers for compact results of n reads"""

    def __init__(self, env, api, n, hits_cap, cig_cap):
        self.k = []
        t, self.st = env.pinned(n, np.uint8); self.k.append(t)
        t, self.ho = env.pinned(n + 1, np.uint32); self.k.append(t)
        t, raw = env.pinned(16 * hits_cap, np.uint8); self.k.append(t)
        self.hits = raw.view(api.HIT16_DTYPE)
        t, self.cig = env.pinned(cig_cap, np.uint8); self.k.append(t)


def e2e_packed(api, env, ix, packed, n, steps, chunk_reads, depth, read_len):
    """host-resident packed reads -> bwt_aligner_run_packed -> compact records in pinned host memory, every step"""
    aligner = api.Aligner(ix, chunk_reads=chunk_reads, chunk_bytes=chunk_reads * (read_len + 16), depth=depth)
    out = CompactOut(env, api, n, 2 * n + 1024, 4 * n + 4096)

    def step():
        return aligner.run_packed_ptr(packed, out.st.ctypes.data, out.ho.ctypes.data, out.hits.ctypes.data, len(out.hits),
                                      out.cig.ctypes.data, len(out.cig))
    step()
    env.barrier()
    launches, cp = 0, [0.0, 0.0]
    e0 = time.perf_counter()
    for _ in range(steps):
        nh, nc = step()
        s = aligner.stats()
        launches += s.launches
        cp[0] += s.ms_h2d; cp[1] += s.ms_d2h
    env.barrier()
    dt = time.perf_counter() - e0
    aligner.close()
    return {"seconds": dt, "h2d": packed.upload_bytes(), "d2h": n + 4 * (n + 1) + 16 * nh + nc, "ms_h2d": cp[0] / steps,
            "ms_d2h": cp[1] / steps, "launches": launches, "n_hits": nh, "cig_bytes": nc}


def e2e_ascii(api, env, ix, reads_ptr, offs_ptr, n, read_bytes, n_hits, cig_bytes, steps, chunk_reads, depth, read_len):
    """the reference-shaped face: ASCII reads + int64 offsets in, 40-byte records out (pinned host buffers)"""
    torch = env.torch
    st_t = torch.empty(n, dtype=torch.int32).pin_memory()
    ho_t = torch.empty(n + 1, dtype=torch.int64).pin_memory()
    hits_t = torch.empty((n_hits + n_hits // 8 + 1024) * 40, dtype=torch.uint8).pin_memory()
    cig_t = torch.empty(cig_bytes + cig_bytes // 8 + 4096, dtype=torch.uint8).pin_memory()
    aligner = api.Aligner(ix, chunk_reads=chunk_reads, chunk_bytes=chunk_reads * (read_len + 16), depth=depth)

    def step():
        return aligner.run_ptr(reads_ptr, offs_ptr, n, st_t.data_ptr(), ho_t.data_ptr(), hits_t.data_ptr(), hits_t.numel() // 40,
                               cig_t.data_ptr(), cig_t.numel())
    step()
    env.barrier()
    e0 = time.perf_counter()
    for _ in range(steps):
        nh, nc = step()
    env.barrier()
    dt = time.perf_counter() - e0
    aligner.close()
    return {"seconds": dt, "h2d": read_bytes + 8 * (n + 1), "d2h": 4 * n + 8 * (n + 1) + 40 * nh + nc, "n_hits": nh, "cig_bytes": nc}


# ---------------------------------------------------------------------------------------------------------------------
_ORC = {}


def cpu_port_setup(ix):
    """Index columns back on the host for the CPU port (oracle/bwt_oracle.c, closed-form LF: same results as the literal scan);
    kept for the other legs of the same run (the export of a 3.1 Gbp index takes seconds)."""
    from oracle.c_oracle import COracle
    key = id(ix)
    if key in _ORC:
        return _ORC[key], 0.0
    t0 = time.time()
    L, SA, ref = ix.export()
    info = ix.info()
    orc = COracle(L, SA, tuple(info.pointers), int(info.shift), ref, mode=1)
    _ORC[key] = orc
    return orc, time.time() - t0


def cpu_port_close():
    for o in _ORC.values():
        o.close()
    _ORC.clear()


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


def cpu_baseline(a, ix, reads, offs, threads, gpu_result, seconds=None, what=""):
    """The CPU port timed on a bounded sample of the step's own reads, and -- with the candidates it produced anyway --
    the in-run parity check: they must equal the GPU hits of the same reads."""
    orc, setup_s = cpu_port_setup(ix)
    n = len(offs) - 1
    seconds = a.cpu_seconds if seconds is None else seconds
    probe = min(n, 20_000 * threads)
    t, _, _, _ = cpu_port_run(orc, reads, offs, 0, probe, threads)
    rate = probe / t
    m = int(min(n, max(probe, rate * seconds)))
    t, st, ncand, cand = cpu_port_run(orc, reads, offs, 0, m, threads)
    out = {"value": m / t, "unit": UNIT, "cores": threads, "kind": "port",
           "sample": "first %d of the step's %d reads%s, same index (exported from the device), %.1f s of CPU work on %d threads; "
                     "C port of the reference's algorithm (oracle/bwt_oracle.c), not BEAM" % (m, n, what, t, threads),
           "lf_steps_per_read": st.lf_steps / m, "cells_per_read": st.cells / m, "setup_s": round(setup_s, 1)}
    st_g, ho, hits, cig = gpu_result
    t0 = time.perf_counter()
    mism, compared = parity_compare(m, ncand, cand, 16, st_g, ho, hits, cig)
    parity = {"reads": m, "candidates": compared, "mismatches": mism, "compare_s": round(time.perf_counter() - t0, 1),
              "what": "CPU port vs GPU hits of the same reads: candidate count, S, D, status, score, Pos, CIGAR text"}
    return out, parity


# ---------------------------------------------------------------------------------------------------------------------
def roofline_objects(a, K, t_ev, agg, read_bytes, cig_bytes, extra):
    hbm_peak, peak_src = load_peaks()
    # algorithmic traffic / work of the kernels that make up the step (DESIGN.md "units"):
    #   sga_kernel : per LF step not served by a k-mer table two 32-byte Occ blocks, per table lookup one 8-byte
    #                entry, per located row 4 bytes of SA, per read its bases once (counted as the bytes of the ASCII
    #                batch, the input of the step)
    #   sw_fast    : per pair Lq + Lr bases in and a 16-byte result + its CIGAR out; cells x 4.25 ALU-pipe
    #                instructions (8.5 per cell pair: PRMT, 2 VIADDMNMX, VIMNMX, VIADD.16x2, 3.5 LOP3 -- from SASS)
    sga_bytes = (64.0 * agg["hbm_lf"] + 8.0 * agg["searches"] + 4.0 * agg["located"] + float(read_bytes * K))
    sga_gbs = sga_bytes / (agg["ms_sga"] * 1e-3) / 1e9 if agg["ms_sga"] > 0 else 0.0
    dp_calls, dp_cells = agg["sw_calls"] - agg["sc_calls"] - agg["band_calls"], agg["cells"] - agg["sc_cells"] - agg["band_cells"]
    sw_bytes = dp_calls * (2.0 * a.read_len + 16.0) + float(cig_bytes * K) * (dp_calls / max(1, agg["sw_calls"]))
    sw_gbs = sw_bytes / (agg["ms_sw"] * 1e-3) / 1e9 if agg["ms_sw"] > 0 else 0.0
    gcups = dp_cells / (agg["ms_sw"] * 1e-3) / 1e9 if agg["ms_sw"] else 0.0
    dominant = "sga_kernel" if agg["ms_sga"] >= agg["ms_sw"] else "sw_fast_kernel"
    share = {"sga_kernel": round(agg["ms_sga"] / t_ev, 3), "sw_fast_kernel": round(agg["ms_sw"] / t_ev, 3),
             "classification + rare tiers (SW stage minus the DP kernel)": round((agg["ms_swstage"] - agg["ms_sw"]) / t_ev, 3),
             "emit": round(agg["ms_emit"] / t_ev, 3)}
    tr = load_traffic(dominant)
    dom_bytes, dom_gbs, dom_ms = (sga_bytes, sga_gbs, agg["ms_sga"]) if dominant == "sga_kernel" else (sw_bytes, sw_gbs, agg["ms_sw"])
    roof = {"kernel": dominant, "bound": "hbm", "achieved": round(dom_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
            "frac": round(dom_gbs / hbm_peak, 4), "traffic": tr["dram_bytes_per_launch"] if tr else None,
            "traffic_source": tr.get("source") if tr else None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": dom_bytes / K, "ms_per_launch": dom_ms / K, "share_of_step": share,
            "dram_gbs_from_traffic": round(tr["dram_bytes_per_launch"] / (dom_ms / K * 1e-3) / 1e9, 1) if tr and dom_ms else None,
            "note": ("the dominant kernel is the integer DP (sw_fast_kernel): it is bound by ALU-pipe issue, not by HBM; "
                     "see roofline_alu for its real ceiling and roofline_sga for the memory-bound kernel") if dominant == "sw_fast_kernel"
            else ("the dominant kernel is sga_kernel: one 64-byte fused table entry (interval + first seed ends) per 15-base subsequence "
                  "of a 25-90 GB index. `achieved` counts ALGORITHMIC bytes (8 B per lookup, 4 B per located row, the read); `traffic` "
                  "(ncu) is what DRAM really moves for them")}
    dpx_peak = extra.get("dpx_gops")
    roof_alu = {"kernel": "sw_fast_kernel", "bound": "alu-pipe issue (DPX VIADDMNMX.S16x2 / packed 16-bit)", "gcups": round(gcups, 1),
                "achieved": round(gcups * 4.25, 1), "peak": dpx_peak, "unit": "Gop/s (ALU-pipe thread instructions)",
                "frac": round(gcups * 4.25 / dpx_peak, 4) if dpx_peak else None,
                "peak_source": "bwt_peak_dpx_gops: VIADDMNMX issue rate measured in this run"}
    tr_sga = load_traffic("sga_kernel")
    roof_sga = {"kernel": "sga_kernel", "bound": "hbm random access", "achieved": round(sga_gbs, 1), "peak": hbm_peak,
                "unit": "GB/s", "frac": round(sga_gbs / hbm_peak, 4), "ms_per_launch": agg["ms_sga"] / K,
                "algorithmic_bytes_per_launch": sga_bytes / K,
                "traffic": tr_sga["dram_bytes_per_launch"] if tr_sga else None,
                "random_gather_peak_gbs": extra.get("gather_gbs"),
                "frac_of_random_gather": round(sga_gbs / extra["gather_gbs"], 4) if extra.get("gather_gbs") else None,
                # the ceiling that applies: random ACCESSES per second -- one per table lookup
                "lookups_per_s": agg["searches"] / (agg["ms_sga"] * 1e-3) if agg["ms_sga"] else None,
                "random_access_peak_per_s": extra["gather_gbs"] * 1e9 / 32 if extra.get("gather_gbs") else None,
                "frac_of_random_access_peak": round(agg["searches"] / (agg["ms_sga"] * 1e-3) / (extra["gather_gbs"] * 1e9 / 32), 4)
                if extra.get("gather_gbs") and agg["ms_sga"] else None,
                "note": "random access on this part is bound by accesses in flight, not bytes: the peak is the measured rate of random "
                        "32-byte sector reads (two 16-byte loads each, L2::64B, as the kernel issues them); the kernel also sorts and votes"}
    return roof, roof_alu, roof_sga, gcups


def work_object(agg, n, K, gcups, n_hits):
    return {"lf_steps_per_s": agg["lf"] / (agg["ms_sga"] * 1e-3) if agg["ms_sga"] else None,
            "hbm_lf_steps_per_read": agg["hbm_lf"] / (n * K), "lf_steps_per_read": agg["lf"] / (n * K),
            "located_rows_per_read": agg["located"] / (n * K), "clusters_per_read": agg["clusters"] / (n * K),
            "sw_calls_per_read": agg["sw_calls"] / (n * K), "cells_per_read": agg["cells"] / (n * K),
            "sw_gcups": gcups,
            "sw_shortcut_pairs_frac": agg["sc_calls"] / max(1, agg["sw_calls"]),
            "sw_band_pairs_frac": agg["band_calls"] / max(1, agg["sw_calls"]),
            "sw_band_cells_filled_per_pair": agg["band_filled"] / max(1, agg["band_calls"]),
            "sw_cells_equivalent_gcups": agg["cells"] / (agg["ms_swstage"] * 1e-3) / 1e9 if agg["ms_swstage"] else None,
            "sw_general_pairs": agg["general"], "hits_per_step": n_hits,
            "ms_sw_fast_kernels": agg["ms_sw"] / K, "ms_sga_kernel": agg["ms_sga"] / K,
            "ms_stage": {"seed": agg["ms_seed"] / K, "sw": agg["ms_swstage"] / K, "emit": agg["ms_emit"] / K}}


# ---------------------------------------------------------------------------------------------------------------------
def run_config3(a, env, api, ix, build_s):
    """BASELINE.json configs[2]: the main line"""
    rank, world, dev = env.rank, env.world, env.dev
    n = a.reads
    reads, offs, keep = sample(api, a, ix, n, a.read_len, a.seed + 1000 * (rank + 1), env)
    read_bytes = int(offs[n])

    # -- untimed side measurements on rank 0: roofline denominators measured on this device, and the FM walk alone
    extra = {}
    if rank == 0 and not a.no_side:
        extra["gather_gbs"] = api.peak_gather_gbs(dev, 8 << 30)          # random 32-byte sectors over 8 GiB (accesses/s x 32 B)
        extra["dpx_gops"] = api.peak_dpx_gops(dev)
        # LF steps/s of bwa:find_seeds alone: 15-mers cut from the reads, 4^8 table only (every later step reads two
        # Occ sectors), the walk the deep tables replace in the aligned-reads path
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
                            "occ_sector_reads_per_s": 2 * hbm_lf * 3 / (ms * 1e-3),
                            "frac_of_random_access_peak": round(2 * hbm_lf * 3 / (ms * 1e-3) / (extra["gather_gbs"] * 1e9 / 32), 4),
                            "note": "BASELINE's 'FM LF-steps/s vs HBM roofline': search_kernel on 15-mers with the 4^8 table only, "
                                    "64 B (two Occ sectors) per step beyond it; the ratio to the random-access peak can exceed 1: the first "
                                    "two levels below the table (16 + 64 MB of Occ blocks) are L2-resident and both bounds of a narrow "
                                    "interval share a sector"}
        # the bounded-edit search of seeds.erl (row A9/F4) on the same index: Defects = 2 on a sample of the reads
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
                                "settled_by_front_end": fss.settled / mq,
                                "note": "fs_search_kernel: per query one warp; pigeonhole front end (15 exact piece searches + bit-parallel "
                                        "window verification) with the depth-first search over gen_pattern/5 (seeds.erl:51-62) as fallback; "
                                        "states = piece searches + verified windows (front end) or search states (fallback); 64 B per LF step"}

    batch = api.Batch(ix, n, read_bytes)
    batch.upload(reads, offs)
    clocks = ClockSampler(dev)
    clocks.start()                    # before the warm-up: nvidia-smi needs a moment to deliver its first sample
    ev_ms, wall_ms, agg, marks = resident_steps(batch, a.steps, a.warmup, env, clocks)
    time.sleep(0.12)                  # let the sample that covers the end of the region arrive
    clk = clocks.stop(*marks)
    n_hits, cig_bytes = batch.result_size()
    gpu_result = batch.download() if (rank == 0 and world == 1 and a.cpu_seconds > 0) else None
    batch.close()

    e2e = e2a = None
    if not a.no_e2e:
        packed = api.PackedReads(reads, offs, threads=min(16, env.threads), pinned=True)          # the ingest's job: outside the timed region
        e2e = e2e_packed(api, env, ix, packed, n, a.steps, a.chunk_reads, a.depth, a.read_len)
        assert e2e["n_hits"] == n_hits, "pipelined and resident paths disagree"
        del packed
        e2a = e2e_ascii(api, env, ix, reads.ctypes.data, offs.ctypes.data, n, read_bytes, n_hits, cig_bytes, max(1, a.steps // 2),
                        a.chunk_reads, a.depth, a.read_len)
        assert (e2a["n_hits"], e2a["cig_bytes"]) == (n_hits, cig_bytes), "pipelined and resident paths disagree"

    t_ev, t_wall, t_e2e, t_e2a = env.max_over_ranks([ev_ms, wall_ms, e2e["seconds"] if e2e else 0.0, e2a["seconds"] if e2a else 0.0])
    total_launches = env.sum_over_ranks([agg["launches"]])[0]
    out = None
    if rank == 0:
        K = a.steps
        roof, roof_alu, roof_sga, gcups = roofline_objects(a, K, t_ev, agg, read_bytes, cig_bytes, extra)
        info = ix.info()
        out = {
            "metric": METRIC, "value": world * n * K / (t_ev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": a.warmup,
            "ms_per_step": t_ev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "reads_per_gpu_per_step": n, "read_len": a.read_len, "ref_bp": a.ref_bp,
                       "edits": a.edits, "parallelism": "read sharding, index replicated x%d" % world,
                       "l2": "inputs larger than L2 (index %.1f GB, reads %.2f GB per step); no flush needed" % (
                           info.device_bytes / 1e9, read_bytes / 1e9),
                       "index": {"device_gb": round(info.device_bytes / 1e9, 1), "kmer_depth": info.kmer_k, "fused_table": bool(info.fused)},
                       "index_build_s": round(build_s, 2), "timer": "CUDA events on the batch stream, max over ranks",
                       "wall_ms_per_step": t_wall / K},
            "clocks": clk, "gpu_launches": total_launches,
            "roofline": roof, "roofline_alu": roof_alu, "roofline_sga": roof_sga,
            "lf_walk": extra.get("lf_walk"), "inexact": extra.get("inexact"),
            "work": work_object(agg, n, K, gcups, n_hits),
        }
        if e2e:
            out["e2e"] = {"value": world * n * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                          "timer": "wall clock around bwt_aligner_run_packed (pinned host buffers in and out), barrier+synchronize both sides, max over ranks",
                          "input": "2-bit packed reads (bwt_pack_reads, as a FASTQ ingest hands them over) + lengths; the packing is the ingest's work, outside the timed region",
                          "output": "uint8 status, uint32 offsets, 16-byte bwt_hit16 records, CIGAR bytes ('<Lq>M' flagged, not stored)",
                          "chunk_reads": a.chunk_reads, "chunks_in_flight": a.depth,
                          "ms_h2d_per_step": e2e["ms_h2d"], "ms_d2h_per_step": e2e["ms_d2h"],
                          "h2d_gbs": e2e["h2d"] / (e2e["ms_h2d"] * 1e-3) / 1e9 if e2e["ms_h2d"] else None,
                          "d2h_gbs": e2e["d2h"] / (e2e["ms_d2h"] * 1e-3) / 1e9 if e2e["ms_d2h"] else None,
                          "note": "copy times are summed over the chunks (CUDA events on each chunk's stream); chunks overlap"}
            out["e2e_ascii"] = {"value": world * n * max(1, a.steps // 2) / t_e2a, "unit": UNIT, "h2d_bytes_per_step": e2a["h2d"],
                                "d2h_bytes_per_step": e2a["d2h"],
                                "note": "bwt_aligner_run: ASCII reads + int64 offsets in, int32 status / int64 offsets / 40-byte bwt_hit records out"}
        if gpu_result is not None:
            out["cpu_baseline"], out["parity_sample"] = cpu_baseline(a, ix, reads, offs, env.threads, gpu_result)
    del reads, offs, keep
    return out


# ---------------------------------------------------------------------------------------------------------------------
def run_config4(a, env, api, ix, steps, warmup):
    """BASELINE.json configs[3]: the same index on every GPU, c4_reads x ~150 bp reads sharded over the ranks (contiguous
    ranges, no collective): strong scaling.  Resident: the rank's reads in slices of 10 M through one batch; e2e: packed."""
    rank, world = env.rank, env.world
    total = a.c4_reads
    lo, hi = total * rank // world, total * (rank + 1) // world
    n = hi - lo
    t0 = time.time()
    reads, offs, keep = sample(api, a, ix, n, 150, a.seed + 4000 + 7 * rank, env, pinned=False)
    sample_s = time.time() - t0
    slice_n = min(n, 10_000_000)
    bounds = list(range(0, n, slice_n)) + [n]
    batch = api.Batch(ix, slice_n, int(max(offs[bounds[k + 1]] - offs[bounds[k]] for k in range(len(bounds) - 1))))
    ev_ms, agg, n_hits = 0.0, new_agg(), 0
    for k in range(len(bounds) - 1):
        s0, s1 = bounds[k], bounds[k + 1]
        batch.upload(reads[offs[s0]:offs[s1]], np.ascontiguousarray(offs[s0:s1 + 1] - offs[s0]))
        e, _, g, _ = resident_steps(batch, steps, warmup if k == 0 else 1, env)
        ev_ms += e
        for key in g:
            agg[key] += g[key]
        n_hits += batch.result_size()[0]
    batch.close()
    t0 = time.time()
    packed = api.PackedReads(reads, offs, threads=min(16, env.threads), pinned=True)
    pack_s = time.time() - t0
    e2e = e2e_packed(api, env, ix, packed, n, steps, a.chunk_reads, a.depth, 150)
    del packed, reads, offs, keep
    t_ev, t_e2e = env.max_over_ranks([ev_ms, e2e["seconds"]])
    tot_hits, tot_launch = env.sum_over_ranks([n_hits, agg["launches"]])
    if rank != 0:
        return None
    return {"workload": "BASELINE.json configs[3]: index replicated x%d, %d synthetic 150 bp reads (%s) sharded %d per rank" % (
                world, total, edits_text(a), n),
            "scaling": "strong", "reads_total": total, "read_len": 150, "steps": steps,
            "value": total * steps / (t_ev * 1e-3), "unit": UNIT, "ms_per_step": t_ev / steps,
            "e2e": {"value": total * steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "ms_h2d_per_step": e2e["ms_h2d"], "ms_d2h_per_step": e2e["ms_d2h"]},
            "hits_per_step": tot_hits, "gpu_launches": tot_launch,
            "sw_shortcut_pairs_frac": agg["sc_calls"] / max(1, agg["sw_calls"]),
            "ms_stage_rank0": {"seed": agg["ms_seed"] / steps, "sw": agg["ms_swstage"] / steps, "emit": agg["ms_emit"] / steps,
                               "sga_kernel": agg["ms_sga"] / steps, "sw_fast_kernel": agg["ms_sw"] / steps},
            "setup_s": {"sample": round(sample_s, 1), "host_pack": round(pack_s, 1)}}


def run_config5(a, env, api, steps, warmup):
    """BASELINE.json configs[4]: the reference in `partitions` _pJ partitions (builder formula, fm_index.erl:113-124), rank r
    holds partitions r, r + world, ...; every partition is aligned against ALL reads (~250 bp, host-resident packed) and the
    records of all partitions are gathered to rank 0 over NCCL (bwt_partition_gather) -- upload, run and gather are inside
    the timed region.  The partition indexes are built one at a time (an offline step in the reference too)."""
    from bwt_b200.partition import partition_slices
    rank, world, dev = env.rank, env.world, env.dev
    P = a.partitions
    n = a.c5_reads
    reads, offs, keep = sample(api, a, None, n, 250, a.seed + 5000, env, pinned=False, synth=(a.ref_bp, a.seed))   # the SAME reads on every rank
    packed = api.PackedReads(reads, offs, threads=min(16, env.threads), pinned=True)
    uid = env.bcast_bytes(api.Comm.unique_id() if rank == 0 else b"\0" * 128, 128) if world > 1 else None
    comm = api.Comm(uid, world, rank, dev)
    slices = partition_slices(a.ref_bp, P)
    rounds = (P + world - 1) // world
    if world < P:
        os.environ["BWT_FUSED"] = "0"        # several partitions per GPU, one after the other: the plain 8-byte depth-15 table
    gather_out = None
    t_total = ms_run = ms_gather = build_s = 0.0
    link_bytes = 0
    fs = {"queries": 0, "kernel_ms": 0.0, "seeds": 0, "lf_steps": 0, "limit": 0, "settled": 0}
    part_hits = []
    for j in range(rounds):
        pidx = j * world + rank
        have = pidx < P
        start0, length, add = slices[min(pidx, P - 1)]
        t0 = time.time()
        px = api.Index.build_synthetic_part(start0, length, seed=a.seed, shift=add, device=dev)
        build_s += time.time() - t0
        batch = api.Batch(px, n, int(offs[n]))
        if rank == 0 and gather_out is None:
            t1, st = env.pinned(world * n, np.uint8); t2, ho = env.pinned(world * (n + 1), np.uint32)
            t3, raw = env.pinned(16 * (2 * world * n + 1024), np.uint8); t4, cg = env.pinned(8 * world * n + 4096, np.uint8)
            gather_out = (st, ho, raw.view(api.HIT16_DTYPE), cg, (t1, t2, t3, t4))

        def step():
            batch.upload_packed(packed)
            batch.run_compact()
            s = batch.stats()
            parts, ms, lb = comm.gather(batch, 0, gather_out[:4] if gather_out else None)
            return s.ms_total, ms, lb, parts
        for _ in range(max(1, warmup)):
            step()
        env.barrier()
        w0 = time.perf_counter()
        for _ in range(steps):
            r_ms, g_ms, lb, parts = step()
            ms_run += r_ms; ms_gather += g_ms; link_bytes += lb
        env.barrier()
        t_total += time.perf_counter() - w0
        if rank == 0:
            part_hits += [len(p[2]) for p in parts][:min(world, P - j * world)]
        # the bounded-edit search of seeds.erl on this partition: Len 244, Defects 2 (kernel time; results stay per partition)
        mq = min(n, a.c5_queries)
        if mq > 0 and have:
            fst, fso, fsd, fss = px.fs_find_seeds(reads[:offs[mq]], offs[:mq + 1], 244, 2, 1024)
            fs["queries"] += mq; fs["kernel_ms"] += fss.ms_kernel; fs["seeds"] += len(fsd); fs["lf_steps"] += fss.lf_steps
            fs["limit"] += int((fst == api.FS_LIMIT).sum()); fs["settled"] += int(fss.settled)
        batch.close()
        px.close()
    os.environ.pop("BWT_FUSED", None)
    comm.close()
    t_all, t_run, t_g, fs_ms = env.max_over_ranks([t_total, ms_run, ms_gather, fs["kernel_ms"]])
    fsq, fsl, fss_, fslim, fsset = env.sum_over_ranks([fs["queries"], fs["lf_steps"], fs["seeds"], fs["limit"], fs["settled"]])
    if rank != 0:
        return None
    return {"workload": "BASELINE.json configs[4]: %.3f Gbp synthetic reference in %d _pJ partitions (fm_index.erl:113-124), %d per rank; %d synthetic "
                        "250 bp reads (%s), every partition aligned against all of them, records gathered to rank 0 over NCCL" % (
                            a.ref_bp / 1e9, P, rounds, n, edits_text(a)),
            "scaling": "strong (partitions over ranks)", "partitions": P, "partitions_per_rank": rounds, "reads": n, "read_len": 250, "steps": steps,
            "value": n * steps / t_all, "unit": UNIT,
            "definition": "reads aligned against ALL partitions with the merged records on rank 0, per second; timed region = packed H2D + run + NCCL gather + D2H at the root, wall clock, max over ranks",
            "ms_per_step": t_all / steps * 1e3, "ms_kernels_per_step": t_run / steps, "ms_gather_per_step": t_g / steps,
            "nvlink_bytes_per_step": link_bytes // max(1, steps), "hits_per_partition": part_hits,
            "index_build_s_per_rank": round(build_s, 1),
            "inexact": {"len": 244, "defects": 2, "queries_x_partitions": fsq, "kernel_ms_max_rank": fs_ms,
                        "query_partition_searches_per_s": fsq / (fs_ms * 1e-3) if fs_ms else None,
                        "seeds": fss_, "lf_steps": fsl, "limit_status": fslim, "settled_by_front_end": fsset / max(1, fsq),
                        "note": "fs_search_kernel per partition (seeds.erl:11-62 semantics, parity unpinned in the reference); seeds stay per partition"}}


def run_config1(a, env, api):
    """BASELINE.json configs[0]: GL000193.1, 10 k reads as 50 worker-sized lists of 200 (master.erl:59) through
    bwt_align_batch, one call per list, beside the CPU port on the same lists."""
    from oracle.c_oracle import COracle
    path = os.path.join(ROOT, "tests", "golden", "GL000193.1_p1.ref")
    if not os.path.exists(path):
        return {"skipped": "tests/golden/GL000193.1_p1.ref missing"}
    with open(path, "rb") as f:
        ref = f.read()
    with api.Index.build(ref, shift=1, device=env.dev) as gx:
        reads, offs, _ = api.sample_reads_edits(gx, 10_000, 100, seed=a.seed, edit_rate=0.01)
        lists = [(np.ascontiguousarray(reads[offs[k]:offs[k + 200]]), np.ascontiguousarray(offs[k:k + 201] - offs[k])) for k in range(0, 10_000, 200)]
        for r, o in lists[:5]:
            gx.align(r, o)
        t0 = time.perf_counter()
        hits = 0
        for _ in range(3):
            for r, o in lists:
                st, ho, h, cig = gx.align(r, o)
                hits += int((h["status"] == 0).sum())
        dt = (time.perf_counter() - t0) / 3
        big = gx.align(reads, offs)                       # the same 10 k reads in one call
        t0 = time.perf_counter()
        for _ in range(5):
            gx.align(reads, offs)
        dt_one = (time.perf_counter() - t0) / 5
        L, SA, _ = gx.export()
        info = gx.info()
        orc = COracle(L, SA, tuple(info.pointers), 1, ref, mode=0)
        t0 = time.perf_counter()
        for r, o in lists:
            orc.align_batch(r, o, nthreads=2, with_sw=True, cap_per_read=16)
        dt_cpu = time.perf_counter() - t0
        ncand, cand, _ = orc.align_batch(reads, offs, nthreads=env.threads, with_sw=True, cap_per_read=16)
        mism, compared = parity_compare(10_000, ncand, cand, 16, big[0], big[1], big[2], big[3])
        orc.close()
    return {"workload": "BASELINE.json configs[0]: GL000193.1 (189,789 bp), 10 k synthetic 100 bp reads, 1 % edits, 50 lists of 200 (master.erl:59)",
            "ms_per_200_read_call": dt / 50 * 1e3, "reads_per_s_list_calls": 10_000 / dt, "reads_per_s_one_call": 10_000 / dt_one,
            "matches": hits // 3, "cpu_port_2_threads_reads_per_s": 10_000 / dt_cpu,
            "parity": {"reads": 10_000, "candidates": compared, "mismatches": mism},
            "note": "wall clock around bwt_align_batch (pageable host buffers, hit tuples back); 2 CPU threads = the 2 local workers of scripts/start_local.sh"}


def run_n_reads(a, env, api, ix):
    """1 % 'N' per base: the reads take the ASCII paths (sga:skip_Ns/1) and every pair the double-precision warp kernel"""
    n = 1_000_000
    reads, offs, keep = sample(api, a, ix, n, a.read_len, a.seed + 77, env, pinned=False, n_rate=0.01)
    batch = api.Batch(ix, n, int(offs[n]))
    batch.upload(reads, offs)
    ev_ms, _, agg, _ = resident_steps(batch, 2, 1, env)
    n_hits = batch.result_size()[0]
    out = {"workload": "%d reads of %d bp, %s, then 1 %% of the bases turned into 'N'" % (n, a.read_len, edits_text(a)),
           "value": n * 2 / (ev_ms * 1e-3), "unit": UNIT, "ms_per_step": ev_ms / 2, "general_pairs_per_step": agg["general"] // 2,
           "sw_calls_per_step": agg["sw_calls"] // 2, "hits_per_step": n_hits,
           "us_per_general_pair": (agg["ms_swstage"] - agg["ms_sw"]) * 1e3 / max(1, agg["general"]),
           "ms_stage": {"seed": agg["ms_seed"] / 2, "sw": agg["ms_swstage"] / 2, "emit": agg["ms_emit"] / 2}}
    if a.cpu_seconds > 0:
        base, par = cpu_baseline(a, ix, reads, offs, env.threads, batch.download(), seconds=3.0, what=" with 'N'")
        out["parity_sample"] = par
        out["cpu_port_reads_per_s"] = base["value"]
    batch.close()
    return out


# ---------------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    env = Env(a)
    if a.impl == "reference" and env.rank != 0:
        return 0
    from bwt_b200 import api          # raises if libbwt_b200.so is missing: no CPU fallback

    if a.impl == "reference":
        ix = api.Index.build_synthetic(a.ref_bp, seed=a.seed, shift=1, device=env.dev)
        return run_reference(a, env, api, ix)

    rc = 0
    out = None
    if a.config == 5:
        leg = run_config5(a, env, api, a.steps, a.warmup)
        if env.rank == 0:
            out = {"metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
                   "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                   "data": "synthetic", "config": {"workload": leg["workload"]}, "gpu_launches": None, "config5": leg,
                   "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                           "note": "the config-5 value is end to end by definition (host packed reads in, merged records on rank 0's host)"}}
    else:
        t0 = time.time()
        ix = api.Index.build_synthetic(a.ref_bp, seed=a.seed, shift=1, device=env.dev)
        build_s = time.time() - t0
        if a.config == 4:
            leg = run_config4(a, env, api, ix, a.steps, a.warmup)
            if env.rank == 0:
                out = {"metric": METRIC, "value": leg["value"], "unit": UNIT, "n_gpus": env.world, "steps": a.steps, "warmup": a.warmup,
                       "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
                       "data": "synthetic", "config": {"workload": leg["workload"], "index_build_s": round(build_s, 2)},
                       "gpu_launches": leg["gpu_launches"], "e2e": leg["e2e"], "config4": leg}
        else:
            out = run_config3(a, env, api, ix, build_s)
            if not a.no_legs:
                legs = {}
                if env.world == 1:
                    legs["n_reads"] = run_n_reads(a, env, api, ix)
                legs["config4"] = run_config4(a, env, api, ix, 2, 1)
                cpu_port_close()
                ix.close()
                ix = None
                legs["config5"] = run_config5(a, env, api, 2, 1)
                if env.world == 1:
                    legs["config1"] = run_config1(a, env, api)
                if env.rank == 0:
                    out.update(legs)
        cpu_port_close()
        if ix is not None:
            ix.close()
    if env.rank == 0:
        print(json.dumps(out), flush=True)
        bad = [k for k in ("parity_sample",) if (out.get(k) or {}).get("mismatches", 0) != 0]
        bad += [k for k in ("n_reads", "config1") if isinstance(out.get(k), dict) and
                (out[k].get("parity_sample") or out[k].get("parity") or {}).get("mismatches", 0) != 0]
        if bad:
            print("bench.py: GPU hits differ from the CPU port on the sampled reads (%s)" % ", ".join(bad), file=sys.stderr)
            rc = 3
    if env.dist is not None:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return rc


def run_reference(a, env, api, ix):
    """--impl reference: the reference's CPU algorithm (C port in oracle/; the Erlang original needs OTP, absent here) on
    all host threads; each step = a bounded sample of the same workload."""
    threads = env.threads
    reads, offs, keep = sample(api, a, ix, a.reads, a.read_len, a.seed + 1000, env, pinned=False)
    orc, setup_s = cpu_port_setup(ix)
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
    erl = subprocess.run(["which", "erl"], capture_output=True, text=True).stdout.strip()
    sample_txt = "each step = %d reads of the %d-read workload, %d threads; C port of the reference (oracle/bwt_oracle.c)%s" % (
        m, a.reads, threads, "; an Erlang runtime is present (%s): scripts/make_reference_golden.escript runs the unmodified reference" % erl if erl else
        "; no Erlang runtime on this box (`which erl` is empty), so the BEAM original cannot be timed")
    _ORC.clear()
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": tot / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample_txt, "index_export_s": round(setup_s, 1)},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_txt},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
    orc.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
