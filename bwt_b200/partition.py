"""Genome-partition mode (BASELINE.json configs[4]): one ``_pJ`` FM-index partition per GPU, every
GPU aligns ALL reads against its partition, and the only exchange step is a gather of the per-(read, partition) hit
records to rank 0 over ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference defines no reduction across partitions -- the worker hard-codes partition 1 (worker_bwt.erl:48,53)
and the master prints every cigar it receives (master.erl:151-155) -- so the merged result of a read is the
concatenation of its per-partition ``Cigars`` lists in partition order, and parity is per partition.

Partition geometry follows the builder's formula generalised to ``Nbr`` chunks (fm_index.erl:113-124):
``Chunk = (Len + Nbr*2) div Nbr``, ``Start_J = 1 + (J-1)(Chunk-2)``, meta shift = contig shift + Start_J.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from . import api
from .fm_index import chunk_plan

MERGED_DTYPE = np.dtype([("pos", "<i8"), ("s", "<i8"), ("cigar_off", "<i8"), ("d", "<i4"), ("score", "<i4"),
                         ("status", "<i4"), ("cigar_len", "<i4"), ("part", "<i4"), ("pad", "<i4")])


def partition_slices(body_len: int, nbr: int) -> List[Tuple[int, int, int]]:
    """-> [(start0 (0-based slice start), length, shift_add (= Start_J, 1-based))] for J = 1..nbr"""
    return [(start - 1, length, start) for (start, length) in chunk_plan(body_len, nbr)]


def build_partition(ref, nbr: int, j: int, contig_shift: int = 0, device: int = 0) -> api.Index:
    """Index of partition j (1-based) of ``ref`` (uint8 array or bytes) on ``device``."""
    start0, length, add = partition_slices(len(ref), nbr)[j - 1]
    chunk = ref[start0:start0 + length]
    return api.Index.build(chunk, shift=contig_shift + add, device=device)


def _as_torch(a: np.ndarray, device):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(device) if str(device) != "cpu" else t


def gather_hits(read_status: np.ndarray, hit_offsets: np.ndarray, hits: np.ndarray, cigars: bytes, group=None,
                device="cpu"):
    """The exchange step.  Every rank passes its partition's result for the SAME reads; rank 0 gets the merged
    result ``(read_status[world, n], hit_offsets[n+1], hits MERGED_DTYPE[total], cigars bytes)``, the others None.

    Two collectives: all_gather of the sizes, then all_gather of the fixed-size records (offsets, 40-byte hits) and
    the CIGAR bytes, each padded to the largest rank.  Payload is a few tens of bytes per read and partition."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = len(hit_offsets) - 1
    nh, nc = len(hits), len(cigars)
    sizes = torch.tensor([nh, nc], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [tuple(int(v) for v in s.cpu()) for s in all_sizes]
    max_h = max(s[0] for s in all_sizes)
    max_c = max(s[1] for s in all_sizes)
    # one byte buffer per rank: status | offsets | hits (padded) | cigars (padded)
    rec = np.zeros(4 * n + 8 * (n + 1) + 40 * max_h + max_c, dtype=np.uint8)
    o = 0
    rec[o:o + 4 * n] = np.ascontiguousarray(read_status, np.int32).view(np.uint8); o += 4 * n
    rec[o:o + 8 * (n + 1)] = np.ascontiguousarray(hit_offsets, np.int64).view(np.uint8); o += 8 * (n + 1)
    rec[o:o + 40 * nh] = np.ascontiguousarray(hits).view(np.uint8); o += 40 * max_h
    rec[o:o + nc] = np.frombuffer(cigars, np.uint8)
    mine = _as_torch(rec, device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    if rank != 0:
        return None
    parts = []
    for p in range(world):
        buf = gathered[p].cpu().numpy()
        ph, pc = all_sizes[p]
        o = 0
        st = buf[o:o + 4 * n].view(np.int32); o += 4 * n
        ho = buf[o:o + 8 * (n + 1)].view(np.int64); o += 8 * (n + 1)
        hh = buf[o:o + 40 * ph].view(api.HIT_DTYPE); o += 40 * max_h
        cg = buf[o:o + pc]
        parts.append((st, ho, hh, cg))
    return merge_partitions(parts)


def merge_partitions(parts):
    """parts[p] = (read_status[n], hit_offsets[n+1], hits[HIT_DTYPE], cigars uint8) -> merged result; per read the hits
    of partition 0, then partition 1, ... (each already in the reference's ascending-position order)."""
    world = len(parts)
    n = len(parts[0][1]) - 1
    counts = np.stack([np.diff(p[1]) for p in parts])                    # [world, n]
    per_read = counts.sum(axis=0)
    ho = np.zeros(n + 1, np.int64)
    np.cumsum(per_read, out=ho[1:])
    total = int(ho[n])
    out = np.zeros(total, MERGED_DTYPE)
    cig_sizes = [len(p[3]) for p in parts]
    cig_base = np.concatenate(([0], np.cumsum(cig_sizes)))[:-1]
    before = np.zeros(n, np.int64)                                       # hits of earlier partitions of the same read
    for p, (st, pho, hh, cg) in enumerate(parts):
        if len(hh):
            read_of = np.repeat(np.arange(n), counts[p])
            dst = ho[read_of] + before[read_of] + (np.arange(len(hh)) - pho[read_of])
            for f in ("pos", "s", "d", "score", "status", "cigar_len"):
                out[f][dst] = hh[f]
            out["cigar_off"][dst] = hh["cigar_off"] + cig_base[p]
            out["part"][dst] = p
        before += counts[p]
    cigars = b"".join(bytes(p[3]) for p in parts)
    status = np.stack([p[0] for p in parts])
    return status, ho, out, cigars


def align_partitioned(index: api.Index, reads, offsets, group=None, device="cpu", aligner: Optional[api.Aligner] = None):
    """This rank's partition against all reads + the gather; rank 0 returns the merged result."""
    res = aligner.run(reads, offsets) if aligner is not None else index.align(reads, offsets)
    return gather_hits(res[0], res[1], res[2], res[3], group=group, device=device)


# ---- bounded-edit search (seeds.erl) across partitions: BASELINE.json configs[4] ----------------------------------------
def merge_seed_partitions(parts):
    """parts[p] = (status int32[n], seed_offsets int64[n+1], seeds int64 (1-based window starts in partition p's chunk,
    descending per query), shift (the partition's meta shift)) -> (status[world, n], offsets[n+1], pos int64[total],
    part int32[total]); per query the seeds of partition 0, then 1, ..., each turned into the 1-based contig coordinate
    (N - 1) + Shift -- the same mapping worker_bwt.erl:94 applies to window starts."""
    n = len(parts[0][1]) - 1
    counts = np.stack([np.diff(p[1]) for p in parts])
    so = np.zeros(n + 1, np.int64)
    np.cumsum(counts.sum(axis=0), out=so[1:])
    pos = np.zeros(int(so[n]), np.int64)
    part = np.zeros(int(so[n]), np.int32)
    before = np.zeros(n, np.int64)
    for p, (st, pso, sd, shift) in enumerate(parts):
        if len(sd):
            q_of = np.repeat(np.arange(n), counts[p])
            dst = so[q_of] + before[q_of] + (np.arange(len(sd)) - pso[q_of])
            pos[dst] = np.asarray(sd, np.int64) - 1 + int(shift)
            part[dst] = p
        before += counts[p]
    return np.stack([p[0] for p in parts]), so, pos, part


def gather_seeds(status: np.ndarray, seed_offsets: np.ndarray, seeds: np.ndarray, shift: int, group=None, device="cpu"):
    """The exchange step of the inexact mode: all_gather of the sizes, then of status | offsets | seeds (padded to the
    largest rank); rank 0 returns merge_seed_partitions(...), the others None."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = len(seed_offsets) - 1
    sizes = torch.tensor([len(seeds), int(shift)], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = [tuple(int(v) for v in s.cpu()) for s in all_sizes]
    max_s = max(s[0] for s in all_sizes)
    rec = np.zeros(4 * n + 8 * (n + 1) + 8 * max_s, dtype=np.uint8)
    o = 0
    rec[o:o + 4 * n] = np.ascontiguousarray(status, np.int32).view(np.uint8); o += 4 * n
    rec[o:o + 8 * (n + 1)] = np.ascontiguousarray(seed_offsets, np.int64).view(np.uint8); o += 8 * (n + 1)
    rec[o:o + 8 * len(seeds)] = np.ascontiguousarray(seeds, np.int64).view(np.uint8)
    mine = _as_torch(rec, device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    if rank != 0:
        return None
    parts = []
    for p in range(world):
        buf = gathered[p].cpu().numpy()
        ns, sh = all_sizes[p]
        o = 0
        st = buf[o:o + 4 * n].view(np.int32); o += 4 * n
        so = buf[o:o + 8 * (n + 1)].view(np.int64); o += 8 * (n + 1)
        sd = buf[o:o + 8 * ns].view(np.int64)
        parts.append((st, so, sd, sh))
    return merge_seed_partitions(parts)


def fs_find_seeds_partitioned(index: api.Index, queries, offsets, length: int, defects: int, max_raw_hits: int = 1024,
                              group=None, device="cpu"):
    """This rank's partition against all queries (bwt_fs_find_seeds_batch) + the gather; rank 0 returns the merged
    (status, offsets, contig positions, partition of each seed) and every rank the stats of its own search."""
    st, so, sd, stats = index.fs_find_seeds(queries, offsets, length, defects, max_raw_hits)
    return gather_seeds(st, so, sd, int(index.info().shift), group=group, device=device), stats
