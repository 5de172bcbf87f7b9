"""N > 1 host logic on CPU: the partition gather (the path's only exchange step) over gloo with world_size 2, and the
merge it feeds.  No GPU, no CUDA library calls."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bwt_b200 import api, partition


def _fake_partition_result(p, n, seed=5):
    """deterministic stand-in for one partition's (status, offsets, hits, cigars)"""
    rng = np.random.default_rng(seed + 17 * p)
    counts = rng.integers(0, 4, size=n)
    ho = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=ho[1:])
    t = int(ho[n])
    hits = np.zeros(t, api.HIT_DTYPE)
    hits["s"] = rng.integers(100, 10 ** 9, size=t)
    hits["d"] = rng.integers(0, 8, size=t)
    hits["pos"] = hits["s"] - 100 - hits["d"] + 1000 * p
    hits["score"] = rng.integers(100, 200, size=t)
    hits["status"] = rng.integers(0, 2, size=t)
    cl = np.where(hits["status"] == 0, rng.integers(4, 12, size=t), 0)
    hits["cigar_len"] = cl
    hits["cigar_off"] = np.concatenate(([0], np.cumsum(cl)))[:-1]
    cig = bytes(rng.integers(48, 58, size=int(cl.sum()), dtype=np.uint8))
    status = rng.integers(0, 2, size=n).astype(np.int32)
    return status, ho, hits, cig


def _check_merged(merged, parts):
    status, ho, hits, cig = merged
    n = len(ho) - 1
    assert status.shape == (len(parts), n)
    for r in range(n):
        want = []
        for p, (st, pho, hh, cg) in enumerate(parts):
            for h in hh[pho[r]:pho[r + 1]]:
                want.append((p, int(h["pos"]), int(h["s"]), int(h["d"]), int(h["score"]), int(h["status"]),
                             bytes(cg[int(h["cigar_off"]):int(h["cigar_off"]) + int(h["cigar_len"])])))
        got = [(int(h["part"]), int(h["pos"]), int(h["s"]), int(h["d"]), int(h["score"]), int(h["status"]),
                cig[int(h["cigar_off"]):int(h["cigar_off"]) + int(h["cigar_len"])]) for h in hits[ho[r]:ho[r + 1]]]
        assert got == want, r


def test_merge_is_concatenation_in_partition_order():
    n = 200
    parts = []
    for p in range(3):
        st, ho, hh, cg = _fake_partition_result(p, n)
        parts.append((st, ho, hh, np.frombuffer(cg, np.uint8)))
    _check_merged(partition.merge_partitions(parts), parts)
    empty = [(np.zeros(0, np.int32), np.zeros(1, np.int64), np.zeros(0, api.HIT_DTYPE), np.zeros(0, np.uint8))] * 2
    st, ho, hits, cig = partition.merge_partitions(empty)
    assert list(ho) == [0] and len(hits) == 0 and cig == b""


def test_partition_slices_follow_builder_formula():
    # fm_index.erl:113-124: Chunk = (Len + Nbr*2) div Nbr, Start_J = 1 + (J-1)(Chunk-2): consecutive chunks overlap by 2
    sl = partition.partition_slices(3_137_454_466, 8)
    assert len(sl) == 8 and sl[0][0] == 0 and sl[0][2] == 1
    for a, b in zip(sl, sl[1:]):
        assert b[0] == a[0] + a[1] - 2 and b[2] == b[0] + 1
    assert sl[-1][0] + sl[-1][1] == 3_137_454_466
    assert partition.partition_slices(100, 1) == [(0, 100, 1)]


def _fake_seed_result(p, n, seed=9):
    rng = np.random.default_rng(seed + 31 * p)
    counts = rng.integers(0, 4, size=n)
    so = np.zeros(n + 1, np.int64)
    np.cumsum(counts, out=so[1:])
    sd = np.concatenate([np.sort(rng.integers(1, 10_000, size=c))[::-1] for c in counts] + [np.zeros(0, np.int64)]).astype(np.int64)
    st = rng.integers(0, 2, size=n).astype(np.int32)
    return st, so, sd, 1 + 9_998 * p


def _check_merged_seeds(merged, parts):
    st, so, pos, part = merged
    n = len(parts[0][1]) - 1
    for r in range(n):
        want = []
        for p, (pst, pso, sd, shift) in enumerate(parts):
            want += [(int(v) - 1 + shift, p) for v in sd[pso[r]:pso[r + 1]]]
        assert list(zip(pos[so[r]:so[r + 1]].tolist(), part[so[r]:so[r + 1]].tolist())) == want
    for p in range(len(parts)):
        assert (st[p] == parts[p][0]).all()


def test_seed_merge_is_concatenation_with_contig_coordinates():
    parts = [_fake_seed_result(p, 200) for p in range(3)]
    _check_merged_seeds(partition.merge_seed_partitions(parts), parts)


def _seed_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st, so, sd, shift = _fake_seed_result(rank, n)
        merged = partition.gather_seeds(st, so, sd, shift, device="cpu")
        if rank == 0:
            _check_merged_seeds(merged, [_fake_seed_result(p, n) for p in range(world)])
            q.put("ok")
        else:
            assert merged is None
    except Exception as e:
        q.put("rank %d: %r" % (rank, e))
        raise
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st, ho, hh, cg = _fake_partition_result(rank, n)
        merged = partition.gather_hits(st, ho, hh, cg, device="cpu")
        if rank == 0:
            parts = []
            for p in range(world):
                s2, h2, hh2, c2 = _fake_partition_result(p, n)
                parts.append((s2, h2, hh2, np.frombuffer(c2, np.uint8)))
            _check_merged(merged, parts)
            q.put("ok")
        else:
            assert merged is None
    except Exception as e:  # surface the failure in the parent
        q.put("rank %d: %r" % (rank, e))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_over_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"


@pytest.mark.timeout(120)
def test_seed_gather_over_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_seed_worker, args=(r, 2, port, 300, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert q.get(timeout=5) == "ok"
