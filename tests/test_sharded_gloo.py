"""
World-size-2 (and 3) gloo tests of the multi-GPU plumbing on CPU: query sharding + all-gather of fixed-size results and
of variable-length results (locate / regex).  The compute callable is the CPU oracle here (tests may use it); on the GPU
box the same ShardedSearcher wraps the GPU searcher.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from findex_b200 import sharded
from oracle import fm_oracle as fo

REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_searcher(o):
    def count_fn(p):
        p = np.ascontiguousarray(p)
        return o.count_batch(p.reshape(-1), np.arange(0, p.size + 1, p.shape[1], dtype=np.int64))

    def locate_fn(sp, ep):
        parts = [o.locate(int(a), int(b)) if b > a else np.zeros(0, np.int64) for a, b in zip(sp, ep)]
        off = np.zeros(len(sp) + 1, np.int64)
        off[1:] = np.cumsum([len(x) for x in parts])
        return off, (np.concatenate(parts) if parts else np.zeros(0, np.int64))

    return sharded.ShardedSearcher(count_fn=count_fn, locate_fn=locate_fn, regex_fn=lambda rxs: [o.regex_match(r) for r in rxs])


def _workload():
    text = open(os.path.join(REF, "test.txt"), "rb").read()
    rng = np.random.default_rng(5)
    m = 1001                                              # not divisible by 2 or 3: uneven shards
    offs = rng.integers(0, len(text) - 4, m)
    pats = np.stack([np.frombuffer(text[s:s + 4][::-1], np.uint8) for s in offs])
    pats[::7] = rng.integers(97, 123, (len(pats[::7]), 4), dtype=np.uint8)
    rxs = ["a(b|c)d", "x[a-f]y", "q.z", "ab?c", "k(l|m)+n", "zz", "a.(b|c)", "qu*"]
    return pats, rxs


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = fo.OracleIndex.load(os.path.join(REF, "test.cmp"), big_endian=False)
        s = _make_searcher(o)
        pats, rxs = _workload()
        sp, ep = s.count(pats)
        off, pos = s.locate(sp[:200], ep[:200])
        rx = s.regex_search(rxs)
        q.put((rank, sp, ep, off, pos, rx))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_results_equal_single_process(world):
    o = fo.OracleIndex.load(os.path.join(REF, "test.cmp"), big_endian=False)
    s1 = _make_searcher(o)
    pats, rxs = _workload()
    sp1, ep1 = s1.count(pats)                             # world size 1 path (no process group)
    off1, pos1 = s1.locate(sp1[:200], ep1[:200])
    rx1 = s1.regex_search(rxs)
    assert (ep1 > sp1).sum() > 500 and len(pos1) > 100 and sum(len(r) for r in rx1) > 5

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, sp, ep, off, pos, rx in got:                # every rank holds the full, identical result
        assert np.array_equal(sp, sp1) and np.array_equal(ep, ep1)
        assert np.array_equal(off, off1) and np.array_equal(pos, pos1)
        assert rx == rx1


def test_shard_bounds_cover_everything():
    for m in (0, 1, 7, 1000, 10_000_019):
        for w in (1, 2, 3, 8):
            b = [sharded.shard_bounds(m, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == m and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
