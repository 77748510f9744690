"""
Device-resident exchange of the multi-GPU path (findex_b200.sharded.GpuExchange): counts, located positions and regex records go from
kernel stores into gathered buffers, never through the host.

  * one GPU playing three ranks (phase by phase, the barrier being the phase boundary) — runs wherever `-m gpu` runs;
  * the real thing under torchrun with NCCL and CUDA-IPC peer memory when the box has >= 2 GPUs: every rank must end up holding exactly
    what a single GPU computes for the whole batch (SURVEY.md §4 plan item 3).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from findex_b200 import fmindex as fx
from findex_b200 import sharded
from oracle import fm_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _batch(text, rng, m, ln):
    offs = rng.integers(0, len(text) - ln, m)
    tarr = np.frombuffer(text, np.uint8)
    pats = np.stack([tarr[s:s + ln][::-1] for s in offs]).copy()
    pats[::7] = rng.integers(97, 123, (len(pats[::7]), ln), dtype=np.uint8)
    return pats


@pytest.mark.parametrize("rate", [0, 8])
def test_exchange_three_virtual_ranks(ref_dir, rate):
    import torch
    from findex_b200 import build
    build.build()
    text = open(os.path.join(ref_dir, "test.txt"), "rb").read()
    o = fo.OracleIndex.load(os.path.join(ref_dir, "test.cmp"), big_endian=False)
    sa = o.sa().astype(np.int64)
    g = fx.GpuFMSearcher(os.path.join(ref_dir, "test.cmp.bwt"), bigEndian=False, sa_sample_rate=rate)
    rng = np.random.default_rng(12)
    world, M, ln = 3, 1000, 2
    pats = _batch(text, rng, M, ln)
    osp, oep = o.count_batch(pats.reshape(-1), np.arange(0, M * ln + 1, ln, dtype=np.int64))
    total = int((oep - osp).sum())
    dev = torch.device("cuda", 0)
    ex = [sharded.GpuExchange(g, r, world, M, total + 16, dev, peers=([], []), barrier=lambda: None) for r in range(world)]
    for e in ex:
        e.set_peers([x.counts.ptr for x in ex], [x.values.ptr for x in ex])
    d_pat = torch.from_numpy(pats).cuda()
    bounds = [sharded.shard_bounds(M, r, world) for r in range(world)]
    for r, (lo, hi) in enumerate(bounds):
        ex[r].locate_count(d_pat[lo:hi].contiguous(), ln, lo, hi)
    torch.cuda.synchronize()
    outs = [ex[r].locate_values(lo, hi, total + 16) for r, (lo, hi) in enumerate(bounds)]
    torch.cuda.synchronize()
    want_off = np.concatenate([[0], np.cumsum(oep - osp)])
    want_pos = np.concatenate([np.sort(sa[a:b]) for a, b in zip(osp, oep)] + [np.zeros(0, np.int64)])
    for r in range(world):
        assert np.array_equal(outs[r][0].cpu().numpy(), want_off)
        got = ex[r].gathered_values(total).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        assert np.array_equal(got, want_pos), r
    # regex records through the same buffers
    rxs = ["%c.%c" % (97 + i % 26, 97 + (i // 26) % 26) for i in range(300)] + ["ab?c", "q(u|a)", "zzzz"]
    Mr = len(rxs)
    want = [o.regex_match(r) for r in rxs]
    nres = sum(len(w) for w in want)
    ex2 = [sharded.GpuExchange(g, r, world, Mr, 4 * nres + 16, dev, peers=([], []), barrier=lambda: None) for r in range(world)]
    for e in ex2:
        e.set_peers([x.counts.ptr for x in ex2], [x.values.ptr for x in ex2])
    rb = [sharded.shard_bounds(Mr, r, world) for r in range(world)]
    sets = [g.regex_set([fx.ReTree(x) for x in rxs[lo:hi]]) for lo, hi in rb]
    for r, (lo, hi) in enumerate(rb):
        ex2[r].regex_count(sets[r], lo, hi, nres + 4)
    torch.cuda.synchronize()
    routs = [ex2[r].regex_values(lo, hi) for r, (lo, hi) in enumerate(rb)]
    torch.cuda.synchronize()
    flat = [(i, l, s, e) for i, w in enumerate(want) for (l, s, e) in w]
    for r in range(world):
        off = routs[r][0].cpu().numpy()
        assert off[-1] == nres and np.array_equal(np.diff(off), [len(w) for w in want])
        rec = ex2[r].gathered_values(4 * nres).cpu().numpy().reshape(-1, 4).astype(np.int64) & 0xFFFFFFFF
        assert [tuple(x) for x in rec.tolist()] == flat
    for s in sets:
        s.close()
    for e in ex + ex2:
        e.close()
    g.close()
    o.close()


def test_exchange_two_gpus_torchrun(tmp_path):
    """the same through NCCL + CUDA IPC on two real GPUs (skipped on a one-GPU box): tests/multi_gpu_worker.py"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "ok"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert out.exists()
