"""2-GPU debugging of sharded.GpuExchange in the order bench.py uses it: a locate exchange (retired, kept alive), then a regex exchange with
large shards; after every step each rank checks the gathered counts against an NCCL all-gather of the local ones."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from findex_b200 import fmindex as fx, sharded, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 20_000_000
    text = bench.make_text(n, "cfg3")
    base = "/tmp/dbg_cfg3"
    if rank == 0 and not os.path.exists(base + ".bwt"):
        fx.build_index_files(text, base, bigEndian=True)
    dist.barrier()
    g = fx.GpuFMSearcher(base + ".bwt", device=local, sa_sample_rate=32)
    # a first exchange that stays alive (as the locate leg's does)
    pats, _ = bench.make_queries(text, 400, 12, 5, 0, workload="cfg3")
    lo, hi = sharded.shard_bounds(400, rank, world)
    sp, ep = g.count_fixed(pats)
    tot = int((ep - sp).sum())
    ex1 = sharded.GpuExchange(g, rank, world, 400, tot + 16, dev)
    off, _ = ex1.locate(torch.from_numpy(np.ascontiguousarray(pats[lo:hi])).to(dev), 12, lo, hi, tot + 16)
    torch.cuda.synchronize()
    print("rank", rank, "locate exchange ok", int(off[-1]) == tot, flush=True)
    ex1.retire()
    mr = 20000
    rxs = synth.regex_templates(text, np.random.default_rng([6, rank]), mr)
    rset = g.regex_set([fx.ReTree(r) for r in rxs])
    cap = 1 << 20
    d_res = torch.zeros((cap, 4), dtype=torch.int32, device=dev)
    d_off = torch.zeros(mr + 1, dtype=torch.int64, device=dev)
    first = rset.search_dev(g, d_res.data_ptr(), cap, d_off.data_ptr())
    t = torch.tensor([first], dtype=torch.float64, device=dev)
    dist.all_reduce(t)
    ex = sharded.GpuExchange(g, rank, world, world * mr, 4 * int(t.item()) + 64, dev)
    print("rank", rank, "sinks", [hex(p) for p in ex.count_sinks], "own", hex(ex.counts.ptr), flush=True)
    for step in range(4):
        goff, tl = ex.regex(rset, rank * mr, (rank + 1) * mr, cap)
        torch.cuda.synchronize()
        mine = ex._cnt.to(torch.int64)
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        want = torch.cat(allc)
        got = ex._view(ex.counts, world * mr).to(torch.int64)
        bad = (got != want).nonzero().flatten()
        print("rank", rank, "step", step, "local", tl, "gathered total", int(goff[-1]), "want", int(want.sum()), "mismatches", bad.numel(),
              ("first at %d (region of rank %d)" % (int(bad[0]), int(bad[0]) // mr)) if bad.numel() else "", flush=True)
        dist.barrier()
    dist.barrier()
    sharded.close_retired(dist.barrier)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
