"""
tools/profile_legs.py — the kernels of the BASELINE configs in isolation, for ncu (run under gpurun):

    ncu --set full --clock-control none --import-source on -k regex:regex_queue_kernel -c 1 --launch-skip 2 -o gpurun_out/rx python tools/profile_legs.py --what regex
    ncu ... -k regex:locate_kernel -c 1 ...  --what locate          ncu ... -k regex:count_fixed_kernel -c 1 --launch-skip 3 ... --what count3 | count2

count2 = cfg-2 index (10^9 uniform bytes), 10 M len-16 queries; count3 / locate / regex = cfg-3 index (10^9 English-like bytes): 4 M len-12
count queries, 4000 len-12 locate queries (~5 x 10^8 occurrences, SA sample rate 32), 100 k template regexes.  Prints the CUDA-event
times of the un-profiled run so that a profile can be related to the bench numbers.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from findex_b200 import build as fbuild  # noqa: E402
from findex_b200 import fmindex as fx  # noqa: E402
from findex_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="regex")
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--rx-local", default="", help="comma list of fmx_set_regex_local_keep values to time the regex search with")
    args = ap.parse_args()
    import torch
    fbuild.build()
    what = set(args.what.split(","))
    n = args.n
    st = torch.cuda.current_stream().cuda_stream
    if "count2" in what:
        text = bench.make_text(n, "cfg2")
        base = bench.index_base(n, "cfg2")
        if not os.path.exists(base + ".bwt"):
            fx.build_index_files(text, base, bigEndian=True)
        g = fx.GpuFMSearcher(base + ".bwt")
        pats, _ = bench.make_queries(text, 10_000_000, 16, 3, 0)
        d_pat = torch.from_numpy(pats).cuda()
        d_sp = torch.zeros(len(pats), dtype=torch.int32, device="cuda")
        d_ep = torch.zeros_like(d_sp)
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.count_fixed_dev(d_pat.data_ptr(), 16, len(pats), d_sp.data_ptr(), d_ep.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
        print(json.dumps({"what": "count2", "kernel_ms": e0.elapsed_time(e1), **g.info()}), flush=True)
        g.close()
        del text, pats
    if what & {"count3", "locate", "regex"}:
        text = bench.make_text(n, "cfg3")
        base = bench.index_base(n, "cfg3")
        if not os.path.exists(base + ".bwt"):
            fx.build_index_files(text, base, bigEndian=True)
        g = fx.GpuFMSearcher(base + ".bwt", sa_sample_rate=32)
        print(json.dumps({"what": "open cfg3", **g.info()}), flush=True)
        if "count3" in what:
            pats, _ = bench.make_queries(text, 4_000_000, 12, 5012, 0, workload="cfg3")
            d_pat = torch.from_numpy(pats).cuda()
            d_sp = torch.zeros(len(pats), dtype=torch.int32, device="cuda")
            d_ep = torch.zeros_like(d_sp)
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.count_fixed_dev(d_pat.data_ptr(), 12, len(pats), d_sp.data_ptr(), d_ep.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
            print(json.dumps({"what": "count3", "kernel_ms": e0.elapsed_time(e1)}), flush=True)
        if "locate" in what:
            pats, _ = bench.make_queries(text, 4000, 12, 5, 0, workload="cfg3")
            sp, ep = g.count_fixed(pats)
            tot = int((ep - sp).sum())
            d_sp = torch.from_numpy(sp.astype(np.uint32).view(np.int32)).cuda()
            d_ep = torch.from_numpy(ep.astype(np.uint32).view(np.int32)).cuda()
            d_off = torch.zeros(len(pats) + 1, dtype=torch.int64, device="cuda")
            d_pos = torch.zeros(tot + 16, dtype=torch.int32, device="cuda")
            for _ in range(max(2, args.reps // 2)):
                g.locate_dev(d_sp.data_ptr(), d_ep.data_ptr(), len(pats), d_off.data_ptr(), d_pos.data_ptr(), tot + 16, st)
            print(json.dumps({"what": "locate", "occurrences": tot, "walk_sort_ms": g.last_locate_ms()}), flush=True)
        if "regex" in what:
            rxs = synth.regex_templates(text, np.random.default_rng([6, 0]), 100_000)
            trees = [fx.ReTree(r) for r in rxs]
            rset = g.regex_set(trees)
            cap = 1 << 22
            d_res = torch.zeros((cap, 4), dtype=torch.int32, device="cuda")
            fx.lib().fmx_set_regex_local_keep.argtypes = [__import__("ctypes").c_int32]
            for keep in ([int(x) for x in args.rx_local.split(",")] if args.rx_local else [64]):
                fx.lib().fmx_set_regex_local_keep(keep)
                ms = []
                for _ in range(args.reps):
                    t0 = time.time()
                    total = rset.search_dev(g, d_res.data_ptr(), cap, 0)
                    wall = time.time() - t0
                    ms.append(g.last_kernel_ms())
                print(json.dumps({"what": "regex", "local_keep": keep, "results": total, "kernel_ms_best": min(ms), "kernel_ms_last": ms[-1], "call_ms": wall * 1e3,
                                  "items": g.last_steps(), "launches": g.last_kernel_launches()}), flush=True)
            fx.lib().fmx_set_regex_local_keep(64)
            rset.close()
        g.close()


if __name__ == "__main__":
    main()
