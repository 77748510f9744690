"""
tools/exp_count.py — count-kernel design sweep on one B200 (run under gpurun): one cfg-2 style index, several accelerator
configurations x lanes-per-query, device-resident timing, request accounting, and a cross-check that every configuration returns
identical (sp,ep) for the whole batch.  Writes JSON lines to gpurun_out/exp_count.jsonl.  Evidence for DESIGN.md, not the bench.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from findex_b200 import build as fbuild  # noqa: E402
from findex_b200 import fmindex as fx  # noqa: E402
import bench  # noqa: E402  (make_text / make_queries: the bench's own workload)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--m", type=int, default=10_000_000)
    ap.add_argument("--len", type=int, default=16)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--configs", default="both,auto,ctx_k3")
    ap.add_argument("--lanes", default="1,2,4")
    ap.add_argument("--out", default="gpurun_out/exp_count.jsonl")
    args = ap.parse_args()
    fbuild.build()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(**kw):
        print(json.dumps(kw), flush=True)
        out.write(json.dumps(kw) + "\n")
        out.flush()

    text = bench.make_text(args.n, args.workload)
    base = "/tmp/exp_index_%s_%d" % (args.workload, args.n)
    t0 = time.time()
    if not os.path.exists(base + ".bwt"):
        fx.build_index_files(text, base, bigEndian=True)
    emit(what="build_files", s=time.time() - t0)
    pats, is_hit, _ = bench.make_queries(text, args.m, args.len, 3, 0, workload=args.workload)
    d_pat = torch.from_numpy(pats).cuda()
    d_sp = torch.zeros(args.m, dtype=torch.int32, device="cuda")
    d_ep = torch.zeros(args.m, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    cfgs = {
        "none": dict(accel=fx.ACCEL_NONE),
        "both": dict(accel=fx.ACCEL_KMER | fx.ACCEL_TEXT),
        "auto": dict(accel=fx.ACCEL_AUTO),
        "ctx_k3": dict(accel=fx.ACCEL_AUTO, kmer_table_bytes=200 << 20),
        "ctx_nokmer": dict(accel=fx.ACCEL_CTX),
    }
    ref = None
    for name in args.configs.split(","):
        t0 = time.time()
        g = fx.GpuFMSearcher(base + ".bwt", bigEndian=True, **cfgs[name])
        info = g.info()
        free, total = torch.cuda.mem_get_info()
        emit(what="open", config=name, open_s=time.time() - t0, free_gb=free / 1e9, **info)
        blocks, steps = g.count_fixed_stats(pats[:2_000_000])
        emit(what="stats", config=name, requests_per_query=blocks / 2e6, steps_per_query=steps / 2e6)
        for lanes in [int(x) for x in args.lanes.split(",")]:
            if lanes:
                g.set_lanes(lanes)
            for _ in range(3):
                g.count_fixed_dev(d_pat.data_ptr(), args.len, args.m, d_sp.data_ptr(), d_ep.data_ptr(), st)
            torch.cuda.synchronize()
            times = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.count_fixed_dev(d_pat.data_ptr(), args.len, args.m, d_sp.data_ptr(), d_ep.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            res = (d_sp.cpu().numpy().copy(), d_ep.cpu().numpy().copy())
            if ref is None:
                ref = res
            same = bool(np.array_equal(ref[0], res[0]) and np.array_equal(ref[1], res[1]))
            ms = float(np.median(times))
            emit(what="count", config=name, lanes=lanes, ms=ms, ms_min=float(min(times)), gqps=args.m / ms / 1e6,
                 greq_per_s=blocks / 2e6 * args.m / ms / 1e6, hits=int((res[1] > res[0]).sum()), same_as_first=same)
            assert same, "configuration %s lanes %d changed the results" % (name, lanes)
        g.close()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
