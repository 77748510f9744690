"""
tools/microbench.py — design-space sweep on a real B200 (run under gpurun): index build time, count throughput per
layout x lanes-per-query, distinct-block accounting, and the K4 random-gather roofline (32/64/128-B gathers).
Writes JSON lines to gpurun_out/microbench.jsonl.  Not part of the bench contract; evidence for DESIGN.md.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--m", type=int, default=10_000_000)
    ap.add_argument("--len", type=int, default=16)
    ap.add_argument("--alphabet", default="bytes255", choices=["bytes255", "dna"])
    ap.add_argument("--layouts", default="wm,planes")
    ap.add_argument("--out", default="gpurun_out/microbench.jsonl")
    ap.add_argument("--gather", type=int, default=1)
    ap.add_argument("--l2fetch", default="0")
    args = ap.parse_args()
    fbuild.build()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    out = open(args.out, "a")

    def emit(**kw):
        print(json.dumps(kw), flush=True)
        out.write(json.dumps(kw) + "\n")
        out.flush()

    rng = np.random.default_rng(2)
    t0 = time.time()
    if args.alphabet == "bytes255":
        text = rng.integers(1, 256, args.n, dtype=np.uint8)
    else:
        text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, args.n, dtype=np.uint8)]
    t1 = time.time()
    base = "/tmp/mb_index"
    fx.build_index_files(text, base, bigEndian=True)
    t2 = time.time()
    emit(what="build", n=args.n, gen_s=t1 - t0, build_files_s=t2 - t1)

    rq = np.random.default_rng(3)
    nh = int(args.m * 0.9)
    offs = rq.integers(0, args.n - args.len, nh)
    pats = np.empty((args.m, args.len), np.uint8)
    idx = offs[:, None] + np.arange(args.len - 1, -1, -1)[None, :]        # reversed substring = what search() wants
    pats[:nh] = text[idx]
    if args.alphabet == "bytes255":
        pats[nh:] = rq.integers(1, 256, (args.m - nh, args.len), dtype=np.uint8)
    else:
        pats[nh:] = np.frombuffer(b"ACGT", np.uint8)[rq.integers(0, 4, (args.m - nh, args.len))]
    perm = rq.permutation(args.m)
    pats = pats[perm]
    d_pat = torch.from_numpy(pats).cuda()
    d_sp = torch.zeros(args.m, dtype=torch.int32, device="cuda")
    d_ep = torch.zeros(args.m, dtype=torch.int32, device="cuda")
    ref = None
    for lay in args.layouts.split(","):
        t0 = time.time()
        g = fx.GpuFMSearcher(base + ".bwt", layout={"wm": fx.LAYOUT_WM, "planes": fx.LAYOUT_PLANES}[lay])
        emit(what="open", open_s=time.time() - t0, **g.info())
        blocks, steps = g.count_fixed_stats(pats[:2_000_000])
        emit(what="stats", layout=lay, queries=2_000_000, blocks=blocks, steps=steps, bytes_per_query=blocks * 64 / 2e6)
        for l2f in [int(x) for x in args.l2fetch.split(",")]:
            eff = fx.set_l2_fetch_granularity(l2f)
            for nbytes, lanes in ((32, 2), (64, 4), (128, 4)):
                gbs, ms = g.gather_bench(nbytes, lanes, 1 << 25, 16, 3)
                emit(what="gather_l2fetch", layout=lay, l2fetch_req=l2f, l2fetch_eff=eff, bytes=nbytes, lanes=lanes, gbs=gbs, ggathers_per_s=gbs / nbytes)
            g.set_lanes(4)
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(2):
                g.count_fixed_dev(d_pat.data_ptr(), args.len, args.m, d_sp.data_ptr(), d_ep.data_ptr(), st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.count_fixed_dev(d_pat.data_ptr(), args.len, args.m, d_sp.data_ptr(), d_ep.data_ptr(), st)
            e1.record()
            torch.cuda.synchronize()
            emit(what="count_l2fetch", layout=lay, l2fetch_req=l2f, l2fetch_eff=eff, ms=e0.elapsed_time(e1) / 5)
        for lanes in (1, 2, 4):
            g.set_lanes(lanes)
            st = torch.cuda.current_stream().cuda_stream
            for _ in range(2):
                g.count_fixed_dev(d_pat.data_ptr(), args.len, args.m, d_sp.data_ptr(), d_ep.data_ptr(), st)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.count_fixed_dev(d_pat.data_ptr(), args.len, args.m, d_sp.data_ptr(), d_ep.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res = (d_sp.cpu().numpy().copy(), d_ep.cpu().numpy().copy())
            if ref is None:
                ref = res
            same = bool(np.array_equal(ref[0], res[0]) and np.array_equal(ref[1], res[1]))
            hits = int((res[1] > res[0]).sum())
            emit(what="count", layout=lay, lanes=lanes, ms=best, qps=args.m / best * 1e3, hits=hits, same_as_first=same,
                 gbs=blocks * 64 / 2e6 * args.m / best * 1e3 / 1e9)
        if args.gather:
            for nbytes, lanes_list in ((32, (1, 2)), (64, (1, 2, 4)), (128, (1, 2, 4, 8))):
                for lanes in lanes_list:
                    for chain in (1, 16):
                        gbs, ms = g.gather_bench(nbytes, lanes, 1 << 25, chain, 3)
                        emit(what="gather", layout=lay, bytes=nbytes, lanes=lanes, chain=chain, gbs=gbs, ms=ms,
                             ggathers_per_s=gbs / nbytes)
        g.close()


if __name__ == "__main__":
    main()
