"""
tools/footprint_curve.py — count throughput versus resident index bytes on one B200 (run under gpurun).

Opens the cfg-2 index (10^9 uniform bytes) and the cfg-3 index (10^9 English-like bytes) under a series of fmx_opts.max_total_bytes caps —
FMX_LAYOUT_AUTO / FMX_ACCEL_AUTO pick the fastest combination that fits — and measures the device-resident count rate of len-16 (cfg 2) and
len-12 (cfg 3) queries, with requests/query and oracle-free parity (every configuration must return the same (sp, ep) checksum).
JSON lines to --out; DESIGN.md §3 quotes the table.
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


class _Cx:
    """the bit of bench.Ctx that device_count_rate needs"""

    def __init__(self):
        import torch
        self.torch, self.dev, self.launches, self.rank, self.world = torch, torch.device("cuda", 0), 0, 0, 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--queries", type=int, default=4_000_000)
    ap.add_argument("--out", default="gpurun_out/footprint.jsonl")
    ap.add_argument("--workloads", default="cfg2,cfg3")
    args = ap.parse_args()
    fbuild.build()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fh = open(args.out, "a")
    cx = _Cx()
    n = args.n
    for w in args.workloads.split(","):
        text = bench.make_text(n, w)
        base = bench.index_base(n, w)
        if not os.path.exists(base + ".bwt"):
            fx.build_index_files(text, base, bigEndian=True)
        ln = 16 if w == "cfg2" else 12
        pats, _ = bench.make_queries(text, args.queries, ln, 3 if w == "cfg2" else 5, 0, workload=w)
        sums = set()
        # caps in bytes per text byte; explicit layouts first (the structures north_star names), then AUTO under growing caps
        plans = [("wm, no accelerators", dict(layout=fx.LAYOUT_WM, accel=fx.ACCEL_NONE)),
                 ("wmx, no accelerators", dict(layout=fx.LAYOUT_WMX, accel=fx.ACCEL_NONE)),
                 ("planes, no accelerators", dict(layout=fx.LAYOUT_PLANES, accel=fx.ACCEL_NONE))]
        for cap in (2.2, 3.3, 4.5, 8, 16, 40, 48, 72, 90, 0):
            plans.append(("auto, cap %.1f n" % cap if cap else "auto, no cap", dict(max_total_bytes=int(cap * n))))
        for label, kw in plans:
            t0 = time.time()
            try:
                g = fx.GpuFMSearcher(base + ".bwt", **kw)
            except fx.FmxError as e:
                rec = {"workload": w, "plan": label, "error": str(e)}
                print(json.dumps(rec), flush=True)
                fh.write(json.dumps(rec) + "\n")
                continue
            info = g.info()
            ms, sp, ep = bench.device_count_rate(cx, g, pats, reps=3)
            req, steps = g.count_fixed_stats(pats)
            cs = int((sp * 1315423911 + ep * 2654435761).sum() & 0xFFFFFFFFFFFF)
            sums.add(cs)
            rec = {"workload": w, "plan": label, "open_s": time.time() - t0, "index_bytes": info["index_bytes"], "bytes_per_text_byte": info["index_bytes"] / n,
                   "layout": info["layout"], "kmer_k": info["kmer_k"], "ctx_entry_bytes": info["ctx_entry_bytes"], "lanes": info["lanes_per_query"],
                   "pattern_len": ln, "queries": args.queries, "kernel_ms": ms, "queries_per_s": args.queries / (ms * 1e-3), "requests_per_query": req / args.queries,
                   "checksum": cs}
            print(json.dumps(rec), flush=True)
            fh.write(json.dumps(rec) + "\n")
            fh.flush()
            g.close()
        assert len(sums) == 1, "configurations disagree on (sp, ep): %s" % sums
        del text, pats


if __name__ == "__main__":
    main()
