"""
tools/exp_dict.py — the wide-interval dictionary on the cfg-3 index (10^9 English-like bytes) under several thresholds (run under gpurun):
for every (dict_min_rows, dict_top_min_rows) pair the index is opened once and counted with 4 M text-substring queries of several
lengths; prints q/s, requests per query (instrumented run), the dictionary's depth / entries / bytes, and oracle parity on a sample.
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--queries", type=int, default=4_000_000)
    ap.add_argument("--lens", default="8,10,12,16,20")
    ap.add_argument("--plans", default="none;8,8;8,1;4,1;2,1", help="semicolon list of min_rows,top_min_rows ('none' = no dictionary)")
    ap.add_argument("--dict-gb", type=float, default=0)
    ap.add_argument("--oracle", type=int, default=20000)
    args = ap.parse_args()
    import torch
    fbuild.build()
    n = args.n
    text = bench.make_text(n, "cfg3")
    base = bench.index_base(n, "cfg3")
    if not os.path.exists(base + ".bwt"):
        fx.build_index_files(text, base, bigEndian=True)
    orc = None
    if args.oracle:
        from oracle import fm_oracle as fo
        orc = fo.OracleIndex.load(base, big_endian=True)
    st = torch.cuda.current_stream().cuda_stream
    lens = [int(x) for x in args.lens.split(",")]
    batches = {ln: bench.make_queries(text, args.queries, ln, 5000 + ln, 0, workload="cfg3")[0] for ln in lens}
    for plan in args.plans.split(";"):
        kw = {}
        if plan == "none":
            kw["accel"] = fx.ACCEL_KMER | fx.ACCEL_CTX
        else:
            a, b = (int(x) for x in plan.split(","))
            kw.update(dict_min_rows=a, dict_top_min_rows=b)
            if args.dict_gb:
                kw["dict_bytes"] = int(args.dict_gb * (1 << 30))
        t0 = time.time()
        g = fx.GpuFMSearcher(base + ".bwt", **kw)
        info = g.info()
        row = {"plan": plan, "open_s": round(time.time() - t0, 2), "index_gb": round(info["index_bytes"] / 1e9, 2), "kmer_k": info["kmer_k"],
               "dict_depth": info["dict_depth"], "dict_entries": info["dict_entries"], "dict_gb": round(info["dict_bytes"] / 1e9, 3)}
        for ln in lens:
            pats = batches[ln]
            d_pat = torch.from_numpy(pats).cuda()
            d_sp = torch.zeros(len(pats), dtype=torch.int32, device="cuda")
            d_ep = torch.zeros_like(d_sp)
            best = 1e9
            for _ in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.count_fixed_dev(d_pat.data_ptr(), ln, len(pats), d_sp.data_ptr(), d_ep.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            blocks, _ = g.count_fixed_stats(pats[:400_000])
            ok = None
            if orc is not None:
                k = args.oracle
                sp = d_sp[:k].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
                ep = d_ep[:k].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
                osp, oep = orc.count_batch(pats[:k].reshape(-1), np.arange(0, k * ln + 1, ln, dtype=np.int64))
                ok = bool(np.array_equal(sp, osp) and np.array_equal(ep, oep))
            row["len%d" % ln] = {"gqps": round(len(pats) / best / 1e6, 2), "ms": round(best, 4), "req_per_q": round(blocks / 400_000, 2), "parity": ok}
        print(json.dumps(row), flush=True)
        g.close()


if __name__ == "__main__":
    main()
