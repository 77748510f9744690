"""
tools/sweep_count.py — count throughput versus pattern length / text / accelerator set on one B200 (run under gpurun).

For every (index, length): device-resident queries/s of fmx_count_fixed_dev (CUDA events, best of a few repeats), requests per query and
skipped steps from the instrumented kernel, and bit-exact parity with the CPU oracle on a sample.  JSON lines to --out.
bench.py carries the same sweep inside its default line (extra.sweep); this script is the stand-alone form with more knobs.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findex_b200 import build as fbuild  # noqa: E402
from findex_b200 import fmindex as fx  # noqa: E402
from findex_b200 import synth  # noqa: E402


def device_rate(g, torch, pats, reps=5):
    m, ln = pats.shape
    d_pat = torch.from_numpy(pats).cuda()
    d_sp = torch.zeros(m, dtype=torch.int32, device="cuda")
    d_ep = torch.zeros(m, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        g.count_fixed_dev(d_pat.data_ptr(), ln, m, d_sp.data_ptr(), d_ep.data_ptr(), st)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.count_fixed_dev(d_pat.data_ptr(), ln, m, d_sp.data_ptr(), d_ep.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sp = d_sp.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    ep = d_ep.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    return best, sp, ep


def sweep(g, torch, text, lens, m, seed, oracle=None, sample=100_000, hit_frac=0.9, alphabet=None, label=""):
    out = []
    for ln in lens:
        rng = np.random.default_rng([seed, ln])
        nh = int(m * hit_frac)
        hits, _ = synth.reversed_substrings(text, nh, ln, rng)
        if alphabet is None:
            rnd = rng.integers(1, 256, (m - nh, ln), dtype=np.uint8)
        else:
            rnd = alphabet[rng.integers(0, len(alphabet), (m - nh, ln))]
        pats = np.concatenate([hits, rnd])[rng.permutation(m)]
        ms, sp, ep = device_rate(g, torch, pats)
        req, steps = g.count_fixed_stats(pats)
        rec = {"what": "count_sweep", "index": label, "len": ln, "queries": m, "kernel_ms": ms, "queries_per_s": m / (ms * 1e-3),
               "requests_per_query": req / m, "steps_per_query": steps / m, "hits": int((ep > sp).sum())}
        if oracle is not None:
            k = min(sample, m)
            osp, oep = oracle.count_batch(pats[:k].reshape(-1), np.arange(0, k * ln + 1, ln, dtype=np.int64), threads=os.cpu_count())
            rec["parity_sample"] = k
            rec["parity"] = bool(np.array_equal(np.where(ep[:k] > sp[:k], sp[:k], 0), osp) and np.array_equal(np.where(ep[:k] > sp[:k], ep[:k], 0), oep))
            assert rec["parity"], (label, ln)
        out.append(rec)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--queries", type=int, default=10_000_000)
    ap.add_argument("--lens", default="8,12,16,20,24,32,64")
    ap.add_argument("--english", action="store_true")
    ap.add_argument("--no-oracle", action="store_true")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    args = ap.parse_args()
    import torch
    fbuild.build()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fh = open(args.out, "a")

    def emit(rec):
        print(json.dumps(rec), flush=True)
        fh.write(json.dumps(rec) + "\n")
        fh.flush()
    lens = [int(x) for x in args.lens.split(",")]
    from oracle import fm_oracle as fo
    fo.build(native=True)

    t0 = time.time()
    text = synth.uniform_bytes(args.n, 2)
    base = "/tmp/fmx_bench_cfg2_%d" % args.n
    if not os.path.exists(base + ".bwt"):
        fx.build_index_files(text, base, bigEndian=True)
    orc = None if args.no_oracle else fo.OracleIndex.load(base)
    g = fx.GpuFMSearcher(base + ".bwt")
    emit(dict(what="open", index="cfg2 auto", secs=time.time() - t0, **g.info()))
    for rec in sweep(g, torch, text, lens, args.queries, 3, orc, label="cfg2 auto"):
        emit(rec)
    g.set_accel_mask(fx.ACCEL_NONE)
    for rec in sweep(g, torch, text, [16], args.queries, 3, orc, label="cfg2 planes, no accelerators"):
        emit(rec)
    g.close()
    g = fx.GpuFMSearcher(base + ".bwt", layout=fx.LAYOUT_WM, accel=fx.ACCEL_NONE)
    emit(dict(what="open", index="cfg2 wm none", **g.info()))
    for rec in sweep(g, torch, text, [16], args.queries // 4, 3, orc, label="cfg2 wavelet matrix, no accelerators"):
        emit(rec)
    g.close()
    if orc is not None:
        orc.close()
    if args.english:
        from tools.bench_configs import english_like
        t0 = time.time()
        text = english_like(args.n)
        t1 = time.time()
        base = "/tmp/fmx_bench_cfg3_%d" % args.n
        if not os.path.exists(base + ".bwt"):
            fx.build_index_files(text, base, bigEndian=True)
        t2 = time.time()
        orc = None if args.no_oracle else fo.OracleIndex.load(base)
        g = fx.GpuFMSearcher(base + ".bwt", sa_sample_rate=32)
        emit(dict(what="open", index="cfg3 english auto", gen_s=t1 - t0, build_s=t2 - t1, open_s=time.time() - t2, **g.info()))
        alpha = np.unique(text[:10_000_000])
        for rec in sweep(g, torch, text, [8, 12, 16, 24, 32], args.queries, 5, orc, hit_frac=1.0, alphabet=alpha, label="cfg3 english auto"):
            emit(rec)
        g.set_accel_mask(fx.ACCEL_NONE)
        for rec in sweep(g, torch, text, [12], args.queries, 5, orc, hit_frac=1.0, alphabet=alpha, label="cfg3 english planes, no accelerators"):
            emit(rec)
        g.close()


if __name__ == "__main__":
    main()
