#!/usr/bin/env python
"""
bench.py — FM-index count throughput on B200 (BASELINE.json metric: FM count queries/s, len-16, 1 GB text).

    python bench.py --gpus N --steps K --warmup W                (this repo's CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference algorithm on the host CPU cores)

Workload (BASELINE.json configs[1], SURVEY.md §8d cfg 2): 10^9 bytes i.i.d. uniform over 1..255 (seed 2), indexed as
the reference does (reverse(text)+'$', .bwt/.aux files); per GPU 10 M len-16 patterns, 90 % substrings of the text
(reversed, as search() consumes them) and 10 % uniform random bytes (seed 3).  A "step" is one pass of the batch
through the count kernel.  With N GPUs the index is replicated, every rank owns its own 10 M-query shard (weak
scaling) and the per-query counts are all-gathered over NCCL; time is the max over ranks of CUDA-event time.

One JSON line on stdout (rank 0).  Everything else goes to stderr.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OUT = sys.stdout
METRIC = "fm_count_queries_per_s_len16_1GB_text"
UNIT = "queries/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ workload
def make_text(n, workload="cfg2"):
    if workload == "cfg5":                                  # 4e9 bytes i.i.d. uniform over ACGT, seed 7 (SURVEY §8d cfg 5)
        return np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(7).integers(0, 4, n, dtype=np.uint8)]
    return np.random.default_rng(2).integers(1, 256, n, dtype=np.uint8)


def make_queries(text, m, ln, seed, rank, out=None, workload="cfg2"):
    """90 % hits (reversed substrings at uniform offsets), 10 % uniform random symbols; shuffled."""
    rng = np.random.default_rng([seed, rank])
    nh = int(m * 0.9)
    pats = out if out is not None else np.empty((m, ln), np.uint8)
    offs = rng.integers(0, len(text) - ln, nh)
    idx = offs[:, None] + np.arange(ln - 1, -1, -1)[None, :]
    hits = text[idx]
    if workload == "cfg5":
        rnd = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (m - nh, ln), dtype=np.uint8)]
    else:
        rnd = rng.integers(1, 256, (m - nh, ln), dtype=np.uint8)
    perm = rng.permutation(m)
    is_hit = np.zeros(m, bool)
    is_hit[:nh] = True
    allp = np.concatenate([hits, rnd])
    pats[:] = allp[perm]
    return pats, is_hit[perm], np.concatenate([offs, np.full(m - nh, -1)])[perm]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 50 ms from before the warm-up until after the timed region;
    the summary uses the samples that fall inside the timed window (all samples under load if the window is shorter
    than the sampling period)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.proc = None
        self.window = [None, None]

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception as e:                       # nvidia-smi missing: report that, do not fail the bench
            log("clock sampler unavailable:", e)
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.window[0] = time.time()

    def mark_end(self):
        self.window[1] = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                parsed.append((ts, float(f[0]), float(f[1]), float(f[2]), [nme for nme, v in zip(names, f[3:7]) if v == "Active"]))
            except ValueError:
                continue
        inside = [p for p in parsed if self.window[0] is not None and self.window[0] <= p[0] <= (self.window[1] or 1e18) + 0.06]
        use = inside if inside else parsed
        reasons = sorted({x for p in use for x in p[4]})
        return {"sm_mhz": float(np.median([p[1] for p in use])) if use else None, "sm_max_mhz": max([p[2] for p in use]) if use else None,
                "power_w_max": max([p[3] for p in use]) if use else None, "reasons": reasons, "samples": len(use),
                "samples_in_timed_window": len(inside)}


def index_base(n, workload="cfg2"):
    return "/tmp/fmx_bench_%s_%d" % (workload, n)


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: NaiveFMSearcher.occ = binary search in the in-memory .fm
    array (bwtmerger.scala:354-375) driving SuffixAlgo.search (findex.scala:15-31) — the C restatement under oracle/
    (no JVM exists in this image), all host threads, a bounded sample of the same workload per step."""
    if rank != 0:
        return
    from oracle import fm_oracle as fo
    cores = os.cpu_count() or 1
    n, m, ln = args.text_bytes, args.queries, args.len
    t0 = time.time()
    text = make_text(n, args.workload)
    base = index_base(n, args.workload)
    if not (os.path.exists(base + ".bwt") and os.path.exists(base + ".aux")):
        # index construction is setup, not the measured path; the device suffix sorter only writes the reference's files
        from findex_b200 import build as fbuild, fmindex as fx
        fbuild.build()
        try:
            fx.build_index_files(text, base, bigEndian=True)
        except fx.FmxError:
            # no CUDA device (CPU-only check of this arm): the oracle's own builder writes the same files; small texts only
            if n > 50_000_000:
                raise
            bwt, eof, cnt = fo.build_bwt(fo.file_to_text_rev(text.tobytes()))
            fo.write_index_files(base, bwt, eof, cnt, big_endian=True)
    fo.build(native=True)
    ix = fo.OracleIndex.load(base)
    log("reference arm: index loaded + fm array materialised in %.1f s" % (time.time() - t0))
    pats, _, _ = make_queries(text, m, ln, 3, 0, workload=args.workload)
    probe = 20000
    t1 = time.time()
    ix.count_batch(pats[:probe].reshape(-1), np.arange(0, probe * ln + 1, ln, dtype=np.int64), threads=cores)
    rate = probe / max(time.time() - t1, 1e-6)
    budget_s = 120.0 / max(args.steps + args.warmup, 1)
    sample = int(min(m, max(probe, rate * min(budget_s, 8.0))))
    off = np.arange(0, sample * ln + 1, ln, dtype=np.int64)
    flat = pats[:sample].reshape(-1)
    for _ in range(args.warmup):
        ix.count_batch(flat, off, threads=cores)
    t1 = time.time()
    for _ in range(args.steps):
        ix.count_batch(flat, off, threads=cores)
    dt = time.time() - t1
    v = sample * args.steps / dt
    sdesc = "first %d of the %d-query batch per step, %d threads" % (sample, m, cores)
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
           "data": "synthetic", "config": workload_config(args),
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sdesc,
                            "note": "JVM unavailable - C restatement of the reference algorithm (binary search in .fm)"},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(out), file=OUT, flush=True)


def workload_config(args):
    what = ("cfg5: %d-byte uniform DNA text (ACGT, seed 7)" % args.text_bytes) if args.workload == "cfg5" else \
        ("cfg2: %d-byte uniform text over bytes 1..255 (seed 2)" % args.text_bytes)
    return {"workload": "%s, %d len-%d count queries per GPU, 90%% hits / 10%% random (seed 3)" % (what, args.queries, args.len),
            "text_bytes": args.text_bytes, "queries_per_gpu": args.queries, "pattern_len": args.len, "parallelism": "dp%d (index replicated, "
            "queries sharded)" % args.gpus, "l2": "inputs (160 MB patterns) and index (GBs) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from findex_b200 import build as fbuild, fmindex as fx
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n, m, ln = args.text_bytes, args.queries, args.len
    if rank == 0:
        fbuild.build()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    text = make_text(n, args.workload)
    base = index_base(n, args.workload)
    if rank == 0 and not (os.path.exists(base + ".bwt") and os.path.exists(base + ".aux")):
        tb = time.time()
        fx.build_index_files(text, base, bigEndian=True)
        log("index files built on the GPU in %.1f s" % (time.time() - tb))
    if world > 1:
        dist.barrier()
    layout = {"auto": fx.LAYOUT_AUTO, "wm": fx.LAYOUT_WM, "planes": fx.LAYOUT_PLANES}[args.layout]
    accel = {"auto": fx.ACCEL_AUTO, "none": fx.ACCEL_NONE, "kmer": fx.ACCEL_KMER, "text": fx.ACCEL_TEXT, "both": fx.ACCEL_KMER | fx.ACCEL_TEXT,
             "ctx": fx.ACCEL_KMER | fx.ACCEL_CTX}[args.accel]
    g = fx.GpuFMSearcher(base + ".bwt", bigEndian=True, device=local_rank, layout=layout, lanes_per_query=args.lanes, accel=accel)
    if args.chunk:
        g.set_chunk(args.chunk)
    info = g.info()
    log("rank %d: index open (%s, %.2f GB on device) after %.1f s" % (rank, info["layout"], info["index_bytes"] / 1e9, time.time() - t0))

    h_pat = fx.PinnedArray((m, ln), np.uint8)
    h_sp = fx.PinnedArray((m,), np.int64)
    h_ep = fx.PinnedArray((m,), np.int64)
    pats, is_hit, offs = make_queries(text, m, ln, 3, rank, out=h_pat.array, workload=args.workload)
    d_pat = torch.from_numpy(pats).to(dev)
    d_sp = torch.zeros(m, dtype=torch.int32, device=dev)
    d_ep = torch.zeros(m, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(m, dtype=torch.int32, device=dev)
    d_all = torch.zeros(m * world, dtype=torch.int32, device=dev) if world > 1 else None
    stream = torch.cuda.current_stream().cuda_stream

    # N > 1: the batch is cut into chunks; the NCCL all-gather of chunk c's hit counts (the one exchange step of the path) runs on a
    # side stream while the count kernel works on chunk c+1, so only the last chunk's gather is exposed.
    nch = max(1, args.gather_chunks) if world > 1 else 1
    csz = (m + nch - 1) // nch
    bounds = [(c * csz, min(m, (c + 1) * csz)) for c in range(nch)]
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    # two result sets, alternated per step, so that step k+1's kernel never waits for step k's gather to finish reading
    sets = [dict(sp=d_sp, ep=d_ep, cnt=d_cnt)]
    if world > 1:
        sets.append(dict(sp=torch.zeros_like(d_sp), ep=torch.zeros_like(d_ep), cnt=torch.zeros_like(d_cnt)))
        for st_ in sets:
            st_["all"] = [torch.zeros((hi - lo) * world, dtype=torch.int32, device=dev) for lo, hi in bounds]
            st_["ev_k"] = [torch.cuda.Event() for _ in range(nch)]
            st_["ev_c"] = [torch.cuda.Event() for _ in range(nch)]
    counter = [0]

    p2p = world > 1 and args.exchange == "p2p"
    if p2p:
        # fused compute + exchange: every rank maps every rank's gathered buffer (CUDA IPC over NVLink/NVSwitch peer memory) and the
        # count kernel stores its hit counts straight into all of them; NCCL is left with one 4-byte all-reduce per step as the
        # cross-rank completion barrier.  Two buffer sets alternate so step k+1 never writes what step k's consumers read.
        gath = [fx.SharedDeviceBuffer(m * world) for _ in range(2)]
        mine = [gb.export_handle() for gb in gath]
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        sinks = [[gath[b].ptr if r == rank else gath[b].import_peer(r, everyone[r][b]) for r in range(world)] for b in range(2)]
        flag = torch.zeros(1, dtype=torch.int32, device=dev)
        pev_k = [torch.cuda.Event() for _ in range(2)]
        pev_c = [torch.cuda.Event() for _ in range(2)]

    def step():
        if world == 1:
            g.count_fixed_dev(d_pat.data_ptr(), ln, m, d_sp.data_ptr(), d_ep.data_ptr(), stream)
            return
        cur = torch.cuda.current_stream()
        if p2p:
            b = counter[0] & 1
            counter[0] += 1
            cur.wait_event(pev_c[b])                    # the barrier of the step that last wrote this buffer set has passed
            g.count_fixed_dev_gather(d_pat.data_ptr(), ln, m, d_sp.data_ptr(), d_ep.data_ptr(), sinks[b], rank * m, cur.cuda_stream)
            pev_k[b].record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(pev_k[b])
                dist.all_reduce(flag)                   # after this, every rank's stores of this step have landed everywhere
                pev_c[b].record(comm)
            return
        b = sets[counter[0] & 1]
        counter[0] += 1
        for c, (lo, hi) in enumerate(bounds):
            cur.wait_event(b["ev_c"][c])                # the gather that last used this result set has read it (two steps ago)
            g.count_fixed_dev(d_pat.data_ptr() + lo * ln, ln, hi - lo, b["sp"].data_ptr() + lo * 4, b["ep"].data_ptr() + lo * 4, stream)
            b["ev_k"][c].record(cur)
            if args.diag == "nocomm":
                continue
            with torch.cuda.stream(comm):
                comm.wait_event(b["ev_k"][c])
                if args.diag != "nosub":
                    torch.sub(b["ep"][lo:hi], b["sp"][lo:hi], out=b["cnt"][lo:hi])
                if args.diag != "nogather":
                    dist.all_gather_into_tensor(b["all"][c], b["cnt"][lo:hi])
                b["ev_c"][c].record(comm)

    def drain():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm)

    # ---- sanity at full size (parity proper lives in tests/): hits are found, a few are verified by brute force
    step()
    drain()
    torch.cuda.synchronize()
    if world > 1:                                       # every rank holds every rank's counts after the exchange
        local = (d_ep - d_sp)
        if p2p:
            dist.barrier()
            got = torch.from_numpy(gath[0].to_host().astype(np.int64)).to(dev)
            assert torch.equal(got[rank * m:(rank + 1) * m], local.to(torch.int64) & 0xFFFFFFFF), "gathered counts differ from the local ones"
            cs = (got * torch.arange(1, got.numel() + 1, device=dev) % 1000003).sum().reshape(1)
            lo_, hi_ = cs.clone(), cs.clone()
            dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
            assert int(lo_) == int(hi_), "ranks disagree on the gathered counts"
        else:
            lo, hi = bounds[0]
            mine_ = sets[0]["all"][0].view(world, hi - lo)[rank]
            assert args.diag != "none" or torch.equal(mine_, local[lo:hi]), "all-gathered counts differ from the local ones"
    sp = d_sp.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    ep = d_ep.cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    cnt = ep - sp
    assert (cnt[is_hit] >= 1).all(), "a text substring was not found"
    if rank == 0:
        tb = text.tobytes() if n <= 1_500_000_000 else text[:1_000_000_000].tobytes()
        for q in (np.flatnonzero(is_hit)[:2] if n <= 1_500_000_000 else []):
            needle = pats[q][::-1].tobytes()
            assert tb.count(needle) == cnt[q], "count mismatch vs brute force"
        del tb
    checksum = int((sp * 1315423911 + ep * 2654435761).sum() & 0xFFFFFFFFFFFF)

    # ---- roofline inputs, outside the timed region
    blocks, steps_exec = g.count_fixed_stats(pats)
    alg_bytes = blocks * 64
    r_rand, _ = g.gather_bench(64, 4, 1 << 25, 16, 3)

    def timed_region(fn, label):
        sampler = ClockSampler(local_rank)
        sampler.start()
        for _ in range(args.warmup):
            fn()
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.mark_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall = time.time()
        e0.record()
        for _ in range(args.steps):
            fn()
        drain()
        e1.record()
        torch.cuda.synchronize()
        wall = time.time() - t_wall
        sampler.mark_end()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        ms = e0.elapsed_time(e1)
        if label == "e2e":
            ms = wall * 1e3                              # host API: the call returns when results are in host memory
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), clocks

    ms_total, clocks = timed_region(step, "device")
    # kernel-only time (same launches, no exchange) for the roofline
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        a.record()
        g.count_fixed_dev(d_pat.data_ptr(), ln, m, d_sp.data_ptr(), d_ep.data_ptr(), stream)
        b.record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    # end to end through the host API: (sp, ep) per query at the reference's own width (Option[(Int, Int)], findex.scala:15-31) ...
    narrow = g.n < 2 ** 31                                  # cfg 5 (n = 4e9) does not fit Int rows: int64 there
    h_sp32 = fx.PinnedArray((m,), np.int32 if narrow else np.int64)
    h_ep32 = fx.PinnedArray((m,), np.int32 if narrow else np.int64)

    def e2e_step():
        g.count_fixed_into(h_pat.array, h_sp32.array, h_ep32.array)

    ms_e2e, clocks_e2e = timed_region(e2e_step, "e2e")
    assert np.array_equal(h_sp32.array, np.where(cnt > 0, sp, 0)) and np.array_equal(h_ep32.array, np.where(cnt > 0, ep, 0)), "host API result differs from device API"

    # ... and with int64 rows (twice the result bytes over PCIe)
    def e2e64_step():
        g.count_fixed_into(h_pat.array, h_sp.array, h_ep.array)

    ms_e2e64, _ = timed_region(e2e64_step, "e2e")
    assert np.array_equal(h_sp.array, np.where(cnt > 0, sp, 0)) and np.array_equal(h_ep.array, np.where(cnt > 0, ep, 0)), "host API (int64) result differs from device API"

    # the same batch through the count-only call (uint32 ep-sp per query: 4 instead of 16 result bytes over PCIe)
    h_cnt = fx.PinnedArray((m,), np.uint32)

    def e2e_count_only_step():
        g.count_only_fixed(h_pat.array, out=h_cnt.array)

    ms_e2e_co, _ = timed_region(e2e_count_only_step, "e2e")
    assert np.array_equal(h_cnt.array.astype(np.int64), np.where(cnt > 0, cnt, 0)), "count-only API differs from ep-sp"

    regex = None if args.regexes <= 0 else regex_leg(args, g, text, world, rank, dev)

    value = world * m * args.steps / (ms_total * 1e-3)
    e2e = world * m * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")          # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            rec = json.load(open(tp)).get("count_fixed_kernel", {})
            if (rec.get("queries") == m and rec.get("layout") == info["layout"] and rec.get("lanes") == info["lanes_per_query"]
                    and rec.get("kmer_k") == info["kmer_k"] and rec.get("text_shortcut") == info["text_shortcut"]
                    and rec.get("ctx_depth", 0) == info["ctx_depth"]):
                traffic = rec.get("dram_bytes_per_launch")
        except Exception:
            pass
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
           "data": "synthetic", "config": dict(workload_config(args), exchange=("none" if world == 1 else ("fused peer stores + 4-byte NCCL barrier" if p2p else "NCCL all_gather_into_tensor on a side stream")), layout=info["layout"], lanes_per_query=info["lanes_per_query"],
                                               index_bytes=info["index_bytes"], kmer_k=info["kmer_k"], text_shortcut=info["text_shortcut"], ctx_depth=info["ctx_depth"], checksum=checksum),
           "clocks": clocks,
           "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": m * ln, "d2h_bytes_per_step": m * (8 if narrow else 16), "ms_per_step": ms_e2e / args.steps,
                   "api": "fmx_count_fixed_i32 (host pinned buffers in/out, (sp, ep) as the reference's 32-bit Int rows)" if narrow else "fmx_count_fixed (host pinned buffers in/out, int64 sp/ep)",
                   "clocks": clocks_e2e,
                   "int64_rows": {"value": world * m * args.steps / (ms_e2e64 * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e64 / args.steps,
                                  "h2d_bytes_per_step": m * ln, "d2h_bytes_per_step": m * 16, "api": "fmx_count_fixed (int64 sp/ep)"},
                   "count_only": {"value": world * m * args.steps / (ms_e2e_co * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e_co / args.steps,
                                  "h2d_bytes_per_step": m * ln, "d2h_bytes_per_step": m * 4, "api": "fmx_count_only_fixed (uint32 ep-sp)"}},
           "gpu_launches": args.steps * (1 if (world == 1 or p2p) else nch),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "peak_source": peak_src, "kernel": "count_fixed_kernel", "kernel_ms": k_ms,
                        "algorithmic_bytes_per_launch": alg_bytes, "distinct_blocks_per_query": blocks / m, "executed_steps_per_query": steps_exec / m,
                        "r_rand_gbs": r_rand, "frac_of_r_rand": achieved / r_rand,
                        "note": "achieved = distinct 64-B rank blocks the batch touches x 64 B / kernel time; r_rand = live K4 random 64-B gather "
                                "bandwidth over the same index (the random-sector HBM roofline of north_star)"}}
    if regex is not None:
        out["regex"] = regex["report"]
    if world == 1 and rank == 0 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(args, base, pats, sp, ep, cnt, regex)
    if rank == 0:
        print(json.dumps(out), file=OUT, flush=True)
    if p2p:
        torch.cuda.synchronize()
        dist.barrier()
        for gb in gath:
            for ptr in list(gb.peers.values()):
                fx.lib().fmx_ipc_close(ptr)
            gb.peers = {}
        dist.barrier()
        for gb in gath:
            gb.close()
    g.close()


def regex_leg(args, g, text, world, rank, dev):
    """Second half of BASELINE.json's metric ("regex queries/s"): `--regexes` template regexes (cfg 4: classes, alternation,
    desugared bounded repeats, \\d, '.') per GPU over the same index, searched through fmx_regex_search_batch with host
    buffers.  Compilation (host, once) is outside the timed region; the search call is timed end to end (wall clock, max over ranks)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from findex_b200 import fmindex as fx, synth
    rxs = synth.regex_templates(text, np.random.default_rng([6, rank]), args.regexes)
    t0 = time.time()
    trees, kept = [], []
    for r in rxs:
        try:
            trees.append(fx.ReTree(r))
            kept.append(r)
        except fx.FmxError:
            pass
    compile_s = time.time() - t0
    mr = len(trees)
    t0 = time.time()
    rset = g.regex_set(trees)                               # concatenated automata, uploaded once (compile once, search many times)
    upload_s = time.time() - t0
    cap = 1 << 22
    off = np.zeros(mr + 1, np.int64)
    ln_, sp_, ep_ = np.zeros(cap, np.int32), np.zeros(cap, np.int64), np.zeros(cap, np.int64)

    def call():
        rc = fx.lib().fmx_regex_set_search(g.h, rset.h, cap, off.ctypes.data_as(C.c_void_p), ln_.ctypes.data_as(C.c_void_p),
                                           sp_.ctypes.data_as(C.c_void_p), ep_.ctypes.data_as(C.c_void_p))
        assert rc == 0, fx.lib().fmx_last_error()
    steps = max(3, min(args.steps, 20))
    for _ in range(3):
        call()
    if world > 1:
        dist.barrier()
    t0 = time.time()
    kms = []
    for _ in range(steps):
        call()
        kms.append(g.last_kernel_ms())
    wall = time.time() - t0
    tt = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    wall = float(tt.item())
    total = int(off[mr])
    res = [(kept[i], sorted(zip(ln_[off[i]:off[i + 1]].tolist(), sp_[off[i]:off[i + 1]].tolist(), ep_[off[i]:off[i + 1]].tolist())))
           for i in np.random.default_rng(9).choice(mr, min(200, mr), replace=False)]
    return {"sample": res,
            "report": {"value": world * mr * steps / wall, "unit": "regexes/s", "what": "fmx_regex_set_search end to end (device-resident regex set, host result buffers), "
                       "Glushkov engine, caps off; device time alone in device_value", "device_value": world * mr / (float(np.mean(kms)) * 1e-3), "regexes_per_gpu": mr,
                       "rejected_by_compiler": len(rxs) - mr, "steps": steps, "ms_per_step": wall / steps * 1e3, "kernel_ms_per_step": float(np.mean(kms)),
                       "traversal_launches_per_step": int(g.last_kernel_launches()), "levels_per_step": int(g.last_regex_levels()), "result_triples": total, "compile_s_once": compile_s, "set_upload_s_once": upload_s}}


def cpu_baseline(args, base, pats, sp, ep, cnt, regex=None):
    """The oracle (C restatement of the reference algorithm) on the box's host cores, on a bounded sample of the same batch;
    doubles as the full-size parity check of that sample."""
    from oracle import fm_oracle as fo
    cores = os.cpu_count() or 1
    t0 = time.time()
    fo.build(native=True)
    ix = fo.OracleIndex.load(base)
    log("cpu_baseline: oracle index ready in %.1f s" % (time.time() - t0))
    ln = args.len
    probe = 20000
    t1 = time.time()
    ix.count_batch(pats[:probe].reshape(-1), np.arange(0, probe * ln + 1, ln, dtype=np.int64), threads=cores)
    rate = probe / max(time.time() - t1, 1e-6)
    sample = int(min(len(pats), max(probe, rate * 15.0)))
    off = np.arange(0, sample * ln + 1, ln, dtype=np.int64)
    t1 = time.time()
    osp, oep = ix.count_batch(pats[:sample].reshape(-1), off, threads=cores)
    dt = time.time() - t1
    want_sp = np.where(cnt[:sample] > 0, sp[:sample], 0)
    want_ep = np.where(cnt[:sample] > 0, ep[:sample], 0)
    parity = bool(np.array_equal(osp, want_sp) and np.array_equal(oep, want_ep))
    assert parity, "GPU (sp,ep) differ from the oracle on the CPU-baseline sample"
    # single-thread figure: the same algorithm on one core, about two seconds of it
    one = int(min(sample, max(2000, rate / max(cores, 1) * 2.0)))
    t1 = time.time()
    ix.count_batch(pats[:one].reshape(-1), off[:one + 1], threads=1)
    dt1 = time.time() - t1
    extra = {"single_thread": {"value": one / max(dt1, 1e-9), "unit": UNIT, "cores": 1, "sample": "first %d queries, %.1f s" % (one, dt1)}}
    if regex is not None:                                 # the oracle's uncapped ReTree._matchSA on a sample of the regex batch (1 thread)
        t1 = time.time()
        ok = all(ix.regex_match(rx, max_expansions=50_000_000) == want for rx, want in regex["sample"])
        dtr = time.time() - t1
        assert ok, "GPU regex results differ from the oracle on the sample"
        extra["regex"] = {"value": len(regex["sample"]) / dtr, "unit": "regexes/s", "cores": 1, "sample": "%d regexes of the batch" % len(regex["sample"]),
                          "parity_on_sample": bool(ok)}
    ix.close()
    return dict({"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                 "sample": "first %d of the %d-query batch, %d threads, %.1f s" % (sample, len(pats), cores, dt),
                 "parity_on_sample": parity, "note": "JVM unavailable - C restatement of the reference algorithm (binary search in .fm)"}, **extra)


def protect_stdout():
    """Libraries (NCCL prints its version) may write to fd 1; the contract is ONE JSON line on stdout.  Point fd 1 at stderr and
    keep the real stdout for the final line."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and allocate its pinned buffers) on the NUMA node its GPU hangs off."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local_rank), "pci_domain_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:00.0/numa_node" % (dom, bus)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception as e:
        log("numa binding skipped:", e)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"], help="cfg2 = the metric's config (default); cfg5 = 4 GB DNA, len-32")
    ap.add_argument("--text-bytes", type=int, default=None)
    ap.add_argument("--queries", type=int, default=None)
    ap.add_argument("--len", type=int, default=None)
    ap.add_argument("--layout", default="auto", choices=["auto", "wm", "planes"])
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--accel", default="auto", choices=["auto", "none", "kmer", "text", "both", "ctx"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer-memory stores (default) or NCCL all-gather")
    ap.add_argument("--gather-chunks", type=int, default=1)
    ap.add_argument("--diag", default="none", choices=["none", "nocomm", "nosub", "nogather"], help="diagnostics only: drop parts of the exchange")
    ap.add_argument("--regexes", type=int, default=100_000, help="regexes per GPU for the secondary regex measurement (0 = skip)")
    ap.add_argument("--chunk", type=int, default=0, help="queries per pipeline chunk of the host-buffer calls (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    dflt = {"cfg2": (1_000_000_000, 10_000_000, 16), "cfg5": (4_000_000_000, 12_500_000, 32)}[args.workload]
    args.text_bytes = args.text_bytes or dflt[0]
    args.queries = args.queries or dflt[1]
    args.len = args.len or dflt[2]
    if args.workload == "cfg5":
        args.regexes = 0
        global METRIC
        METRIC = "fm_count_queries_per_s_len32_4GB_dna_text"
    if args.warmup < 3:
        log("note: contract asks for >= 3 warm-up steps; got", args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global OUT
    OUT = protect_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        node = bind_to_gpu_numa_node(local_rank)
        log("rank %d: bound to NUMA node %s" % (rank, node))
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
