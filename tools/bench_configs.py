"""
tools/bench_configs.py — the non-headline BASELINE configs on a real B200 (run under gpurun):

  cfg3  1e9-byte English-like text (Zipf over the words.txt vocabulary), 1 M len-12 locate queries, SA sample rate 32
        (and the full-SA accelerator for comparison); every located position is verified against the text.
  cfg4  100 k generated regexes (classes, alternation, desugared bounded repeats, \\d, .) over the cfg-3 index;
        every result of a sample is checked for soundness (bit-exact parity vs the oracle is in tests/).
  cfg5s 4e9/--dna-scale-byte DNA text, len-32 count queries on one GPU (the 8-GPU run is bench.py --gpus 8 territory).

Writes JSON lines to gpurun_out/configs.jsonl.  Evidence for DESIGN.md; bench.py remains the contract benchmark (cfg 2).
"""
import argparse
import json
import lzma
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findex_b200 import build as fbuild  # noqa: E402
from findex_b200 import fmindex as fx  # noqa: E402
from findex_b200 import synth  # noqa: E402


def vocabulary():
    """words.txt vocabulary recovered from the committed words.bwt fixture through the GPU searcher itself
    (prevSubstr(eof, n) walks the whole file: '\\0' + file bytes, T/Indexer.scala:1120)."""
    d = "/tmp/fmx_words"
    os.makedirs(d, exist_ok=True)
    with lzma.open(os.path.join(ROOT, "tests", "golden", "ref", "words.bwt.xz"), "rb") as f:
        open(os.path.join(d, "words.bwt"), "wb").write(f.read())
    open(os.path.join(d, "words.aux"), "wb").write(open(os.path.join(ROOT, "tests", "golden", "ref", "words.aux"), "rb").read())
    g = fx.GpuFMSearcher(os.path.join(d, "words.bwt"), accel=fx.ACCEL_NONE)
    text = g.prevSubstr(g.eof, g.n)[1:]
    g.close()
    return [w for w in text.split(b"\r\n") if w]


def english_like(n_bytes, seed=4):
    return synth.english_like(vocabulary(), n_bytes, seed)


def emit(fh, **kw):
    print(json.dumps(kw), flush=True)
    fh.write(json.dumps(kw) + "\n")
    fh.flush()


regex_templates = synth.regex_templates


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000_000)
    ap.add_argument("--locate-queries", type=int, default=1_000_000)
    ap.add_argument("--regexes", type=int, default=100_000)
    ap.add_argument("--oracle-sample", type=int, default=300)
    ap.add_argument("--dna-n", type=int, default=0)
    ap.add_argument("--out", default="gpurun_out/configs.jsonl")
    args = ap.parse_args()
    fbuild.build()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fh = open(args.out, "a")

    # ------------------------------------------------------------------ cfg 3: text + index
    t0 = time.time()
    text = english_like(args.n)
    t1 = time.time()
    base = "/tmp/fmx_cfg3_%d" % args.n
    fx.build_index_files(text, base, bigEndian=True)
    t2 = time.time()
    emit(fh, what="cfg3_build", n=args.n, gen_s=t1 - t0, build_files_s=t2 - t1, sigma=int(len(np.unique(text[:50_000_000]))))

    rng = np.random.default_rng(5)
    m, ln = args.locate_queries, 12
    offs = rng.integers(0, args.n - ln, m)
    pats = text[offs[:, None] + np.arange(ln - 1, -1, -1)[None, :]]           # reversed substrings (what search() consumes)
    for label, kw in (("rate32", dict(sa_sample_rate=32, accel=fx.ACCEL_KMER)), ("fullsa", dict(sa_sample_rate=0, accel=fx.ACCEL_AUTO))):
        t0 = time.time()
        g = fx.GpuFMSearcher(base + ".bwt", **kw)
        info = g.info()
        emit(fh, what="cfg3_open", mode=label, open_s=time.time() - t0, **info)
        sp, ep = g.count_fixed(pats)
        count_ms = g.last_kernel_ms()
        occ = ep - sp
        assert (occ >= 1).all()
        # locate in slices of <= 2^30 occurrences
        order = np.argsort(occ, kind="stable")
        keep = order[np.cumsum(occ[order]) < (1 << 30)]                       # drop only the few patterns with astronomically many hits
        keep.sort()
        t0 = time.time()
        off, pos = g.locate_batch(sp[keep], ep[keep])
        wall = time.time() - t0
        k_ms = g.last_kernel_ms()
        total = int(off[-1])
        # size-independent parity: every position is a real occurrence, positions ascending and distinct, count = ep - sp
        n1 = g.n
        chk = rng.choice(len(keep), min(20000, len(keep)), replace=False)
        bad = 0
        for j in chk:
            q = pos[off[j]:off[j + 1]]
            if len(q) != occ[keep[j]] or (len(q) > 1 and not (np.diff(q) > 0).all()):
                bad += 1
                continue
            qq = q[:50]
            fo_ = (n1 - 1) - qq - ln                                          # file offset = (n-1) - pos - len
            want = pats[keep[j]][::-1]
            got = text[fo_[:, None] + np.arange(ln)[None, :]]
            if not (got == want[None, :]).all():
                bad += 1
        emit(fh, what="cfg3_locate", mode=label, queries=int(len(keep)), dropped=int(m - len(keep)), occurrences=total, count_kernel_ms=count_ms,
             locate_kernel_ms=k_ms, wall_s=wall, queries_per_s=len(keep) / (k_ms * 1e-3), positions_per_s=total / (k_ms * 1e-3),
             e2e_queries_per_s=len(keep) / wall, verified_queries=int(len(chk)), mismatches=bad)
        assert bad == 0
        if label == "fullsa":
            # ------------------------------------------------------------ cfg 4 on the same index
            rxs = regex_templates(text, np.random.default_rng(6), args.regexes)
            t0 = time.time()
            trees, kept = [], []
            for r in rxs:
                try:
                    trees.append(fx.ReTree(r))
                    kept.append(r)
                except fx.FmxError:
                    pass
            compile_s = time.time() - t0
            t0 = time.time()
            res = g.regex_search_batch(trees, cap_total=1 << 24)
            wall = time.time() - t0
            k_ms, launches = g.last_kernel_ms(), g.last_kernel_launches()
            nres = sum(len(r) for r in res)
            emit(fh, what="cfg4_regex", regexes=len(kept), rejected=len(rxs) - len(kept), compile_s=compile_s, kernel_ms=k_ms, launches=launches,
                 wall_s=wall, regexes_per_s=len(kept) / (k_ms * 1e-3), e2e_regexes_per_s=len(kept) / wall, results=nres,
                 occurrences=int(sum(e - s for r in res for _, s, e in r)))
            # size-independent soundness check (the bit-exact comparison with the oracle lives in tests/): every result
            # (len,sp,ep) must be exactly the interval of the length-len literal found at that row
            pick = np.random.default_rng(9).choice(len(kept), min(args.oracle_sample, len(kept)), replace=False)
            bad = 0
            for i in pick:
                for (ln_, s_, e_) in res[i][:5]:
                    lit = g.nextSubstr(s_, ln_)[::-1]                             # the matched string as search() consumes it
                    if len(lit) != ln_ or g.search(lit) != (s_, e_):
                        bad += 1
            emit(fh, what="cfg4_soundness", regex_sample=int(len(pick)), violations=bad)
            assert bad == 0
        g.close()

    if args.dna_n:
        rng = np.random.default_rng(7)
        text = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, args.dna_n, dtype=np.uint8)]
        base = "/tmp/fmx_cfg5_%d" % args.dna_n
        t0 = time.time()
        fx.build_index_files(text, base, bigEndian=True)
        emit(fh, what="cfg5_build", n=args.dna_n, build_files_s=time.time() - t0)
        rq = np.random.default_rng(8)
        m, ln = 12_500_000, 32
        nh = int(m * 0.9)
        offs = rq.integers(0, args.dna_n - ln, nh)
        pats = np.empty((m, ln), np.uint8)
        pats[:nh] = text[offs[:, None] + np.arange(ln - 1, -1, -1)[None, :]]
        pats[nh:] = np.frombuffer(b"ACGT", np.uint8)[rq.integers(0, 4, (m - nh, ln))]
        pats = pats[rq.permutation(m)]
        for lay in ("wm", "planes"):
            g = fx.GpuFMSearcher(base + ".bwt", layout={"wm": fx.LAYOUT_WM, "planes": fx.LAYOUT_PLANES}[lay])
            info = g.info()
            g.count_fixed(pats[:1000])
            sp, ep = g.count_fixed(pats)
            sp, ep = g.count_fixed(pats)
            ms = g.last_kernel_ms()
            blocks, steps = g.count_fixed_stats(pats)
            emit(fh, what="cfg5_count_1gpu", queries=m, pipeline_ms=ms, hits=int((ep > sp).sum()), blocks_per_query=blocks / m,
                 steps_per_query=steps / m, **info)
            g.close()


if __name__ == "__main__":
    main()
