import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from findex_b200 import fmindex as fx, build
build.build()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
text = bench.make_text(n, "cfg3")
base = bench.index_base(n, "cfg3")
if not os.path.exists(base + ".bwt"):
    fx.build_index_files(text, base, bigEndian=True)
for accel in (fx.ACCEL_KMER, fx.ACCEL_AUTO):
    g = fx.GpuFMSearcher(base + ".bwt", sa_sample_rate=32, accel=accel)
    print("open", accel, g.info(), flush=True)
    pats, _ = bench.make_queries(text, 200_000, 12, 5, 0, workload="cfg3")
    sp, ep = g.count_fixed(pats)
    occ = ep - sp
    print("occ total", occ.sum(), "max", occ.max(), "top", np.sort(occ)[-5:], flush=True)
    for k in (1000, 10000, 50000, 200000):
        for stats in (False, True):
            g.set_stats(stats)
            d_sp = torch.from_numpy(sp[:k].astype(np.uint32).view(np.int32)).cuda()
            d_ep = torch.from_numpy(ep[:k].astype(np.uint32).view(np.int32)).cuda()
            tot = int(occ[:k].sum())
            d_off = torch.zeros(k + 1, dtype=torch.int64, device="cuda")
            d_pos = torch.zeros(tot + 16, dtype=torch.int32, device="cuda")
            try:
                t = g.locate_dev(d_sp.data_ptr(), d_ep.data_ptr(), k, d_off.data_ptr(), d_pos.data_ptr(), tot + 16, torch.cuda.current_stream().cuda_stream)
                torch.cuda.synchronize()
                print("ok", k, stats, t, g.last_locate_ms(), g.last_steps(), flush=True)
            except Exception as e:
                print("FAIL", k, stats, e, flush=True)
                sys.exit(1)
    g.close()
