"""torchrun worker of tests/test_gpu_exchange.py::test_exchange_two_gpus_torchrun — also runnable by hand:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_worker.py /tmp/ok
Every rank opens the replicated index, takes its shard of a count+locate batch and of a regex batch, exchanges through GpuExchange
(kernel stores into CUDA-IPC peer buffers + a 4-byte NCCL barrier) and checks the gathered result against the single-GPU answer."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from findex_b200 import fmindex as fx  # noqa: E402
from findex_b200 import sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ref = os.path.join(ROOT, "tests", "golden", "ref")
    text = open(os.path.join(ref, "test.txt"), "rb").read()
    g = fx.GpuFMSearcher(os.path.join(ref, "test.cmp.bwt"), bigEndian=False, device=local, sa_sample_rate=8)
    rng = np.random.default_rng(12)
    M, ln = 5000, 2
    offs = rng.integers(0, len(text) - ln, M)
    tarr = np.frombuffer(text, np.uint8)
    pats = np.stack([tarr[s:s + ln][::-1] for s in offs]).copy()
    pats[::7] = rng.integers(97, 123, (len(pats[::7]), ln), dtype=np.uint8)
    sp, ep = g.count_fixed(pats)                              # the single-GPU answer for the whole batch
    off1, pos1 = g.locate_batch(sp, ep)
    total = int(off1[-1])
    ex = sharded.GpuExchange(g, rank, world, M, total + 16, dev)
    lo, hi = sharded.shard_bounds(M, rank, world)
    d_pat = torch.from_numpy(pats[lo:hi]).to(dev)
    for _ in range(3):                                        # repeated steps reuse the buffers
        off, _ = ex.locate(d_pat, ln, lo, hi, total + 16)
        torch.cuda.synchronize()
        assert np.array_equal(off.cpu().numpy(), off1)
        got = ex.gathered_values(total).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
        assert np.array_equal(got, pos1), "rank %d: gathered positions differ from the single-GPU answer" % rank
    rxs = ["%c.%c" % (97 + i % 26, 97 + (i // 26) % 26) for i in range(600)] + ["ab?c", "q(u|a)", "zzzz"]
    Mr = len(rxs)
    want = g.regex_search_batch([fx.ReTree(x) for x in rxs])
    nres = sum(len(w) for w in want)
    ex2 = sharded.GpuExchange(g, rank, world, Mr, 4 * nres + 16, dev)
    lo, hi = sharded.shard_bounds(Mr, rank, world)
    rset = g.regex_set([fx.ReTree(x) for x in rxs[lo:hi]])
    off, _ = ex2.regex(rset, lo, hi, nres + 4)
    torch.cuda.synchronize()
    assert np.array_equal(np.diff(off.cpu().numpy()), [len(w) for w in want])
    rec = ex2.gathered_values(4 * nres).cpu().numpy().reshape(-1, 4).astype(np.int64) & 0xFFFFFFFF
    assert [tuple(x) for x in rec.tolist()] == [(i, l, s, e) for i, w in enumerate(want) for (l, s, e) in w]
    dist.barrier()
    rset.close()
    ex.close_peers()
    ex2.close_peers()
    dist.barrier()
    ex.close()
    ex2.close()
    g.close()
    if rank == 0:
        open(sys.argv[1], "w").write("ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
