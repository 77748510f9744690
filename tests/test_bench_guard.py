"""bench.py's RunGuard on CPU (gloo, world size 2): a leg that fails on one rank counts as failed on every rank and the later legs are
skipped everywhere (ranks stay in step); a rank stuck behind a peer that is gone is ended by the watchdog, and rank 0 still prints the
ONE JSON line with what had been measured.  (What went wrong on 8 x B200: two ranks lost a file race in a leg's set-up, the others
waited for them inside a collective until the box's time limit.)"""
import json
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


WORKER = textwrap.dedent("""
    import json, os, sys, time
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    import bench
    rank, world, mode = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), sys.argv[1]
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bench.OUT = sys.stdout
    G = bench.GUARD
    G.attach(rank, world)
    G.out = {"value": 1.0}
    flag = torch.zeros(1)

    def healthy():
        dist.all_reduce(flag)
        return {"ok": True}

    def fails_on_rank1_after_its_collectives():
        dist.all_reduce(flag)
        if rank == 1:
            raise RuntimeError("boom")
        return {"ok": True}

    def rank1_never_arrives():
        if rank == 1:
            raise RuntimeError("set-up failed before the collective")
        dist.all_reduce(flag)                      # rank 0 waits here for a partner that is gone
        return {"ok": True}

    res = {"a": bench.guarded("a", healthy)}
    if mode == "agree":
        res["b"] = bench.guarded("b", fails_on_rank1_after_its_collectives)
        res["c"] = bench.guarded("c", healthy)      # must be skipped on BOTH ranks
        print("RESULT %%d %%s" %% (rank, json.dumps(res)), flush=True)
        os._exit(0)
    else:
        res["b"] = bench.guarded("b", rank1_never_arrives, limit_s=4)
        print("RESULT %%d %%s" %% (rank, json.dumps(res)), flush=True)
        os._exit(0)
""") % ROOT


def _launch(tmp_path, mode):
    script = tmp_path / "guard_worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script), mode], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT))
    return [p.communicate(timeout=300) + (p.returncode,) for p in procs]


def test_failed_leg_is_failed_everywhere_and_later_legs_are_skipped(tmp_path):
    outs = _launch(tmp_path, "agree")
    for r, (so, se, rc) in enumerate(outs):
        assert rc == 0, se[-2000:]
        line = [ln for ln in so.splitlines() if ln.startswith("RESULT %d " % r)][0]
        res = json.loads(line.split(" ", 2)[2])
        assert res["a"] == {"ok": True}
        assert "error" in res["b"], res                     # rank 0's own run of the leg was fine, the leg still counts as failed
        assert "skipped" in res["c"], res


def test_watchdog_ends_a_rank_whose_partner_is_gone(tmp_path):
    outs = _launch(tmp_path, "watchdog")
    so0, se0, rc0 = outs[0]
    so1, se1, rc1 = outs[1]
    assert rc0 == 0 and rc1 == 0, (se0[-1500:], se1[-1500:])
    js = [ln for ln in so0.splitlines() if ln.startswith("{")]
    assert len(js) == 1, so0                                # rank 0: the partial line, printed by the watchdog
    d = json.loads(js[0])
    assert d["value"] == 1.0 and "watchdog" in d
    assert "watchdog" in se0
    assert not [ln for ln in so1.splitlines() if ln.startswith("{")]
